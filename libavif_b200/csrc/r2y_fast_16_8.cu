/* specialised RGB->YUV kernels, 16-bit-container pixels -> 8-bit-container planes; see r2y_fast.cuh */
#include "r2y_fast.cuh"

extern "C" int abR2YFast_16_8(const AbR2YJob * jobs, uint32_t count, void * stream)
{
    return abfast::launch_r2y_fast_typed<uint16_t, uint8_t>(jobs, count, (cudaStream_t)stream);
}
