/* specialised RGB->YUV kernels, 8-bit-container pixels -> 16-bit-container planes; see r2y_fast.cuh */
#include "r2y_fast.cuh"

extern "C" int abR2YFast_8_16(const AbR2YJob * jobs, uint32_t count, void * stream)
{
    return abfast::launch_r2y_fast_typed<uint8_t, uint16_t>(jobs, count, (cudaStream_t)stream);
}
