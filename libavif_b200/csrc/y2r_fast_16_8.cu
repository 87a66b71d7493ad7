/* specialised YUV->RGB kernels, 16-bit-container planes -> 8-bit-container pixels (one translation unit per
 * type pair so the instantiations compile in parallel); see y2r_fast.cuh */
#include "y2r_fast.cuh"

extern "C" int abY2RFast_16_8(const AbY2RJob * jobs, uint32_t count, void * stream)
{
    return abfast::launch_y2r_fast_typed<uint16_t, uint8_t>(jobs, count, (cudaStream_t)stream);
}
