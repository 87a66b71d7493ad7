/* Kernel-variant experiments (selected with avifB200SetKernelPolicy bits 3-4; results in profiles/r01_variants.md): the two
 * headline 16-bit 4:4:4 configurations instantiated with the non-production store / range-check variants
 * (policy 8 -> VARIANT 0 staged, 16 -> VARIANT 1 direct only, 24 -> VARIANT 2 staged + uniform tiles). */
#include "y2r_fast.cuh"

using namespace abfast;

extern "C" int abY2RExperiment(const AbY2RJob * job, void * stream, int variant)
{
    /* variant 1: VARIANT 3 (one tile per warp, no prefetch) of the two headline 16-bit 4:4:4 configurations and of the
     * 4:2:0 bilinear RGB16 / RGBA16+unmultiply ones, to compare with the persistent production kernels */
    const AbY2RJob & J = *job;
    cudaStream_t st = (cudaStream_t)stream;
    if (variant != 1) return 0;
    if (J.yuvBytes != 2 || J.rgbBytes != 2 || J.alphaFirst || J.swapRB) return 0;
    if (J.mode != AB_MODE_COEFF || J.toF16 || J.alphaRescale || !J.rangeY.exact || !J.rangeUV.exact || !J.kg.exact || !J.rgbMaxDiv.exact || !J.yuvMaxDiv.exact) return 0;
    if ((((uintptr_t)J.y | (uintptr_t)J.yRB | (uintptr_t)J.rgb | (uintptr_t)J.rgbRB) & 15) != 0) return 0;
    if ((((uintptr_t)J.u | (uintptr_t)J.uRB | (uintptr_t)J.v | (uintptr_t)J.vRB) & 15) != 0) return 0;
    if (J.a && ((((uintptr_t)J.a | (uintptr_t)J.aRB) & 15) != 0)) return 0;
    if (J.chroma == AB_CHROMA_FULL && J.nChan == 4) {
        if (J.alphaWrite == AB_ALPHA_FILL && J.alphaMul == AB_MUL_NONE) return launch_cfg<uint16_t, uint16_t, CH_FULL, 4, false, AB_ALPHA_FILL, AB_MUL_NONE, 3>(J, st);
        if (J.alphaWrite == AB_ALPHA_COPY && J.alphaMul == AB_MUL_INT_PREMUL && J.a) return launch_cfg<uint16_t, uint16_t, CH_FULL, 4, false, AB_ALPHA_COPY, AB_MUL_INT_PREMUL, 3>(J, st);
    }
    if (J.chroma == AB_CHROMA_BILINEAR && J.vertBlend) {
        if (J.nChan == 3 && J.alphaMul == AB_MUL_NONE) return launch_cfg<uint16_t, uint16_t, CH_BIL420, 3, false, AB_ALPHA_KEEP, AB_MUL_NONE, 3>(J, st);
        if (J.nChan == 4 && J.alphaWrite == AB_ALPHA_COPY && J.alphaMul == AB_MUL_FLOAT_UNMUL && J.a)
            return launch_cfg<uint16_t, uint16_t, CH_BIL420, 4, false, AB_ALPHA_COPY, AB_MUL_FLOAT_UNMUL, 3>(J, st);
    }
    return 0;
}
