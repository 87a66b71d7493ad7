/*
 * scale_kernels.cu -- plane scaling with libyuv's integer arithmetic, one thread per destination sample.
 *
 * avifImageScale (src/scale.c:24-201) hands every plane to libyuv's ScalePlane / ScalePlane_12 with kFilterBox
 * (third_party/libyuv/source/scale.c:829-1007 in the reference tree).  After ScaleFilterReduce that is one of seven loops; each
 * destination sample is a closed-form integer function of a few source samples and of the 16.16 positions the loop was at, so
 * the host tabulates the positions (rows[], cols[]: csrc/scale.c) and this kernel does the sample arithmetic:
 *   VERTICAL        ScalePlaneVertical     scale_common.c:348-426   InterpolateRow of rows yi, yi+1
 *   SIMPLE          ScalePlaneSimple       scale.c:763-827          point sampling
 *   BILINEAR_DOWN   ScalePlaneBilinearDown scale.c:258-369          InterpolateRow, then ScaleFilterCols
 *   BILINEAR_UP     ScalePlaneBilinearUp   scale.c:371-466,669-761  ScaleFilterCols of two rows, then InterpolateRow
 *   UP2_LINEAR      ScalePlaneUp2_Linear   scale.c:467-501 ...      (3,1)/4 kernel + edge columns (scale_any.c:27-40)
 *   UP2_BILINEAR    ScalePlaneUp2_Bilinear scale.c:503-538 ...      (9,3,3,1)/16 kernel + edges (scale_any.c:54-78)
 *   BOX             ScalePlaneBox          scale.c:148-256          ScaleAddRow sums (u16 / u32 wrap kept), ScaleAddCols
 * Bit-exact with the reference build (tests/test_gpu_scale.py).
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace {

template <typename T>
__device__ __forceinline__ int ld(const AbScaleJob & J, int y, int x)
{
    /* the reference never reads outside the plane with a non-zero weight; clamping keeps zero-weight taps in bounds */
    y = min(max(y, 0), J.srcH - 1);
    x = min(max(x, 0), J.srcW - 1);
    return (int)*reinterpret_cast<const T *>(J.src + (int64_t)y * J.srcRB + (int64_t)x * sizeof(T));
}

/* InterpolateRow_C / _16_C, row_common.c:46-105 */
__device__ __forceinline__ int interp(int a, int b, int yf)
{
    if (yf == 0) return a;
    if (yf == 128) return (a + b + 1) >> 1;
    return (a * (256 - yf) + b * yf + 128) >> 8;
}

/* ScaleFilterCols' BLENDER, scale_common.c:188-196 (8-bit, non-ARM form) and :253-256 (16-bit) */
template <typename T>
__device__ __forceinline__ int blend(int a, int b, int f)
{
    if (sizeof(T) == 1) return (int)(uint8_t)(a + ((((f >> 9) * (b - a)) + 0x40) >> 7));
    return (int)(uint16_t)(a + (int)((((int64_t)f * ((int64_t)b - a)) + 0x8000) >> 16));
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(const __grid_constant__ AbScaleJob J)
{
    const int64_t total = (int64_t)J.dstW * J.dstH;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int j = (int)(idx / J.dstW), i = (int)(idx - (int64_t)j * J.dstW);
        const int32_t * R = J.rows + 4 * j;
        const int32_t * C = J.cols + 4 * i;
        int v = 0;
        switch (J.mode) {
            case AB_SCALE_VERTICAL: { /* rows: yi, yf */
                const int yi = R[0], yf = R[1];
                v = interp(ld<T>(J, yi, i), (yf == 0) ? 0 : ld<T>(J, yi + 1, i), yf);
                break;
            }
            case AB_SCALE_SIMPLE: /* rows: yi ; cols: xi */
                v = ld<T>(J, R[0], C[0]);
                break;
            case AB_SCALE_BILINEAR_DOWN: { /* rows: yi, yf ; cols: xi, xf */
                const int yi = R[0], yf = J.linearOnly ? 0 : R[1], xi = C[0], xf = C[1];
                const int a = interp(ld<T>(J, yi, xi), yf ? ld<T>(J, yi + 1, xi) : 0, yf);
                const int b = interp(ld<T>(J, yi, xi + 1), yf ? ld<T>(J, yi + 1, xi + 1) : 0, yf);
                v = blend<T>(a, b, xf);
                break;
            }
            case AB_SCALE_BILINEAR_UP: { /* rows: ra, rb, yf ; cols: xi, xf */
                const int ra = R[0], rb = R[1], yf = J.linearOnly ? 0 : R[2], xi = C[0], xf = C[1];
                const int hA = blend<T>(ld<T>(J, ra, xi), ld<T>(J, ra, xi + 1), xf);
                const int hB = yf ? blend<T>(ld<T>(J, rb, xi), ld<T>(J, rb, xi + 1), xf) : 0;
                v = interp(hA, hB, yf);
                break;
            }
            case AB_SCALE_UP2_LINEAR: { /* rows: yi */
                const int yi = R[0];
                if (i == 0) v = ld<T>(J, yi, 0);
                else if (i == J.dstW - 1) v = ld<T>(J, yi, (J.dstW - 1) / 2);
                else {
                    const int k = (i - 1) >> 1, a = ld<T>(J, yi, k), b = ld<T>(J, yi, k + 1);
                    v = ((i - 1) & 1) ? ((a + 3 * b + 2) >> 2) : ((3 * a + b + 2) >> 2);
                }
                break;
            }
            case AB_SCALE_UP2_BILINEAR: { /* scale.c:503-538: row 0 and the last row of an even height come from one source row */
                int rs, rt; /* near row gets weights (9,3), far row (3,1) */
                if (j == 0) rs = rt = 0;
                else {
                    const int r = (j - 1) >> 1;
                    if (r >= J.srcH - 1) rs = rt = J.srcH - 1;
                    else if ((j - 1) & 1) { rs = r + 1; rt = r; }
                    else { rs = r; rt = r + 1; }
                }
                if (i == 0 || i == J.dstW - 1) {
                    const int k = (i == 0) ? 0 : (J.dstW - 1) / 2;
                    v = (3 * ld<T>(J, rs, k) + ld<T>(J, rt, k) + 2) >> 2;
                } else {
                    const int k = (i - 1) >> 1;
                    const int kn = ((i - 1) & 1) ? k + 1 : k, kf = ((i - 1) & 1) ? k : k + 1; /* near / far column */
                    v = (ld<T>(J, rs, kn) * 9 + ld<T>(J, rs, kf) * 3 + ld<T>(J, rt, kn) * 3 + ld<T>(J, rt, kf) * 1 + 8) >> 4;
                }
                break;
            }
            case AB_SCALE_BOX: { /* rows: iy, boxheight ; cols: ix, boxwidth, table index */
                const int iy = R[0], bh = R[1], ix = C[0], bw = C[1];
                uint32_t sum = 0;
                for (int c = 0; c < bw; ++c) {
                    uint32_t col = 0; /* ScaleAddRow: one accumulator per source column */
                    for (int r = 0; r < bh; ++r) col += (uint32_t)ld<T>(J, iy + r, ix + c);
                    if (sizeof(T) == 1) col &= 0xFFFFu; /* uint16_t accumulators wrap (scale_common.c:316-328) */
                    sum += col;
                }
                int wdiv = J.minBoxWidth + C[2];
                if (wdiv < 1) wdiv = 1;
                const int scale = 65536 / (wdiv * bh); /* scaletbl[], scale.c:58-59 */
                v = (int)((sum * (uint32_t)scale) >> 16);
                break;
            }
            default: break;
        }
        *reinterpret_cast<T *>(J.dst + (int64_t)j * J.dstRB + (int64_t)i * sizeof(T)) = (T)v;
    }
}

} // namespace

extern "C" int abLaunchScale(const AbScaleJob * job, void * stream)
{
    const int64_t total = (int64_t)job->dstW * job->dstH;
    if (total <= 0) return 0;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (job->bps == 1) scale_kernel<uint8_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*job);
    else scale_kernel<uint16_t><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*job);
    return (int)cudaGetLastError();
}
