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

/* one destination sample (i, j) of plane job J */
template <typename T, int MODE>
__device__ __forceinline__ int scale_sample(const AbScaleJob & J, const int32_t * R, int i, int j)
{
    const int32_t * C = J.cols + 4 * i;
    switch (MODE ? MODE : (int)J.mode) { /* MODE != 0: the loop is a compile-time constant and the other cases fold away */
        case AB_SCALE_VERTICAL: { /* rows: yi, yf */
            const int yi = R[0], yf = R[1];
            return interp(ld<T>(J, yi, i), (yf == 0) ? 0 : ld<T>(J, yi + 1, i), yf);
        }
        case AB_SCALE_SIMPLE: /* rows: yi ; cols: xi */
            return ld<T>(J, R[0], C[0]);
        case AB_SCALE_BILINEAR_DOWN: { /* rows: yi, yf ; cols: xi, xf */
            const int yi = R[0], yf = J.linearOnly ? 0 : R[1], xi = C[0], xf = C[1];
            const int a = interp(ld<T>(J, yi, xi), yf ? ld<T>(J, yi + 1, xi) : 0, yf);
            const int b = interp(ld<T>(J, yi, xi + 1), yf ? ld<T>(J, yi + 1, xi + 1) : 0, yf);
            return blend<T>(a, b, xf);
        }
        case AB_SCALE_BILINEAR_UP: { /* rows: ra, rb, yf ; cols: xi, xf */
            const int ra = R[0], rb = R[1], yf = J.linearOnly ? 0 : R[2], xi = C[0], xf = C[1];
            const int hA = blend<T>(ld<T>(J, ra, xi), ld<T>(J, ra, xi + 1), xf);
            const int hB = yf ? blend<T>(ld<T>(J, rb, xi), ld<T>(J, rb, xi + 1), xf) : 0;
            return interp(hA, hB, yf);
        }
        case AB_SCALE_UP2_LINEAR: { /* rows: yi */
            const int yi = R[0];
            if (i == 0) return ld<T>(J, yi, 0);
            if (i == J.dstW - 1) return ld<T>(J, yi, (J.dstW - 1) / 2);
            const int k = (i - 1) >> 1, a = ld<T>(J, yi, k), b = ld<T>(J, yi, k + 1);
            return ((i - 1) & 1) ? ((a + 3 * b + 2) >> 2) : ((3 * a + b + 2) >> 2);
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
                return (3 * ld<T>(J, rs, k) + ld<T>(J, rt, k) + 2) >> 2;
            }
            const int k = (i - 1) >> 1;
            const int kn = ((i - 1) & 1) ? k + 1 : k, kf = ((i - 1) & 1) ? k : k + 1; /* near / far column */
            return (ld<T>(J, rs, kn) * 9 + ld<T>(J, rs, kf) * 3 + ld<T>(J, rt, kn) * 3 + ld<T>(J, rt, kf) * 1 + 8) >> 4;
        }
        case AB_SCALE_BOX: { /* rows: iy, boxheight ; cols: ix, boxwidth, table index */
            const int iy = R[0], bh = R[1], ix = C[0], bw = C[1];
            uint32_t sum = 0;
            if (iy >= 0 && ix >= 0 && iy + bh <= J.srcH && ix + bw <= J.srcW) {
                /* the box lies inside the plane (always, for the positions libyuv generates): plain pointer walk instead of
                 * a clamped address computation per tap */
                const T * p0 = reinterpret_cast<const T *>(J.src + (int64_t)iy * J.srcRB) + ix;
                for (int c = 0; c < bw; ++c) {
                    uint32_t col = 0; /* ScaleAddRow: one accumulator per source column */
                    const uint8_t * q = reinterpret_cast<const uint8_t *>(p0 + c);
                    for (int r = 0; r < bh; ++r, q += J.srcRB) col += (uint32_t)*reinterpret_cast<const T *>(q);
                    if (sizeof(T) == 1) col &= 0xFFFFu; /* uint16_t accumulators wrap (scale_common.c:316-328) */
                    sum += col;
                }
            } else {
                for (int c = 0; c < bw; ++c) {
                    uint32_t col = 0;
                    for (int r = 0; r < bh; ++r) col += (uint32_t)ld<T>(J, iy + r, ix + c);
                    if (sizeof(T) == 1) col &= 0xFFFFu;
                    sum += col;
                }
            }
            int wdiv = J.minBoxWidth + C[2];
            if (wdiv < 1) wdiv = 1;
            const int scale = 65536 / (wdiv * bh); /* scaletbl[], scale.c:58-59 */
            return (int)((sum * (uint32_t)scale) >> 16);
        }
        default: return 0;
    }
}

/* ScalePlaneUp2_Linear / _Bilinear (scale.c:467-538, scale_any.c:27-78) for a thread's 8 (u8) or 4 (u16) consecutive outputs that
 * all lie strictly inside the row (i0 >= 1 .. i0 + VEC - 1 <= dstW - 2): outputs i0..i0+VEC-1 read source columns
 * (i0 - 1) / 2 .. (i0 + VEC - 2) / 2 + 1, i.e. VEC / 2 + 2 columns when i0 is even -- loaded once per row instead of four
 * clamped byte loads per output. */
template <typename T, bool BILINEAR>
__device__ __forceinline__ void up2_group(const AbScaleJob & J, int i0, int j, int v[8 / sizeof(T)])
{
    constexpr int VEC = 8 / (int)sizeof(T);
    constexpr int NC = VEC / 2 + 2;
    int rs = 0, rt = 0; /* near row gets weights (9,3) / (3,1), far row (3,1); linear: one row */
    if (BILINEAR) {
        if (j != 0) {
            const int r = (j - 1) >> 1;
            if (r >= J.srcH - 1) rs = rt = J.srcH - 1;
            else if ((j - 1) & 1) { rs = r + 1; rt = r; }
            else { rs = r; rt = r + 1; }
        }
    } else {
        rs = rt = j; /* the caller passes the source row */
    }
    const int k0 = (i0 - 1) >> 1; /* i0 even and >= 2: k0 = i0/2 - 1 >= 0 */
    const T * ps = reinterpret_cast<const T *>(J.src + (int64_t)rs * J.srcRB) + k0;
    const T * pt = reinterpret_cast<const T *>(J.src + (int64_t)rt * J.srcRB) + k0;
    int a[NC], b[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        a[k] = (int)ps[k];
        b[k] = BILINEAR ? (int)pt[k] : 0;
    }
#pragma unroll
    for (int o = 0; o < VEC; ++o) {
        const int i = i0 + o;
        const int k = ((i - 1) >> 1) - k0;
        const int kn = ((i - 1) & 1) ? k + 1 : k, kf = ((i - 1) & 1) ? k : k + 1;
        if (BILINEAR) v[o] = (a[kn] * 9 + a[kf] * 3 + b[kn] * 3 + b[kf] + 8) >> 4;
        else v[o] = ((i - 1) & 1) ? ((a[k] + 3 * a[k + 1] + 2) >> 2) : ((3 * a[k] + a[k + 1] + 2) >> 2);
    }
}

/* All planes of one image in one launch.  A thread produces 8 bytes of one destination row (8 u8 or 4 u16 samples) and stores
 * them as one vector when the row is 8-byte aligned; consecutive threads take consecutive groups, so a warp writes 256
 * contiguous bytes.  Blocks are shared out over the planes by blockStart[]. */
template <typename T>
__global__ void __launch_bounds__(256) scale_batch_kernel(const __grid_constant__ AbScaleBatch B)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    uint32_t p = 0;
    while (p + 1 < B.count && blockIdx.x >= B.blockStart[p + 1]) ++p;
    const AbScaleJob & J = B.job[p];
    const int64_t groupsPerRow = (J.dstW + VEC - 1) / VEC;
    const int64_t total = groupsPerRow * J.dstH;
    const int64_t stride = (int64_t)(B.blockStart[p + 1] - B.blockStart[p]) * blockDim.x;
    const bool aligned = ((((uintptr_t)J.dst) | (uintptr_t)J.dstRB) & 7u) == 0;
    for (int64_t u = (int64_t)(blockIdx.x - B.blockStart[p]) * blockDim.x + threadIdx.x; u < total; u += stride) {
        const int j = (int)(u / groupsPerRow), i0 = (int)(u - (int64_t)j * groupsPerRow) * VEC;
        const int32_t * R = J.rows + 4 * j;
        int v[VEC];
        /* the loop is a per-plane (block-uniform) constant: one specialised body per loop, the generic one as fallback */
        const bool inner = (i0 >= 2) && (i0 + VEC <= J.dstW - 1);
        if (J.mode == AB_SCALE_UP2_BILINEAR && inner) {
            up2_group<T, true>(J, i0, j, v);
        } else if (J.mode == AB_SCALE_UP2_LINEAR && inner) {
            up2_group<T, false>(J, i0, R[0], v);
        } else {
#define AB_SCALE_CASE(M)                                                                                     \
    case M:                                                                                                  \
        _Pragma("unroll") for (int k = 0; k < VEC; ++k) v[k] = (i0 + k < J.dstW) ? scale_sample<T, M>(J, R, i0 + k, j) : 0; \
        break;
            switch (J.mode) {
                AB_SCALE_CASE(AB_SCALE_VERTICAL)
                AB_SCALE_CASE(AB_SCALE_SIMPLE)
                AB_SCALE_CASE(AB_SCALE_BILINEAR_DOWN)
                AB_SCALE_CASE(AB_SCALE_BILINEAR_UP)
                AB_SCALE_CASE(AB_SCALE_UP2_LINEAR)
                AB_SCALE_CASE(AB_SCALE_UP2_BILINEAR)
                AB_SCALE_CASE(AB_SCALE_BOX)
                default:
#pragma unroll
                    for (int k = 0; k < VEC; ++k) v[k] = 0;
            }
#undef AB_SCALE_CASE
        }
        T * out = reinterpret_cast<T *>(J.dst + (int64_t)j * J.dstRB) + i0;
        if (aligned && i0 + VEC <= J.dstW) {
            uint2 w;
            if (sizeof(T) == 1) {
                w.x = (uint32_t)(v[0] & 0xFF) | ((uint32_t)(v[1] & 0xFF) << 8) | ((uint32_t)(v[2] & 0xFF) << 16) | ((uint32_t)(v[3] & 0xFF) << 24);
                w.y = (uint32_t)(v[4 % VEC] & 0xFF) | ((uint32_t)(v[5 % VEC] & 0xFF) << 8) | ((uint32_t)(v[6 % VEC] & 0xFF) << 16) |
                      ((uint32_t)(v[7 % VEC] & 0xFF) << 24);
            } else {
                w.x = (uint32_t)(v[0] & 0xFFFF) | ((uint32_t)(v[1] & 0xFFFF) << 16);
                w.y = (uint32_t)(v[2] & 0xFFFF) | ((uint32_t)(v[3] & 0xFFFF) << 16);
            }
            *reinterpret_cast<uint2 *>(out) = w;
        } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (i0 + k < J.dstW) out[k] = (T)v[k];
        }
    }
}

} // namespace

extern "C" int abLaunchScaleBatch(AbScaleBatch * batch, void * stream)
{
    if (batch->count == 0) return 0;
    /* blocks in proportion to the work of each plane, 16 resident waves' worth in all */
    int64_t units[AB_SCALE_MAX_PLANES], all = 0;
    const int vec = 8 / batch->job[0].bps;
    for (uint32_t p = 0; p < batch->count; ++p) {
        units[p] = (int64_t)((batch->job[p].dstW + vec - 1) / vec) * batch->job[p].dstH;
        all += units[p];
    }
    if (all <= 0) return 0;
    const int64_t budget = 148 * 16;
    uint32_t start = 0;
    for (uint32_t p = 0; p < batch->count; ++p) {
        int64_t need = (units[p] + 255) / 256;
        int64_t share = (units[p] * budget + all - 1) / all;
        if (share < 1) share = 1;
        if (share > need) share = need;
        if (share < 1) share = 1;
        batch->blockStart[p] = start;
        start += (uint32_t)share;
    }
    batch->blockStart[batch->count] = start;
    if (batch->job[0].bps == 1) scale_batch_kernel<uint8_t><<<start, 256, 0, (cudaStream_t)stream>>>(*batch);
    else scale_batch_kernel<uint16_t><<<start, 256, 0, (cudaStream_t)stream>>>(*batch);
    return (int)cudaGetLastError();
}
