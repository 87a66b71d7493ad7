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

#include <type_traits>

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

/* ---- row-staged loops --------------------------------------------------------------------------------------------------
 * A block produces one SEGMENT of one destination row at a time.  Phase 1: the source columns the segment needs are read with
 * 16-byte loads, combined VERTICALLY the way the loop does it first (box: ScaleAddRow's column sums; bilinear down:
 * InterpolateRow; bilinear up / point sampling: the raw rows) and parked in shared memory, so every source byte is read from
 * HBM once and in full sectors.  Phase 2: each thread produces 8 bytes of the destination row from shared memory with the
 * loop's HORIZONTAL arithmetic and stores them as one vector.  Units that do not fit (ratios beyond the staging buffer, boxes
 * hanging over the plane) take the table-driven sampler above. */
enum { kStageBytes = 24576, kScaleThreads = 128 };

template <typename T> struct Vec16; /* 16 bytes of samples */
template <> struct Vec16<uint8_t> { enum { N = 16 }; };
template <> struct Vec16<uint16_t> { enum { N = 8 }; };

template <typename T>
__device__ __forceinline__ void unpack16(const uint4 & v, int s[16 / sizeof(T)])
{
    const uint32_t w[4] = { v.x, v.y, v.z, v.w };
    if (sizeof(T) == 1) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s[k % (16 / sizeof(T))] = (int)((w[k >> 2] >> (8 * (k & 3))) & 0xFFu);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = (int)((w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu);
    }
}

template <typename T>
__device__ __forceinline__ const T * src_row(const AbScaleJob & J, int y)
{
    y = min(max(y, 0), J.srcH - 1);
    return reinterpret_cast<const T *>(J.src + (int64_t)y * J.srcRB);
}

/* 8 bytes of destination row j starting at sample i (i + VEC may run past i1: the tail is stored sample by sample) */
template <typename T>
__device__ __forceinline__ void store_group(const AbScaleJob & J, int j, int i, int i1, const int v[8 / sizeof(T)], bool aligned)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    T * out = reinterpret_cast<T *>(J.dst + (int64_t)j * J.dstRB) + i;
    if (aligned && i + VEC <= i1) {
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
            if (i + k < i1) out[k] = (T)v[k];
    }
}

/* Phase 1 for NR source rows y[0..NR): samples a0 .. a0 + span of each, handed VL at a time to `sink(k, rows)`, rows[r][e] =
 * sample a0 + k + e of row r.  Vector loads where the rows are 16-byte aligned and the vector lies inside the plane. */
template <typename T, int NR, typename Sink>
__device__ __forceinline__ void stage_rows(const AbScaleJob & J, const int y[NR], int nr, int a0, int span, bool srcAligned, Sink sink)
{
    constexpr int VL = Vec16<T>::N;
    const T * row[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) row[r] = src_row<T>(J, y[r < nr ? r : 0]);
    for (int k = (int)threadIdx.x * VL; k < span; k += kScaleThreads * VL) {
        int s[NR][VL];
        if (srcAligned && a0 + k + VL <= J.srcW) {
            uint4 v[NR];
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if (r < nr) v[r] = __ldg(reinterpret_cast<const uint4 *>(row[r] + a0 + k));
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if (r < nr) unpack16<T>(v[r], s[r]);
        } else {
#pragma unroll
            for (int r = 0; r < NR; ++r)
#pragma unroll
                for (int e = 0; e < VL; ++e) s[r][e] = (r < nr) ? (int)row[r][min(a0 + k + e, J.srcW - 1)] : 0;
        }
        sink(k, s);
    }
}

/* ScalePlaneBox (scale.c:148-256): column sums of the box rows (ScaleAddRow; 8-bit sums live in uint16_t and wrap), then
 * ScaleAddCols0/1/2 (scale_common.c) over them */
template <typename T>
__device__ bool unit_box(const AbScaleJob & J, const int32_t * R, int j, int i0, int i1, uint8_t * stage, bool srcAligned, bool dstAligned)
{
    typedef typename std::conditional<sizeof(T) == 1, uint16_t, uint32_t>::type Acc;
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T), MAXBH = 8;
    const int iy = R[0], bh = R[1];
    const int c0 = J.cols[4 * i0], cl = J.cols[4 * (i1 - 1)] + J.cols[4 * (i1 - 1) + 1];
    const int a0 = srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    if (iy < 0 || c0 < 0 || iy + bh > J.srcH || cl > J.srcW || span > kStageBytes / (int)sizeof(Acc) - VL) return false;
    Acc * cs = reinterpret_cast<Acc *>(stage);
    const int spanUp = (span + VL - 1) / VL * VL;
    for (int k = (int)threadIdx.x * VL; k < spanUp; k += kScaleThreads * VL) {
        uint32_t acc[VL];
#pragma unroll
        for (int e = 0; e < VL; ++e) acc[e] = 0;
        const bool vec = srcAligned && a0 + k + VL <= J.srcW;
        for (int r0 = 0; r0 < bh; r0 += MAXBH) { /* MAXBH independent loads in flight */
            if (vec) {
                uint4 v[MAXBH];
#pragma unroll
                for (int r = 0; r < MAXBH; ++r)
                    if (r0 + r < bh) v[r] = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const T *>(J.src + (int64_t)(iy + r0 + r) * J.srcRB) + a0 + k));
#pragma unroll
                for (int r = 0; r < MAXBH; ++r)
                    if (r0 + r < bh) {
                        int s[VL];
                        unpack16<T>(v[r], s);
#pragma unroll
                        for (int e = 0; e < VL; ++e) acc[e] += (uint32_t)s[e];
                    }
            } else {
                for (int r = r0; r < min(bh, r0 + MAXBH); ++r) {
                    const T * q = reinterpret_cast<const T *>(J.src + (int64_t)(iy + r) * J.srcRB);
#pragma unroll
                    for (int e = 0; e < VL; ++e) acc[e] += (uint32_t)q[min(a0 + k + e, J.srcW - 1)];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < VL; ++e) cs[k + e] = (Acc)acc[e]; /* the compiler merges these into vector stores */
    }
    __syncthreads();
    /* scaletbl[0], [1] (scale.c:58-59 / scale_common.c ScaleAddCols2): box widths minBoxWidth and minBoxWidth + 1 */
    const uint32_t scale0 = (uint32_t)(65536 / (max(J.minBoxWidth, 1) * bh)), scale1 = (uint32_t)(65536 / (max(J.minBoxWidth + 1, 1) * bh));
    for (int i = i0 + (int)threadIdx.x * VEC; i < i1; i += kScaleThreads * VEC) {
        int v[VEC];
#pragma unroll
        for (int o = 0; o < VEC; ++o) {
            v[o] = 0;
            if (i + o < i1) {
                const int4 C = __ldg(reinterpret_cast<const int4 *>(J.cols) + i + o);
                uint32_t sum = 0;
                const Acc * q = cs + (C.x - a0);
                for (int c = 0; c < C.y; ++c) sum += (uint32_t)q[c];
                const uint32_t scale = (C.z == 0) ? scale0 : ((C.z == 1) ? scale1 : (uint32_t)(65536 / (max(J.minBoxWidth + C.z, 1) * bh)));
                v[o] = (int)((sum * scale) >> 16);
            }
        }
        store_group<T>(J, j, i, i1, v, dstAligned);
    }
    __syncthreads();
    return true;
}

/* ScalePlaneBilinearDown (scale.c:258-369): InterpolateRow of rows yi, yi + 1 into a row buffer, then ScaleFilterCols */
template <typename T>
__device__ bool unit_bilinear_down(const AbScaleJob & J, const int32_t * R, int j, int i0, int i1, uint8_t * stage, bool srcAligned, bool dstAligned)
{
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T);
    const int yi = R[0], yf = J.linearOnly ? 0 : R[1];
    const int c0 = min(max(J.cols[4 * i0], 0), J.srcW - 1), cl = min(J.cols[4 * (i1 - 1)] + 2, J.srcW);
    const int a0 = srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    if (span <= 0 || span > kStageBytes / (int)sizeof(T) - VL) return false;
    T * sv = reinterpret_cast<T *>(stage);
    const int y[2] = { yi, yi + 1 };
    stage_rows<T, 2>(J, y, yf ? 2 : 1, a0, span, srcAligned, [&](int k, int s[2][VL]) {
#pragma unroll
        for (int e = 0; e < VL; ++e) sv[k + e] = (T)interp(s[0][e], yf ? s[1][e] : 0, yf);
    });
    __syncthreads();
    const int xmax = J.srcW - 1;
    for (int i = i0 + (int)threadIdx.x * VEC; i < i1; i += kScaleThreads * VEC) {
        int v[VEC];
#pragma unroll
        for (int o = 0; o < VEC; ++o) {
            v[o] = 0;
            if (i + o < i1) {
                const int2 C = __ldg(reinterpret_cast<const int2 *>(J.cols + 4 * (i + o)));
                const int xa = min(max(C.x, 0), xmax), xb = min(max(C.x + 1, 0), xmax);
                v[o] = blend<T>((int)sv[xa - a0], (int)sv[xb - a0], C.y);
            }
        }
        store_group<T>(J, j, i, i1, v, dstAligned);
    }
    __syncthreads();
    return true;
}

/* ScalePlaneBilinearUp (scale.c:371-466): ScaleFilterCols of the two cached rows, then InterpolateRow between them */
template <typename T>
__device__ bool unit_bilinear_up(const AbScaleJob & J, const int32_t * R, int j, int i0, int i1, uint8_t * stage, bool srcAligned, bool dstAligned)
{
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T);
    const int ra = R[0], rb = R[1], yf = J.linearOnly ? 0 : R[2];
    const int c0 = min(max(J.cols[4 * i0], 0), J.srcW - 1), cl = min(J.cols[4 * (i1 - 1)] + 2, J.srcW);
    const int a0 = srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    constexpr int half = kStageBytes / 2 / (int)sizeof(T);
    if (span <= 0 || span > half - VL) return false;
    T * sa = reinterpret_cast<T *>(stage);
    T * sb = sa + half;
    const int y[2] = { ra, rb };
    stage_rows<T, 2>(J, y, yf ? 2 : 1, a0, span, srcAligned, [&](int k, int s[2][VL]) {
#pragma unroll
        for (int e = 0; e < VL; ++e) {
            sa[k + e] = (T)s[0][e];
            sb[k + e] = (T)s[1][e];
        }
    });
    __syncthreads();
    const int xmax = J.srcW - 1;
    for (int i = i0 + (int)threadIdx.x * VEC; i < i1; i += kScaleThreads * VEC) {
        int v[VEC];
#pragma unroll
        for (int o = 0; o < VEC; ++o) {
            v[o] = 0;
            if (i + o < i1) {
                const int2 C = __ldg(reinterpret_cast<const int2 *>(J.cols + 4 * (i + o)));
                const int xa = min(max(C.x, 0), xmax) - a0, xb = min(max(C.x + 1, 0), xmax) - a0;
                const int hA = blend<T>((int)sa[xa], (int)sa[xb], C.y);
                const int hB = yf ? blend<T>((int)sb[xa], (int)sb[xb], C.y) : 0;
                v[o] = interp(hA, hB, yf);
            }
        }
        store_group<T>(J, j, i, i1, v, dstAligned);
    }
    __syncthreads();
    return true;
}

/* ScalePlaneVertical (scale_common.c:348-426): InterpolateRow straight from the two source rows, no staging */
template <typename T>
__device__ void unit_vertical(const AbScaleJob & J, const int32_t * R, int j, int i0, int i1, bool srcAligned, bool dstAligned)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    const int yi = R[0], yf = R[1];
    const T * pa = src_row<T>(J, yi);
    const T * pb = src_row<T>(J, yi + 1);
    const bool vec = srcAligned && dstAligned; /* 16-byte aligned rows are 8-byte aligned too; i is a multiple of VEC */
    for (int i = i0 + (int)threadIdx.x * VEC; i < i1; i += kScaleThreads * VEC) {
        int v[VEC];
        if (vec && i + VEC <= i1) {
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(pa + i));
            const uint2 b = yf ? __ldg(reinterpret_cast<const uint2 *>(pb + i)) : make_uint2(0u, 0u);
            const uint32_t wa[2] = { a.x, a.y }, wb[2] = { b.x, b.y };
#pragma unroll
            for (int o = 0; o < VEC; ++o) {
                const int sh = (int)(8 * sizeof(T)) * (o % (4 / (int)sizeof(T)));
                const uint32_t m = (sizeof(T) == 1) ? 0xFFu : 0xFFFFu;
                v[o] = interp((int)((wa[o / (4 / (int)sizeof(T))] >> sh) & m), (int)((wb[o / (4 / (int)sizeof(T))] >> sh) & m), yf);
            }
        } else {
#pragma unroll
            for (int o = 0; o < VEC; ++o) v[o] = (i + o < i1) ? interp((int)pa[i + o], yf ? (int)pb[i + o] : 0, yf) : 0;
        }
        store_group<T>(J, j, i, i1, v, dstAligned);
    }
}

/* the table-driven sampler for a whole unit (any loop; also the 2x up-scale groups) */
template <typename T>
__device__ void unit_generic(const AbScaleJob & J, const int32_t * R, int j, int i0, int i1, bool dstAligned)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    for (int i = i0 + (int)threadIdx.x * VEC; i < i1; i += kScaleThreads * VEC) {
        int v[VEC];
        const bool inner = (i >= 2) && (i + VEC <= J.dstW - 1);
        if (J.mode == AB_SCALE_UP2_BILINEAR && inner) {
            up2_group<T, true>(J, i, j, v);
        } else if (J.mode == AB_SCALE_UP2_LINEAR && inner) {
            up2_group<T, false>(J, i, R[0], v);
        } else {
#define AB_SCALE_CASE(M)                                                                                     \
    case M:                                                                                                  \
        _Pragma("unroll") for (int k = 0; k < VEC; ++k) v[k] = (i + k < i1) ? scale_sample<T, M>(J, R, i + k, j) : 0; \
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
        store_group<T>(J, j, i, i1, v, dstAligned);
    }
}

/* All planes of one image in one launch.  Blocks are shared out over the planes by blockStart[]; a block walks the units
 * (destination row, segment of segLen samples) of its plane. */
template <typename T>
__global__ void __launch_bounds__(kScaleThreads) scale_batch_kernel(const __grid_constant__ AbScaleBatch B)
{
    __shared__ __align__(16) uint8_t stage[kStageBytes];
    uint32_t p = 0;
    while (p + 1 < B.count && blockIdx.x >= B.blockStart[p + 1]) ++p;
    const AbScaleJob & J = B.job[p];
    const int segs = (J.dstW + J.segLen - 1) / J.segLen;
    const int64_t total = (int64_t)segs * J.dstH;
    const int64_t stride = (int64_t)(B.blockStart[p + 1] - B.blockStart[p]);
    const bool dstAligned = ((((uintptr_t)J.dst) | (uintptr_t)J.dstRB) & 7u) == 0;
    const bool srcAligned = ((((uintptr_t)J.src) | (uintptr_t)J.srcRB) & 15u) == 0;
    for (int64_t u = (int64_t)(blockIdx.x - B.blockStart[p]); u < total; u += stride) {
        const int j = (int)(u / segs), i0 = (int)(u - (int64_t)j * segs) * J.segLen, i1 = min(J.dstW, i0 + J.segLen);
        const int32_t * R = J.rows + 4 * j;
        bool done = false; /* block-uniform */
        switch (J.mode) {
            case AB_SCALE_BOX: done = unit_box<T>(J, R, j, i0, i1, stage, srcAligned, dstAligned); break;
            case AB_SCALE_BILINEAR_DOWN: done = unit_bilinear_down<T>(J, R, j, i0, i1, stage, srcAligned, dstAligned); break;
            case AB_SCALE_BILINEAR_UP: done = unit_bilinear_up<T>(J, R, j, i0, i1, stage, srcAligned, dstAligned); break;
            case AB_SCALE_VERTICAL: unit_vertical<T>(J, R, j, i0, i1, srcAligned, dstAligned); done = true; break;
            default: break;
        }
        if (!done) unit_generic<T>(J, R, j, i0, i1, dstAligned);
    }
}

} // namespace

extern "C" int abLaunchScaleBatch(AbScaleBatch * batch, void * stream)
{
    if (batch->count == 0) return 0;
    /* units = (destination row, segment); a segment is at most one pass of the block (kScaleThreads x 8 bytes), rows are cut into
     * equal segments, and a segment's source span has to fit the staging buffer (else the unit falls back to the sampler) */
    int64_t units[AB_SCALE_MAX_PLANES], work[AB_SCALE_MAX_PLANES], all = 0;
    const int vec = 8 / batch->job[0].bps;
    for (uint32_t p = 0; p < batch->count; ++p) {
        AbScaleJob * j = &batch->job[p];
        int maxSeg = kScaleThreads * vec;
        if (j->dstW > 0 && j->srcW > j->dstW) { /* down-scaling: keep the span (~ seg * srcW / dstW columns) inside the buffer */
            const int64_t elem = (j->mode == AB_SCALE_BOX) ? 2 * j->bps : ((j->mode == AB_SCALE_BILINEAR_UP) ? 2 * j->bps : j->bps);
            const int64_t cap = kStageBytes / elem - 64;
            while (maxSeg > vec && ((int64_t)maxSeg * j->srcW + j->dstW - 1) / j->dstW + 4 > cap) maxSeg /= 2;
        }
        const int nseg = (j->dstW + maxSeg - 1) / maxSeg;
        int seg = (j->dstW + nseg - 1) / (nseg > 0 ? nseg : 1);
        seg = (seg + vec - 1) / vec * vec;
        if (seg < vec) seg = vec;
        j->segLen = seg;
        units[p] = (int64_t)((j->dstW + seg - 1) / seg) * j->dstH;
        work[p] = units[p] * ((int64_t)seg + (int64_t)seg * (j->srcW > j->dstW ? j->srcW / j->dstW : 1)); /* ~ bytes touched per unit */
        all += work[p];
    }
    if (all <= 0) return 0;
    const int64_t budget = 148 * 8; /* resident blocks: 8 per SM (24 KB of staging each) */
    uint32_t start = 0;
    for (uint32_t p = 0; p < batch->count; ++p) {
        int64_t share = (work[p] * budget + all - 1) / all;
        if (share > units[p]) share = units[p];
        if (share < 1) share = 1;
        batch->blockStart[p] = start;
        start += (uint32_t)share;
    }
    batch->blockStart[batch->count] = start;
    if (batch->job[0].bps == 1) scale_batch_kernel<uint8_t><<<start, kScaleThreads, 0, (cudaStream_t)stream>>>(*batch);
    else scale_batch_kernel<uint16_t><<<start, kScaleThreads, 0, (cudaStream_t)stream>>>(*batch);
    return (int)cudaGetLastError();
}
