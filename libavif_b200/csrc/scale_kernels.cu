/*
 * scale_kernels.cu -- plane scaling with libyuv's integer arithmetic.  A warp produces (segment of 32 x 8 bytes,
 * block of destination rows) units through per-loop code that stages source rows in shared memory (further down); the table-driven
 * one-sample-at-a-time form that follows first is the fallback and the second implementation.
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

/* ---- row-staged loops --------------------------------------------------------------------------------------------------
 * A WARP produces one segment (32 lanes x 8 bytes) of one destination row at a time, independently of the other warps of its
 * block (its own slice of shared memory, __syncwarp only), so an SM keeps 40 units in flight.  Phase 1: the source columns the
 * segment needs are read with 16-byte loads, combined VERTICALLY the way the loop does it first (box: ScaleAddRow's column
 * sums; bilinear down: InterpolateRow; bilinear up, point sampling, 2x up-scaling: the raw rows) and parked in shared memory, so
 * every source byte is read from HBM once and in full sectors.  Phase 2: each lane produces 8 bytes of the destination row from
 * shared memory with the loop's HORIZONTAL arithmetic and stores them as one vector.  Units that do not fit (ratios beyond the
 * staging slice, boxes hanging over the plane) take the table-driven sampler above. */
enum { kScaleWarps = 8, kScaleThreads = kScaleWarps * 32, kStageBytes = 3072 /* per warp */ };

template <typename T> struct Vec16; /* 16 bytes of samples */
template <> struct Vec16<uint8_t> { enum { N = 16 }; };
template <> struct Vec16<uint16_t> { enum { N = 8 }; };

/* sample e of a 16-byte vector */
template <typename T>
__device__ __forceinline__ uint32_t vec_sample(const uint4 & v, int e)
{
    const uint32_t w = (sizeof(T) == 1) ? ((e < 4) ? v.x : (e < 8) ? v.y : (e < 12) ? v.z : v.w) : ((e < 2) ? v.x : (e < 4) ? v.y : (e < 6) ? v.z : v.w);
    return (sizeof(T) == 1) ? ((w >> (8 * (e & 3))) & 0xFFu) : ((w >> (16 * (e & 1))) & 0xFFFFu);
}

template <typename T>
__device__ __forceinline__ const T * src_row(const AbScaleJob & J, int y)
{
    y = min(max(y, 0), J.srcH - 1);
    return reinterpret_cast<const T *>(J.src + (int64_t)y * J.srcRB);
}

/* samples a .. a + VL of a source row as one vector (columns past the plane replicate the last one: they only ever meet zero
 * weights).  The element-wise form is rare (unaligned planes, the last vector of a row) and kept out of line: these loops run
 * once per unit, so their code has to stay small enough for the instruction caches (ncu: no_instruction was the top stall) */
template <typename T>
__device__ __noinline__ uint4 load_vec_slow(const T * row, int a, int last)
{
    constexpr int VL = Vec16<T>::N;
    uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll 1
    for (int e = 0; e < VL; ++e) {
        const uint32_t s = (uint32_t)row[min(a + e, last)];
        const int word = (sizeof(T) == 1) ? (e >> 2) : (e >> 1);
        const uint32_t v = (sizeof(T) == 1) ? (s << (8 * (e & 3))) : (s << (16 * (e & 1)));
        if (word == 0) w[0] |= v;
        else if (word == 1) w[1] |= v;
        else if (word == 2) w[2] |= v;
        else w[3] |= v;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
template <typename T>
__device__ __forceinline__ uint4 load_vec(const AbScaleJob & J, const T * row, int a, bool srcAligned)
{
    constexpr int VL = Vec16<T>::N;
    if (srcAligned && a + VL <= J.srcW) return __ldg(reinterpret_cast<const uint4 *>(row + a));
    return load_vec_slow<T>(row, a, J.srcW - 1);
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

/* raw samples a0 .. a0 + span of one source row into a staging array of T */
template <typename T>
__device__ __forceinline__ void stage_raw(const AbScaleJob & J, int y, int a0, int span, bool srcAligned, T * st, uint32_t lane)
{
    constexpr int VL = Vec16<T>::N;
    const T * row = src_row<T>(J, y);
    for (int k = (int)lane * VL; k < span; k += 32 * VL) *reinterpret_cast<uint4 *>(st + k) = load_vec<T>(J, row, a0 + k, srcAligned);
}

/* A unit of work: destination rows j0..j1 of the segment i0..i1 of one plane.  Whatever depends on the COLUMN only -- the source
 * span of the segment, each output's staging offsets and fraction -- is derived once per unit and kept in one packed register per
 * output; the row loop then only stages and filters. */
struct Seg {
    int i0, i1, j0, j1;
    bool srcAligned, dstAligned;
    uint32_t lane;
};

template <typename T>
__device__ __noinline__ void rows_generic(const AbScaleJob & J, const Seg & S, int ja, int jb);

/* column of the 16.16 position of destination sample i: from the closed form x0 + i * dx when the host vouches for it
 * (AbScaleJob::colsArith -- no dependent table load on the unit's critical path), else from cols[] */
__device__ __forceinline__ uint32_t col_pos(const AbScaleJob & J, int i) { return (uint32_t)J.x0 + (uint32_t)i * (uint32_t)J.dx; } /* < 2^31: csrc/scale.c checks before it sets colsArith */

/* ScalePlaneBox (scale.c:148-256): column sums of the box rows (ScaleAddRow; 8-bit sums live in uint16_t and wrap), then
 * ScaleAddCols0/1/2 (scale_common.c) over them.  A box is minBoxWidth or minBoxWidth + 1 columns wide (x advances by dx). */
template <typename T>
__device__ bool seg_box(const AbScaleJob & J, const Seg & S, uint8_t * stage)
{
    typedef typename std::conditional<sizeof(T) == 1, uint16_t, uint32_t>::type Acc;
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T);
    const int m = J.minBoxWidth;
    const bool arith = J.colsArith != 0;
    /* the first row's table entry: asked for before anything else so that its latency hides behind the column arithmetic */
    int iyNext = J.rows[4 * S.j0], bhNext = J.rows[4 * S.j0 + 1];
    const int c0 = arith ? (int)(col_pos(J, S.i0) >> 16) : J.cols[4 * S.i0];
    const int cl = arith ? (int)(col_pos(J, S.i1) >> 16) : J.cols[4 * (S.i1 - 1)] + J.cols[4 * (S.i1 - 1) + 1];
    const int a0 = S.srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    if (m < 1 || c0 < 0 || cl > J.srcW || span > kStageBytes / (int)sizeof(Acc) - VL) return false;
    Acc * cs = reinterpret_cast<Acc *>(stage);
    const int i = S.i0 + (int)S.lane * VEC;
    /* staging offset of the lane's box o | (one column more than minBoxWidth) << 16 | (inside the segment) << 17; derived where it is
     * used (a few integer operations from the closed form) instead of being held across the loads of the row */
    auto box_info = [&](int o, bool & odd) -> uint32_t {
        if (i + o >= S.i1) return 0u;
        int ix, bw;
        if (arith) { /* widths are minBoxWidth or one more by construction: (i + 1) * dx >> 16 minus i * dx >> 16, verified on the host */
            const uint32_t t = col_pos(J, i + o);
            ix = (int)(t >> 16);
            bw = (int)((t + (uint32_t)J.dx) >> 16) - ix;
        } else {
            const int4 C = __ldg(reinterpret_cast<const int4 *>(J.cols) + i + o);
            ix = C.x;
            bw = C.y;
            odd = odd || (C.y != m + C.z) || (bw != m && bw != m + 1) || (ix < a0);
        }
        return (uint32_t)((ix - a0) & 0xFFFF) | ((uint32_t)((bw - m) & 1) << 16) | (1u << 17);
    };
    bool odd = false;
    if (!arith) {
#pragma unroll 1
        for (int o = 0; o < VEC; ++o) (void)box_info(o, odd);
    }
    if (__any_sync(0xFFFFFFFFu, odd)) return false; /* never for the tables of csrc/scale.c; the sampler would handle it */
    int lastBh = -1;
    uint32_t scale0 = 0, scale1 = 0;
    for (int j = S.j0; j < S.j1; ++j) {
        const int iy = iyNext, bh = bhNext;
        if (j + 1 < S.j1) {
            iyNext = J.rows[4 * (j + 1)];
            bhNext = J.rows[4 * (j + 1) + 1];
        }
        if (iy < 0 || bh < 1 || iy + bh > J.srcH || (sizeof(T) == 1 && bh > 257)) { /* 8-bit sums past 2^16 wrap (scale_common.c:316-328) */
            rows_generic<T>(J, S, j, j + 1);
            continue;
        }
        /* two vectors per lane and pass, two rows at a time: four loads in flight per lane */
        for (int k = (int)S.lane * VL; k < span; k += 64 * VL) {
            const int k1 = k + 32 * VL;
            const bool second = k1 < span;
            /* 8 words of sums per vector, in memory order: 8-bit samples as pairs of 16-bit sums (no carry between the halves:
             * bh * 255 < 2^16), 16-bit samples as 32-bit sums */
            uint32_t accA[8], accB[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) accA[e] = accB[e] = 0;
            const uint8_t * row = J.src + (int64_t)iy * J.srcRB;
            for (int r0 = 0; r0 < bh; r0 += 2) {
                uint4 va[2], vb[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    va[r] = vb[r] = make_uint4(0u, 0u, 0u, 0u);
                    if (r0 + r < bh) {
                        const T * q = reinterpret_cast<const T *>(row + (int64_t)(r0 + r) * J.srcRB);
                        va[r] = load_vec<T>(J, q, a0 + k, S.srcAligned);
                        if (second) vb[r] = load_vec<T>(J, q, a0 + k1, S.srcAligned);
                    }
                }
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const uint32_t wa[4] = { va[r].x, va[r].y, va[r].z, va[r].w }, wb[4] = { vb[r].x, vb[r].y, vb[r].z, vb[r].w };
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (sizeof(T) == 1) {
                            accA[2 * q] += __byte_perm(wa[q], 0u, 0x4140u);     /* bytes 0, 1 -> two 16-bit lanes */
                            accA[2 * q + 1] += __byte_perm(wa[q], 0u, 0x4342u); /* bytes 2, 3 */
                            accB[2 * q] += __byte_perm(wb[q], 0u, 0x4140u);
                            accB[2 * q + 1] += __byte_perm(wb[q], 0u, 0x4342u);
                        } else {
                            accA[2 * q] += wa[q] & 0xFFFFu;
                            accA[2 * q + 1] += wa[q] >> 16;
                            accB[2 * q] += wb[q] & 0xFFFFu;
                            accB[2 * q + 1] += wb[q] >> 16;
                        }
                    }
                }
            }
            *reinterpret_cast<uint4 *>(cs + k) = make_uint4(accA[0], accA[1], accA[2], accA[3]);
            *reinterpret_cast<uint4 *>(cs + k + VL / 2) = make_uint4(accA[4], accA[5], accA[6], accA[7]);
            if (second) {
                *reinterpret_cast<uint4 *>(cs + k1) = make_uint4(accB[0], accB[1], accB[2], accB[3]);
                *reinterpret_cast<uint4 *>(cs + k1 + VL / 2) = make_uint4(accB[4], accB[5], accB[6], accB[7]);
            }
        }
        if (bh != lastBh) { /* scaletbl[0], [1] (scale.c:58-59 / ScaleAddCols2) */
            scale0 = (uint32_t)(65536 / (m * bh));
            scale1 = (uint32_t)(65536 / ((m + 1) * bh));
            lastBh = bh;
        }
        __syncwarp();
        if (i < S.i1) {
            /* rolled on purpose (code size): the 8 outputs are packed into two words as they are produced */
            uint32_t w0 = 0, w1 = 0;
#pragma unroll 1
            for (int o = 0; o < VEC; ++o) {
                bool ignored = false;
                const uint32_t inf = box_info(o, ignored);
                const Acc * q = cs + (inf & 0xFFFFu);
                uint32_t sum = ((inf >> 16) & 1u) ? (uint32_t)q[m] : 0u;
                for (int c = 0; c < m; ++c) sum += (uint32_t)q[c]; /* m is warp-uniform */
                const uint32_t val = (inf >> 17) ? (((sum * (((inf >> 16) & 1u) ? scale1 : scale0)) >> 16) & (sizeof(T) == 1 ? 0xFFu : 0xFFFFu)) : 0u;
                const int sh = (int)(8 * sizeof(T)) * (o % (VEC / 2));
                if (o < VEC / 2) w0 |= val << sh;
                else w1 |= val << sh;
            }
            T * out = reinterpret_cast<T *>(J.dst + (int64_t)j * J.dstRB) + i;
            if (S.dstAligned && i + VEC <= S.i1) {
                *reinterpret_cast<uint2 *>(out) = make_uint2(w0, w1);
            } else {
#pragma unroll 1
                for (int k = 0; k < VEC; ++k)
                    if (i + k < S.i1) out[k] = (T)(((k < VEC / 2) ? w0 : w1) >> ((int)(8 * sizeof(T)) * (k % (VEC / 2))));
            }
        }
        __syncwarp();
    }
    return true;
}

/* ScalePlaneBilinearDown (scale.c:258-369): InterpolateRow of rows yi, yi + 1 into a row buffer, then ScaleFilterCols;
 * ScalePlaneSimple (scale.c:763-827) is the same walk without either filter */
template <typename T, bool POINT>
__device__ bool seg_bilinear_down(const AbScaleJob & J, const Seg & S, uint8_t * stage)
{
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T);
    const int xmax = J.srcW - 1;
    const bool arith = J.colsArith != 0;
    int yiNext = J.rows[4 * S.j0], yfNext = (POINT || J.linearOnly) ? 0 : J.rows[4 * S.j0 + 1];
    const int xFirst = arith ? (int)(col_pos(J, S.i0) >> 16) : J.cols[4 * S.i0];
    const int xLast = arith ? (int)(col_pos(J, S.i1 - 1) >> 16) : J.cols[4 * (S.i1 - 1)];
    const int c0 = min(max(xFirst, 0), xmax), cl = min(xLast + (POINT ? 1 : 2), J.srcW);
    const int a0 = S.srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    if (span <= 0 || span > kStageBytes / (int)sizeof(T) - VL || xLast < xFirst) return false;
    T * sv = reinterpret_cast<T *>(stage);
    const int i = S.i0 + (int)S.lane * VEC;
    uint32_t info[VEC]; /* fraction | staging offset << 16 | (right neighbour is one further) << 31 */
#pragma unroll
    for (int o = 0; o < VEC; ++o) {
        info[o] = 0;
        if (i + o < S.i1) {
            int xi, xf;
            if (arith) {
                const uint32_t t = col_pos(J, i + o);
                xi = (int)(t >> 16);
                xf = (int)(t & 0xFFFFu);
            } else {
                const int2 C = __ldg(reinterpret_cast<const int2 *>(J.cols + 4 * (i + o)));
                xi = C.x;
                xf = C.y & 0xFFFF;
            }
            const int xa = min(max(xi, 0), xmax);
            info[o] = (uint32_t)xf | ((uint32_t)(xa - a0) << 16) | ((xa < xmax) ? 0x80000000u : 0u);
        }
    }
    for (int j = S.j0; j < S.j1; ++j) {
        const int yi = yiNext, yf = yfNext;
        if (j + 1 < S.j1) {
            yiNext = J.rows[4 * (j + 1)];
            yfNext = (POINT || J.linearOnly) ? 0 : J.rows[4 * (j + 1) + 1];
        }
        const T * ra = src_row<T>(J, yi);
        const T * rb = src_row<T>(J, yi + 1);
        for (int k = (int)S.lane * VL; k < span; k += 64 * VL) { /* two vectors per lane and pass: every load of the row in flight at once */
            const int k1 = k + 32 * VL;
            const bool second = k1 < span;
            uint4 a[2], b[2];
            a[0] = load_vec<T>(J, ra, a0 + k, S.srcAligned);
            a[1] = second ? load_vec<T>(J, ra, a0 + k1, S.srcAligned) : make_uint4(0u, 0u, 0u, 0u);
            if (yf) {
                b[0] = load_vec<T>(J, rb, a0 + k, S.srcAligned);
                b[1] = second ? load_vec<T>(J, rb, a0 + k1, S.srcAligned) : make_uint4(0u, 0u, 0u, 0u);
                auto lerp_vec = [&](const uint4 & va, const uint4 & vb) -> uint4 {
                    uint32_t w[4] = { 0, 0, 0, 0 };
#pragma unroll
                    for (int e = 0; e < VL; ++e) {
                        const uint32_t t = (uint32_t)interp((int)vec_sample<T>(va, e), (int)vec_sample<T>(vb, e), yf);
                        if (sizeof(T) == 1) w[e >> 2] |= t << (8 * (e & 3));
                        else w[e >> 1] |= t << (16 * (e & 1));
                    }
                    return make_uint4(w[0], w[1], w[2], w[3]);
                };
                a[0] = lerp_vec(a[0], b[0]);
                if (second) a[1] = lerp_vec(a[1], b[1]);
            }
            *reinterpret_cast<uint4 *>(sv + k) = a[0];
            if (second) *reinterpret_cast<uint4 *>(sv + k1) = a[1];
        }
        __syncwarp();
        if (i < S.i1) {
            int v[VEC];
#pragma unroll
            for (int o = 0; o < VEC; ++o) {
                const int xa = (int)((info[o] >> 16) & 0x7FFFu);
                if (POINT) v[o] = (int)sv[xa];
                else v[o] = blend<T>((int)sv[xa], (int)sv[xa + (int)(info[o] >> 31)], (int)(info[o] & 0xFFFFu));
            }
            store_group<T>(J, j, i, S.i1, v, S.dstAligned);
        }
        __syncwarp();
    }
    return true;
}

/* Up-scaling loops.  The reference filters every SOURCE row horizontally once (ScaleFilterCols into the two row buffers of
 * ScalePlaneBilinearUp, scale.c:371-466; the (3,1) taps of ScaleRowUp2_(Bi)linear, scale.c:467-538 / scale_any.c:27-78) and then
 * combines two such rows per destination row.  Same here: the lane keeps the horizontally filtered values of its 8 bytes of
 * output for two source rows in registers (two slots tagged with their row number), re-filters a slot only when the walk down
 * the destination rows needs a row neither slot holds, and each destination row costs one vertical combination per sample.
 *   KIND 0  bilinear up     H = BLENDER(row[xi], row[xi+1], xf);  out = InterpolateRow(H(ra), H(rb), yf)
 *   KIND 1  2x bilinear     H = 3 row[near] + row[far];           out = (3 H(near row) + H(far row) + 8) >> 4
 *   KIND 2  2x linear       H = 3 row[near] + row[far];           out = (H(row) + 2) >> 2
 * (edge columns of the 2x loops use one source column: near == far gives 4 row[k], and (12 a + 4 b + 8) >> 4 == (3 a + b + 2) >> 2,
 * (4 a + 2) >> 2 == a.) */
template <typename T, int KIND>
__device__ bool seg_up(const AbScaleJob & J, const Seg & S, uint8_t * stage)
{
    constexpr int VL = Vec16<T>::N, VEC = 8 / (int)sizeof(T);
    const int xmax = J.srcW - 1;
    int c0, cl;
    if (KIND == 0) {
        const int xFirst = J.colsArith ? (int)(col_pos(J, S.i0) >> 16) : J.cols[4 * S.i0];
        const int xLast = J.colsArith ? (int)(col_pos(J, S.i1 - 1) >> 16) : J.cols[4 * (S.i1 - 1)];
        if (xLast < xFirst) return false;
        c0 = min(max(xFirst, 0), xmax);
        cl = min(xLast + 2, J.srcW);
    } else {
        c0 = max((S.i0 - 1) >> 1, 0);
        cl = min(max(((S.i1 - 2) >> 1) + 2, 1), J.srcW);
        if (S.i1 == J.dstW) cl = max(cl, min((J.dstW - 1) / 2 + 1, J.srcW)); /* the last column's own source column */
    }
    const int a0 = S.srcAligned ? (c0 & ~(VL - 1)) : c0;
    const int span = cl - a0;
    if (span <= 0 || span > kStageBytes / (int)sizeof(T) - VL) return false;
    T * st = reinterpret_cast<T *>(stage);
    const int i = S.i0 + (int)S.lane * VEC;
    uint32_t info[VEC]; /* KIND 0: fraction | offset << 16 | (neighbour one further) << 31;  else near offset | far offset << 16 */
#pragma unroll
    for (int o = 0; o < VEC; ++o) {
        info[o] = 0;
        const int x = i + o;
        if (x < S.i1) {
            if (KIND == 0) {
                int xi, xf;
                if (J.colsArith) {
                    const uint32_t t = col_pos(J, x);
                    xi = (int)(t >> 16);
                    xf = (int)(t & 0xFFFFu);
                } else {
                    const int2 C = __ldg(reinterpret_cast<const int2 *>(J.cols + 4 * x));
                    xi = C.x;
                    xf = C.y & 0xFFFF;
                }
                const int xa = min(max(xi, 0), xmax);
                info[o] = (uint32_t)xf | ((uint32_t)(xa - a0) << 16) | ((xa < xmax) ? 0x80000000u : 0u);
            } else if (x == 0 || x == J.dstW - 1) {
                const int k = ((x == 0) ? 0 : (J.dstW - 1) / 2) - a0;
                info[o] = (uint32_t)k | ((uint32_t)k << 16);
            } else {
                const int k = (x - 1) >> 1, k1 = min(k + 1, xmax);
                const int kn = ((x - 1) & 1) ? k1 : k, kf = ((x - 1) & 1) ? k : k1; /* near / far column */
                info[o] = (uint32_t)(kn - a0) | ((uint32_t)(kf - a0) << 16);
            }
        }
    }
    int h0[VEC], h1[VEC]; /* the two slots */
    int id0 = -1, id1 = -1;
#pragma unroll
    for (int o = 0; o < VEC; ++o) h0[o] = h1[o] = 0;
    /* the walk goes down the source rows one at a time: right after a row has been filtered the lane asks for its 16 bytes of the
     * NEXT row (one vector per lane covers the span of an up-scaling segment), so that the load is long done when that row is
     * needed -- a destination row then costs arithmetic only */
    const bool canPre = span <= 32 * VL;
    uint4 pre = make_uint4(0u, 0u, 0u, 0u);
    int preRow = -1;
    for (int j = S.j0; j < S.j1; ++j) {
        int p, q, yf = 0; /* rows of the first / second operand */
        if (KIND == 0) {
            p = min(max(J.rows[4 * j], 0), J.srcH - 1);
            yf = J.linearOnly ? 0 : J.rows[4 * j + 2];
            q = yf ? min(max(J.rows[4 * j + 1], 0), J.srcH - 1) : p;
        } else if (KIND == 1) { /* scale.c:503-538: row 0 and the last row of an even height come from one source row */
            p = q = 0;
            if (j != 0) {
                const int r = (j - 1) >> 1;
                if (r >= J.srcH - 1) p = q = J.srcH - 1;
                else if ((j - 1) & 1) { p = r + 1; q = r; }
                else { p = r; q = r + 1; }
            }
        } else {
            p = q = min(max(J.rows[4 * j], 0), J.srcH - 1);
        }
        /* which slot holds p, which q (all warp-uniform); a missing row replaces the slot the other operand does not need.  One
         * rolled loop over the two operands: the filter below exists once in the code */
        int sp = 0, sq = 0;
#pragma unroll 1
        for (int op = 0; op < 2; ++op) {
            const int r = op ? q : p;
            int slot;
            if (op == 1 && q == p) slot = sp;
            else if (id0 == r) slot = 0;
            else if (id1 == r) slot = 1;
            else {
                slot = op ? (1 - sp) : ((id0 == q) ? 1 : 0);
                if (canPre && preRow == r) {
                    if ((int)S.lane * VL < span) *reinterpret_cast<uint4 *>(st + (int)S.lane * VL) = pre;
                } else {
                    stage_raw<T>(J, r, a0, span, S.srcAligned, st, S.lane);
                }
                __syncwarp();
                auto hfilter = [&](int o) -> int {
                    if (KIND == 0) {
                        const int xa = (int)((info[o] >> 16) & 0x7FFFu);
                        return blend<T>((int)st[xa], (int)st[xa + (int)(info[o] >> 31)], (int)(info[o] & 0xFFFFu));
                    }
                    return 3 * (int)st[info[o] & 0xFFFFu] + (int)st[info[o] >> 16];
                };
                if (slot == 0) { /* warp-uniform: two copies of the loop instead of a select per output */
#pragma unroll
                    for (int o = 0; o < VEC; ++o) h0[o] = hfilter(o);
                } else {
#pragma unroll
                    for (int o = 0; o < VEC; ++o) h1[o] = hfilter(o);
                }
                __syncwarp();
                if (slot == 0) id0 = r;
                else id1 = r;
            }
            if (op == 0) sp = slot;
            else sq = slot;
        }
        if (canPre) { /* ask for the row below everything the slots hold, once */
            const int nxt = max(id0, id1) + 1;
            if (nxt < J.srcH && nxt != preRow) {
                preRow = nxt;
                if ((int)S.lane * VL < span) pre = load_vec<T>(J, src_row<T>(J, nxt), a0 + (int)S.lane * VL, S.srcAligned);
            }
        }
        if (i < S.i1) {
            int v[VEC];
            /* the vertical combination, written out per (slot of p, slot of q) and per class of the row's fraction -- all warp-uniform
             * -- so that the 8 outputs are straight arithmetic without selects or per-sample branches */
            auto combine = [&](const int (&a)[VEC], const int (&b)[VEC]) {
                if (KIND == 0) { /* InterpolateRow_C, row_common.c:46-105 */
                    if (yf == 0) {
#pragma unroll
                        for (int o = 0; o < VEC; ++o) v[o] = a[o];
                    } else if (yf == 128) {
#pragma unroll
                        for (int o = 0; o < VEC; ++o) v[o] = (a[o] + b[o] + 1) >> 1;
                    } else {
#pragma unroll
                        for (int o = 0; o < VEC; ++o) v[o] = (a[o] * (256 - yf) + b[o] * yf + 128) >> 8;
                    }
                } else if (KIND == 1) {
#pragma unroll
                    for (int o = 0; o < VEC; ++o) v[o] = (3 * a[o] + b[o] + 8) >> 4;
                } else {
#pragma unroll
                    for (int o = 0; o < VEC; ++o) v[o] = (a[o] + 2) >> 2;
                }
            };
            if (sp == 0 && sq == 1) combine(h0, h1);
            else if (sp == 1 && sq == 0) combine(h1, h0);
            else if (sp == 0) combine(h0, h0);
            else combine(h1, h1);
            store_group<T>(J, j, i, S.i1, v, S.dstAligned);
        }
    }
    return true;
}

/* ScalePlaneVertical (scale_common.c:348-426): InterpolateRow straight from the two source rows, no staging */
template <typename T>
__device__ void seg_vertical(const AbScaleJob & J, const Seg & S)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    const int i = S.i0 + (int)S.lane * VEC;
    if (i >= S.i1) return;
    for (int j = S.j0; j < S.j1; ++j) {
        const int yi = J.rows[4 * j], yf = J.rows[4 * j + 1];
        const T * pa = src_row<T>(J, yi);
        const T * pb = src_row<T>(J, yi + 1);
        int v[VEC];
        if (S.srcAligned && i + VEC <= S.i1) { /* 16-byte aligned rows are 8-byte aligned too; i is a multiple of VEC */
            const uint2 a = __ldg(reinterpret_cast<const uint2 *>(pa + i));
            const uint2 b = yf ? __ldg(reinterpret_cast<const uint2 *>(pb + i)) : make_uint2(0u, 0u);
            const uint4 va = make_uint4(a.x, a.y, 0u, 0u), vb = make_uint4(b.x, b.y, 0u, 0u);
#pragma unroll
            for (int o = 0; o < VEC; ++o) v[o] = interp((int)vec_sample<T>(va, o), (int)vec_sample<T>(vb, o), yf);
        } else {
#pragma unroll
            for (int o = 0; o < VEC; ++o) v[o] = (i + o < S.i1) ? interp((int)pa[i + o], yf ? (int)pb[i + o] : 0, yf) : 0;
        }
        store_group<T>(J, j, i, S.i1, v, S.dstAligned);
    }
}

/* the table-driven sampler for rows ja..jb of a segment (any loop): the fallback, sample by sample */
template <typename T>
__device__ __noinline__ void rows_generic(const AbScaleJob & J, const Seg & S, int ja, int jb)
{
    constexpr int VEC = 8 / (int)sizeof(T);
    const int i = S.i0 + (int)S.lane * VEC;
    for (int j = ja; j < jb; ++j) {
        const int32_t * R = J.rows + 4 * j;
        T * out = reinterpret_cast<T *>(J.dst + (int64_t)j * J.dstRB);
#pragma unroll 1
        for (int k = i; k < min(i + VEC, S.i1); ++k) out[k] = (T)scale_sample<T, 0>(J, R, k, j);
    }
}

/* All planes of one image in one launch.  Blocks are shared out over the planes by blockStart[]; the warps of a plane's blocks
 * walk its units: (segment of segLen <= 32 x 8 bytes, block of rowsPerUnit destination rows), segments fastest. */
template <typename T>
__global__ void __launch_bounds__(kScaleThreads, 3) scale_batch_kernel(const __grid_constant__ AbScaleBatch B)
{
    __shared__ __align__(16) uint8_t stageAll[kScaleWarps][kStageBytes];
    const uint32_t warp = threadIdx.x >> 5;
    uint8_t * stage = stageAll[warp];
    uint32_t p = 0;
    while (p + 1 < B.count && blockIdx.x >= B.blockStart[p + 1]) ++p;
    const AbScaleJob & J = B.job[p];
    const uint32_t segs = (uint32_t)((J.dstW + J.segLen - 1) / J.segLen);
    const uint32_t rowBlocks = (uint32_t)((J.dstH + J.rowsPerUnit - 1) / J.rowsPerUnit);
    const uint32_t total = segs * rowBlocks; /* < 2^31: the launcher checked */
    const uint32_t stride = (B.blockStart[p + 1] - B.blockStart[p]) * kScaleWarps;
    Seg S;
    S.lane = threadIdx.x & 31;
    S.dstAligned = ((((uintptr_t)J.dst) | (uintptr_t)J.dstRB) & 7u) == 0;
    S.srcAligned = ((((uintptr_t)J.src) | (uintptr_t)J.srcRB) & 15u) == 0;
    for (uint32_t u = (blockIdx.x - B.blockStart[p]) * kScaleWarps + warp; u < total; u += stride) {
        const uint32_t rb = u / segs, sg = u - rb * segs;
        S.i0 = (int)sg * J.segLen;
        S.i1 = min(J.dstW, S.i0 + J.segLen);
        S.j0 = (int)rb * J.rowsPerUnit;
        S.j1 = min(J.dstH, S.j0 + J.rowsPerUnit);
        bool done = false; /* warp-uniform */
        switch (J.mode) {
            case AB_SCALE_BOX: done = seg_box<T>(J, S, stage); break;
            case AB_SCALE_BILINEAR_DOWN: done = seg_bilinear_down<T, false>(J, S, stage); break;
            case AB_SCALE_SIMPLE: done = seg_bilinear_down<T, true>(J, S, stage); break;
            case AB_SCALE_BILINEAR_UP: done = seg_up<T, 0>(J, S, stage); break;
            case AB_SCALE_UP2_BILINEAR: done = seg_up<T, 1>(J, S, stage); break;
            case AB_SCALE_UP2_LINEAR: done = seg_up<T, 2>(J, S, stage); break;
            case AB_SCALE_VERTICAL: seg_vertical<T>(J, S); done = true; break;
            default: break;
        }
        if (!done) rows_generic<T>(J, S, S.j0, S.j1);
    }
}

} // namespace

extern "C" int abLaunchScaleBatch(AbScaleBatch * batch, void * stream)
{
    if (batch->count == 0) return 0;
    /* units = (segment, block of destination rows).  A segment is at most one pass of a warp (32 lanes x 8 bytes), rows are cut
     * into equal segments, and a segment's source span has to fit the warp's staging slice (else the unit falls back to the
     * sampler).  Up-scaling loops reuse filtered source rows down a block of rows; the blocks shrink until there are enough units
     * to fill the device a few times over. */
    int64_t units[AB_SCALE_MAX_PLANES], work[AB_SCALE_MAX_PLANES], all = 0;
    const int vec = 8 / batch->job[0].bps;
    const int64_t residentWarps = 148 * 3 * kScaleWarps;
    int64_t outputs = 0;
    for (uint32_t p = 0; p < batch->count; ++p) outputs += (int64_t)batch->job[p].dstW * batch->job[p].dstH;
    for (uint32_t p = 0; p < batch->count; ++p) {
        AbScaleJob * j = &batch->job[p];
        int maxSeg = 32 * vec;
        if (j->dstW > 0 && j->srcW > j->dstW) { /* down-scaling: keep the span (~ seg * srcW / dstW columns) inside the slice */
            const int64_t elem = (j->mode == AB_SCALE_BOX) ? 2 * j->bps : j->bps;
            const int64_t cap = kStageBytes / elem - 48;
            while (maxSeg > vec && ((int64_t)maxSeg * j->srcW + j->dstW - 1) / j->dstW + 4 > cap) maxSeg /= 2;
        }
        const int nseg = (j->dstW + maxSeg - 1) / maxSeg;
        int seg = (j->dstW + nseg - 1) / (nseg > 0 ? nseg : 1);
        seg = (seg + vec - 1) / vec * vec;
        if (seg < vec) seg = vec;
        j->segLen = seg;
        const int64_t segs = (j->dstW + seg - 1) / seg;
        const bool up = (j->mode == AB_SCALE_BILINEAR_UP || j->mode == AB_SCALE_UP2_BILINEAR || j->mode == AB_SCALE_UP2_LINEAR);
        int rb = up ? 16 : 2;
        /* this plane's share of the resident warps, two units each */
        const int64_t wantUnits = 2 * residentWarps * ((int64_t)j->dstW * j->dstH) / (outputs > 0 ? outputs : 1);
        while (rb > 1 && segs * ((j->dstH + rb - 1) / rb) < wantUnits) rb /= 2;
        j->rowsPerUnit = rb;
        units[p] = segs * ((j->dstH + rb - 1) / rb);
        if (units[p] >= (1ll << 31)) return (int)cudaErrorInvalidValue;
        work[p] = (int64_t)j->dstW * j->dstH + (int64_t)j->dstH * (j->srcW > j->dstW ? j->srcW : 0); /* ~ bytes touched */
        if (work[p] < 1) work[p] = 1;
        all += work[p];
    }
    const int64_t budget = 148 * 3 * 2; /* two rounds of resident blocks (3 per SM): the second hides the ragged end of the first */
    uint32_t start = 0;
    for (uint32_t p = 0; p < batch->count; ++p) {
        int64_t share = (work[p] * budget + all - 1) / all;
        const int64_t need = (units[p] + kScaleWarps - 1) / kScaleWarps;
        if (share > need) share = need;
        if (share < 1) share = 1;
        batch->blockStart[p] = start;
        start += (uint32_t)share;
    }
    batch->blockStart[batch->count] = start;
    if (batch->job[0].bps == 1) scale_batch_kernel<uint8_t><<<start, kScaleThreads, 0, (cudaStream_t)stream>>>(*batch);
    else scale_batch_kernel<uint16_t><<<start, kScaleThreads, 0, (cudaStream_t)stream>>>(*batch);
    return (int)cudaGetLastError();
}
