/*
 * r2y_fast.cuh -- compile-time specialised RGB->YUV kernels for the hot configurations (coefficients mode,
 * colour input, verified divisors, 16-byte aligned buffers); everything else runs on kernels.cu's r2y_block_kernel.
 *
 * A lane owns 8 consecutive pixels of a 256-pixel warp tile -- of TWO rows for 4:2:0, so that the reference's 2x2
 * block walk (src/reformat.c:295-470) stays inside one thread: the box filter is an ordered register sum, no
 * shuffle needed.  Interleaved pixels are read with 128-bit loads (64-bit for 3-byte pixels), planes are written
 * with one 128/64-bit store per lane and row, all lane-contiguous.  Arithmetic: pixel_math.cuh, operator for
 * operator the reference's.
 */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "pixel_math.cuh"
#include "plan.h"

namespace abfast {

using namespace ab;

enum { SUB_444 = 0, SUB_422 = 1, SUB_420 = 2, SUB_400 = 3 };

template <typename RT_, typename YT_, int SUB_, int NCH_, bool AFIRST_, int AM_, bool IDENT_ = false>
struct RCfg {
    static constexpr bool IDENT = IDENT_; /* identity matrix (lossless GBR): Y = G, U = B, V = R, all quantised with the luma range (src/reformat.c:361-365, :212-213) */
    using RT = RT_;
    using YT = YT_;
    static constexpr int SUB = SUB_;
    static constexpr int NCH = NCH_;
    static constexpr bool AFIRST = AFIRST_;
    static constexpr int AM = AM_; /* AB_MUL_NONE / FLOAT_MUL / FLOAT_UNMUL */
};

/* Loop shape per tile class, from the sweeps in profiles/r01_variants.md (tools/variant_sweep.sh; -D switches for
 * `make variant`).  The chroma box sums are accumulated as the pixels are produced, so even the two-row 4:2:0 tiles fit
 * 80 registers:
 *   one-row tiles (4:4:4, 4:2:2, 4:0:0)   prefetched second register set, 4 blocks/SM (64 registers) -- 7 % faster than 3;
 *                                         5 blocks/SM spill (+60 %)
 *   two-row tiles, 8-bit channels         prefetch, 3 blocks/SM   -- 10 % faster than 2; plain loop at 4 measured equal
 *   two-row tiles, 16-bit channels        3 blocks/SM (-13 % on 8K RGBA16); RGBA16 runs the plain loop: its tile alone is 32
 *                                         registers of input per lane and a second set does not fit */
#ifndef AB_R2Y_UNROLL2 /* 1: the prefetching loop written out twice (ping-pong register sets, no copies) */
#define AB_R2Y_UNROLL2 0
#endif
#ifndef AB_R2Y_NOPREFETCH420
#define AB_R2Y_NOPREFETCH420 0 /* 1: plain loop for every two-row tile */
#endif
constexpr bool kR2YNoPrefetch420 = (AB_R2Y_NOPREFETCH420 != 0);
#ifndef AB_R2Y_PLAIN444
#define AB_R2Y_PLAIN444 0
#endif
#ifndef AB_R2Y_MINB444
#define AB_R2Y_MINB444 4
#endif
#ifndef AB_R2Y_MINB420_8
#define AB_R2Y_MINB420_8 3
#endif
#ifndef AB_R2Y_MINB420_16
#define AB_R2Y_MINB420_16 3
#endif
#ifndef AB_R2Y_PLAIN420_8
#define AB_R2Y_PLAIN420_8 0
#endif
template <class C>
struct RLoop {
    static constexpr bool SMALL420 = (C::SUB == SUB_420) && (sizeof(typename C::RT) == 1);
    static constexpr bool PLAIN = SMALL420 ? (AB_R2Y_PLAIN420_8 != 0)
                                           : ((C::SUB == SUB_420) ? (kR2YNoPrefetch420 || (sizeof(typename C::RT) == 2 && C::NCH == 4)) : (AB_R2Y_PLAIN444 != 0));
    static constexpr int MIN_BLOCKS = SMALL420 ? AB_R2Y_MINB420_8 : ((C::SUB == SUB_420) ? AB_R2Y_MINB420_16 : AB_R2Y_MINB444);
};
constexpr int kRTile = 256;
constexpr int kRWarps = 8;

/* Instruction diet of round 2 (each switch = one exact rewrite; 0 restores the round-1 form for A/B timing):
 *   AB_R2Y_QMERGE  the rounding constant of the quantiser folded into the bias: floor(RN(p + (bias + 0.5))) equals the reference's
 *                  floor(RN(RN(p + bias) + 0.5)) after the clamp for EVERY binary32 p = RN(v*range) and every (range, bias) pair of the
 *                  path (tools/verify_quant.c `quant`: 16 pairs x 2^32 values, bad = 0) -- one FADD less per sample;
 *   AB_R2Y_DIV2    x / d for the verified chroma divisors as fma(x, rh, RN(x*rl)) (AbDiv.exact2: tools/verify_quant.c `div`, all 2^32
 *                  numerators per divisor) -- one operation less than Markstein's form;
 *   (no switch)    R/B order and alpha copy-vs-fill resolved by warp-uniform branches around the tile body instead of per-pixel selects;
 *   AB_R2Y_LIM     unmultiply: the three `a < 1 ? q : c` selects become one clamp limit (1, or +inf when a >= 1). */
#ifndef AB_R2Y_QMERGE
#define AB_R2Y_QMERGE 1
#endif
#ifndef AB_R2Y_DIV2
#define AB_R2Y_DIV2 1
#endif
#ifndef AB_R2Y_LIM
#define AB_R2Y_LIM 1
#endif

__device__ __forceinline__ float rdivx(float x, const AbDiv & dv)
{
#if AB_R2Y_DIV2
    return __fmaf_rn(x, dv.rh, __fmul_rn(x, dv.rl)); /* launcher guarantees dv.exact2 */
#else
    const float q = __fmul_rn(x, dv.r);
    return __fmaf_rn(__fmaf_rn(-q, dv.d, x), dv.r, q);
#endif
}

/* channel value `v` (memory order, 0 .. 8*NCH-1) of the lane's packed pixels as a magic float 2^23 + code */
template <typename RT>
__device__ __forceinline__ float chan_magic(const uint32_t * w, int v, uint32_t K)
{
    if (sizeof(RT) == 2) return __uint_as_float(__byte_perm(w[v >> 1], K, (v & 1) ? 0x7432u : 0x7410u));
    return __uint_as_float(__byte_perm(w[v >> 2], K, 0x7440u | (uint32_t)(v & 3)));
}

/* ref: src/reformat.c:197-219: clamp((int)floorf((v*range + bias) + 0.5f), 0, yuvMax), returned as 2^23 + code */
__device__ __forceinline__ float quant_yuv_magic(float v, float range, float bias, float hiLimit)
{
#if AB_R2Y_QMERGE
    const float t = __fadd_rn(__fmul_rn(v, range), bias); /* bias + 0.5, see AB_R2Y_QMERGE */
#else
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(v, range), bias), 0.5f);
#endif
    return __fadd_rz(fminf(fmaxf(t, 0.0f), hiLimit), 8388608.0f); /* floor+clamp == trunc(clamp(t, 0, max + 0.5)) */
}

/* luma: every term of Y = kr*R + kg*G + kb*B is >= 0 for the matrices the specialised kernels accept (their divisors are
 * whitelisted, all with kr, kg, kb >= 0), so t >= bias + 0.5 > 0 and only the upper clamp can act */
__device__ __forceinline__ float quant_luma_magic(float v, float range, float bias, float hiLimit)
{
#if AB_R2Y_QMERGE
    const float t = __fadd_rn(__fmul_rn(v, range), bias); /* bias + 0.5, see AB_R2Y_QMERGE */
#else
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(v, range), bias), 0.5f);
#endif
    return __fadd_rz(fminf(t, hiLimit), 8388608.0f);
}

#ifndef AB_R2Y_SATCVT /* tuning switch: 0 = clamp in float (2 FMNMX) and take the code from FADD.RZ's mantissa */
#define AB_R2Y_SATCVT 1
#endif
/* The same quantisation through the saturating float -> integer conversion: cvt.rzi truncates (== floor for t >= 0), maps
 * negatives and NaN to 0 and saturates at the container maximum -- the reference's clamp(floor(t), 0, max) for 8-bit samples;
 * 16-bit samples of depth 10/12 still need min(code, yuvMax), which the store does on packed pairs (CLAMP below).
 * Returned as a bit container (code in the low half, upper half zero) so that both forms pack alike. */
template <typename YT>
__device__ __forceinline__ float quant_code_sat(float v, float range, float bias)
{
#if AB_R2Y_QMERGE
    const float t = __fadd_rn(__fmul_rn(v, range), bias); /* `bias` is bias + 0.5 here (see AB_R2Y_QMERGE) */
#else
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(v, range), bias), 0.5f);
#endif
    uint32_t c;
    if (sizeof(YT) == 1) {
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(c) : "f"(t));
    } else {
        unsigned short h;
        asm("cvt.rzi.u16.f32 %0, %1;" : "=h"(h) : "f"(t));
        c = h;
    }
    return __uint_as_float(c);
}
template <typename YT>
__device__ __forceinline__ float quant_yuv(float v, float range, float bias, float hiLimit)
{
    return AB_R2Y_SATCVT ? quant_code_sat<YT>(v, range, bias) : quant_yuv_magic(v, range, bias, hiLimit);
}
template <typename YT>
__device__ __forceinline__ float quant_luma(float v, float range, float bias, float hiLimit)
{
    return AB_R2Y_SATCVT ? quant_code_sat<YT>(v, range, bias) : quant_luma_magic(v, range, bias, hiLimit);
}
/* bias == 0 (full-range luma): floor(RN(fma(v, range, 0.5))) equals the reference's floor((v*range + 0) + 0.5) after the clamp for
 * every binary32 v (tools/verify_quant.c `quant`, the "single fma" column: bad = 0 exactly for the four bias-0 pairs) */
template <typename YT>
__device__ __forceinline__ float quant_luma_fullrange(float v, float range)
{
    const float t = __fmaf_rn(v, range, 0.5f);
    uint32_t c;
    if (sizeof(YT) == 1) {
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(c) : "f"(t));
    } else {
        unsigned short h;
        asm("cvt.rzi.u16.f32 %0, %1;" : "=h"(h) : "f"(t));
        c = h;
    }
    return __uint_as_float(c);
}

/* CLAMP: the containers hold saturated 16-bit codes that still have to be limited to yuvMax (mm = yuvMax in both halves) */
template <typename YT, bool CLAMP = false>
__device__ __forceinline__ void store_samples8(uint8_t * row, uint32_t x, uint32_t n, const float q[8], uint32_t mm = 0, bool vec = true)
{
    if (n >= 8 && vec) {
        if (sizeof(YT) == 2) {
            uint4 v;
            v.x = __byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x5410u);
            v.y = __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x5410u);
            v.z = __byte_perm(__float_as_uint(q[4]), __float_as_uint(q[5]), 0x5410u);
            v.w = __byte_perm(__float_as_uint(q[6]), __float_as_uint(q[7]), 0x5410u);
            if (CLAMP) {
                v.x = __vminu2(v.x, mm); v.y = __vminu2(v.y, mm); v.z = __vminu2(v.z, mm); v.w = __vminu2(v.w, mm);
            }
            *reinterpret_cast<uint4 *>(row + (size_t)x * 2) = v;
        } else {
            uint2 v;
            v.x = __byte_perm(__byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x0040u), __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x0040u), 0x5410u);
            v.y = __byte_perm(__byte_perm(__float_as_uint(q[4]), __float_as_uint(q[5]), 0x0040u), __byte_perm(__float_as_uint(q[6]), __float_as_uint(q[7]), 0x0040u), 0x5410u);
            *reinterpret_cast<uint2 *>(row + x) = v;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t c = __float_as_uint(q[k]) & 0xFFFFu;
            if (CLAMP) c = min(c, mm & 0xFFFFu);
            if ((uint32_t)k < n) reinterpret_cast<YT *>(row)[x + k] = (YT)c;
        }
    }
}

template <typename YT, bool CLAMP = false>
__device__ __forceinline__ void store_samples4(uint8_t * row, uint32_t x, uint32_t n, const float q[4], uint32_t mm = 0, bool vec = true)
{
    if (n >= 4 && vec) {
        if (sizeof(YT) == 2) {
            uint2 v;
            v.x = __byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x5410u);
            v.y = __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x5410u);
            if (CLAMP) {
                v.x = __vminu2(v.x, mm); v.y = __vminu2(v.y, mm);
            }
            *reinterpret_cast<uint2 *>(row + (size_t)x * 2) = v;
        } else {
            *reinterpret_cast<uint32_t *>(row + x) =
                __byte_perm(__byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x0040u), __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x0040u), 0x5410u);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t c = __float_as_uint(q[k]) & 0xFFFFu;
            if (CLAMP) c = min(c, mm & 0xFFFFu);
            if ((uint32_t)k < n) reinterpret_cast<YT *>(row)[x + k] = (YT)c;
        }
    }
}

template <class C>
struct RTileIn {
    static constexpr int ROWS = (C::SUB == SUB_420) ? 2 : 1;
    static constexpr int NW = 8 * C::NCH * (int)sizeof(typename C::RT) / 4; /* words per lane and row */
    uint32_t w[ROWS][NW];
};

/* lane-specific addresses of one tile: the lane's first pixel in the RGB buffer, its first sample in every plane row of the
 * tile's first luma row (subsampled chroma: of the chroma row the tile's first luma row maps to) */
struct RPtrs {
    const uint8_t * rgb;
    uint8_t * y;
    uint8_t * u;
    uint8_t * v;
    uint8_t * a;
};

template <class C>
__device__ __forceinline__ RPtrs rtile_ptrs(const AbR2YFrame & F, uint32_t x0, uint32_t j0, uint32_t lane)
{
    using RT = typename C::RT;
    using YT = typename C::YT;
    constexpr int PIXB = C::NCH * (int)sizeof(RT);
    const uint32_t px0 = x0 + 8 * lane;
    RPtrs P;
    P.rgb = F.rgb + (int64_t)j0 * F.rgbRB + (int64_t)px0 * PIXB;
    P.y = F.y + (int64_t)j0 * F.yRB + (size_t)px0 * sizeof(YT);
    P.a = F.a ? F.a + (int64_t)j0 * F.aRB + (size_t)px0 * sizeof(YT) : nullptr;
    P.u = P.v = nullptr;
    if (C::SUB == SUB_444) {
        P.u = F.u + (int64_t)j0 * F.uRB + (size_t)px0 * sizeof(YT);
        P.v = F.v + (int64_t)j0 * F.vRB + (size_t)px0 * sizeof(YT);
    } else if (C::SUB != SUB_400) {
        const uint32_t uj = (C::SUB == SUB_420) ? (j0 >> 1) : j0;
        P.u = F.u + (int64_t)uj * F.uRB + (size_t)(px0 >> 1) * sizeof(YT);
        P.v = F.v + (int64_t)uj * F.vRB + (size_t)(px0 >> 1) * sizeof(YT);
    }
    return P;
}

/* tile = 256 pixels x ROWS rows; `rgb` = the lane's first pixel of the first row, nValid = pixels of the lane inside the
 * picture (8 for FULL tiles), nRows = rows of the tile inside the picture */
template <class C, bool FULL>
__device__ __forceinline__ void load_rtile_p(const AbR2YFrame & F, const uint8_t * rgb, uint32_t nValid, uint32_t nRows, RTileIn<C> & in)
{
    using RT = typename C::RT;
    constexpr int ROWS = RTileIn<C>::ROWS;
    constexpr int PIXB = C::NCH * (int)sizeof(RT);
    constexpr int NW = RTileIn<C>::NW;
    constexpr bool fullTile = FULL;
    /* ---- loads: 8 pixels x ROWS rows, issued before any arithmetic ---- */
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
#pragma unroll
        for (int k = 0; k < NW; ++k)
            if (!FULL || r > 0) in.w[r][k] = 0; /* row 0 of a full tile is loaded unconditionally */
        if (r == 0 || (uint32_t)r < nRows) { /* the tile's first row always exists */
            const uint8_t * p = rgb + (int64_t)r * F.rgbRB;
            if (fullTile || (nValid >= 8 && F.vecWidth != 0)) { /* rows that are only sample-aligned: element by element */
                if ((8 * PIXB) % 16 == 0) {
#pragma unroll
                    for (int q = 0; q < NW / 4; ++q) {
                        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p) + q);
                        in.w[r][4 * q] = v.x; in.w[r][4 * q + 1] = v.y; in.w[r][4 * q + 2] = v.z; in.w[r][4 * q + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < NW / 2; ++q) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p) + q);
                        in.w[r][2 * q] = v.x; in.w[r][2 * q + 1] = v.y;
                    }
                }
            } else {
#pragma unroll
                for (int v = 0; v < 8 * C::NCH; ++v) {
                    if ((uint32_t)(v / C::NCH) < nValid) {
                        const uint32_t s = (uint32_t)__ldg(reinterpret_cast<const RT *>(p) + v);
                        if (sizeof(RT) == 2) in.w[r][v >> 1] |= s << ((v & 1) * 16);
                        else in.w[r][v >> 2] |= s << ((v & 3) * 8);
                    }
                }
            }
        }
    }
}

template <class C, bool FULL>
__device__ __forceinline__ void load_rtile(const AbR2YJob & J, const AbR2YFrame & F, uint32_t x0, uint32_t j0, uint32_t lane, RTileIn<C> & in)
{
    constexpr int ROWS = RTileIn<C>::ROWS;
    const uint32_t px0 = x0 + 8 * lane;
    const uint32_t nValid = FULL ? 8u : ((px0 < F.width) ? min(8u, F.width - px0) : 0u);
    const uint32_t nRows = min((uint32_t)ROWS, F.height - j0);
    const RPtrs P = rtile_ptrs<C>(F, x0, j0, lane);
    load_rtile_p<C, FULL>(F, P.rgb, nValid, nRows, in);
}

/* SWAP: B precedes R in memory; COPYA: the alpha plane receives the pixel's alpha (else it is filled with yuvMax) -- both are
 * call constants, resolved by the warp-uniform dispatch in convert_store_rtile so that the pixel loop carries no selects */
template <class C, bool FULL, bool SWAP, bool COPYA, bool YFULL>
__device__ __forceinline__ void convert_store_rtile_p(const AbR2YJob & J, const AbR2YFrame & F, uint32_t px0, uint32_t j0, uint32_t nValidIn, uint32_t nRows,
                                                      const RTileIn<C> & in)
{
    using RT = typename C::RT;
    using YT = typename C::YT;
    constexpr int ROWS = RTileIn<C>::ROWS;
    constexpr bool fullTile = FULL;
    const uint32_t nValid = FULL ? 8u : nValidIn;
    if (nValid == 0) return; /* no cross-lane traffic in this kernel */

    /* ---- per pixel: normalise, alpha, matrix (ref: src/reformat.c:311-387) ---- */
    const uint32_t K = J.magicK;
    const bool writeAlpha = (J.alphaWrite != AB_ALPHA_KEEP);
    const float magicOne = 8388608.0f;
    const float hiLimit = __fadd_rn(J.yuvMaxF, 0.5f);
    constexpr bool QC = (AB_R2Y_SATCVT != 0) && (sizeof(YT) == 2); /* saturated 16-bit codes: min with yuvMax at the store */
    const uint32_t mm = J.yuvMax * 0x10001u;
    /* the quantiser's additive constant: bias, or bias + 0.5 with the rounding constant folded in (exact, AB_R2Y_QMERGE) */
    const float qbY = AB_R2Y_QMERGE ? __fadd_rn(J.biasY, 0.5f) : J.biasY;
    const float qbUV = AB_R2Y_QMERGE ? __fadd_rn(J.biasUV, 0.5f) : J.biasUV;
    float su[4], sv[4]; /* subsampled chroma: running sums of each 2x2 (4:2:0) / 2x1 (4:2:2) box, in the reference's order */
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if ((uint32_t)r >= nRows) continue;
        float yq[8], aq[8], uq[8], vq[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            constexpr int c0 = C::AFIRST ? 1 : 0;
            const float x0n = div_int(__fsub_rn(chan_magic<RT>(in.w[r], t * C::NCH + c0, K), magicOne), J.rgbMaxDiv);
            const float gg = div_int(__fsub_rn(chan_magic<RT>(in.w[r], t * C::NCH + c0 + 1, K), magicOne), J.rgbMaxDiv);
            const float x2n = div_int(__fsub_rn(chan_magic<RT>(in.w[r], t * C::NCH + c0 + 2, K), magicOne), J.rgbMaxDiv);
            float R = SWAP ? x2n : x0n, G = gg, B = SWAP ? x0n : x2n;
            float am = 0.0f;
            if (C::NCH == 4 && (COPYA || C::AM != AB_MUL_NONE)) am = chan_magic<RT>(in.w[r], t * C::NCH + (C::AFIRST ? 0 : 3), K);
            if (C::AM != AB_MUL_NONE) { /* ref: :325-358 */
                const float a = div_int(__fsub_rn(am, magicOne), J.rgbMaxDiv);
                if (C::AM == AB_MUL_FLOAT_MUL) {
                    /* a >= 1 leaves the colour untouched (x*1 == x), a == 0 zeroes it through the multiply */
                    const float am1 = fminf(a, 1.0f);
                    R = __fmul_rn(R, am1);
                    G = __fmul_rn(G, am1);
                    B = __fmul_rn(B, am1);
                } else {
                    /* branch-free so that the 8 pixels of the lane interleave freely: a == 0 -> 0, 0 < a < 1 -> min(c/a, 1),
                     * a >= 1 -> untouched (the channel may hold more than 1 there: codes above max are not clamped).
                     * One quotient path for all three cases: the divisor is a when 0 < a < 1, else 1 (c/1 == c exactly, and
                     * rcp_refined(1) == 1); a == 0 uses a zero "reciprocal", for which div_by_rcp returns +0 (q0 = c*0 = 0,
                     * remainder c, 0*c + 0 = 0); the clamp limit is 1 below a == 1 and +inf from there on. */
                    const bool zero = (a == 0.0f), below = (a < 1.0f);
                    const float d = (below && !zero) ? a : 1.0f;
                    const float r1 = zero ? 0.0f : rcp_refined(d);
#if AB_R2Y_LIM
                    const float lim = below ? 1.0f : __uint_as_float(0x7F800000u);
                    R = fminf(div_by_rcp(R, d, r1), lim);
                    G = fminf(div_by_rcp(G, d, r1), lim);
                    B = fminf(div_by_rcp(B, d, r1), lim);
#else
                    const float qR = min1(div_by_rcp(R, d, r1)), qG = min1(div_by_rcp(G, d, r1)), qB = min1(div_by_rcp(B, d, r1));
                    R = below ? qR : R;
                    G = below ? qG : G;
                    B = below ? qB : B;
#endif
                }
            }
            const float Y = C::IDENT ? G : __fadd_rn(__fadd_rn(__fmul_rn(J.kr, R), __fmul_rn(J.kg, G)), __fmul_rn(J.kb, B)); /* ref: :383-386, :361-365 */
            yq[t] = (YFULL && AB_R2Y_QMERGE && AB_R2Y_SATCVT) ? quant_luma_fullrange<YT>(Y, J.rangeY) : quant_luma<YT>(Y, J.rangeY, qbY, hiLimit);
            if (C::SUB != SUB_400) {
                const float U = C::IDENT ? B : rdivx(__fsub_rn(B, Y), J.du);
                const float V = C::IDENT ? R : rdivx(__fsub_rn(R, Y), J.dv);
                if (C::SUB == SUB_444) {
                    uq[t] = quant_yuv<YT>(U, J.rangeUV, qbUV, hiLimit);
                    vq[t] = quant_yuv<YT>(V, J.rangeUV, qbUV, hiLimit);
                } else {
                    /* ref: :413-467 sums the samples that exist row by row, left to right, starting from 0 + first */
                    const int k = t >> 1;
                    if ((t & 1) == 0) {
                        su[k] = (r == 0) ? U : __fadd_rn(su[k], U);
                        sv[k] = (r == 0) ? V : __fadd_rn(sv[k], V);
                    } else if (fullTile || (uint32_t)t < nValid) {
                        su[k] = __fadd_rn(su[k], U);
                        sv[k] = __fadd_rn(sv[k], V);
                    }
                }
            }
            if (COPYA) aq[t] = am; /* ref: :545-569 */
        }
        const uint32_t j = j0 + r;
        const uint32_t ns = fullTile ? 8u : nValid;
        const bool vec = fullTile || (F.vecWidth != 0);
        /* addresses are formed where they are used: carrying five row pointers would cost ten registers */
        store_samples8<YT, QC>(F.y + (int64_t)j * F.yRB, px0, ns, yq, mm, vec);
        if (C::SUB == SUB_444) {
            store_samples8<YT, QC>(F.u + (int64_t)j * F.uRB, px0, ns, uq, mm, vec);
            store_samples8<YT, QC>(F.v + (int64_t)j * F.vRB, px0, ns, vq, mm, vec);
        }
        if (COPYA) {
            store_samples8<YT>(F.a + (int64_t)j * F.aRB, px0, ns, aq, 0, vec);
        } else if (writeAlpha) { /* warp-uniform */
            const float fillAlpha = __uint_as_float(J.yuvMax);
#pragma unroll
            for (int t = 0; t < 8; ++t) aq[t] = fillAlpha;
            store_samples8<YT>(F.a + (int64_t)j * F.aRB, px0, ns, aq, 0, vec);
        }
    }

    /* ---- subsampled chroma: average of the box (a division by 1, 2 or 4: exact scaling), quantise, store ---- */
    if (C::SUB == SUB_420 || C::SUB == SUB_422) {
        const uint32_t nc = (nValid + 1) >> 1;
        float uq[4], vq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool two = fullTile || ((uint32_t)(2 * k + 1) < nValid); /* block width 2 */
            const float inv = (two ? 0.5f : 1.0f) * ((C::SUB == SUB_420 && nRows == 2) ? 0.5f : 1.0f);
            uq[k] = quant_yuv<YT>(__fmul_rn(su[k], inv), J.rangeUV, qbUV, hiLimit);
            vq[k] = quant_yuv<YT>(__fmul_rn(sv[k], inv), J.rangeUV, qbUV, hiLimit);
        }
        const uint32_t uj = (C::SUB == SUB_420) ? (j0 >> 1) : j0;
        store_samples4<YT, QC>(F.u + (int64_t)uj * F.uRB, px0 >> 1, nc, uq, mm, fullTile || (F.vecWidth != 0));
        store_samples4<YT, QC>(F.v + (int64_t)uj * F.vRB, px0 >> 1, nc, vq, mm, fullTile || (F.vecWidth != 0));
    }
}

template <class C, bool FULL>
__device__ __forceinline__ void convert_store_rtile(const AbR2YJob & J, const AbR2YFrame & F, uint32_t x0, uint32_t j0, uint32_t lane, const RTileIn<C> & in)
{
    /* warp-uniform (block-uniform, launch-uniform) dispatch on three call constants; ragged edge tiles (FULL == false) are rare
     * and keep one body per channel order */
    constexpr int ROWS = RTileIn<C>::ROWS;
    const uint32_t px0 = x0 + 8 * lane;
    const uint32_t nValid = FULL ? 8u : ((px0 < F.width) ? min(8u, F.width - px0) : 0u);
    const uint32_t nRows = min((uint32_t)ROWS, F.height - j0);
    const bool copyA = (C::NCH == 4) && (J.alphaWrite == AB_ALPHA_COPY);
    const bool yFull = FULL && (J.biasY == 0.0f);
#define AB_R2Y_BODY(SW, CA, YF) convert_store_rtile_p<C, FULL, SW, CA, YF>(J, F, px0, j0, nValid, nRows, in)
    if (J.swapRB) {
        if (copyA) AB_R2Y_BODY(true, (C::NCH == 4), false);
        else AB_R2Y_BODY(true, false, false);
    } else if (copyA) {
        if (yFull) AB_R2Y_BODY(false, (C::NCH == 4), FULL);
        else AB_R2Y_BODY(false, (C::NCH == 4), false);
    } else {
        if (yFull) AB_R2Y_BODY(false, false, FULL);
        else AB_R2Y_BODY(false, false, false);
    }
#undef AB_R2Y_BODY
}

/* ---- strip kernel ---------------------------------------------------------------------------------------------------
 * Round-2 loop shape.  The group kernel below pays ~14 instructions per pixel for bookkeeping (ncu r02c, per source line: tile ->
 * (row, column) by division, 64-bit address arithmetic for five streams at every tile, the register moves of "next tile becomes
 * current"): a sixth of everything it executes.  Here a warp owns a CHUNK = one 256-pixel column strip x kRChunk consecutive
 * row units: positions are derived once per chunk, then every stream advances by its row stride (one 64-bit add each), loads
 * and stores use constant offsets from those pointers, and the prefetching loop is written out twice (two register sets that
 * swap roles) so that nothing is moved.  Chunks are numbered row-major (adjacent strips of one row band are adjacent ids) and
 * dealt to the warps exactly like the groups of tiles before: group g = chunks 8g..8g+7, one per warp, blocks stride over the
 * groups of their frame. */
#ifndef AB_R2Y_STRIP /* one-row tiles (4:4:4, 4:2:2, 4:0:0) run the strip kernel; two-row tiles keep the group kernel (their
                      * register sets leave no room: the strip form spills and measured 9 % slower, profiles/r02_variants.md) */
#define AB_R2Y_STRIP 1
#endif
#ifndef AB_R2Y_STRIP_MINB /* 80 registers: at 4 blocks/SM (64) the strip loop spills ~28 registers and loses (26.7 vs 22.5 us on C2 back) */
#define AB_R2Y_STRIP_MINB 3
#endif
template <class C>
struct RStrip {
    static constexpr bool ON = (AB_R2Y_STRIP != 0) && (C::SUB != SUB_420);
};
#ifndef AB_R2Y_CHUNK
#define AB_R2Y_CHUNK 8
#endif
#ifndef AB_R2Y_PINGPONG /* 1: the prefetching loop written out twice (register sets swap roles); 0: "next becomes current" by register moves */
#define AB_R2Y_PINGPONG 0
#endif
constexpr uint32_t kRChunk = AB_R2Y_CHUNK;

template <class C, bool SWAP, bool COPYA, bool YFULL>
__device__ __forceinline__ void r2y_strip_run(const AbR2YJob & J, const AbR2YFrame & F, uint32_t g0, uint32_t gStride, uint32_t warp, uint32_t lane)
{
    constexpr int ROWS = RTileIn<C>::ROWS;
    const uint32_t units = (F.rowEnd - F.rowBegin + ROWS - 1) / ROWS;
#pragma unroll 1
    for (uint32_t g = g0; g < F.groups; g += gStride) {
        const uint32_t c = g * kRWarps + warp;
        if (c >= F.tiles) continue; /* warp-uniform */
        const uint32_t cy = c / F.tilesX, sx = c - cy * F.tilesX;
        const uint32_t x0 = sx * kRTile;
        const uint32_t u0 = cy * kRChunk;
        const uint32_t jEnd = F.rowBegin + min(u0 + kRChunk, units) * ROWS; /* first luma row after the chunk (may exceed the picture by one) */
        uint32_t j = F.rowBegin + u0 * ROWS;
        /* the position is (x0, j): addresses are re-derived where they are used (two multiply-adds each) instead of being
         * carried in ten registers -- the two tile register sets already fill the budget of 4 blocks per SM */
        auto rowsAt = [&](uint32_t jj) { return (ROWS == 1) ? 1u : min((uint32_t)ROWS, F.height - jj); };
        auto rgbAt = [&](uint32_t jj) { return F.rgb + (int64_t)jj * F.rgbRB + (int64_t)(x0 + 8 * lane) * (C::NCH * (int)sizeof(typename C::RT)); };
        RTileIn<C> a, b;
        if (x0 + kRTile <= F.vecWidth && RLoop<C>::PLAIN) { /* tiles whose second register set does not fit: plain loop */
            for (; j < jEnd; j += ROWS) {
                load_rtile_p<C, true>(F, rgbAt(j), 8u, rowsAt(j), a);
                convert_store_rtile_p<C, true, SWAP, COPYA, YFULL>(J, F, x0 + 8 * lane, j, 8u, rowsAt(j), a);
            }
        } else if (x0 + kRTile <= F.vecWidth) {
            load_rtile_p<C, true>(F, rgbAt(j), 8u, rowsAt(j), a);
#if AB_R2Y_PINGPONG
            while (true) {
                bool more = (j + ROWS < jEnd);
                if (more) load_rtile_p<C, true>(F, rgbAt(j + ROWS), 8u, rowsAt(j + ROWS), b);
                convert_store_rtile_p<C, true, SWAP, COPYA, YFULL>(J, F, x0 + 8 * lane, j, 8u, rowsAt(j), a);
                if (!more) break;
                j += ROWS;
                more = (j + ROWS < jEnd);
                if (more) load_rtile_p<C, true>(F, rgbAt(j + ROWS), 8u, rowsAt(j + ROWS), a);
                convert_store_rtile_p<C, true, SWAP, COPYA, YFULL>(J, F, x0 + 8 * lane, j, 8u, rowsAt(j), b);
                if (!more) break;
                j += ROWS;
            }
#else
#pragma unroll 1
            for (;;) {
                const bool more = (j + ROWS < jEnd);
                if (more) load_rtile_p<C, true>(F, rgbAt(j + ROWS), 8u, rowsAt(j + ROWS), b);
                convert_store_rtile_p<C, true, SWAP, COPYA, YFULL>(J, F, x0 + 8 * lane, j, 8u, rowsAt(j), a);
                if (!more) break;
                j += ROWS;
                a = b;
            }
#endif
        } else { /* ragged last strip of the picture: no prefetch */
            const uint32_t px0 = x0 + 8 * lane;
            const uint32_t nValid = (px0 < F.width) ? min(8u, F.width - px0) : 0u;
            for (; j < jEnd; j += ROWS) {
                load_rtile_p<C, false>(F, rgbAt(j), nValid, rowsAt(j), a);
                convert_store_rtile_p<C, false, SWAP, COPYA, false>(J, F, px0, j, nValid, rowsAt(j), a);
            }
        }
    }
}

template <class C>
__global__ void __launch_bounds__(kRWarps * 32, AB_R2Y_STRIP_MINB) r2y_strip_kernel(const __grid_constant__ AbR2YBatch B)
{
    const AbR2YJob & J = B.common;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t f = 0;
    while (f + 1 < B.nFrames && blockIdx.x >= B.frame[f + 1].blockStart) ++f; /* block-uniform, once per block */
    const AbR2YFrame & F = B.frame[f];
    const uint32_t gStride = F.blocks;
    const uint32_t g0 = blockIdx.x - F.blockStart;
    if (g0 >= F.groups) return;
    /* launch-uniform dispatch on three call constants: each combination is its own straight-line loop */
    const bool copyA = (C::NCH == 4) && (J.alphaWrite == AB_ALPHA_COPY);
    const bool yFull = (J.biasY == 0.0f);
    if (J.swapRB) {
        if (copyA) r2y_strip_run<C, true, (C::NCH == 4), false>(J, F, g0, gStride, warp, lane);
        else r2y_strip_run<C, true, false, false>(J, F, g0, gStride, warp, lane);
    } else if (copyA) {
        if (yFull) r2y_strip_run<C, false, (C::NCH == 4), true>(J, F, g0, gStride, warp, lane);
        else r2y_strip_run<C, false, (C::NCH == 4), false>(J, F, g0, gStride, warp, lane);
    } else {
        if (yFull) r2y_strip_run<C, false, false, true>(J, F, g0, gStride, warp, lane);
        else r2y_strip_run<C, false, false, false>(J, F, g0, gStride, warp, lane);
    }
}

/* Persistent, software-pipelined and batched like y2r_fast_kernel: a block strides over groups of 8 consecutive tiles of one
 * frame (one tile per warp); each warp loads its tile of the next group into a second register set before converting the
 * current one. */
template <class C>
__global__ void __launch_bounds__(kRWarps * 32, RLoop<C>::MIN_BLOCKS) r2y_fast_kernel(const __grid_constant__ AbR2YBatch B)
{
    constexpr int ROWS = RTileIn<C>::ROWS;
    const AbR2YJob & J = B.common;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t f = 0;
    while (f + 1 < B.nFrames && blockIdx.x >= B.frame[f + 1].blockStart) ++f; /* block-uniform, once per block */
    const AbR2YFrame & F = B.frame[f];                                         /* invariant for the life of the block */
    const uint32_t gStride = F.blocks;
    uint32_t g = blockIdx.x - F.blockStart;
    if (g >= F.groups) return;
    if (RLoop<C>::PLAIN) { /* two-row tiles: a second register set costs a resident block per SM */
        for (; g < F.groups; g += gStride) {
            const uint32_t t = g * kRWarps + warp;
            if (t >= F.tiles) continue;
            const uint32_t unit = t / F.tilesX, tx0 = (t - unit * F.tilesX) * kRTile, tj0 = F.rowBegin + unit * ROWS;
            RTileIn<C> in;
            if (tx0 + kRTile <= F.vecWidth) {
                load_rtile<C, true>(J, F, tx0, tj0, lane, in);
                convert_store_rtile<C, true>(J, F, tx0, tj0, lane, in);
            } else {
                load_rtile<C, false>(J, F, tx0, tj0, lane, in);
                convert_store_rtile<C, false>(J, F, tx0, tj0, lane, in);
            }
        }
        return;
    }
#if AB_R2Y_UNROLL2
    /* ping-pong register sets: see y2r_fast_kernel */
    struct Slot {
        RTileIn<C> in;
        uint32_t x0, j0;
        bool live, full;
    };
    auto fetch = [&](Slot & s, uint32_t grp) {
        const uint32_t t = grp * kRWarps + warp;
        s.live = t < F.tiles;
        s.x0 = s.j0 = 0;
        s.full = false;
        if (s.live) {
            const uint32_t unit = t / F.tilesX;
            s.x0 = (t - unit * F.tilesX) * kRTile;
            s.j0 = F.rowBegin + unit * ROWS;
            s.full = (s.x0 + kRTile <= F.vecWidth);
            if (s.full) load_rtile<C, true>(J, F, s.x0, s.j0, lane, s.in);
            else load_rtile<C, false>(J, F, s.x0, s.j0, lane, s.in);
        }
    };
    auto run = [&](const Slot & s) {
        if (s.live) {
            if (s.full) convert_store_rtile<C, true>(J, F, s.x0, s.j0, lane, s.in);
            else convert_store_rtile<C, false>(J, F, s.x0, s.j0, lane, s.in);
        }
    };
    Slot a, b;
    fetch(a, g);
    while (true) {
        uint32_t ng = g + gStride;
        bool more = ng < F.groups; /* block-uniform */
        if (more) fetch(b, ng);
        run(a);
        if (!more) break;
        g = ng;
        ng = g + gStride;
        more = ng < F.groups;
        if (more) fetch(a, ng);
        run(b);
        if (!more) break;
        g = ng;
    }
#else
    RTileIn<C> cur, nxt;
    uint32_t tile = g * kRWarps + warp;
    bool live = tile < F.tiles;
    uint32_t x0 = 0, j0 = 0;
    bool full = false;
    if (live) {
        const uint32_t unit = tile / F.tilesX;
        x0 = (tile - unit * F.tilesX) * kRTile;
        j0 = F.rowBegin + unit * ROWS;
        full = (x0 + kRTile <= F.vecWidth);
        if (full) load_rtile<C, true>(J, F, x0, j0, lane, cur);
        else load_rtile<C, false>(J, F, x0, j0, lane, cur);
    }
    while (true) {
        const uint32_t ng = g + gStride;
        const bool more = ng < F.groups; /* block-uniform */
        uint32_t nx0 = 0, nj0 = 0;
        bool nlive = false, nfull = false;
        if (more) {
            const uint32_t ntile = ng * kRWarps + warp;
            nlive = ntile < F.tiles;
            if (nlive) {
                const uint32_t nunit = ntile / F.tilesX;
                nx0 = (ntile - nunit * F.tilesX) * kRTile;
                nj0 = F.rowBegin + nunit * ROWS;
                nfull = (nx0 + kRTile <= F.vecWidth);
                if (nfull) load_rtile<C, true>(J, F, nx0, nj0, lane, nxt);
                else load_rtile<C, false>(J, F, nx0, nj0, lane, nxt);
            }
        }
        if (live) {
            if (full) convert_store_rtile<C, true>(J, F, x0, j0, lane, cur);
            else convert_store_rtile<C, false>(J, F, x0, j0, lane, cur);
        }
        if (!more) break;
        cur = nxt;
        g = ng;
        live = nlive;
        x0 = nx0;
        j0 = nj0;
        full = nfull;
    }
#endif
}

template <class FrameT>
static uint32_t rassign_blocks(FrameT * frames, uint32_t n, uint32_t totalGroups, uint32_t blocksWanted)
{
    /* floor shares (>= 1 block per frame, <= its group count), then hand the remaining blocks to the frames with the most
     * groups per block; the total never exceeds blocksWanted (as long as n <= blocksWanted), so every block of a persistent
     * launch is resident from the start */
    if (totalGroups == 0) totalGroups = 1;
    uint32_t used = 0;
    for (uint32_t f = 0; f < n; ++f) {
        uint64_t share = (uint64_t)frames[f].groups * blocksWanted / totalGroups;
        if (share < 1) share = 1;
        if (share > frames[f].groups && frames[f].groups) share = frames[f].groups;
        frames[f].blocks = (uint32_t)share;
        used += (uint32_t)share;
    }
    while (used < blocksWanted) {
        uint32_t best = n;
        uint64_t bestLoad = 0;
        for (uint32_t f = 0; f < n; ++f) {
            if (frames[f].blocks >= frames[f].groups) continue;
            const uint64_t load = ((uint64_t)frames[f].groups << 16) / frames[f].blocks;
            if (best == n || load > bestLoad) {
                best = f;
                bestLoad = load;
            }
        }
        if (best == n) break;
        ++frames[best].blocks;
        ++used;
    }
    uint32_t start = 0;
    for (uint32_t f = 0; f < n; ++f) {
        frames[f].blockStart = start;
        start += frames[f].blocks;
    }
    return start;
}

template <typename RT, typename YT, int SUB, int NCH, bool AFIRST, int AM, bool IDENT = false>
static int rlaunch_cfg(AbR2YBatch & B, cudaStream_t st)
{
    using C = RCfg<RT, YT, SUB, NCH, AFIRST, AM, IDENT>;
    constexpr int ROWS = (SUB == SUB_420) ? 2 : 1;
    int dev = 0;
    cudaGetDevice(&dev);
    /* form 1 = strip kernel (chunks of kRChunk row units walked by one warp), form 0 = group kernel (independent units).  A lone small
     * picture does not have enough chunks to fill the device and would wait behind those serial chains: the strip form needs at
     * least two rounds of resident blocks, else the launch falls back to the group form. */
    for (int form = RStrip<C>::ON ? 1 : 0; form >= 0; --form) {
        uint32_t groups = 0;
        for (uint32_t f = 0; f < B.nFrames; ++f) {
            AbR2YFrame & F = B.frame[f];
            uint32_t units = (F.rowEnd - F.rowBegin + ROWS - 1) / ROWS;
            if (form) units = (units + kRChunk - 1) / kRChunk; /* work items are chunks of kRChunk row units */
            F.tilesX = (F.width + kRTile - 1) / kRTile;
            F.tiles = F.tilesX * units;
            F.groups = (F.tiles + kRWarps - 1) / kRWarps;
            groups += F.groups;
        }
        /* blocks that fit on the whole device at once: per instantiation, form and device (a process may drive several GPUs) */
        static int residentOf[2][16] = { { 0 }, { 0 } };
        int resident = (dev >= 0 && dev < 16) ? residentOf[form][dev] : 0;
        if (!resident) {
            int sms = 0, perSm = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (form) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, r2y_strip_kernel<C>, kRWarps * 32, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, r2y_fast_kernel<C>, kRWarps * 32, 0);
            { /* never more blocks per SM than the loop shape was tuned for (see y2r_fast.cuh) */
                const int tuned = form ? AB_R2Y_STRIP_MINB : RLoop<C>::MIN_BLOCKS;
                if (perSm > tuned) perSm = tuned;
            }
            resident = (sms > 0 && perSm > 0) ? sms * perSm : 148 * 2;
            if (dev >= 0 && dev < 16) residentOf[form][dev] = resident;
        }
        if (form && groups < 2u * (uint32_t)resident) continue;
        const uint32_t wanted = (groups >= 4u * (uint32_t)resident) ? (uint32_t)resident : groups;
        B.totalBlocks = rassign_blocks(B.frame, B.nFrames, groups, wanted);
        if (form) r2y_strip_kernel<C><<<B.totalBlocks, kRWarps * 32, 0, st>>>(B);
        else r2y_fast_kernel<C><<<B.totalBlocks, kRWarps * 32, 0, st>>>(B);
        break;
    }
    return 1;
}

template <typename RT, typename YT, int SUB>
static int rlaunch_sub(AbR2YBatch & B, cudaStream_t st)
{
    const AbR2YJob & J = B.common;
    if (J.nChan == 3) return (J.alphaMul == AB_MUL_NONE) ? rlaunch_cfg<RT, YT, SUB, 3, false, AB_MUL_NONE>(B, st) : 0;
    if (J.alphaFirst) {
        switch (J.alphaMul) {
            case AB_MUL_NONE: return rlaunch_cfg<RT, YT, SUB, 4, true, AB_MUL_NONE>(B, st);
            case AB_MUL_FLOAT_MUL: return rlaunch_cfg<RT, YT, SUB, 4, true, AB_MUL_FLOAT_MUL>(B, st);
            default: return rlaunch_cfg<RT, YT, SUB, 4, true, AB_MUL_FLOAT_UNMUL>(B, st);
        }
    }
    switch (J.alphaMul) {
        case AB_MUL_NONE: return rlaunch_cfg<RT, YT, SUB, 4, false, AB_MUL_NONE>(B, st);
        case AB_MUL_FLOAT_MUL: return rlaunch_cfg<RT, YT, SUB, 4, false, AB_MUL_FLOAT_MUL>(B, st);
        default: return rlaunch_cfg<RT, YT, SUB, 4, false, AB_MUL_FLOAT_UNMUL>(B, st);
    }
}

/* 2: vector-aligned rows everywhere; 1: only sample-aligned (the specialised kernels run their ragged, element-wise code on every
 * tile: AbR2YFrame::vecWidth == 0); 0: not even that.  See y2r_alignment. */
template <typename RT, typename YT>
static int r2y_alignment(const AbR2YJob & J)
{
    const uintptr_t rbits = (uintptr_t)J.rgb | (uintptr_t)J.rgbRB;
    uintptr_t ybits = (uintptr_t)J.y | (uintptr_t)J.yRB;
    if (J.alphaWrite != AB_ALPHA_KEEP) ybits |= (uintptr_t)J.a | (uintptr_t)J.aRB;
    const uintptr_t cbits = J.hasChroma ? ((uintptr_t)J.u | (uintptr_t)J.uRB | (uintptr_t)J.v | (uintptr_t)J.vRB) : 0;
    if ((rbits & (sizeof(RT) - 1)) || ((ybits | cbits) & (sizeof(YT) - 1))) return 0;
    const uint32_t pixb = J.nChan * (uint32_t)sizeof(RT);
    const uintptr_t srcAlign = ((8 * pixb) % 16 == 0) ? 15 : 7;
    const uintptr_t lumaAlign = 8 * sizeof(YT) - 1;
    const bool vec = !(rbits & srcAlign) && !(ybits & lumaAlign) && !(cbits & (J.shiftX ? (4 * sizeof(YT) - 1) : lumaAlign));
    return vec ? 2 : 1;
}

template <typename RT, typename YT>
static bool rfast_eligible(const AbR2YJob & J)
{
    const bool identity = (J.mode == AB_MODE_IDENTITY);
    if ((J.mode != AB_MODE_COEFF && !identity) || J.isGray) return false;
    if (identity && (!J.hasChroma || J.shiftX || J.alphaMul != AB_MUL_NONE)) return false; /* 4:4:4, straight alpha */
    if (J.nChan != 3 && J.nChan != 4) return false;
    if (!J.rgbMaxDiv.intExact) return false;
    if (!identity && J.hasChroma && (!J.du.exact || !J.dv.exact)) return false;
    if (!identity && AB_R2Y_DIV2 && J.hasChroma && (!J.du.exact2 || !J.dv.exact2)) return false;
    if (J.alphaWrite == AB_ALPHA_COPY && (J.alphaRescale || J.nChan != 4)) return false;
    if (J.alphaMul != AB_MUL_NONE && J.nChan != 4) return false;
    if (J.width == 0 || J.rowEnd <= J.rowBegin) return false;
    if (r2y_alignment<RT, YT>(J) == 0) return false; /* not even sample-aligned: the byte-addressed kernel */
    return true;
}

static inline bool rsame_constants(const AbR2YJob & a, const AbR2YJob & b)
{
    const size_t from = offsetof(AbR2YJob, yuvMax), to = offsetof(AbR2YJob, srcPairAligned);
    return memcmp(reinterpret_cast<const char *>(&a) + from, reinterpret_cast<const char *>(&b) + from, to - from) == 0;
}

/* jobs[0..n) (n <= AB_MAX_BATCH) in ONE specialised launch if all are eligible and share their constants; else 0 */
template <typename RT, typename YT>
static int launch_r2y_fast_typed(const AbR2YJob * jobs, uint32_t n, cudaStream_t st)
{
    if (n == 0 || n > AB_MAX_BATCH) return 0;
    for (uint32_t k = 0; k < n; ++k) {
        if (!rfast_eligible<RT, YT>(jobs[k])) return 0;
        if (k && !rsame_constants(jobs[0], jobs[k])) return 0;
    }
    AbR2YBatch B;
    B.common = jobs[0];
    B.nFrames = n;
    for (uint32_t k = 0; k < n; ++k) {
        AbR2YFrame & F = B.frame[k];
        const AbR2YJob & S = jobs[k];
        F.rgb = S.rgb; F.y = S.y; F.u = S.u; F.v = S.v; F.a = S.a;
        F.rgbRB = S.rgbRB; F.yRB = S.yRB; F.uRB = S.uRB; F.vRB = S.vRB; F.aRB = S.aRB;
        F.width = S.width; F.height = S.height; F.rowBegin = S.rowBegin; F.rowEnd = S.rowEnd;
        F.tilesX = F.tiles = F.groups = F.blockStart = F.blocks = 0;
        F.vecWidth = (r2y_alignment<RT, YT>(S) == 2) ? S.width : 0;
    }
    const AbR2YJob & J = B.common;
    if (J.mode == AB_MODE_IDENTITY) { /* lossless GBR: 4:4:4, no alpha arithmetic */
        if (J.nChan == 3) return rlaunch_cfg<RT, YT, SUB_444, 3, false, AB_MUL_NONE, true>(B, st);
        return J.alphaFirst ? rlaunch_cfg<RT, YT, SUB_444, 4, true, AB_MUL_NONE, true>(B, st) : rlaunch_cfg<RT, YT, SUB_444, 4, false, AB_MUL_NONE, true>(B, st);
    }
    if (!J.hasChroma) return rlaunch_sub<RT, YT, SUB_400>(B, st);
    if (!J.shiftX) return rlaunch_sub<RT, YT, SUB_444>(B, st);
    return J.shiftY ? rlaunch_sub<RT, YT, SUB_420>(B, st) : rlaunch_sub<RT, YT, SUB_422>(B, st);
}

} // namespace abfast
