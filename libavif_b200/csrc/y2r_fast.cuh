/*
 * y2r_fast.cuh -- compile-time specialised YUV->RGB kernels for the hot configurations.
 *
 * The general kernel (kernels.cu) decides everything from runtime flags and spends ~3/4 of its issue slots on
 * uniform branches and constant loads (ncu, profiles/r01a).  Here every policy decision is a template parameter,
 * so the inner loop is straight-line code:
 *   - a lane owns 8 consecutive pixels of a 256-pixel warp tile: one 128-bit (16-bit samples) or 64-bit (8-bit
 *     samples) load per plane row, 64/32-bit for horizontally subsampled chroma; every load of the tile is issued
 *     before any arithmetic (64..150 B in flight per thread);
 *   - sample -> float is one PRMT (as_float(code | 0x4B000000) == 2^23 + code), the high-bit clamp one FMNMX,
 *     float -> code one FADD.RZ; division by call constants is the verified 3-op sequence (pixel_math.cuh);
 *   - 4:2:0 bilinear: a warp converts TWO luma rows (2r-1, 2r) that share chroma rows r-1 and r; the 9/3/3/1
 *     products of every chroma sample are formed once and reused by both rows; the only cross-lane traffic is one
 *     warp shuffle per side per (chroma row, plane) -- the neighbour column of a lane's first / last sample;
 *   - the lane's 8 packed pixels (24..64 B) leave as 16-byte (8-byte for 3-byte pixels) vector stores straight from
 *     registers; lanes are contiguous, so a warp instruction covers one 0.75..2 KB span.  (A shared-memory staged
 *     variant with XOR-swizzled, fully linear write-back measured 6 % slower on NS -- profiles/r01_variants.md --
 *     the __syncwarp / LDS / STG chain at the end of each tile costs more issue slots than the partially filled
 *     sectors cost L2 bandwidth; VARIANT 0 keeps it for reference.)
 *   - full 256-pixel tiles take a warp-uniform path with no per-lane range checks;
 *   - warps are persistent and software-pipelined: the loads of the next tile are issued before the current tile is
 *     converted (see y2r_fast_kernel).
 * Arithmetic is the reference's, operator for operator (pixel_math.cuh); eligibility (coefficients mode, aligned
 * planes, verified divisors, ...) is checked by the launcher, everything else runs on the general kernel.
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

enum { CH_FULL = 1, CH_NEAREST = 2, CH_BIL420 = 3, CH_BIL422 = 4 };

enum { OUT_INT = 0, OUT_F16 = 1, OUT_565 = 2 };
/* AW value of the kernel templates only: alpha copied from a plane of another DEPTH in the same container size (10 / 12-bit planes
 * into 16-bit channels, e.g. HDR images into half floats): src/alpha.c:93-97 rescale.  Containers of different size always rescale. */
constexpr int AW_COPY_RESC = 3;

enum { MX_COEFF = 0, MX_IDENTITY = 1, MX_IDENTITY8 = 2 };

template <typename YT_, typename RT_, int CHROMA_, int NCH_, bool AFIRST_, int AW_, int AM_, int OUT_ = OUT_INT, int MX_ = MX_COEFF>
struct Cfg {
    using YT = YT_;
    using RT = RT_;
    static constexpr int CHROMA = CHROMA_; /* CH_* */
    static constexpr int NCH = NCH_;       /* 3 or 4 channels (RGB565: 3 computed, one 16-bit word stored) */
    static constexpr bool AFIRST = AFIRST_; /* alpha stored before colour */
    static constexpr int AW = AW_;         /* AB_ALPHA_FILL / AB_ALPHA_COPY (NCH == 4) */
    static constexpr int AM = AM_;         /* AB_MUL_* */
    static constexpr int MX = MX_;         /* MX_COEFF: the kr/kg/kb matrix; MX_IDENTITY: G=Y, B=U, R=V through the luma table (src/reformat.c:846-851,
                                              :587-589: lossless GBR); MX_IDENTITY8: the 8-bit full-range byte shuffle (:1278-1309) */
    static constexpr int OUT = OUT_;       /* OUT_INT: integer codes; OUT_F16: half floats (src/reformat.c:1419-1443, RT = u16);
                                              OUT_565: RGB565 words (src/reformat.c:619, RT = u8, NCH = 3) */
};

constexpr int kTile = 256;
constexpr int kWarps = 8;

/* ---- sample vectors -------------------------------------------------------------------------- */

/* magic float (2^23 + code) of sample k of a packed vector; k is a compile-time constant after unrolling */
template <typename T>
__device__ __forceinline__ float magic_sample(const uint32_t * w, int k, uint32_t K)
{
    if (sizeof(T) == 2) return __uint_as_float(__byte_perm(w[k >> 1], K, (k & 1) ? 0x7432u : 0x7410u));
    return __uint_as_float(__byte_perm(w[k >> 2], K, 0x7440u | (uint32_t)(k & 3)));
}

#ifndef AB_Y2R_PACKED_CLAMP /* tuning switch: 0 = one FMNMX per sample after the conversion to float */
#define AB_Y2R_PACKED_CLAMP 1
#endif
/* the same with the code clamped to yuvMax first (ref: AVIF_MIN(ptr[i], yuvMaxChannel), src/reformat.c:583-598): 16-bit
 * containers may hold garbage above the depth.  mm = yuvMax in both halves: one VIMNMX.U16x2 per pair of samples (the two
 * samples of a word share it), instead of one FMNMX per sample. */
template <typename T>
__device__ __forceinline__ float magic_sample_c(const uint32_t * w, int k, uint32_t K, uint32_t mm)
{
    if (sizeof(T) == 2) return __uint_as_float(__byte_perm(AB_Y2R_PACKED_CLAMP ? __vminu2(w[k >> 1], mm) : w[k >> 1], K, (k & 1) ? 0x7432u : 0x7410u));
    return __uint_as_float(__byte_perm(w[k >> 2], K, 0x7440u | (uint32_t)(k & 3)));
}

/* 8 consecutive samples starting at x (x % 8 == 0, row 16/8-byte aligned); n = how many of them exist */
template <typename T>
__device__ __forceinline__ void load8(const uint8_t * row, uint32_t x, uint32_t n, uint32_t w[4], bool vec = true)
{
    w[0] = w[1] = w[2] = w[3] = 0;
    if (n >= 8 && vec) { /* vec == false: rows that are only sample-aligned are read element by element */
        if (sizeof(T) == 2) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row + (size_t)x * 2));
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        } else {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row + x));
            w[0] = v.x; w[1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((uint32_t)k < n) {
                const uint32_t s = (uint32_t)__ldg(reinterpret_cast<const T *>(row) + x + k);
                if (sizeof(T) == 2) w[k >> 1] |= s << ((k & 1) * 16);
                else w[k >> 2] |= s << ((k & 3) * 8);
            }
        }
    }
}

/* 4 consecutive chroma samples starting at column m (m % 4 == 0), columns >= cw replicate column cw-1 */
template <typename T>
__device__ __forceinline__ void load4_clamped(const uint8_t * row, uint32_t m, uint32_t cw, uint32_t w[2], bool vec = true)
{
    w[0] = w[1] = 0;
    if (m + 3 < cw && vec) {
        if (sizeof(T) == 2) {
            const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row + (size_t)m * 2));
            w[0] = v.x; w[1] = v.y;
        } else {
            w[0] = __ldg(reinterpret_cast<const uint32_t *>(row + m));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t s = (uint32_t)__ldg(reinterpret_cast<const T *>(row) + min(m + k, cw - 1));
            if (sizeof(T) == 2) w[k >> 1] |= s << ((k & 1) * 16);
            else w[0] |= s << (k * 8);
        }
    }
}

template <typename T>
__device__ __forceinline__ float magic_scalar(const uint8_t * row, uint32_t col)
{
    return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const T *>(row) + col) | 0x4B000000u);
}

/* ---- arithmetic on magic floats ------------------------------------------------------------------ */

/* Instruction diet of round 2 (0 restores the round-1 form for A/B timing):
 *   AB_Y2R_DIV2  x / d and 2x / d for verified divisors as fma(x, rh, RN(x*rl)) (AbDiv.exact2; proof: tools/verify_quant.c `div`, all
 *                2^32 numerators per divisor, re-run on the device by tests/test_gpu_division.py) -- one operation less than Markstein's;
 *   AB_Y2R_QFMA  (T)(0.5f + c*rgbMaxF) with the product left unrounded: floor(RN(fma(c, max, 0.5))) == floor(RN(0.5 + RN(c*max))) for
 *                EVERY float c in [0,1] and max = 2^8-1, 2^10-1, 2^12-1, 2^16-1 (tools/verify_quant.c `rgbquant`, 1.07e9 values each). */
#ifndef AB_Y2R_DIV2
#define AB_Y2R_DIV2 1
#endif
#ifndef AB_Y2R_QFMA
#define AB_Y2R_QFMA 1
#endif
#ifndef AB_Y2R_P1FMA /* bilinear blend: the diagonal tap's (exact) 1/16 product folded into an FFMA */
#define AB_Y2R_P1FMA 1
#endif
#ifndef AB_Y2R_TMA /* bilinear 4:2:0: input rows staged in shared memory by the TMA engine (cp.async.bulk + mbarrier), see y2r_fast_staged_kernel.
                    * Off by default: measured 25-30 % SLOWER than the plain loop (profiles/r02_variants.md); `make variant VFLAGS=-DAB_Y2R_TMA=1` builds it */
#define AB_Y2R_TMA 0
#endif
#ifndef AB_Y2R_TMA_STAGES
#define AB_Y2R_TMA_STAGES 3
#endif

/* verified constant division, unconditional fast form (the launcher guarantees dv.exact, and dv.exact2 for the 2-op form) */
__device__ __forceinline__ float divx(float x, const AbDiv & dv)
{
#if AB_Y2R_DIV2
    return __fmaf_rn(x, dv.rh, __fmul_rn(x, dv.rl));
#else
    const float q = __fmul_rn(x, dv.r);
    return __fmaf_rn(__fmaf_rn(-q, dv.d, x), dv.r, q);
#endif
}

/* RN(2x / d) for a verified divisor with the doubling folded into the constants (r2 = 2r, dh = d/2): every intermediate of
 * divx is scaled by exactly 2 or unchanged (q2 = 2q, the residual x - q2*dh = x - q*d, result 2q'), and RN(2x/d) = 2 RN(x/d)
 * because doubling commutes with rounding away from the subnormal range -- |x| is 0 or far above 2^-60 here. */
__device__ __forceinline__ float div2x(float x, float r2, float dh)
{
#if AB_Y2R_DIV2
    return __fmaf_rn(x, r2, __fmul_rn(x, dh)); /* r2 = 2 rh, dh = 2 rl of the divisor (convert_store_tile) */
#else
    const float q2 = __fmul_rn(x, r2);
    return __fmaf_rn(__fmaf_rn(-q2, dh, x), r2, q2);
#endif
}

/* normalised value of a magic-float code: ((float)min(code, yuvMax) - bias) / range  (ref: src/reformat.c:583,:598,:1005) */
template <typename T, bool CLAMP = true>
__device__ __forceinline__ float norm_magic(float mf, float magicMax, float magicBias, const AbDiv & range)
{
    if (CLAMP && sizeof(T) == 2) mf = fminf(mf, magicMax); /* clamp garbage high bits of 16-bit containers */
    return div_int(__fsub_rn(mf, magicBias), range); /* integer numerator: 2-op exact quotient */
}

/* (T)(0.5f + c*max) as the bit pattern of 2^23 + code (low 16 bits = code) */
__device__ __forceinline__ float quant_magic(float c, float maxF)
{
#if AB_Y2R_QFMA
    return __fadd_rz(__fmaf_rn(c, maxF, 0.5f), 8388608.0f); /* c in [0,1]: see AB_Y2R_QFMA */
#else
    return __fadd_rz(__fadd_rn(0.5f, __fmul_rn(c, maxF)), 8388608.0f);
#endif
}

__device__ __forceinline__ float clamp01f(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); } /* == AVIF_CLAMP for non-NaN */

/* ---- one warp tile: loads, then arithmetic + stores --------------------------------------------------- */

template <class C>
struct TileTraits {
    using YT = typename C::YT;
    using RT = typename C::RT;
    static constexpr bool SUB = (C::CHROMA != CH_FULL);
    static constexpr bool BIL = (C::CHROMA == CH_BIL420 || C::CHROMA == CH_BIL422);
    static constexpr int ROWS = (C::CHROMA == CH_BIL420) ? 2 : 1;
    static constexpr int CROWS = ROWS;
    static constexpr int PIXB = (C::OUT == OUT_565) ? 2 : C::NCH * (int)sizeof(RT);
    static constexpr bool HAS_A = (C::NCH == 4) && (C::AW == AB_ALPHA_COPY || C::AW == AW_COPY_RESC);
    /* alpha plane depth != channel depth (src/alpha.c:93-97): always when the containers differ (8-bit planes into 16-bit channels
     * or the reverse), by request (AW_COPY_RESC) for equal containers */
    static constexpr bool RESC = HAS_A && (sizeof(YT) != sizeof(RT) || C::AW == AW_COPY_RESC);
    static constexpr bool NEED_A = HAS_A || (C::AM != AB_MUL_NONE);
    static constexpr int CW = SUB ? 2 : 4; /* words of one chroma row of the lane */
};

/* everything a lane loads for one tile: packed sample words, plus the bilinear halo columns of the edge lanes */
template <class C>
struct TileIn {
    using T = TileTraits<C>;
    uint32_t yw[T::ROWS][4];
    uint32_t aw[T::NEED_A ? T::ROWS : 1][4];
    uint32_t uw[T::CROWS][T::CW], vw[T::CROWS][T::CW];
    float hU[T::BIL ? T::CROWS : 1], hV[T::BIL ? T::CROWS : 1]; /* bilinear: the column left of lane 0's / right of lane 31's (magic floats) */
};

/* rows of a tile: `unit` counts luma rows, or row pairs (2r-1, 2r) for 4:2:0 bilinear */
template <class C>
struct TileRows {
    int jrow[TileTraits<C>::ROWS];
    bool rowOn[TileTraits<C>::ROWS];
    int cA, cB; /* chroma rows: r-1 and r of a bilinear 4:2:0 pair (missing ones replicate the other), else the row's own */
    bool any;
};

template <class C>
__device__ __forceinline__ TileRows<C> tile_rows(const AbY2RJob & J, const AbY2RFrame & F, uint32_t unit)
{
    constexpr int ROWS = TileTraits<C>::ROWS;
    TileRows<C> R;
    if (ROWS == 2) {
        const int ru = (int)(F.rowBegin >> 1) + (int)unit;
        R.jrow[0] = 2 * ru - 1;
        R.jrow[ROWS - 1] = 2 * ru;
        const int chromaH = (int)((F.height + 1) >> 1);
        R.cA = (ru >= 1) ? ru - 1 : ru;
        R.cB = (ru < chromaH) ? ru : ru - 1;
    } else {
        R.jrow[0] = (int)(F.rowBegin + unit);
        R.cA = R.cB = R.jrow[0] >> J.shiftY;
    }
    R.any = false;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        R.rowOn[r] = (R.jrow[r] >= (int)F.rowBegin) && (R.jrow[r] < (int)F.rowEnd);
        R.any = R.any || R.rowOn[r];
    }
    return R;
}

template <class C, bool FULL>
__device__ __forceinline__ void load_tile(const AbY2RJob & J, const AbY2RFrame & F, const TileRows<C> & R, uint32_t x0, uint32_t lane, TileIn<C> & in)
{
    using T = TileTraits<C>;
    using YT = typename T::YT;
    const uint32_t px0 = x0 + 8 * lane;
    const uint32_t nValid = FULL ? 8u : ((px0 < F.width) ? min(8u, F.width - px0) : 0u);
    const bool vec = FULL || (F.vecWidth != 0);
#pragma unroll
    for (int r = 0; r < T::ROWS; ++r) {
        in.yw[r][0] = in.yw[r][1] = in.yw[r][2] = in.yw[r][3] = 0;
        if (T::NEED_A) in.aw[r][0] = in.aw[r][1] = in.aw[r][2] = in.aw[r][3] = 0;
        if (R.rowOn[r] && (FULL || nValid)) {
            load8<YT>(F.y + (int64_t)R.jrow[r] * F.yRB, px0, nValid, in.yw[r], vec);
            if (T::NEED_A) load8<YT>(F.a + (int64_t)R.jrow[r] * F.aRB, px0, nValid, in.aw[r], vec);
        }
    }
    const uint32_t cw = T::SUB ? ((F.width + 1) >> 1) : F.width;
    const uint32_t m0 = T::SUB ? (px0 >> 1) : px0;
#pragma unroll
    for (int c = 0; c < T::CROWS; ++c) {
        const int crow = (c == 0) ? R.cA : R.cB;
        const uint8_t * ur = F.u + (int64_t)crow * F.uRB;
        const uint8_t * vr = F.v + (int64_t)crow * F.vRB;
        if (T::SUB) {
            if (FULL) { /* the lane's 4 columns exist: m0 + 3 <= 127 + 128*tile < cw */
                load4_clamped<YT>(ur, m0, m0 + 4, in.uw[c]);
                load4_clamped<YT>(vr, m0, m0 + 4, in.vw[c]);
            } else {
                load4_clamped<YT>(ur, m0, cw, in.uw[c], vec);
                load4_clamped<YT>(vr, m0, cw, in.vw[c], vec);
            }
            if (T::BIL) {
                /* only the two edge lanes need a column outside the warp's span (clamped at the picture edge): one
                 * divergent region for both, one register per plane */
                in.hU[c] = in.hV[c] = 0.0f;
                if (lane == 0 || lane == 31) {
                    const uint32_t col = (lane == 0) ? ((m0 > 0) ? m0 - 1 : 0) : min(m0 + 4, cw - 1);
                    in.hU[c] = magic_scalar<YT>(ur, col);
                    in.hV[c] = magic_scalar<YT>(vr, col);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < T::CW; ++k) in.uw[c][k] = in.vw[c][k] = 0;
            if (FULL || nValid) {
                load8<YT>(ur, px0, nValid, in.uw[c], vec);
                load8<YT>(vr, px0, nValid, in.vw[c], vec);
            }
        }
    }
}

/* CARRY (bilinear 4:2:0 strips): chroma row A of this unit is row B of the unit above, whose normalised columns (own 4 + the two
 * neighbours) arrive in carry[plane][6] instead of being loaded and normalised again; either way carry leaves holding row B */
/* LATE (fed strips, AB_Y2R_LATE): the luma / alpha words of a row are fetched from the warp's ring stage (lateStage: the lane's
 * 16/8 bytes of row 0; rows are lateRow bytes apart, alpha lateAlpha bytes behind luma) only when that row is converted, and the
 * blended chroma of the second row waits in the lane's spill slots (lateSpill, shared memory) while the first row is converted:
 * ~24 registers less are live across the tile, which is what lets a fourth block onto the SM. */
__device__ __forceinline__ void spill_st(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float spill_ld(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
struct LateFeed {
    const uint8_t * stage; /* the lane's luma bytes of row 0 in the ring stage */
    uint32_t row, alpha;   /* byte distance between the two rows / from a luma row to its alpha row */
    uint32_t spill;        /* shared-space address of the lane's first spill slot; slot t is 128 bytes further */
};
template <class C, bool FULL, bool CARRY, bool LATE = false>
__device__ __forceinline__ void convert_store_tile_x(const AbY2RJob & J, const AbY2RFrame & F, const TileRows<C> & R, uint32_t x0, uint32_t lane, const TileIn<C> & in,
                                                     float (&carry)[2][6], const LateFeed & late = LateFeed{ nullptr, 0, 0, 0 })
{
    using T = TileTraits<C>;
    using YT = typename T::YT;
    using RT = typename T::RT;
    constexpr int ROWS = T::ROWS, CROWS = T::CROWS, PIXB = T::PIXB;
    constexpr bool SUB = T::SUB, BIL = T::BIL, HAS_A = T::HAS_A, NEED_A = T::NEED_A;
    const uint32_t px0 = x0 + 8 * lane;
    const uint32_t nValid = FULL ? 8u : ((px0 < F.width) ? min(8u, F.width - px0) : 0u);

    /* ------------------------------ chroma ------------------------------ */
    const uint32_t K = J.magicK;
    const float magicMax = __uint_as_float(0x4B000000u | J.yuvMax);
    const uint32_t mm = J.yuvMax * 0x10001u; /* yuvMax in both 16-bit halves (unused for 8-bit samples) */
    float cb[ROWS][8], cr[ROWS][8];
    if (!SUB) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            cb[0][t] = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(in.uw[0], t, K, mm), magicMax, J.magicBiasUV, J.rangeUV);
            cr[0][t] = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(in.vw[0], t, K, mm), magicMax, J.magicBiasUV, J.rangeUV);
        }
    } else if (!BIL) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float nu = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(in.uw[0], k, K, mm), magicMax, J.magicBiasUV, J.rangeUV);
            const float nv = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(in.vw[0], k, K, mm), magicMax, J.magicBiasUV, J.rangeUV);
            cb[0][2 * k] = cb[0][2 * k + 1] = nu;
            cr[0][2 * k] = cr[0][2 * k + 1] = nv;
        }
    } else {
        /* bilinear: per plane, normalise the lane's 4 columns of each chroma row, fetch the two neighbour columns
         * from the adjacent lanes (edge lanes: halo loads), form the 9/16, 3/16, 1/16 products once and blend.
         * ref: src/reformat.c:764-838; blend order ((near*9/16 + adjCol*3/16) + adjRow*3/16) + diag*1/16 */
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
            float p9[CROWS][4], p3[CROWS][6], p1[CROWS][6]; /* index 0 = left halo, 1..4 = own columns, 5 = right halo */
            float nn[CROWS][6];
#pragma unroll
            for (int c = 0; c < CROWS; ++c) {
                float n[6];
                if (CARRY && c == 0) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) n[k] = carry[plane][k];
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        n[k + 1] = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(plane ? in.vw[c] : in.uw[c], k, K, mm), magicMax, J.magicBiasUV, J.rangeUV);
                    }
                    float left = __shfl_up_sync(0xFFFFFFFFu, n[4], 1);
                    float right = __shfl_down_sync(0xFFFFFFFFu, n[1], 1);
                    const float halo = norm_magic<YT>(plane ? in.hV[c] : in.hU[c], magicMax, J.magicBiasUV, J.rangeUV);
                    left = (lane == 0) ? halo : left;
                    right = (lane == 31) ? halo : right;
                    n[0] = left;
                    n[5] = right;
                }
                if (c == CROWS - 1) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) carry[plane][k] = n[k];
                }
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    p3[c][k] = __fmul_rn(n[k], 0.1875f);
                    nn[c][k] = n[k];
                    p1[c][k] = AB_Y2R_P1FMA ? 0.0f : __fmul_rn(n[k], 0.0625f);
                    if (k >= 1 && k <= 4) p9[c][k - 1] = __fmul_rn(n[k], 0.5625f);
                }
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                /* luma row 2r-1 (r == 0 of a 4:2:0 unit): near = chroma row A, adjacent = B; row 2r: near = B, adjacent = A.
                 * 4:2:2 (one chroma row): near == adjacent. */
                const int cn = (ROWS == 2) ? r : 0, ca = (ROWS == 2) ? (1 - r) : 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int k = 1 + (t >> 1);                 /* own column index in the 6-wide arrays */
                    const int ka = (t & 1) ? k + 1 : k - 1;     /* adjacent column: right of odd pixels, left of even ones */
#if AB_Y2R_P1FMA
                    /* the 1/16 product is exact (a power of two times a normal float), so adding it is one FFMA: RN(s + n/16) */
                    const float val = __fmaf_rn(nn[ca][ka], 0.0625f, __fadd_rn(__fadd_rn(p9[cn][k - 1], p3[cn][ka]), p3[ca][k]));
#else
                    const float val = __fadd_rn(__fadd_rn(__fadd_rn(p9[cn][k - 1], p3[cn][ka]), p3[ca][k]), p1[ca][ka]);
#endif
                    if (LATE && r == 1) spill_st(late.spill + (uint32_t)(plane * 8 + t) * 128u, val);
                    else if (plane) cr[r][t] = val;
                    else cb[r][t] = val;
                }
            }
        }
    }

    /* ------------------------------ pixels ------------------------------ */
    const float magicOne = 8388608.0f;
#if AB_Y2R_DIV2
    /* RN(2x / kg) = 2 RN(x / kg) = fma(x, 2rh, RN(x * 2rl)): every intermediate of the 2-op quotient doubles exactly */
    const float kgR2 = __fmul_rn(J.kg.rh, 2.0f), kgDh = __fmul_rn(J.kg.rl, 2.0f);
#else
    const float kgR2 = __fmul_rn(J.kg.r, 2.0f), kgDh = __fmul_rn(J.kg.d, 0.5f); /* exact scalings: see div2x */
#endif
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (!R.rowOn[r]) continue; /* warp-uniform */
        uint32_t ywr[4], awr[4];
        if (LATE) {
            const uint8_t * ys = late.stage + r * late.row;
            if (sizeof(YT) == 2) {
                const uint4 v = *reinterpret_cast<const uint4 *>(ys);
                ywr[0] = v.x; ywr[1] = v.y; ywr[2] = v.z; ywr[3] = v.w;
                if (NEED_A) {
                    const uint4 a = *reinterpret_cast<const uint4 *>(ys + late.alpha);
                    awr[0] = a.x; awr[1] = a.y; awr[2] = a.z; awr[3] = a.w;
                }
            } else {
                const uint2 v = *reinterpret_cast<const uint2 *>(ys);
                ywr[0] = v.x; ywr[1] = v.y; ywr[2] = ywr[3] = 0;
                if (NEED_A) {
                    const uint2 a = *reinterpret_cast<const uint2 *>(ys + late.alpha);
                    awr[0] = a.x; awr[1] = a.y; awr[2] = awr[3] = 0;
                }
            }
            if (r == 1) {
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    cb[1][t] = spill_ld(late.spill + (uint32_t)t * 128u);
                    cr[1][t] = spill_ld(late.spill + (uint32_t)(8 + t) * 128u);
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ywr[k] = in.yw[r][k];
                awr[k] = NEED_A ? in.aw[r][k] : 0u;
            }
        }
        float ch[8][4];            /* per pixel, memory-order channel values as magic-float bit patterns (low 16 bits = code) */
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float Y = norm_magic<YT, !AB_Y2R_PACKED_CLAMP>(magic_sample_c<YT>(ywr, t, K, mm), magicMax, J.magicBiasY, J.rangeY);
            const float Cb = cb[SUB && !BIL ? 0 : r][t], Cr = cr[SUB && !BIL ? 0 : r][t];
            float Rv, Gv, Bv;
            if (C::MX == MX_COEFF) { /* ref: src/reformat.c:1014-1016 == :874-876 */
                Rv = __fadd_rn(Y, __fmul_rn(J.c1, Cr));
                Bv = __fadd_rn(Y, __fmul_rn(J.c2, Cb));
                Gv = __fsub_rn(Y, div2x(__fadd_rn(__fmul_rn(J.c3, Cr), __fmul_rn(J.c4, Cb)), kgR2, kgDh));
            } else { /* ref: :846-851 (chroma was normalised with the luma bias / range: the host passes them as the UV pair) */
                Gv = Y;
                Bv = Cb;
                Rv = Cr;
            }
            Rv = clamp01f(Rv);
            Gv = clamp01f(Gv);
            Bv = clamp01f(Bv);
            float am = 0.0f; /* alpha sample as magic float */
            if (NEED_A) am = magic_sample<YT>(awr, t, K);
            float qaCopy = am; /* what the alpha channel receives when it is copied */
            if (T::RESC) {
                /* ref: src/alpha.c:93-97 -- clamp((int)(0.5f + (a / srcMaxF) * dstMaxF), dstMax) on the raw sample.  8-bit samples
                 * give a / 255 in [0,1], where the fused form is the proven one (AB_Y2R_QFMA); 16-bit containers may hold more than
                 * srcMax, so the product keeps its own rounding there */
                const float an = div_int(__fsub_rn(am, magicOne), J.yuvMaxDiv);
                const float tt = (AB_Y2R_QFMA && sizeof(YT) == 1) ? __fmaf_rn(an, J.rgbMaxF, 0.5f) : __fadd_rn(0.5f, __fmul_rn(an, J.rgbMaxF));
                qaCopy = __fadd_rz(fminf(tt, __fadd_rn(J.rgbMaxF, 0.5f)), magicOne);
            }
            if (C::AM == AB_MUL_FLOAT_MUL || C::AM == AB_MUL_FLOAT_UNMUL) { /* ref: src/reformat.c:894-947 */
                const float ac = (sizeof(YT) == 2) ? fminf(am, magicMax) : am;
                const float Ac = clamp01f(div_int(__fsub_rn(ac, magicOne), J.yuvMaxDiv));
                if (C::AM == AB_MUL_FLOAT_MUL) {
                    /* Ac is clamped to [0,1]: Ac == 1 multiplies exactly (the reference skips it) and Ac == 0 gives exact
                     * zeros (the reference stores them): no selects */
                    Rv = __fmul_rn(Rv, Ac);
                    Gv = __fmul_rn(Gv, Ac);
                    Bv = __fmul_rn(Bv, Ac);
                } else {
                    /* branch- and select-free per channel: Ac == 0 -> 0, 0 < Ac < 1 -> min(c/Ac, 1), Ac == 1 -> untouched.
                     * The colour values are already clamped to [0,1], so with the divisor forced to 1 the quotient (x/1 == x,
                     * exactly) and its min with 1 leave them as they are; a zero "reciprocal" makes div_by_rcp return +0
                     * (q0 = x*0 = 0, remainder x, 0*x + 0 = 0). */
                    const bool zero = (Ac == 0.0f);
                    const float d = zero ? 1.0f : Ac;
                    const float r1 = zero ? 0.0f : rcp_refined(d);
                    Rv = min1(div_by_rcp(Rv, d, r1));
                    Gv = min1(div_by_rcp(Gv, d, r1));
                    Bv = min1(div_by_rcp(Bv, d, r1));
                }
            }
            float qr = quant_magic(Rv, J.rgbMaxF), qg = quant_magic(Gv, J.rgbMaxF), qb = quant_magic(Bv, J.rgbMaxF);
            if (C::MX == MX_IDENTITY8) { /* ref: :1278-1309 -- the samples themselves, no arithmetic (4:4:4 only) */
                qg = magic_sample<YT>(in.yw[r], t, K);
                qb = magic_sample<YT>(in.uw[0], t, K);
                qr = magic_sample<YT>(in.vw[0], t, K);
            }
            if (C::AM == AB_MUL_INT_PREMUL) { /* ref: src/alpha.c:179-192 on the quantised codes */
                /* a >= max leaves the code untouched and a == 0 zeroes it (src/alpha.c:180-186).  Both fall out of the
                 * arithmetic with a clamped to max: floor(RN(RN(c*max)/max) + 0.5) == c for every code c (the product's
                 * rounding error is < 2^-9 of a code even at 16 bits), and c*0 is 0 -- no per-channel selects needed. */
                const float ac = fminf(__fsub_rn(qaCopy, magicOne), J.rgbMaxF); /* the alpha code the channel holds */
                qr = __fadd_rz(__fadd_rn(divx(__fmul_rn(__fsub_rn(qr, magicOne), ac), J.rgbMaxDiv), 0.5f), magicOne);
                qg = __fadd_rz(__fadd_rn(divx(__fmul_rn(__fsub_rn(qg, magicOne), ac), J.rgbMaxDiv), 0.5f), magicOne);
                qb = __fadd_rz(__fadd_rn(divx(__fmul_rn(__fsub_rn(qb, magicOne), ac), J.rgbMaxDiv), 0.5f), magicOne);
            } else if (C::AM == AB_MUL_INT_PREMUL_FMA) { /* the same function of (c, a) for depths <= 12: pixel_math.cuh */
                const float ac = fminf(__fsub_rn(qaCopy, magicOne), J.rgbMaxF);
                const float negA23 = __fmul_rn(ac, -8388608.0f);
                qr = premul_fma_magic(qr, ac, negA23, J.pmR, J.pmK);
                qg = premul_fma_magic(qg, ac, negA23, J.pmR, J.pmK);
                qb = premul_fma_magic(qb, ac, negA23, J.pmR, J.pmK);
            } else if (C::AM == AB_MUL_INT_UNPREMUL) { /* ref: src/alpha.c:366-381 */
                const float af = __fsub_rn(qaCopy, magicOne);
                const bool opaque = (af >= J.rgbMaxF), zero = (af == 0.0f);
                const float lim = __fadd_rn(J.rgbMaxF, 0.5f);
                const float d = (opaque || zero) ? 1.0f : af;
                const float r1 = rcp_refined(d); /* codes * max over a code: exponents within +-33, inside the guarded range */
                const float pr = __fadd_rz(fminf(__fadd_rn(div_by_rcp(__fmul_rn(__fsub_rn(qr, magicOne), J.rgbMaxF), d, r1), 0.5f), lim), magicOne);
                const float pg = __fadd_rz(fminf(__fadd_rn(div_by_rcp(__fmul_rn(__fsub_rn(qg, magicOne), J.rgbMaxF), d, r1), 0.5f), lim), magicOne);
                const float pb = __fadd_rz(fminf(__fadd_rn(div_by_rcp(__fmul_rn(__fsub_rn(qb, magicOne), J.rgbMaxF), d, r1), 0.5f), lim), magicOne);
                qr = opaque ? qr : (zero ? magicOne : pr);
                qg = opaque ? qg : (zero ? magicOne : pg);
                qb = opaque ? qb : (zero ? magicOne : pb);
            }
            const float qa = HAS_A ? qaCopy : __uint_as_float(0x4B000000u | J.rgbMax);
            if (C::NCH == 3) {
                ch[t][0] = qr; ch[t][1] = qg; ch[t][2] = qb; ch[t][3] = 0.0f;
            } else if (C::AFIRST) {
                ch[t][0] = qa; ch[t][1] = qr; ch[t][2] = qg; ch[t][3] = qb;
            } else {
                ch[t][0] = qr; ch[t][1] = qg; ch[t][2] = qb; ch[t][3] = qa;
            }
        }
        /* pack the lane's 8 pixels (8*PIXB bytes) and store them: lanes are contiguous, 16-byte (8-byte for 3-byte pixels) vectors */
        constexpr int NW = 8 * PIXB / 4; /* 32-bit words of the lane */
        uint32_t words[NW];
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) {
            if (C::OUT == OUT_F16) {
                /* ref: src/reformat.c:1419-1443 -- every channel (alpha included): bits((float)code * (2^-112 / max)) >> 13, the
                 * multiplier itself a denormal float (so -ftz=false matters); two halves per word */
                const int v0 = 2 * wi, v1 = 2 * wi + 1;
                const uint32_t b0 = __float_as_uint(__fmul_rn(__fsub_rn(ch[v0 / C::NCH][v0 % C::NCH], magicOne), J.f16mul));
                const uint32_t b1 = __float_as_uint(__fmul_rn(__fsub_rn(ch[v1 / C::NCH][v1 % C::NCH], magicOne), J.f16mul));
                words[wi] = ((b0 >> 13) & 0xFFFFu) | ((b1 << 3) & 0xFFFF0000u);
            } else if (C::OUT == OUT_565) {
                /* ref: RGB565 macro, src/reformat.c:619: ((B >> 3) | ((G >> 2) << 5) | ((R >> 3) << 11)); two pixels per word */
                uint32_t px2[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t R8 = __float_as_uint(ch[2 * wi + h][0]), G8 = __float_as_uint(ch[2 * wi + h][1]), B8 = __float_as_uint(ch[2 * wi + h][2]);
                    px2[h] = ((B8 >> 3) & 0x1Fu) | ((G8 << 3) & 0x7E0u) | ((R8 << 8) & 0xF800u);
                }
                words[wi] = px2[0] | (px2[1] << 16);
            } else if (sizeof(RT) == 2) {
                const int v0 = 2 * wi, v1 = 2 * wi + 1;
                words[wi] = __byte_perm(__float_as_uint(ch[v0 / C::NCH][v0 % C::NCH]), __float_as_uint(ch[v1 / C::NCH][v1 % C::NCH]), 0x5410u);
            } else {
                const int v0 = 4 * wi;
                const uint32_t lo = __byte_perm(__float_as_uint(ch[v0 / C::NCH][v0 % C::NCH]), __float_as_uint(ch[(v0 + 1) / C::NCH][(v0 + 1) % C::NCH]), 0x0040u);
                const uint32_t hi = __byte_perm(__float_as_uint(ch[(v0 + 2) / C::NCH][(v0 + 2) % C::NCH]), __float_as_uint(ch[(v0 + 3) / C::NCH][(v0 + 3) % C::NCH]), 0x0040u);
                words[wi] = __byte_perm(lo, hi, 0x5410u);
            }
        }
        uint8_t * d = F.rgb + (int64_t)R.jrow[r] * F.rgbRB + (int64_t)px0 * PIXB;
        if (FULL || (nValid == 8 && F.vecWidth != 0)) {
            if ((8 * PIXB) % 16 == 0) {
#pragma unroll
                for (int q = 0; q < NW / 4; ++q) reinterpret_cast<uint4 *>(d)[q] = make_uint4(words[4 * q], words[4 * q + 1], words[4 * q + 2], words[4 * q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < NW / 2; ++q) reinterpret_cast<uint2 *>(d)[q] = make_uint2(words[2 * q], words[2 * q + 1]);
            }
        } else if (nValid) { /* ragged last lane of a row: 16-bit (or byte) granular */
            if (PIXB % 2 == 0) {
#pragma unroll
                for (int hw = 0; hw < 2 * NW; ++hw)
                    if ((uint32_t)(2 * hw) < nValid * PIXB) reinterpret_cast<uint16_t *>(d)[hw] = (uint16_t)(words[hw >> 1] >> ((hw & 1) * 16));
            } else {
#pragma unroll
                for (int b = 0; b < 4 * NW; ++b)
                    if ((uint32_t)b < nValid * PIXB) d[b] = (uint8_t)(words[b >> 2] >> ((b & 3) * 8));
            }
        }
    }
}

template <class C, bool FULL>
__device__ __forceinline__ void convert_store_tile(const AbY2RJob & J, const AbY2RFrame & F, const TileRows<C> & R, uint32_t x0, uint32_t lane, const TileIn<C> & in)
{
    float unused[2][6];
    convert_store_tile_x<C, FULL, false>(J, F, R, x0, lane, in, unused);
}

/* ---- the kernel ------------------------------------------------------------------------------------- */

/* Persistent, software-pipelined, batched.  Work unit = GROUP of 8 consecutive tiles of one frame (one per warp), so
 * the frame is block-uniform.  A block strides over the groups of all frames of the launch; each warp loads its tile
 * of group g+stride into a second register set before converting its tile of group g, so it always has 50-150 B of
 * loads in flight while it computes (at 25-50 % occupancy the scheduler alone cannot hide HBM latency: ncu showed 33 %
 * long_scoreboard + 20 % store-drain stalls on the 8K case with one tile per warp).  Tiles are row-major, so concurrently
 * running warps touch adjacent 256-pixel spans.  (Shared-memory staged stores measured slower: profiles/r01_variants.md.) */
/* Loop shape per chroma class, from the sweep in profiles/r01_variants.md (tools/variant_sweep.sh; -D switches for `make variant`):
 *   4:4:4            prefetched second register set, 3 blocks/SM (80 registers)   -- 4 blocks or no prefetch: 4-20 % slower
 *   nearest 4:2:x    plain loop, 5 blocks/SM (48 registers)                        -- 6-15 % faster than prefetch at 4 blocks
 *   bilinear 4:2:0   plain loop, 3 blocks/SM (the two-row tile keeps ~100 values live; 4 blocks spill)
 *   bilinear 4:2:2   prefetch, 3 blocks/SM */
#ifndef AB_Y2R_MINB
#define AB_Y2R_MINB 3
#endif
#ifndef AB_Y2R_MINB_BIL420
#define AB_Y2R_MINB_BIL420 3
#endif
#ifndef AB_Y2R_MINB_NEAREST
#define AB_Y2R_MINB_NEAREST 5
#endif
#ifndef AB_Y2R_PLAIN_NEAREST
#define AB_Y2R_PLAIN_NEAREST 1
#endif
#ifndef AB_Y2R_PREFETCH_BIL420 /* 1: second register set for the two-row bilinear tile too (needs AB_Y2R_MINB_BIL420=2: ~120 registers) */
#define AB_Y2R_PREFETCH_BIL420 0
#endif
#ifndef AB_Y2R_UNROLL2 /* 1: the prefetching loop written out twice (ping-pong register sets, no copies) */
#define AB_Y2R_UNROLL2 0
#endif
#ifndef AB_Y2R_PLAIN /* 1: no prefetched register set in any kernel */
#define AB_Y2R_PLAIN 0
#endif
template <class C>
struct YLoop {
    static constexpr bool PLAIN = (C::CHROMA == CH_BIL420 && AB_Y2R_PREFETCH_BIL420 == 0) || (AB_Y2R_PLAIN != 0) || (C::CHROMA == CH_NEAREST && AB_Y2R_PLAIN_NEAREST != 0);
    static constexpr int MIN_BLOCKS = (C::CHROMA == CH_BIL420) ? AB_Y2R_MINB_BIL420 : ((C::CHROMA == CH_NEAREST) ? AB_Y2R_MINB_NEAREST : AB_Y2R_MINB);
};
template <class C>
__global__ void __launch_bounds__(kWarps * 32, YLoop<C>::MIN_BLOCKS) y2r_fast_kernel(const __grid_constant__ AbY2RBatch B)
{
    const AbY2RJob & J = B.common;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t f = 0;
    while (f + 1 < B.nFrames && blockIdx.x >= B.frame[f + 1].blockStart) ++f; /* block-uniform, once per block */
    const AbY2RFrame & F = B.frame[f];                                         /* invariant for the life of the block */
    const uint32_t gStride = F.blocks;
    uint32_t g = blockIdx.x - F.blockStart;
    if (g >= F.groups) return;
    if (YLoop<C>::PLAIN) {
        /* the two-row bilinear tile already keeps ~100 values live: a second (prefetch) register set would cost a whole
         * resident block per SM and measured no gain (profiles/r01_variants.md) -- plain persistent loop */
        for (; g < F.groups; g += gStride) {
            const uint32_t t = g * kWarps + warp;
            if (t >= F.tiles) continue; /* warp-uniform */
            const uint32_t unit = t / F.tilesX, tx0 = (t - unit * F.tilesX) * kTile;
            const TileRows<C> TR = tile_rows<C>(J, F, unit);
            if (!TR.any) continue;
            TileIn<C> in;
            if (tx0 + kTile <= F.vecWidth) {
                load_tile<C, true>(J, F, TR, tx0, lane, in);
                convert_store_tile<C, true>(J, F, TR, tx0, lane, in);
            } else {
                load_tile<C, false>(J, F, TR, tx0, lane, in);
                convert_store_tile<C, false>(J, F, TR, tx0, lane, in);
            }
        }
        return;
    }
#if AB_Y2R_UNROLL2
    /* Two register sets that swap roles: the loop body is written out twice so that "the prefetched tile becomes the current
     * one" is a change of names, not ~20 register moves per tile. */
    struct Slot {
        TileIn<C> in;
        TileRows<C> R;
        uint32_t x0;
        bool live, full;
    };
    auto fetch = [&](Slot & s, uint32_t grp) {
        const uint32_t t = grp * kWarps + warp;
        s.live = t < F.tiles;
        s.x0 = 0;
        s.full = false;
        s.R.any = false;
        if (s.live) {
            const uint32_t unit = t / F.tilesX;
            s.x0 = (t - unit * F.tilesX) * kTile;
            s.R = tile_rows<C>(J, F, unit);
            s.full = (s.x0 + kTile <= F.vecWidth);
            if (s.R.any) {
                if (s.full) load_tile<C, true>(J, F, s.R, s.x0, lane, s.in);
                else load_tile<C, false>(J, F, s.R, s.x0, lane, s.in);
            }
        }
    };
    auto run = [&](const Slot & s) {
        if (s.live && s.R.any) {
            if (s.full) convert_store_tile<C, true>(J, F, s.R, s.x0, lane, s.in);
            else convert_store_tile<C, false>(J, F, s.R, s.x0, lane, s.in);
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
    TileIn<C> cur, nxt;
    uint32_t tile = g * kWarps + warp;
    bool live = tile < F.tiles; /* the last group of a frame may be partial */
    uint32_t x0 = 0;
    TileRows<C> R;
    R.any = false;
    bool full = false;
    if (live) {
        const uint32_t unit = tile / F.tilesX;
        x0 = (tile - unit * F.tilesX) * kTile;
        R = tile_rows<C>(J, F, unit);
        full = (x0 + kTile <= F.vecWidth);
        if (R.any) {
            if (full) load_tile<C, true>(J, F, R, x0, lane, cur);
            else load_tile<C, false>(J, F, R, x0, lane, cur);
        }
    }
    while (true) {
        const uint32_t ng = g + gStride;
        const bool more = ng < F.groups; /* block-uniform */
        uint32_t nx0 = 0;
        bool nlive = false, nfull = false;
        TileRows<C> NR;
        NR.any = false;
        if (more) {
            const uint32_t ntile = ng * kWarps + warp;
            nlive = ntile < F.tiles;
            if (nlive) {
                const uint32_t nunit = ntile / F.tilesX;
                nx0 = (ntile - nunit * F.tilesX) * kTile;
                NR = tile_rows<C>(J, F, nunit);
                nfull = (nx0 + kTile <= F.vecWidth);
                if (NR.any) {
                    if (nfull) load_tile<C, true>(J, F, NR, nx0, lane, nxt);
                    else load_tile<C, false>(J, F, NR, nx0, lane, nxt);
                }
            }
        }
        if (live && R.any) {
            if (full) convert_store_tile<C, true>(J, F, R, x0, lane, cur);
            else convert_store_tile<C, false>(J, F, R, x0, lane, cur);
        }
        if (!more) break;
        cur = nxt;
        g = ng;
        live = nlive;
        x0 = nx0;
        R = NR;
        full = nfull;
    }
#endif
}

/* ---- strip kernel (bilinear 4:2:0) -----------------------------------------------------------------------------------
 * A warp walks DOWN a 256-pixel column strip, kYChunk row units at a time.  Unit r converts luma rows 2r-1 and 2r from chroma rows
 * r-1 and r; unit r+1 needs rows r and r+1 -- so the lower chroma row of a unit, already loaded, clamped, normalised and joined
 * with its two neighbour columns, is handed to the next unit in 12 registers instead of being fetched and prepared again (a quarter
 * of the chroma work of the group kernel, which visits units in row-major order and shares nothing), and positions advance by
 * rows instead of being re-derived from a tile number.  Units that touch the picture's top / bottom edge, a band limit of the
 * host pipeline, or a ragged last strip take the general tile code. */
#ifndef AB_Y2R_STRIP_BIL420
#define AB_Y2R_STRIP_BIL420 1
#endif
#ifndef AB_Y2R_CHUNK
#define AB_Y2R_CHUNK 8
#endif
#ifndef AB_Y2R_MINB_STRIP /* resident blocks per SM the strip kernel is compiled for */
#define AB_Y2R_MINB_STRIP AB_Y2R_MINB_BIL420
#endif
#ifndef AB_Y2R_STRIP_ALPHA /* 1: tiles with alpha rows take the strip kernel too (they stayed on the group kernel while their rows lived in registers) */
#define AB_Y2R_STRIP_ALPHA 0
#endif
constexpr uint32_t kYChunk = AB_Y2R_CHUNK;

/* AB_Y2R_FEED = D > 0: the interior units of a strip are fed through shared memory.  Each lane issues the 16/8-byte asynchronous
 * copies (cp.async, SASS LDGSTS) of ITS OWN luma / alpha / chroma-row-B bytes of the unit D steps ahead into a per-warp ring of
 * D + 1 stages, and reads its own bytes back D units later: no register holds a sample while it is in flight, no lane reads
 * another lane's bytes (so there is no barrier and nothing to poll: cp.async.wait_group is a per-thread scoreboard wait), and
 * a warp has D + 1 units of loads in flight instead of one.  The two halo columns of the edge lanes travel one unit ahead in
 * two registers. */
#ifndef AB_Y2R_FEED
#define AB_Y2R_FEED 0
#endif
#ifndef AB_Y2R_LATE /* needs AB_Y2R_FEED: rows leave the ring only when they are converted; second-row chroma parked in shared memory.
                    * Compiled and inspected with ptxas only (it did not reach 64 registers): NEVER run on a GPU, no parity result exists */
#define AB_Y2R_LATE 0
#endif

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t smemAddr, const void * gptr)
{
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smemAddr), "l"(gptr) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr), "l"(gptr) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smemAddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

/* layout of one ring stage of one warp: luma rows 2r-1, 2r | alpha rows (if the tile needs them) | chroma row r of U, of V */
template <class C>
struct FeedLayout {
    using T = TileTraits<C>;
    static constexpr int LB = 8 * (int)sizeof(typename T::YT);  /* bytes of a lane's 8 luma samples */
    static constexpr int CBL = 4 * (int)sizeof(typename T::YT); /* bytes of a lane's 4 chroma samples */
    static constexpr int YROW = 32 * LB, CROW = 32 * CBL;
    static constexpr int OFF_A = 2 * YROW;
    static constexpr int OFF_U = (T::NEED_A ? 4 : 2) * YROW;
    static constexpr int OFF_V = OFF_U + CROW;
    static constexpr int STAGE = OFF_V + CROW;
    static constexpr int STAGES = AB_Y2R_FEED + 1;
    static constexpr int SPILL = AB_Y2R_LATE ? 16 * 128 : 0; /* 16 floats per lane, slot-major (AB_Y2R_LATE) */
    static constexpr int WARP_BYTES = STAGES * STAGE + SPILL;
    static constexpr int BLOCK_BYTES = kWarps * WARP_BYTES;
};

template <class C>
__global__ void __launch_bounds__(kWarps * 32, AB_Y2R_MINB_STRIP) y2r_bil420_strip_kernel(const __grid_constant__ AbY2RBatch B)
{
    using T = TileTraits<C>;
    using YT = typename T::YT;
    static_assert(C::CHROMA == CH_BIL420, "strip kernel: bilinear 4:2:0");
    const AbY2RJob & J = B.common;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if AB_Y2R_FEED
    extern __shared__ __align__(16) uint8_t feedSmem[];
    const uint32_t feedBase = (uint32_t)__cvta_generic_to_shared(feedSmem);
#endif
    uint32_t f = 0;
    while (f + 1 < B.nFrames && blockIdx.x >= B.frame[f + 1].blockStart) ++f;
    const AbY2RFrame & F = B.frame[f];
    const uint32_t gStride = F.blocks;
    const uint32_t g0 = blockIdx.x - F.blockStart;
    if (g0 >= F.groups) return;
    const uint32_t units = (F.rowEnd >> 1) - (F.rowBegin >> 1) + 1;
    const uint32_t cw = (F.width + 1) >> 1;
#pragma unroll 1
    for (uint32_t g = g0; g < F.groups; g += gStride) {
        const uint32_t c = g * kWarps + warp;
        if (c >= F.tiles) continue; /* warp-uniform */
        const uint32_t cy = c / F.tilesX, sx = c - cy * F.tilesX;
        const uint32_t x0 = sx * kTile, px0 = x0 + 8 * lane, m0 = px0 >> 1;
        const uint32_t u0 = cy * kYChunk, u1 = min(u0 + kYChunk, units);
        const bool full = (x0 + kTile <= F.vecWidth);
        float carry[2][6];
        bool have = false;
#if AB_Y2R_FEED
        using L = FeedLayout<C>;
        const uint32_t ring = feedBase + warp * (uint32_t)L::WARP_BYTES;
        /* the copies of unit `pu` into stage pu % STAGES (nothing if the unit is not an interior one); always one commit group */
        auto feed = [&](uint32_t pu) {
            if (pu < u1) {
                const TileRows<C> PR = tile_rows<C>(J, F, pu);
                const int pru = (int)(F.rowBegin >> 1) + (int)pu;
                if (full && PR.rowOn[0] && PR.rowOn[1] && (PR.cA == pru - 1) && (PR.cB == pru)) {
                    const uint32_t st = ring + (pu % (uint32_t)L::STAGES) * (uint32_t)L::STAGE;
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        cp_async<L::LB>(st + r * L::YROW + lane * L::LB, F.y + (int64_t)PR.jrow[r] * F.yRB + (size_t)px0 * sizeof(YT));
                        if (T::NEED_A) cp_async<L::LB>(st + L::OFF_A + r * L::YROW + lane * L::LB, F.a + (int64_t)PR.jrow[r] * F.aRB + (size_t)px0 * sizeof(YT));
                    }
                    cp_async<L::CBL>(st + L::OFF_U + lane * L::CBL, F.u + (int64_t)PR.cB * F.uRB + (size_t)m0 * sizeof(YT));
                    cp_async<L::CBL>(st + L::OFF_V + lane * L::CBL, F.v + (int64_t)PR.cB * F.vRB + (size_t)m0 * sizeof(YT));
                }
            }
            cp_async_commit();
        };
        /* halo columns (edge lanes) of chroma row `crow`, as magic floats */
        const uint32_t hcol = (lane == 0) ? ((m0 > 0) ? m0 - 1 : 0) : min(m0 + 4, cw - 1);
        auto halo = [&](int crow, float & hu, float & hv) {
            hu = hv = 0.0f;
            if (lane == 0 || lane == 31) {
                hu = magic_scalar<YT>(F.u + (int64_t)crow * F.uRB, hcol);
                hv = magic_scalar<YT>(F.v + (int64_t)crow * F.vRB, hcol);
            }
        };
#pragma unroll
        for (int d = 0; d < AB_Y2R_FEED; ++d) feed(u0 + (uint32_t)d);
        float hUn = 0.0f, hVn = 0.0f; /* halos of chroma row B of the unit about to run (loaded one unit ahead) */
        {
            const int chromaH = (int)((F.height + 1) >> 1);
            const int ru0 = (int)(F.rowBegin >> 1) + (int)u0;
            if (full && ru0 < chromaH) halo(ru0, hUn, hVn);
        }
#endif
#pragma unroll 1
        for (uint32_t unit = u0; unit < u1; ++unit) {
#if AB_Y2R_FEED
            feed(unit + AB_Y2R_FEED);
            const float hUc = hUn, hVc = hVn;
            {
                const int chromaH = (int)((F.height + 1) >> 1);
                const int run = (int)(F.rowBegin >> 1) + (int)unit + 1;
                if (full && unit + 1 < u1 && run < chromaH) halo(run, hUn, hVn);
            }
#endif
            const TileRows<C> R = tile_rows<C>(J, F, unit);
            if (!R.any) {
                have = false;
                continue;
            }
            const int ru = (int)(F.rowBegin >> 1) + (int)unit;
            const bool interior = full && R.rowOn[0] && R.rowOn[1] && (R.cA == ru - 1) && (R.cB == ru);
            TileIn<C> in;
            if (!interior) { /* picture / band edge or ragged strip */
                if (full) {
                    load_tile<C, true>(J, F, R, x0, lane, in);
                    convert_store_tile_x<C, true, false>(J, F, R, x0, lane, in, carry);
                    have = (R.cB == ru); /* carry holds chroma row ru: what the unit below calls its row A */
                } else {
                    load_tile<C, false>(J, F, R, x0, lane, in);
                    convert_store_tile_x<C, false, false>(J, F, R, x0, lane, in, carry);
                    have = false;
                }
                continue;
            }
            /* interior unit of a full strip: two luma (and alpha) rows, chroma row B; row A only if the unit above did not leave it */
#if AB_Y2R_FEED
            const uint8_t * lateStage = nullptr;
            cp_async_wait<AB_Y2R_FEED>(); /* this unit's group (and every older one) has landed */
            {
                const uint8_t * st = feedSmem + (size_t)warp * L::WARP_BYTES + (size_t)(unit % (uint32_t)L::STAGES) * L::STAGE;
                lateStage = st;
#pragma unroll
                for (int r = 0; r < (AB_Y2R_LATE ? 0 : 2); ++r) {
                    if (sizeof(YT) == 2) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(st + r * L::YROW + lane * 16);
                        in.yw[r][0] = v.x; in.yw[r][1] = v.y; in.yw[r][2] = v.z; in.yw[r][3] = v.w;
                        if (T::NEED_A) {
                            const uint4 a = *reinterpret_cast<const uint4 *>(st + L::OFF_A + r * L::YROW + lane * 16);
                            in.aw[r][0] = a.x; in.aw[r][1] = a.y; in.aw[r][2] = a.z; in.aw[r][3] = a.w;
                        }
                    } else {
                        const uint2 v = *reinterpret_cast<const uint2 *>(st + r * L::YROW + lane * 8);
                        in.yw[r][0] = v.x; in.yw[r][1] = v.y; in.yw[r][2] = in.yw[r][3] = 0;
                        if (T::NEED_A) {
                            const uint2 a = *reinterpret_cast<const uint2 *>(st + L::OFF_A + r * L::YROW + lane * 8);
                            in.aw[r][0] = a.x; in.aw[r][1] = a.y; in.aw[r][2] = in.aw[r][3] = 0;
                        }
                    }
                }
                if (sizeof(YT) == 2) {
                    const uint2 u = *reinterpret_cast<const uint2 *>(st + L::OFF_U + lane * 8);
                    const uint2 v = *reinterpret_cast<const uint2 *>(st + L::OFF_V + lane * 8);
                    in.uw[1][0] = u.x; in.uw[1][1] = u.y; in.vw[1][0] = v.x; in.vw[1][1] = v.y;
                } else {
                    in.uw[1][0] = *reinterpret_cast<const uint32_t *>(st + L::OFF_U + lane * 4);
                    in.vw[1][0] = *reinterpret_cast<const uint32_t *>(st + L::OFF_V + lane * 4);
                    in.uw[1][1] = in.vw[1][1] = 0;
                }
                in.hU[1] = hUc;
                in.hV[1] = hVc;
            }
#else
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                load8<YT>(F.y + (int64_t)R.jrow[r] * F.yRB, px0, 8u, in.yw[r]);
                if (T::NEED_A) load8<YT>(F.a + (int64_t)R.jrow[r] * F.aRB, px0, 8u, in.aw[r]);
            }
#endif
#pragma unroll
            for (int cr = 0; cr < (AB_Y2R_FEED ? 1 : 2); ++cr) {
                if (cr == 0 && have) { /* warp-uniform */
                    in.uw[0][0] = in.uw[0][1] = in.vw[0][0] = in.vw[0][1] = 0;
                    in.hU[0] = in.hV[0] = 0.0f;
                    continue;
                }
                const int crow = (cr == 0) ? R.cA : R.cB;
                const uint8_t * ur = F.u + (int64_t)crow * F.uRB;
                const uint8_t * vr = F.v + (int64_t)crow * F.vRB;
                load4_clamped<YT>(ur, m0, m0 + 4, in.uw[cr]);
                load4_clamped<YT>(vr, m0, m0 + 4, in.vw[cr]);
                in.hU[cr] = in.hV[cr] = 0.0f;
                if (lane == 0 || lane == 31) {
                    const uint32_t col = (lane == 0) ? ((m0 > 0) ? m0 - 1 : 0) : min(m0 + 4, cw - 1);
                    in.hU[cr] = magic_scalar<YT>(ur, col);
                    in.hV[cr] = magic_scalar<YT>(vr, col);
                }
            }
#if AB_Y2R_FEED && AB_Y2R_LATE
            {
                const LateFeed late = { lateStage + lane * L::LB, (uint32_t)L::YROW, (uint32_t)L::OFF_A,
                                        ring + (uint32_t)(L::STAGES * L::STAGE) + lane * 4u };
                if (have) convert_store_tile_x<C, true, true, true>(J, F, R, x0, lane, in, carry, late);
                else convert_store_tile_x<C, true, false, true>(J, F, R, x0, lane, in, carry, late);
            }
#else
            if (have) convert_store_tile_x<C, true, true>(J, F, R, x0, lane, in, carry);
            else convert_store_tile_x<C, true, false>(J, F, R, x0, lane, in, carry);
#endif
            have = true;
        }
    }
}

/* ---- TMA-staged variant (bilinear 4:2:0) -------------------------------------------------------------------
 * The plain loop above stalls whenever all resident warps of an SM sub-partition reach the loads of their next tile together
 * (ncu r01s: 2.4 eligible warps per cycle on average, yet 26 % of the cycles without any: long_scoreboard in bursts), and a
 * second register set for prefetching does not fit next to the two-row tile.  Here the TMA engine does the prefetching: one
 * lane of each warp issues 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of the NEXT tiles' plane rows -- two luma rows (and
 * alpha rows), two chroma rows per plane, 256 / 128 samples each -- into a ring of AB_Y2R_TMA_STAGES warp-private shared-memory
 * stages, each guarded by an mbarrier that counts the bytes landed; the warp then reads a stage with one conflict-free
 * LDS.128 / LDS.64 per row into the same TileIn registers and runs the unchanged convert_store_tile.  No registers are spent on
 * data in flight, no __syncthreads: warps stay independent.  The filter's halo columns (one chroma sample left of lane 0 and
 * right of lane 31) are two scalar loads by the edge lanes, issued one tile ahead into four registers: replicated at the picture
 * edge as the reference does (src/reformat.c:768-795) -- which is also why a 2-D tensor map is not used: it would zero-fill
 * there, and a batch of frames has no room for eight descriptors per frame in the kernel parameters.  Ragged last tiles of a
 * row (fewer than 256 pixels) and planes whose rows are not 16-byte aligned take the plain loads. */
namespace tma {
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * dstSmem, const void * srcGlobal, uint32_t bytes, uint64_t * bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dstSmem)), "l"(srcGlobal),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
} // namespace tma

template <class C>
struct StageLayout {
    using T = TileTraits<C>;
    using YT = typename T::YT;
    static constexpr int YB = kTile * (int)sizeof(YT);       /* bytes of one luma / alpha row segment */
    static constexpr int CB = (kTile / 2) * (int)sizeof(YT); /* bytes of one chroma row segment */
    static constexpr int OFF_A = T::ROWS * YB;
    static constexpr int OFF_U = OFF_A + (T::NEED_A ? T::ROWS * YB : 0);
    static constexpr int OFF_V = OFF_U + T::CROWS * CB;
    static constexpr int BYTES = OFF_V + T::CROWS * CB;
    static constexpr int STAGES = AB_Y2R_TMA_STAGES;
    static constexpr int WARP_BYTES = STAGES * BYTES;
    static constexpr int BLOCK_BYTES = kWarps * WARP_BYTES + kWarps * STAGES * 8;
};

/* position of the i-th tile of this warp (tiles of a block's groups, one per warp) */
struct WarpTile {
    uint32_t unit, x0;
    bool live, full;
};
__device__ __forceinline__ WarpTile warp_tile(const AbY2RFrame & F, uint32_t g, uint32_t warp)
{
    WarpTile w;
    const uint32_t t = g * kWarps + warp;
    w.live = (g < F.groups) && (t < F.tiles);
    w.unit = w.live ? t / F.tilesX : 0;
    w.x0 = w.live ? (t - w.unit * F.tilesX) * kTile : 0;
    w.full = w.live && (w.x0 + kTile <= F.vecWidth);
    return w;
}

template <class C>
__global__ void __launch_bounds__(kWarps * 32, YLoop<C>::MIN_BLOCKS) y2r_fast_staged_kernel(const __grid_constant__ AbY2RBatch B)
{
    using T = TileTraits<C>;
    using YT = typename T::YT;
    using L = StageLayout<C>;
    static_assert(C::CHROMA == CH_BIL420, "staged kernel: bilinear 4:2:0 tiles");
    extern __shared__ __align__(128) uint8_t smemRaw[];
    const AbY2RJob & J = B.common;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t f = 0;
    while (f + 1 < B.nFrames && blockIdx.x >= B.frame[f + 1].blockStart) ++f;
    const AbY2RFrame & F = B.frame[f];
    const uint32_t gStride = F.blocks;
    const uint32_t g0 = blockIdx.x - F.blockStart;
    if (g0 >= F.groups) return;
    uint8_t * ring = smemRaw + warp * L::WARP_BYTES;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smemRaw + kWarps * L::WARP_BYTES) + warp * L::STAGES;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < L::STAGES; ++k) tma::mbar_init(&bars[k], 1);
        tma::fence_proxy_async(); /* make the initialised barriers visible to the async proxy */
    }
    __syncwarp();
    const uint32_t cw = (F.width + 1) >> 1;

    /* issue the bulk copies of tile (g) into stage `s` (lane 0); returns nothing: the consumer re-derives what was issued */
    auto issue = [&](uint32_t g, int s) {
        const WarpTile w = warp_tile(F, g, warp);
        if (!w.full) return;
        const TileRows<C> R = tile_rows<C>(J, F, w.unit);
        if (!R.any) return;
        if (lane == 0) {
            uint8_t * st = ring + s * L::BYTES;
            uint32_t bytes = 2u * T::CROWS * L::CB;
#pragma unroll
            for (int r = 0; r < T::ROWS; ++r)
                if (R.rowOn[r]) bytes += (T::NEED_A ? 2u : 1u) * L::YB;
            tma::mbar_expect_tx(&bars[s], bytes);
#pragma unroll
            for (int r = 0; r < T::ROWS; ++r) {
                if (!R.rowOn[r]) continue;
                tma::bulk_g2s(st + r * L::YB, F.y + (int64_t)R.jrow[r] * F.yRB + (size_t)w.x0 * sizeof(YT), L::YB, &bars[s]);
                if (T::NEED_A) tma::bulk_g2s(st + L::OFF_A + r * L::YB, F.a + (int64_t)R.jrow[r] * F.aRB + (size_t)w.x0 * sizeof(YT), L::YB, &bars[s]);
            }
#pragma unroll
            for (int c = 0; c < T::CROWS; ++c) {
                const int crow = (c == 0) ? R.cA : R.cB;
                tma::bulk_g2s(st + L::OFF_U + c * L::CB, F.u + (int64_t)crow * F.uRB + (size_t)(w.x0 >> 1) * sizeof(YT), L::CB, &bars[s]);
                tma::bulk_g2s(st + L::OFF_V + c * L::CB, F.v + (int64_t)crow * F.vRB + (size_t)(w.x0 >> 1) * sizeof(YT), L::CB, &bars[s]);
            }
        }
    };
    /* halo columns of a full tile: magic floats, only meaningful in lanes 0 and 31 */
    auto halo = [&](const WarpTile & w, const TileRows<C> & R, float hU[2], float hV[2]) {
        hU[0] = hU[1] = hV[0] = hV[1] = 0.0f;
        if (w.full && R.any && (lane == 0 || lane == 31)) {
            const uint32_t m0 = (w.x0 + 8 * lane) >> 1;
            const uint32_t col = (lane == 0) ? ((m0 > 0) ? m0 - 1 : 0) : min(m0 + 4, cw - 1);
#pragma unroll
            for (int c = 0; c < T::CROWS; ++c) {
                const int crow = (c == 0) ? R.cA : R.cB;
                hU[c] = magic_scalar<YT>(F.u + (int64_t)crow * F.uRB, col);
                hV[c] = magic_scalar<YT>(F.v + (int64_t)crow * F.vRB, col);
            }
        }
    };

    /* prologue: fill the ring */
#pragma unroll
    for (int k = 0; k < L::STAGES; ++k) issue(g0 + (uint32_t)k * gStride, k);
    uint32_t phase = 0; /* bit s = parity the consumer waits for on stage s */
    float nhU[2], nhV[2];
    {
        const WarpTile w = warp_tile(F, g0, warp);
        const TileRows<C> R0 = tile_rows<C>(J, F, w.unit);
        halo(w, R0, nhU, nhV);
    }
    int s = 0;
    for (uint32_t g = g0; g < F.groups; g += gStride) { /* block-uniform trip count */
        const WarpTile w = warp_tile(F, g, warp);
        const TileRows<C> R = tile_rows<C>(J, F, w.unit);
        const bool staged = w.full && R.any;
        TileIn<C> in;
        if (staged) {
            tma::mbar_wait(&bars[s], (phase >> s) & 1u);
            phase ^= 1u << s;
            const uint8_t * st = ring + s * L::BYTES;
#pragma unroll
            for (int r = 0; r < T::ROWS; ++r) {
                in.yw[r][0] = in.yw[r][1] = in.yw[r][2] = in.yw[r][3] = 0;
                if (T::NEED_A) in.aw[r][0] = in.aw[r][1] = in.aw[r][2] = in.aw[r][3] = 0;
                if (R.rowOn[r]) {
                    if (sizeof(YT) == 2) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(st + r * L::YB + lane * 16);
                        in.yw[r][0] = v.x; in.yw[r][1] = v.y; in.yw[r][2] = v.z; in.yw[r][3] = v.w;
                        if (T::NEED_A) {
                            const uint4 a = *reinterpret_cast<const uint4 *>(st + L::OFF_A + r * L::YB + lane * 16);
                            in.aw[r][0] = a.x; in.aw[r][1] = a.y; in.aw[r][2] = a.z; in.aw[r][3] = a.w;
                        }
                    } else {
                        const uint2 v = *reinterpret_cast<const uint2 *>(st + r * L::YB + lane * 8);
                        in.yw[r][0] = v.x; in.yw[r][1] = v.y;
                        if (T::NEED_A) {
                            const uint2 a = *reinterpret_cast<const uint2 *>(st + L::OFF_A + r * L::YB + lane * 8);
                            in.aw[r][0] = a.x; in.aw[r][1] = a.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < T::CROWS; ++c) {
                if (sizeof(YT) == 2) {
                    const uint2 u = *reinterpret_cast<const uint2 *>(st + L::OFF_U + c * L::CB + lane * 8);
                    const uint2 v = *reinterpret_cast<const uint2 *>(st + L::OFF_V + c * L::CB + lane * 8);
                    in.uw[c][0] = u.x; in.uw[c][1] = u.y;
                    in.vw[c][0] = v.x; in.vw[c][1] = v.y;
                } else {
                    in.uw[c][0] = *reinterpret_cast<const uint32_t *>(st + L::OFF_U + c * L::CB + lane * 4);
                    in.vw[c][0] = *reinterpret_cast<const uint32_t *>(st + L::OFF_V + c * L::CB + lane * 4);
                    in.uw[c][1] = in.vw[c][1] = 0;
                }
                in.hU[c] = nhU[c];
                in.hV[c] = nhV[c];
            }
        }
        /* halo columns of the next tile: in flight while this one converts */
        {
            const WarpTile wn = warp_tile(F, g + gStride, warp);
            const TileRows<C> Rn = tile_rows<C>(J, F, wn.unit);
            halo(wn, Rn, nhU, nhV);
        }
        if (staged) {
            convert_store_tile<C, true>(J, F, R, w.x0, lane, in);
        } else if (w.live && R.any) { /* ragged last tile of a row: plain loads */
            load_tile<C, false>(J, F, R, w.x0, lane, in);
            convert_store_tile<C, false>(J, F, R, w.x0, lane, in);
        }
        /* every lane has consumed its shared-memory reads (their values fed the arithmetic above; an unstaged tile never used the
         * stage): refill it with the tile STAGES steps ahead */
        __syncwarp();
        if (lane == 0) tma::fence_proxy_async();
        issue(g + (uint32_t)L::STAGES * gStride, s);
        s = (s + 1 == L::STAGES) ? 0 : s + 1;
    }
}

/* share `blocksWanted` blocks among the frames in proportion to their group counts (every frame gets at least one) */
template <class FrameT>
static uint32_t assign_blocks(FrameT * frames, uint32_t n, uint32_t totalGroups, uint32_t blocksWanted)
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

/* ---- eligibility + dispatch for one (YT, RT) pair ------------------------------------------------ */

template <typename YT, typename RT, int CHROMA, int NCH, bool AFIRST, int AW, int AM, int OUT = OUT_INT, int MX = MX_COEFF, bool ALLOW_STRIP = true>
static int launch_cfg(AbY2RBatch & B, cudaStream_t st)
{
    using C = Cfg<YT, RT, CHROMA, NCH, AFIRST, AW, AM, OUT, MX>;
    /* the strip kernel pays where the tile needs no alpha rows (C5 -3 %, C1 -7 %); with them its 12 carried registers spill and it
     * measured 14 % SLOWER than the group kernel (C3 110 vs 96.6 us, profiles/r02_variants.md) */
    constexpr bool STRIP = ALLOW_STRIP && (CHROMA == CH_BIL420) && (AB_Y2R_STRIP_BIL420 != 0) && (AB_Y2R_STRIP_ALPHA != 0 || !TileTraits<C>::NEED_A);
    constexpr size_t stripSmem = (STRIP && AB_Y2R_FEED) ? (size_t)FeedLayout<C>::BLOCK_BYTES : 0; /* the cp.async ring of the strip kernel */
    uint32_t groups = 0;
    for (uint32_t f = 0; f < B.nFrames; ++f) {
        AbY2RFrame & F = B.frame[f];
        const uint32_t rows = F.rowEnd - F.rowBegin;
        uint32_t units = rows;
        if (CHROMA == CH_BIL420) units = (F.rowEnd >> 1) - (F.rowBegin >> 1) + 1; /* row pairs (2r-1, 2r) touching [rowBegin,rowEnd) */
        if (STRIP) units = (units + kYChunk - 1) / kYChunk;                        /* work items: chunks of kYChunk units of one strip */
        F.tilesX = (F.width + kTile - 1) / kTile;
        F.tiles = F.tilesX * units;
        F.groups = (F.tiles + kWarps - 1) / kWarps;
        groups += F.groups;
    }
    /* blocks that fit on the whole device at once: per instantiation and per device (a process may drive several GPUs) */
    static int residentOf[16] = { 0 };
    int dev = 0;
    cudaGetDevice(&dev);
    int resident = (dev >= 0 && dev < 16) ? residentOf[dev] : 0;
    if (!resident) {
        int sms = 0, perSm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if constexpr (STRIP) {
            if (stripSmem > 48 * 1024) cudaFuncSetAttribute(y2r_bil420_strip_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stripSmem);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, y2r_bil420_strip_kernel<C>, kWarps * 32, stripSmem);
        } else {
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, y2r_fast_kernel<C>, kWarps * 32, 0);
        }
        /* the loop shapes were tuned at YLoop::MIN_BLOCKS blocks per SM; when the compiler needs fewer registers than the bound
         * allows, more blocks would fit -- and measured slower (4:2:0 nearest: 6 blocks 14.8 us, 5 blocks 13.3 us per 4K frame) */
        constexpr int tuned = STRIP ? AB_Y2R_MINB_STRIP : YLoop<C>::MIN_BLOCKS;
        if (perSm > tuned) perSm = tuned;
        resident = (sms > 0 && perSm > 0) ? sms * perSm : 148 * 2;
        if (dev >= 0 && dev < 16) residentOf[dev] = resident;
    }
    if constexpr (STRIP) {
        /* a strip chunk is kYChunk units walked one after the other by one warp: a lone small picture (one 1080p frame: 68 groups for
         * 444 resident blocks) leaves most of the device idle behind those serial chains -- 22.6 us against 11 us for the group kernel,
         * whose units are independent.  The strip form needs at least two rounds of resident blocks. */
        if (groups < 2u * (uint32_t)resident) return launch_cfg<YT, RT, CHROMA, NCH, AFIRST, AW, AM, OUT, MX, false>(B, st);
    }
    const uint32_t wanted = (groups >= 4u * (uint32_t)resident) ? (uint32_t)resident : groups;
    B.totalBlocks = assign_blocks(B.frame, B.nFrames, groups, wanted);
    if constexpr (STRIP) y2r_bil420_strip_kernel<C><<<B.totalBlocks, kWarps * 32, stripSmem, st>>>(B);
    else y2r_fast_kernel<C><<<B.totalBlocks, kWarps * 32, 0, st>>>(B);
    return 1;
}

/* bilinear 4:2:0 through the TMA-staged kernel: every plane row segment must be a legal bulk-copy source (16-byte aligned) */
static inline bool tma_aligned(const AbY2RBatch & B, bool needA)
{
    for (uint32_t f = 0; f < B.nFrames; ++f) {
        const AbY2RFrame & F = B.frame[f];
        uintptr_t bits = (uintptr_t)F.y | (uintptr_t)F.yRB | (uintptr_t)F.u | (uintptr_t)F.uRB | (uintptr_t)F.v | (uintptr_t)F.vRB;
        if (needA) bits |= (uintptr_t)F.a | (uintptr_t)F.aRB;
        if (bits & 15) return false;
    }
    return true;
}

template <typename YT, typename RT, int NCH, bool AFIRST, int AW, int AM, int OUT = OUT_INT>
static int launch_cfg_staged(AbY2RBatch & B, cudaStream_t st)
{
    using C = Cfg<YT, RT, CH_BIL420, NCH, AFIRST, AW, AM, OUT>;
    using L = StageLayout<C>;
    if (!AB_Y2R_TMA || OUT != OUT_INT || !tma_aligned(B, TileTraits<C>::NEED_A)) return launch_cfg<YT, RT, CH_BIL420, NCH, AFIRST, AW, AM, OUT>(B, st);
    uint32_t groups = 0;
    for (uint32_t f = 0; f < B.nFrames; ++f) {
        AbY2RFrame & F = B.frame[f];
        const uint32_t units = (F.rowEnd >> 1) - (F.rowBegin >> 1) + 1; /* row pairs (2r-1, 2r) touching [rowBegin,rowEnd) */
        F.tilesX = (F.width + kTile - 1) / kTile;
        F.tiles = F.tilesX * units;
        F.groups = (F.tiles + kWarps - 1) / kWarps;
        groups += F.groups;
    }
    static int residentOf[16] = { 0 };
    int dev = 0;
    cudaGetDevice(&dev);
    int resident = (dev >= 0 && dev < 16) ? residentOf[dev] : 0;
    if (!resident) {
        int sms = 0, perSm = 0;
        cudaFuncSetAttribute(y2r_fast_staged_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BLOCK_BYTES);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, y2r_fast_staged_kernel<C>, kWarps * 32, L::BLOCK_BYTES);
        resident = (sms > 0 && perSm > 0) ? sms * perSm : 148 * 2;
        if (dev >= 0 && dev < 16) residentOf[dev] = resident;
    }
    /* the ring only pays when a warp sees several tiles: small launches keep one group per block on the plain kernel */
    if (groups < 4u * (uint32_t)resident) return launch_cfg<YT, RT, CH_BIL420, NCH, AFIRST, AW, AM, OUT>(B, st);
    B.totalBlocks = assign_blocks(B.frame, B.nFrames, groups, (uint32_t)resident);
    y2r_fast_staged_kernel<C><<<B.totalBlocks, kWarps * 32, L::BLOCK_BYTES, st>>>(B);
    return 1;
}

template <typename YT, typename RT, int CHROMA, int NCH, bool AFIRST, int AW, int AM, int OUT = OUT_INT>
static int launch_any(AbY2RBatch & B, cudaStream_t st)
{
    if constexpr (CHROMA == CH_BIL420 && OUT == OUT_INT && AB_Y2R_TMA) return launch_cfg_staged<YT, RT, NCH, AFIRST, AW, AM, OUT>(B, st);
    else return launch_cfg<YT, RT, CHROMA, NCH, AFIRST, AW, AM, OUT>(B, st);
}

/* identity matrix (lossless GBR, 4:4:4): straight alpha or none, integer outputs */
template <typename YT, typename RT>
static int launch_identity(AbY2RBatch & B, cudaStream_t st)
{
    const AbY2RJob & J = B.common;
    if (J.alphaMul != AB_MUL_NONE) return 0;
    if (J.alphaWrite == AB_ALPHA_COPY && J.alphaRescale && sizeof(YT) == sizeof(RT)) return 0;
    const bool afirst = J.alphaFirst != 0;
#define ABF_ID(MXV)                                                                                                                     \
    do {                                                                                                                                \
        if (J.nChan == 3) return launch_cfg<YT, RT, CH_FULL, 3, false, AB_ALPHA_KEEP, AB_MUL_NONE, OUT_INT, MXV>(B, st);                 \
        if (J.alphaWrite == AB_ALPHA_FILL)                                                                                              \
            return afirst ? launch_cfg<YT, RT, CH_FULL, 4, true, AB_ALPHA_FILL, AB_MUL_NONE, OUT_INT, MXV>(B, st)                       \
                          : launch_cfg<YT, RT, CH_FULL, 4, false, AB_ALPHA_FILL, AB_MUL_NONE, OUT_INT, MXV>(B, st);                     \
        if (J.alphaWrite == AB_ALPHA_COPY)                                                                                              \
            return afirst ? launch_cfg<YT, RT, CH_FULL, 4, true, AB_ALPHA_COPY, AB_MUL_NONE, OUT_INT, MXV>(B, st)                       \
                          : launch_cfg<YT, RT, CH_FULL, 4, false, AB_ALPHA_COPY, AB_MUL_NONE, OUT_INT, MXV>(B, st);                     \
        return 0;                                                                                                                       \
    } while (0)
    if (J.mode == AB_MODE_IDENTITY8_SHUFFLE) {
        if constexpr (sizeof(YT) == 1 && sizeof(RT) == 1) ABF_ID(MX_IDENTITY8);
        else return 0;
    }
    ABF_ID(MX_IDENTITY);
#undef ABF_ID
}

/* half-float and RGB565 outputs (Android's RGBA_F16 / RGB_565 bitmaps, android_jni/.../libavif_jni.cc:206-223): the combinations
 * real callers use -- colour-first layouts, alpha filled, copied or (pre)multiplied the way the reference's variant requires */
template <typename YT, typename RT, int CHROMA>
static int launch_special(AbY2RBatch & B, cudaStream_t st)
{
    const AbY2RJob & J = B.common;
    constexpr bool slow = (CHROMA == CH_BIL420 || CHROMA == CH_BIL422);
    if constexpr (sizeof(RT) == 1) {
        if (J.is565) {
            if (J.alphaMul == AB_MUL_NONE) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_NONE, OUT_565>(B, st);
            if (J.alphaMul == AB_MUL_FLOAT_MUL) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_FLOAT_MUL, OUT_565>(B, st);
        }
        return 0;
    } else {
        if (!J.toF16 || J.alphaFirst) return 0;
        if (J.nChan == 3) {
            if (J.alphaMul == AB_MUL_NONE) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_NONE, OUT_F16>(B, st);
            if (J.alphaMul == AB_MUL_FLOAT_MUL) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_FLOAT_MUL, OUT_F16>(B, st);
            return 0;
        }
        if (J.alphaWrite == AB_ALPHA_FILL) return (J.alphaMul == AB_MUL_NONE) ? launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_FILL, AB_MUL_NONE, OUT_F16>(B, st) : 0;
        if (J.alphaWrite != AB_ALPHA_COPY) return 0;
        if constexpr (sizeof(YT) == 2) {
            if (J.alphaRescale) { /* 10 / 12-bit image into half floats: the usual HDR case */
                switch (J.alphaMul) {
                    case AB_MUL_NONE: return launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_NONE, OUT_F16>(B, st);
                    case AB_MUL_INT_PREMUL: return slow ? 0 : launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_INT_PREMUL, OUT_F16>(B, st);
                    case AB_MUL_FLOAT_MUL: return !slow ? 0 : launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_FLOAT_MUL, OUT_F16>(B, st);
                    default: return 0;
                }
            }
        }
        switch (J.alphaMul) {
            case AB_MUL_NONE: return launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_NONE, OUT_F16>(B, st);
            case AB_MUL_INT_PREMUL: return slow ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_INT_PREMUL, OUT_F16>(B, st); /* depth 16: no fma form */
            case AB_MUL_FLOAT_MUL: return !slow ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_FLOAT_MUL, OUT_F16>(B, st);
            default: return 0;
        }
    }
}

template <typename YT, typename RT, int CHROMA, int NCH, bool AFIRST>
static int launch_alpha(AbY2RBatch & B, cudaStream_t st)
{
    const AbY2RJob & J = B.common;
    constexpr bool slow = (CHROMA == CH_BIL420 || CHROMA == CH_BIL422); /* general loop: alpha math in float; fast loops: integer post-pass */
    if (NCH == 3) {
        if (J.alphaMul == AB_MUL_NONE) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_NONE>(B, st);
        if (J.alphaMul == AB_MUL_FLOAT_MUL) return launch_any<YT, RT, CHROMA, 3, false, AB_ALPHA_KEEP, AB_MUL_FLOAT_MUL>(B, st);
        return 0;
    }
    if (J.alphaWrite == AB_ALPHA_FILL) {
        if (J.alphaMul == AB_MUL_NONE) return launch_any<YT, RT, CHROMA, 4, AFIRST, AB_ALPHA_FILL, AB_MUL_NONE>(B, st);
        return 0;
    }
    if (J.alphaWrite != AB_ALPHA_COPY) return 0;
    if constexpr (sizeof(YT) == 2 && sizeof(RT) == 2 && !AFIRST) {
        if (J.alphaRescale) { /* 10 / 12-bit alpha plane into deeper 16-bit-container channels (or the reverse) */
            switch (J.alphaMul) {
                case AB_MUL_NONE: return launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_NONE>(B, st);
                case AB_MUL_INT_PREMUL:
                    if (slow) return 0;
                    if (J.pmR != 0.0f) return launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_INT_PREMUL_FMA>(B, st);
                    return launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_INT_PREMUL>(B, st);
                case AB_MUL_FLOAT_MUL: return !slow ? 0 : launch_any<YT, RT, CHROMA, 4, false, AW_COPY_RESC, AB_MUL_FLOAT_MUL>(B, st);
                default: return 0;
            }
        }
    }
    switch (J.alphaMul) {
        case AB_MUL_NONE: return launch_any<YT, RT, CHROMA, 4, AFIRST, AB_ALPHA_COPY, AB_MUL_NONE>(B, st);
        case AB_MUL_INT_PREMUL:
            if (slow || AFIRST) return 0;
            if (J.pmR != 0.0f) return launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_INT_PREMUL_FMA>(B, st);
            /* 8-bit channels always have verified constants: the general form is only instantiated for 16-bit ones */
            return (sizeof(RT) == 1) ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_INT_PREMUL>(B, st);
        case AB_MUL_INT_UNPREMUL: return (slow || AFIRST) ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_INT_UNPREMUL>(B, st);
        case AB_MUL_FLOAT_MUL: return (!slow || AFIRST) ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_FLOAT_MUL>(B, st);
        case AB_MUL_FLOAT_UNMUL: return (!slow || AFIRST) ? 0 : launch_any<YT, RT, CHROMA, 4, false, AB_ALPHA_COPY, AB_MUL_FLOAT_UNMUL>(B, st);
        default: return 0;
    }
}

/* 2: every plane row and every pixel row starts where the lanes' 16 / 8-byte vectors need it; 1: rows are only sample-aligned
 * (views into larger images, odd strides, tightly packed 257-pixel rows): the specialised kernels still apply, every tile through
 * their ragged code with element-wise loads and stores (AbY2RFrame::vecWidth == 0); 0: not even that */
template <typename YT, typename RT>
static int y2r_alignment(const AbY2RJob & J)
{
    const bool needA = (J.alphaWrite == AB_ALPHA_COPY) || (J.alphaMul != AB_MUL_NONE);
    const bool sub = (J.chroma != AB_CHROMA_FULL);
    uintptr_t ybits = (uintptr_t)J.y | (uintptr_t)J.yRB;
    if (needA) ybits |= (uintptr_t)J.a | (uintptr_t)J.aRB;
    const uintptr_t cbits = (uintptr_t)J.u | (uintptr_t)J.uRB | (uintptr_t)J.v | (uintptr_t)J.vRB;
    const uintptr_t rbits = (uintptr_t)J.rgb | (uintptr_t)J.rgbRB;
    const uintptr_t pixb = J.is565 ? 2 : (uintptr_t)J.nChan * sizeof(RT);
    const uintptr_t rgbGran = (pixb % 2 == 0) ? 2 : 1; /* the ragged store is 16-bit granular for even pixel sizes, else byte-wise */
    if (((ybits | cbits) & (sizeof(YT) - 1)) || (rbits & (rgbGran - 1))) return 0;
    /* a lane reads 8 luma / 4 subsampled chroma samples per vector and stores its 8 pixels as 16-byte vectors (3-byte pixels, 24
     * bytes per lane: 8-byte ones) */
    const uintptr_t lumaAlign = 8 * sizeof(YT) - 1, subAlign = 4 * sizeof(YT) - 1;
    const bool vec = !(ybits & lumaAlign) && !(cbits & (sub ? subAlign : lumaAlign)) && !(rbits & (((8 * pixb) % 16 == 0) ? 15 : 7));
    return vec ? 2 : 1;
}

/* can this job run on a specialised kernel? (constants + alignment of its own planes) */
template <typename YT, typename RT>
static bool fast_eligible(const AbY2RJob & J)
{
    const bool identity = (J.mode == AB_MODE_IDENTITY || J.mode == AB_MODE_IDENTITY8_SHUFFLE);
    if ((J.mode != AB_MODE_COEFF && !identity) || J.isGray || J.chroma == AB_CHROMA_NONE) return false;
    if (identity && (J.chroma != AB_CHROMA_FULL || J.is565 || J.toF16 || J.alphaMul != AB_MUL_NONE)) return false;
    if (J.is565 && sizeof(RT) != 1) return false;
    if (J.toF16 && (sizeof(RT) != 2 || J.alphaFirst)) return false;
    if (!J.is565 && J.nChan != 3 && J.nChan != 4) return false;
    if (J.hasAlphaChan && J.alphaWrite == AB_ALPHA_KEEP) return false;
    /* alpha depth rescale: only the 8 <-> 16-bit container pairs (TileTraits::RESC), whose depths always differ */
    if (J.alphaWrite == AB_ALPHA_COPY && sizeof(YT) != sizeof(RT) && !J.alphaRescale) return false; /* cannot happen: depths differ */
    if (J.alphaWrite == AB_ALPHA_COPY && J.alphaRescale && sizeof(YT) == sizeof(RT) && (J.alphaFirst || sizeof(YT) == 1)) return false;
    if (!J.rangeY.intExact || !J.rangeUV.intExact) return false;
    if (!identity && (!J.kg.exact || (AB_Y2R_DIV2 && !J.kg.exact2))) return false;
    if ((J.alphaMul == AB_MUL_INT_PREMUL) && J.pmR == 0.0f && (!J.rgbMaxDiv.exact || (AB_Y2R_DIV2 && !J.rgbMaxDiv.exact2))) return false;
    if ((J.alphaMul == AB_MUL_FLOAT_MUL || J.alphaMul == AB_MUL_FLOAT_UNMUL) && !J.yuvMaxDiv.intExact) return false;
    if (J.alphaWrite == AB_ALPHA_COPY && J.alphaRescale && !J.yuvMaxDiv.intExact) return false;
    const bool needA = (J.alphaWrite == AB_ALPHA_COPY) || (J.alphaMul != AB_MUL_NONE);
    if (needA && !J.a) return false;
    if (y2r_alignment<YT, RT>(J) == 0) return false; /* not even sample-aligned: the byte-addressed kernel */
    if (J.width == 0 || J.rowEnd <= J.rowBegin) return false;
    return true;
}

/* every conversion constant equal (the per-frame members -- pointers, strides, extents -- lead the struct) */
static inline bool same_constants(const AbY2RJob & a, const AbY2RJob & b)
{
    const size_t from = offsetof(AbY2RJob, yuvMax), to = offsetof(AbY2RJob, pairAligned);
    return memcmp(reinterpret_cast<const char *>(&a) + from, reinterpret_cast<const char *>(&b) + from, to - from) == 0;
}

/* Launches jobs[0..n) (n <= AB_MAX_BATCH) in ONE specialised kernel if all are eligible and share their constants.
 * Returns 1 if launched, 0 if the caller must take the general path. */
template <typename YT, typename RT>
static int launch_y2r_fast_typed(const AbY2RJob * jobs, uint32_t n, cudaStream_t st)
{
    if (n == 0 || n > AB_MAX_BATCH) return 0;
    for (uint32_t k = 0; k < n; ++k) {
        if (!fast_eligible<YT, RT>(jobs[k])) return 0;
        if (k && !same_constants(jobs[0], jobs[k])) return 0;
    }
    AbY2RBatch B;
    B.common = jobs[0];
    B.nFrames = n;
    AbY2RJob & J = B.common;
    /* B-first layouts: the matrix is symmetric in (R,Cr,c1,c3) <-> (B,Cb,c2,c4) and float addition commutes, so BGR order
     * is the RGB kernel with the chroma planes and their constants exchanged. */
    const bool swap = J.swapRB != 0;
    if (swap) {
        float tf = J.c1; J.c1 = J.c2; J.c2 = tf;
        tf = J.c3; J.c3 = J.c4; J.c4 = tf;
        J.swapRB = 0;
    }
    for (uint32_t k = 0; k < n; ++k) {
        AbY2RFrame & F = B.frame[k];
        const AbY2RJob & S = jobs[k];
        F.y = S.y; F.a = S.a; F.rgb = S.rgb;
        F.u = swap ? S.v : S.u;
        F.v = swap ? S.u : S.v;
        F.yRB = S.yRB; F.aRB = S.aRB; F.rgbRB = S.rgbRB;
        F.uRB = swap ? S.vRB : S.uRB;
        F.vRB = swap ? S.uRB : S.vRB;
        F.width = S.width; F.height = S.height; F.rowBegin = S.rowBegin; F.rowEnd = S.rowEnd;
        F.tilesX = F.tiles = F.groups = F.blockStart = F.blocks = 0;
        F.vecWidth = (y2r_alignment<YT, RT>(S) == 2) ? S.width : 0;
    }
    const bool afirst = J.alphaFirst != 0;
    const int nch = J.nChan;
    const bool special = J.is565 || J.toF16;
    if (J.mode == AB_MODE_IDENTITY || J.mode == AB_MODE_IDENTITY8_SHUFFLE) return launch_identity<YT, RT>(B, st);
#define ABF_DISPATCH(CH)                                                              \
    do {                                                                              \
        if (special) return launch_special<YT, RT, CH>(B, st);                        \
        if (nch == 3) return launch_alpha<YT, RT, CH, 3, false>(B, st);               \
        return afirst ? launch_alpha<YT, RT, CH, 4, true>(B, st) : launch_alpha<YT, RT, CH, 4, false>(B, st); \
    } while (0)
    if (J.chroma == AB_CHROMA_FULL) ABF_DISPATCH(CH_FULL);
    if (J.chroma == AB_CHROMA_NEAREST) ABF_DISPATCH(CH_NEAREST);
    if (J.chroma == AB_CHROMA_BILINEAR) {
        if (J.vertBlend) ABF_DISPATCH(CH_BIL420);
        ABF_DISPATCH(CH_BIL422);
    }
#undef ABF_DISPATCH
    return 0;
}

} // namespace abfast
