/*
 * tonemap.cu -- gain-map tone mapping of an RGB image, one thread per pixel.
 *
 * Mirrors the per-pixel body of avifRGBImageApplyGainMap (src/gainmap.c:140-168 without a gain, :229-296 with one):
 *   read RGBA (src/reformat.c:1856-1897) -> gamma-to-linear (src/colr.c:216-467) -> optional primaries change in double
 *   (src/colrconvert.c:178-186) -> per channel  (base + baseOffset) * exp2f(lerp(min, max, powf(g, 1/gamma)) * weight) - altOffset
 *   -> optional primaries change -> linear-to-gamma -> clamp [0,1] -> write (src/reformat.c:1899-1939),
 *   plus the two reductions behind the content light level (src/gainmap.c:265-270, :293, :297-306).
 * Every C operator of the reference is one explicitly rounded operation here (-fmad=false); the transcendental calls are
 * CUDA's powf / exp2f / expf / logf / log10f / sqrtf where the reference calls libm's -- equal to about an ulp, which is why this
 * row is compared to +-1 output code instead of bit for bit.
 * Compute-bound by design (up to 9 powf + 3 exp2f per pixel): the bound that matters is the SFU/FMA issue rate, not HBM.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <mutex>

#include "plan.h"

namespace {

__device__ __forceinline__ float clamp01(float v) { return (v < 0.0f) ? 0.0f : ((1.0f < v) ? 1.0f : v); } /* AVIF_CLAMP */

/* The reference calls the host libm (glibc: powf / exp2f / expf / logf / log10f, each correctly rounded in all but rare cases);
 * CUDA's single-precision versions are good to 1-2 ulp, which moves an 8..12-bit output code by at most one but a 16-bit code
 * by several (PQ raises to the 78.84th power).  P = true evaluates the function in double and rounds once: the correctly rounded
 * float except within ~2^-29 of a rounding boundary, i.e. the value glibc returns.  Used for the per-code tables (65536 entries at
 * most: free) and, when the output has 16-bit channels, for the per-pixel output transfer function. */
template <bool P> __device__ __forceinline__ float xpow(float a, float b) { return P ? (float)pow((double)a, (double)b) : powf(a, b); }
template <bool P> __device__ __forceinline__ float xexp2(float a) { return P ? (float)exp2((double)a) : exp2f(a); }
template <bool P> __device__ __forceinline__ float xexp(float a) { return P ? (float)exp((double)a) : expf(a); }
template <bool P> __device__ __forceinline__ float xlog(float a) { return P ? (float)log((double)a) : logf(a); }
template <bool P> __device__ __forceinline__ float xlog10(float a) { return P ? (float)log10((double)a) : log10f(a); }

/* src/colr.c:216-467; id = avifTransferCharacteristics; unknown ids take BT.709 (:497, :508) */
template <bool P>
__device__ float to_linear(int id, float g)
{
    switch (id) {
        case 4: return xpow<P>(clamp01(g), 2.2f);
        case 5: return xpow<P>(clamp01(g), 2.8f);
        case 7:
            if (g < 0.0f) return 0.0f;
            if (g < __fmul_rn(4.0f, 0.022821585529445f)) return __fdiv_rn(g, 4.0f);
            if (g < 1.0f) return xpow<P>(__fdiv_rn(__fadd_rn(g, 0.111572195921731f), 1.111572195921731f), __fdiv_rn(1.0f, 0.45f));
            return 1.0f;
        case 8: return clamp01(g);
        case 9: return (g <= 0.0f) ? __fdiv_rn(0.01f, 2.0f) : xpow<P>(10.0f, __fmul_rn(2.0f, __fsub_rn(fminf(g, 1.0f), 1.0f)));
        case 10: return (g <= 0.0f) ? __fdiv_rn(0.00316227766f, 2.0f) : xpow<P>(10.0f, __fmul_rn(2.5f, __fsub_rn(fminf(g, 1.0f), 1.0f)));
        case 11:
            if (g < __fmul_rn(-4.5f, 0.018053968510807f)) return -xpow<P>(__fdiv_rn(__fsub_rn(g, 0.09929682680944f), -1.09929682680944f), __fdiv_rn(1.0f, 0.45f));
            if (g < __fmul_rn(4.5f, 0.018053968510807f)) return __fdiv_rn(g, 4.5f);
            return xpow<P>(__fdiv_rn(__fadd_rn(g, 0.09929682680944f), 1.09929682680944f), __fdiv_rn(1.0f, 0.45f));
        case 12:
            if (g < -0.25f) return -0.25f;
            if (g < 0.0f) return __fdiv_rn(xpow<P>(__fdiv_rn(__fsub_rn(g, 0.02482420670236f), -0.27482420670236f), __fdiv_rn(1.0f, 0.45f)), -4.0f);
            if (g < __fmul_rn(4.5f, 0.018053968510807f)) return __fdiv_rn(g, 4.5f);
            if (g < 1.0f) return xpow<P>(__fdiv_rn(__fadd_rn(g, 0.09929682680944f), 1.09929682680944f), __fdiv_rn(1.0f, 0.45f));
            return 1.0f;
        case 13:
            if (g < 0.0f) return 0.0f;
            if (g < __fmul_rn(12.92f, 0.0030412825601275209f)) return __fdiv_rn(g, 12.92f);
            if (g < 1.0f) return xpow<P>(__fdiv_rn(__fadd_rn(g, 0.0550107189475866f), 1.0550107189475866f), 2.4f);
            return 1.0f;
        case 16: {
            if (!(g > 0.0f)) return 0.0f;
            const float pg = xpow<P>(g, __fdiv_rn(1.0f, 78.84375f));
            const float num = fmaxf(__fsub_rn(pg, 0.8359375f), 0.0f);
            const float den = fmaxf(__fsub_rn(18.8515625f, __fmul_rn(18.6875f, pg)), 1.17549435e-38f);
            const float lin = xpow<P>(__fdiv_rn(num, den), __fdiv_rn(1.0f, 0.1593017578125f));
            return __fdiv_rn(__fmul_rn(lin, 10000.0f), 203.0f);
        }
        case 17: return __fdiv_rn(xpow<P>(fmaxf(g, 0.0f), 2.6f), 0.91655527974030934f);
        case 18: {
            if (g < 0.0f) return 0.0f;
            float lin;
            if (g <= 0.5f) lin = xpow<P>(__fmul_rn(__fmul_rn(g, g), __fdiv_rn(1.0f, 3.0f)), 1.2f);
            else lin = xpow<P>(__fdiv_rn(__fadd_rn(xexp<P>(__fdiv_rn(__fsub_rn(g, 0.55991073f), 0.17883277f)), 0.28466892f), 12.0f), 1.2f);
            return __fdiv_rn(__fmul_rn(lin, 1000.0f), 203.0f);
        }
        default: /* 1, 6, 14, 15 and everything unknown: BT.709 */
            if (g < 0.0f) return 0.0f;
            if (g < __fmul_rn(4.5f, 0.018053968510807f)) return __fdiv_rn(g, 4.5f);
            if (g < 1.0f) return xpow<P>(__fdiv_rn(__fadd_rn(g, 0.09929682680944f), 1.09929682680944f), __fdiv_rn(1.0f, 0.45f));
            return 1.0f;
    }
}

template <bool P>
__device__ float to_gamma(int id, float l)
{
    switch (id) {
        case 4: return xpow<P>(clamp01(l), __fdiv_rn(1.0f, 2.2f));
        case 5: return xpow<P>(clamp01(l), __fdiv_rn(1.0f, 2.8f));
        case 7:
            if (l < 0.0f) return 0.0f;
            if (l < 0.022821585529445f) return __fmul_rn(l, 4.0f);
            if (l < 1.0f) return __fsub_rn(__fmul_rn(1.111572195921731f, xpow<P>(l, 0.45f)), 0.111572195921731f);
            return 1.0f;
        case 8: return clamp01(l);
        case 9: return (l <= 0.01f) ? 0.0f : __fadd_rn(1.0f, __fdiv_rn(xlog10<P>(fminf(l, 1.0f)), 2.0f));
        case 10: return (l <= 0.00316227766f) ? 0.0f : __fadd_rn(1.0f, __fdiv_rn(xlog10<P>(fminf(l, 1.0f)), 2.5f));
        case 11:
            if (l < -0.018053968510807f) return __fadd_rn(__fmul_rn(-1.09929682680944f, xpow<P>(-l, 0.45f)), 0.09929682680944f);
            if (l < 0.018053968510807f) return __fmul_rn(l, 4.5f);
            return __fsub_rn(__fmul_rn(1.09929682680944f, xpow<P>(l, 0.45f)), 0.09929682680944f);
        case 12:
            if (l < -0.25f) return -0.25f;
            if (l < 0.0f) return __fadd_rn(__fmul_rn(-0.27482420670236f, xpow<P>(__fmul_rn(-4.0f, l), 0.45f)), 0.02482420670236f);
            if (l < 0.018053968510807f) return __fmul_rn(l, 4.5f);
            if (l < 1.0f) return __fsub_rn(__fmul_rn(1.09929682680944f, xpow<P>(l, 0.45f)), 0.09929682680944f);
            return 1.0f;
        case 13:
            if (l < 0.0f) return 0.0f;
            if (l < 0.0030412825601275209f) return __fmul_rn(l, 12.92f);
            if (l < 1.0f) return __fsub_rn(__fmul_rn(1.0550107189475866f, xpow<P>(l, __fdiv_rn(1.0f, 2.4f))), 0.0550107189475866f);
            return 1.0f;
        case 16: {
            if (!(l > 0.0f)) return 0.0f;
            const float s = clamp01(__fdiv_rn(__fmul_rn(l, 203.0f), 10000.0f));
            const float pl = xpow<P>(s, 0.1593017578125f);
            const float num = __fsub_rn(__fmul_rn(0.1640625f, pl), 0.1640625f);
            const float den = __fadd_rn(1.0f, __fmul_rn(18.6875f, pl));
            return xpow<P>(__fadd_rn(1.0f, __fdiv_rn(num, den)), 78.84375f);
        }
        case 17: return xpow<P>(__fmul_rn(0.91655527974030934f, fmaxf(l, 0.0f)), __fdiv_rn(1.0f, 2.6f));
        case 18: {
            float s = clamp01(__fdiv_rn(__fmul_rn(l, 203.0f), 1000.0f));
            s = xpow<P>(s, __fdiv_rn(1.0f, 1.2f));
            if (s < 0.0f) return 0.0f;
            if (s <= __fdiv_rn(1.0f, 12.0f)) return sqrtf(__fmul_rn(3.0f, s));
            return __fadd_rn(__fmul_rn(0.17883277f, xlog<P>(__fsub_rn(__fmul_rn(12.0f, s), 0.28466892f))), 0.55991073f);
        }
        default:
            if (l < 0.0f) return 0.0f;
            if (l < 0.018053968510807f) return __fmul_rn(l, 4.5f);
            if (l < 1.0f) return __fsub_rn(__fmul_rn(1.09929682680944f, xpow<P>(l, 0.45f)), 0.09929682680944f);
            return 1.0f;
    }
}

#define AB_F16_MULTIPLIER 1.9259299444e-34f /* src/reformat.c:1411 */

/* src/reformat.c:1856-1897 split in two: the stored codes of a pixel (alpha = the maximum code for formats without alpha)
 * and the code -> float step.  Keeping the codes lets the kernel index per-call tables of every function that depends on a
 * code alone (see tone_lut_kernel). */
__device__ __forceinline__ void load_codes(const uint8_t * row, uint32_t x, const AbPixFmt & f, uint32_t c[4])
{
    const uint8_t * p = row + (size_t)x * f.pixBytes;
    if (f.chanBytes > 1) {
        c[0] = *reinterpret_cast<const uint16_t *>(p + f.offR);
        c[1] = *reinterpret_cast<const uint16_t *>(p + f.offG);
        c[2] = *reinterpret_cast<const uint16_t *>(p + f.offB);
        c[3] = f.hasAlpha ? (uint32_t)*reinterpret_cast<const uint16_t *>(p + f.offA) : 0xFFFFFFFFu;
    } else if (f.is565) {
        const uint32_t v = *reinterpret_cast<const uint16_t *>(p + f.offR); /* :637-648 */
        const uint32_t r5 = (v & 0xF800u) >> 11, g6 = (v & 0x07E0u) >> 5, b5 = v & 0x001Fu;
        c[0] = ((r5 << 3) | (r5 >> 2)) & 0xFFu;
        c[1] = ((g6 << 2) | (g6 >> 4)) & 0xFFu;
        c[2] = ((b5 << 3) | (b5 >> 2)) & 0xFFu;
        c[3] = 0xFFFFFFFFu;
    } else {
        c[0] = p[f.offR];
        c[1] = p[f.offG];
        c[2] = p[f.offB];
        c[3] = f.hasAlpha ? (uint32_t)p[f.offA] : 0xFFFFFFFFu;
    }
}
/* code -> the float avifGetRGBAPixel hands out: code / maxF, or the half float's value (:1847-1853) */
__device__ __forceinline__ float code_value(uint32_t code, const AbPixFmt & f)
{
    if (f.isFloat) return __fdiv_rn(__uint_as_float(code << 13), AB_F16_MULTIPLIER);
    return __fdiv_rn((float)code, f.maxF);
}
__device__ __forceinline__ float alpha_value(uint32_t code, const AbPixFmt & f)
{
    if (code == 0xFFFFFFFFu) return 1.0f; /* no alpha channel: max / max, or the literal 1.0f of the 8-bit branches */
    return code_value(code, f);
}

/* src/reformat.c:1899-1939; the float -> integer casts truncate.  Split in two -- the codes of a pixel, then the store -- so that
 * the colour codes can also come from the threshold table (see tone_thresh_kernel). */
__device__ __forceinline__ uint32_t quant_code(float v, float maxF) { return __float2uint_rz(__fadd_rn(0.5f, __fmul_rn(v, maxF))); }

__device__ __forceinline__ void store_codes(uint8_t * row, uint32_t x, const AbPixFmt & f, const uint32_t q[4])
{
    uint8_t * p = row + (size_t)x * f.pixBytes;
    if (f.chanBytes > 1) {
        *reinterpret_cast<uint16_t *>(p + f.offR) = (uint16_t)q[0];
        *reinterpret_cast<uint16_t *>(p + f.offG) = (uint16_t)q[1];
        *reinterpret_cast<uint16_t *>(p + f.offB) = (uint16_t)q[2];
        if (f.hasAlpha) *reinterpret_cast<uint16_t *>(p + f.offA) = (uint16_t)q[3];
    } else {
        const uint32_t R = q[0] & 0xFFu, G = q[1] & 0xFFu, B = q[2] & 0xFFu;
        if (f.is565) {
            *reinterpret_cast<uint16_t *>(p + f.offR) = (uint16_t)((B >> 3) | ((G >> 2) << 5) | ((R >> 3) << 11)); /* :619 */
            return; /* 565 has no alpha */
        }
        p[f.offR] = (uint8_t)R;
        p[f.offG] = (uint8_t)G;
        p[f.offB] = (uint8_t)B;
        if (f.hasAlpha) p[f.offA] = (uint8_t)(q[3] & 0xFFu);
    }
}

__device__ __forceinline__ void set_rgba(uint8_t * row, uint32_t x, const AbPixFmt & f, const float px[4])
{
    if (f.chanBytes > 1 && f.isFloat) {
        uint8_t * p = row + (size_t)x * f.pixBytes;
        *reinterpret_cast<uint16_t *>(p + f.offR) = (uint16_t)(__float_as_uint(__fmul_rn(px[0], AB_F16_MULTIPLIER)) >> 13);
        *reinterpret_cast<uint16_t *>(p + f.offG) = (uint16_t)(__float_as_uint(__fmul_rn(px[1], AB_F16_MULTIPLIER)) >> 13);
        *reinterpret_cast<uint16_t *>(p + f.offB) = (uint16_t)(__float_as_uint(__fmul_rn(px[2], AB_F16_MULTIPLIER)) >> 13);
        if (f.hasAlpha) *reinterpret_cast<uint16_t *>(p + f.offA) = (uint16_t)(__float_as_uint(__fmul_rn(px[3], AB_F16_MULTIPLIER)) >> 13);
        return;
    }
    const uint32_t q[4] = { quant_code(px[0], f.maxF), quant_code(px[1], f.maxF), quant_code(px[2], f.maxF), f.hasAlpha ? quant_code(px[3], f.maxF) : 0u };
    store_codes(row, x, f, q);
}

/* ---- output transfer function as a step function -----------------------------------------------------------------------
 * For integer outputs the stored code is a monotone step function of the linear value:  code(x) = (T)(0.5 + clamp(g(x)) * max)
 * never decreases and is 0 for x <= 0 (all transfer functions but BT.1361 extended: kAbNotStepFunctions below).  So code(x) =
 * number of thresholds t_k <= x, t_k = the smallest float whose
 * code is >= k (k = 1..max).  The thresholds are found once per (device, transfer function, max) by bisection over the float bit
 * patterns with g evaluated in double and rounded once (what glibc's correctly rounded powf / logf return), kept for the life of
 * the process, copied into shared memory by every block and searched in log2(max + 1) steps: 8 / 10 / 12 shared loads and compares
 * per channel instead of up to two powf.  Used for channel depths <= 12. */
__device__ __forceinline__ uint32_t exact_code(int tf, float x, float maxF)
{
    const float g = fminf(1.0f, fmaxf(0.0f, to_gamma<true>(tf, x)));
    return quant_code(g, maxF);
}

__global__ void __launch_bounds__(128) tone_thresh_kernel(float * T, uint32_t n, int tf, float maxF)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x + 1u; /* the code this threshold opens */
    if (k > n) return;
    float t;
    if (exact_code(tf, 0.0f, maxF) >= k) {
        t = -__int_as_float(0x7F800000); /* (no transfer function of the table does this) */
    } else if (exact_code(tf, __int_as_float(0x7F800000), maxF) < k) {
        t = __int_as_float(0x7FC00000); /* never reached: compares false against everything */
    } else {
        uint32_t lo = 0u, hi = 0x7F800000u; /* positive floats order like their bit patterns; code(lo) < k <= code(hi) */
        while (hi - lo > 1u) {
            const uint32_t mid = lo + ((hi - lo) >> 1);
            if (exact_code(tf, __uint_as_float(mid), maxF) >= k) hi = mid;
            else lo = mid;
        }
        t = __uint_as_float(hi);
    }
    T[k - 1u] = t;
}

/* number of thresholds <= x among 2^STEPS - 1 sorted ones.  The block keeps them in LEVEL ORDER (Eytzinger: node i has children 2i and
 * 2i + 1, level k occupies words 2^k .. 2^(k+1) - 1): the lanes of a warp then read neighbouring words at every level.  In sorted
 * order the candidates of a level all sit a multiple of 32 words apart -- one shared-memory bank -- and the middle levels of the
 * search serialised up to 32-fold (ncu r02z: short_scoreboard 11 warps per issue, issue utilisation 46 %).  STEPS loads at a running
 * index, fully unrolled; the leaf index reached is the rank. */
template <int STEPS>
__device__ __forceinline__ uint32_t thresh_code(const float * sE, float x)
{
    uint32_t i = 1;
#pragma unroll
    for (int k = 0; k < STEPS; ++k) i = 2u * i + ((sE[i] <= x) ? 1u : 0u);
    return i - (1u << STEPS);
}
/* sorted index of level-order node i (1-based) of a complete tree with `steps` levels */
__device__ __forceinline__ uint32_t eytzinger_source(uint32_t i, int steps)
{
    const int k = 31 - __clz((int)i);       /* depth of the node */
    const uint32_t j = i - (1u << k);        /* its index within the level */
    return ((2u * j + 1u) << (steps - 1 - k)) - 1u;
}


/* src/colrconvert.c:70-75, :178-186: y = M.x in double, left-to-right sums, back to float */
__device__ __forceinline__ void convert_primaries(float px[4], const double * M)
{
    const double x0 = (double)px[0], x1 = (double)px[1], x2 = (double)px[2];
    px[0] = (float)__dadd_rn(__dadd_rn(__dmul_rn(M[0], x0), __dmul_rn(M[1], x1)), __dmul_rn(M[2], x2));
    px[1] = (float)__dadd_rn(__dadd_rn(__dmul_rn(M[3], x0), __dmul_rn(M[4], x1)), __dmul_rn(M[5], x2));
    px[2] = (float)__dadd_rn(__dadd_rn(__dmul_rn(M[6], x0), __dmul_rn(M[7], x1)), __dmul_rn(M[8], x2));
}

/* Everything the per-pixel work computes from ONE stored code -- to_linear(code/max) of a base channel, and for a gain-map
 * channel the whole factor exp2f(lerp(min, max, powf(code/max, 1/gamma)) * weight) -- is evaluated once per possible code
 * (256 or 65536 of them) by the same device functions, then looked up: 9 of the 12 transcendental calls per pixel disappear
 * and not one result bit changes.  Half-float bases are not tabulated. */
__global__ void __launch_bounds__(256) tone_lut_kernel(const __grid_constant__ AbToneMapJob J)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < J.baseLutN) {
        J.baseLut[i] = to_linear<true>(J.tfIn, code_value(i, J.baseFmt));
        return;
    }
    const uint32_t k = i - J.baseLutN;
    if (k < 3u * J.gainLutN) {
        const uint32_t c = k / J.gainLutN, code = k - c * J.gainLutN;
        const float w = xpow<true>(code_value(code, J.gainFmt), J.gammaInv[c]);
        const float log2g = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, w), J.gmMin[c]), __fmul_rn(w, J.gmMax[c])); /* lerp, src/gainmap.c:66-69 */
        J.gainLut[k] = xexp2<true>(__fmul_rn(log2g, J.weight));
    }
}

/* STEPS > 0: integer outputs of STEPS-bit channels through the threshold table (2^STEPS - 1 entries) */
template <bool P, int STEPS>
__global__ void __launch_bounds__(256) tone_map_kernel(const __grid_constant__ AbToneMapJob J, const float * __restrict__ thresh, uint32_t threshN)
{
    constexpr bool TAB = (STEPS > 0);
    extern __shared__ float sThresh[];
    if (TAB) { /* word 0 is unused; words 1 .. threshN hold the thresholds in level order */
        for (uint32_t i = 1u + threadIdx.x; i <= threshN; i += blockDim.x) sThresh[i] = thresh[eytzinger_source(i, STEPS)];
        __syncthreads();
    }
    const uint64_t total = (uint64_t)J.width * J.height;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    float localMax = 0.0f;  /* src/gainmap.c:213: starts at 0, only larger values replace it */
    double localSum = 0.0;
    uint32_t localNan = 0xFFFFFFFFu;
    const bool baseTab = (J.baseLutN != 0);
    /* (x, y) advance by the grid stride incrementally: a 64-bit division per pixel costs more than a transfer function */
    const uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t stepY = (uint32_t)(stride / J.width), stepX = (uint32_t)(stride - (uint64_t)stepY * J.width);
    uint32_t y = (uint32_t)(i0 / J.width), x = (uint32_t)(i0 - (uint64_t)y * J.width);
    for (uint64_t i = i0; i < total; i += stride, x += stepX, y += stepY) {
        if (x >= J.width) { /* at most one wrap: stepX < width */
            x -= J.width;
            ++y;
        }
        uint32_t bc[4];
        load_codes(J.base + (int64_t)y * J.baseRB, x, J.baseFmt, bc);
        float base[4], out[4];
        bool stored = false; /* TAB: the pixel has been written from the threshold table */
        base[3] = alpha_value(bc[3], J.baseFmt);
        if (!J.applyGain) { /* src/gainmap.c:152-167 */
            if (J.transcode) {
#pragma unroll
                for (int c = 0; c < 3; ++c) base[c] = baseTab ? J.baseLut[bc[c]] : to_linear<P>(J.tfIn, code_value(bc[c], J.baseFmt));
                if (J.inConv) convert_primaries(base, J.inM);
                if (TAB) { /* the codes straight from the step function */
                    const uint32_t q[4] = { thresh_code<STEPS>(sThresh, base[0]), thresh_code<STEPS>(sThresh, base[1]), thresh_code<STEPS>(sThresh, base[2]),
                                            J.outFmt.hasAlpha ? quant_code(base[3], J.outFmt.maxF) : 0u };
                    store_codes(J.out + (int64_t)y * J.outRB, x, J.outFmt, q);
                    stored = true;
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) base[c] = fminf(1.0f, fmaxf(0.0f, to_gamma<P>(J.tfOut, base[c])));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) base[c] = code_value(bc[c], J.baseFmt);
            }
            out[0] = base[0]; out[1] = base[1]; out[2] = base[2]; out[3] = base[3];
        } else { /* src/gainmap.c:231-294 */
            uint32_t gc[4];
            load_codes(J.gain + (int64_t)y * J.gainRB, x, J.gainFmt, gc);
#pragma unroll
            for (int c = 0; c < 3; ++c) base[c] = baseTab ? J.baseLut[bc[c]] : to_linear<P>(J.tfIn, code_value(bc[c], J.baseFmt));
            if (J.inConv) convert_primaries(base, J.inM);
            float pixelMax = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float factor = J.gainLut[(uint32_t)c * J.gainLutN + gc[c]];
                const float t = __fsub_rn(__fmul_rn(__fadd_rn(base[c], J.baseOff[c]), factor), J.altOff[c]);
                if (t > localMax) localMax = t;
                if (t > pixelMax) pixelMax = t;
                out[c] = t;
            }
            if (J.outConv) convert_primaries(out, J.outM);
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 3; ++c) bad = bad || isnan(out[c]);
            if (bad && (uint32_t)i < localNan) localNan = (uint32_t)i;
            out[3] = base[3];
            localSum += (double)pixelMax;
            if (TAB) {
                const uint32_t q[4] = { thresh_code<STEPS>(sThresh, out[0]), thresh_code<STEPS>(sThresh, out[1]), thresh_code<STEPS>(sThresh, out[2]),
                                        J.outFmt.hasAlpha ? quant_code(out[3], J.outFmt.maxF) : 0u };
                store_codes(J.out + (int64_t)y * J.outRB, x, J.outFmt, q);
                stored = true;
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c) out[c] = fminf(1.0f, fmaxf(0.0f, to_gamma<P>(J.tfOut, out[c])));
            }
        }
        if (!stored) set_rgba(J.out + (int64_t)y * J.outRB, x, J.outFmt, out);
    }
    if (!J.applyGain) return;
    /* block reduction, then one atomic each */
    __shared__ float sMax[8];
    __shared__ double sSum[8];
    __shared__ uint32_t sNan[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        localMax = fmaxf(localMax, __shfl_xor_sync(0xFFFFFFFFu, localMax, o));
        localSum += __shfl_xor_sync(0xFFFFFFFFu, localSum, o);
        localNan = min(localNan, __shfl_xor_sync(0xFFFFFFFFu, localNan, o));
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sMax[warp] = localMax;
        sSum[warp] = localSum;
        sNan[warp] = localNan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.0f;
        double s = 0.0;
        uint32_t n = 0xFFFFFFFFu;
        for (uint32_t k = 0; k < (blockDim.x >> 5); ++k) {
            m = fmaxf(m, sMax[k]);
            s += sSum[k];
            n = min(n, sNan[k]);
        }
        atomicMax(J.maxBits, __float_as_uint(m)); /* m >= 0: unsigned order of the bits == float order */
        atomicAdd(J.sum, s);
        atomicMin(J.nanIndex, n);
    }
}

} // namespace

/* Transfer functions whose stored code is NOT a non-decreasing function of the linear value with code(x <= 0) == 0 -- the premise of
 * the threshold table.  BT.1361 extended (12): its negative branch -0.2748 (-4 l)^0.45 + 0.0248 is POSITIVE for -0.0012 < l < 0 (up to
 * code 6 / 25 / 102 at 8 / 10 / 12 bits) and drops back to 0 at l = 0.  tests/test_proofs.py scans the compiled reference's thirteen
 * functions and requires this list to be exactly the set that fails the premise; those keep the per-pixel evaluation. */
static const int kAbNotStepFunctions[] = { 12 };

/* AVIF_B200_TONEMAP_TABLE=0 evaluates the output transfer function per pixel as before (A/B timing, tests of both forms) */
static bool abToneMapNoThresholds()
{
    static int off = -1;
    if (off < 0) {
        const char * e = getenv("AVIF_B200_TONEMAP_TABLE");
        off = (e && e[0] == '0') ? 1 : 0;
    }
    return off == 1;
}

/* thresholds of (device, transfer function, channel maximum): built on first use, immutable afterwards, never freed */
struct ThreshEntry {
    int device, tf;
    uint32_t n;
    float * d;
};
static ThreshEntry gThresh[64];
static int gThreshCount = 0;
static std::mutex gThreshLock;

static const float * thresholds_for(int tf, uint32_t n, float maxF, cudaStream_t st)
{
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> hold(gThreshLock);
    for (int k = 0; k < gThreshCount; ++k)
        if (gThresh[k].device == device && gThresh[k].tf == tf && gThresh[k].n == n) return gThresh[k].d;
    if (gThreshCount >= 64) return nullptr;
    float * d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)n * sizeof(float)) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    tone_thresh_kernel<<<(n + 127u) / 128u, 128, 0, st>>>(d, n, tf, maxF);
    /* other threads (other streams) may pick the entry up as soon as it is published: finish it first, once */
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        (void)cudaGetLastError();
        (void)cudaFree(d);
        return nullptr;
    }
    gThresh[gThreshCount++] = ThreshEntry{ device, tf, n, d };
    return d;
}

extern "C" int abLaunchToneMap(const AbToneMapJob * job, void * stream)
{
    const uint64_t total = (uint64_t)job->width * job->height;
    if (total == 0) return 0;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148u * 16u) blocks = 148u * 16u; /* grid-stride: 8 resident blocks per SM, two waves */
    const uint32_t lutN = job->baseLutN + 3u * job->gainLutN;
    if (lutN) tone_lut_kernel<<<(lutN + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*job);
    /* 16-bit integer channels resolve the last ulp of the output transfer function: evaluate it in double there */
    const bool precise = (job->outFmt.chanBytes > 1 && !job->outFmt.isFloat && job->outFmt.maxF > 4095.5f);
    /* integer channels of at most 12 bits: the output transfer function as a searched step function (when a transfer function runs at all) */
    const float * thresh = nullptr;
    uint32_t threshN = 0;
    bool stepFunction = true;
    for (size_t k = 0; k < sizeof(kAbNotStepFunctions) / sizeof(kAbNotStepFunctions[0]); ++k)
        if (job->tfOut == kAbNotStepFunctions[k]) stepFunction = false;
    if (stepFunction && !job->outFmt.isFloat && job->outFmt.maxF < 4095.5f && (job->applyGain || job->transcode) && !abToneMapNoThresholds()) {
        threshN = (uint32_t)job->outFmt.maxF;
        if (((threshN + 1u) & threshN) == 0u) thresh = thresholds_for(job->tfOut, threshN, job->outFmt.maxF, (cudaStream_t)stream);
    }
    if (thresh && threshN == 255u) tone_map_kernel<false, 8><<<(unsigned)blocks, 256, (threshN + 1u) * sizeof(float), (cudaStream_t)stream>>>(*job, thresh, threshN);
    else if (thresh && threshN == 1023u) tone_map_kernel<false, 10><<<(unsigned)blocks, 256, (threshN + 1u) * sizeof(float), (cudaStream_t)stream>>>(*job, thresh, threshN);
    else if (thresh && threshN == 4095u) tone_map_kernel<false, 12><<<(unsigned)blocks, 256, (threshN + 1u) * sizeof(float), (cudaStream_t)stream>>>(*job, thresh, threshN);
    else if (precise) tone_map_kernel<true, 0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*job, nullptr, 0u);
    else tone_map_kernel<false, 0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*job, nullptr, 0u);
    return (int)cudaGetLastError();
}
