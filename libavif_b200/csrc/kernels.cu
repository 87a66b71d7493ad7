/*
 * kernels.cu -- hand-written sm_100a kernels of the reformat + alpha path and their launchers.
 *
 * Every kernel here is a per-pixel affine map: HBM-bound byte/short traffic, no tensor cores.  Design rules:
 *   - one fused pass per call: alpha channel, colour conversion, (un)premultiply, F16 and packing all happen
 *     in registers between one read of the source planes and one write of the destination (the reference
 *     makes up to four passes over the RGB buffer, src/reformat.c:1445-1593);
 *   - planar inputs are read with lane-contiguous (coalesced) accesses, all loads of a warp tile are issued
 *     before any arithmetic so that >= 48 B per thread are in flight;
 *   - interleaved outputs (1,2,3,4,6,8 bytes per pixel) are packed into a per-warp shared-memory tile and
 *     leave the SM as 16-byte, fully coalesced vector stores regardless of the pixel format;
 *   - arbitrary pointer alignment / row strides (view rectangles) are handled by the same kernels through a
 *     narrower copy-out granularity; a byte-addressed one-thread-per-pixel kernel exists as an independent
 *     second implementation for cross-checking (avifB200SetKernelPolicy(1)).
 *
 * Arithmetic lives in pixel_math.cuh and mirrors the reference bit for bit.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "pixel_math.cuh"
#include "plan.h"

namespace ab {

static thread_local uint64_t g_launches = 0;

/* ---------------------------------------------------------------------------------------------- */
/* sample access                                                                                   */

template <typename T>
__device__ __forceinline__ uint32_t ld_sample(const uint8_t * row, uint32_t x)
{
    return (uint32_t)__ldg(reinterpret_cast<const T *>(row) + x);
}
/* YUV/alpha codes feeding arithmetic: 16-bit containers are clamped to yuvMax (ref: src/reformat.c:712, :1005) */
template <typename YT>
__device__ __forceinline__ uint32_t ld_code(const uint8_t * row, uint32_t x, uint32_t yuvMax)
{
    const uint32_t v = ld_sample<YT>(row, x);
    return (sizeof(YT) == 2) ? min(v, yuvMax) : v;
}

template <typename T>
__device__ __forceinline__ void st_chan(uint8_t * p, uint32_t v)
{
    *reinterpret_cast<T *>(p) = (T)v;
}

/* ---------------------------------------------------------------------------------------------- */
/* YUV -> RGB: one pixel                                                                           */

struct PixelIn {
    uint32_t y;          /* luma code (clamped) */
    float cb, cr;        /* normalised (and, for bilinear, already blended) chroma */
    uint32_t u, v;       /* raw chroma codes (identity8 shuffle only) */
    uint32_t aRaw;       /* source alpha sample, unclamped (alpha channel copy, ref: src/alpha.c:56) */
    uint32_t aDst;       /* alpha already present in the destination (AB_ALPHA_KEEP) */
};

struct PixelOut {
    uint32_t c0, c1, c2; /* R,G,B codes -- or gray in c0 -- or the RGB565 word in c0 */
    uint32_t a;
};

template <bool PLAIN>
__device__ __forceinline__ PixelOut convert_pixel(const PixelIn & in, bool hasChroma, const AbY2RJob & J)
{
    PixelOut o;
    /* alpha channel value (ref: src/reformat.c:1464-1486 -> src/alpha.c:9-149) */
    if (J.alphaWrite == AB_ALPHA_COPY) {
        o.a = J.alphaRescale ? alpha_rescale(in.aRaw, J.yuvMaxDiv, J.rgbMaxF, J.rgbMax) : in.aRaw;
    } else if (J.alphaWrite == AB_ALPHA_FILL) {
        o.a = J.rgbMax;
    } else {
        o.a = in.aDst;
    }

    if (!PLAIN && J.mode == AB_MODE_IDENTITY8_SHUFFLE) { /* ref: src/reformat.c:1278-1309 */
        o.c0 = in.v;
        o.c1 = in.y;
        o.c2 = in.u;
    } else {
        const float Y = norm_y(in.y, J);
        if (!PLAIN && J.isGray) { /* ref: src/reformat.c:888-892 */
            float g = clamp01(Y), g1 = g, g2 = g;
            if (J.alphaMul == AB_MUL_FLOAT_MUL || J.alphaMul == AB_MUL_FLOAT_UNMUL) alpha_float(g, g1, g2, min(in.aRaw, J.yuvMax), J);
            o.c0 = quant_rgb(g, J.rgbMaxF);
            o.c1 = o.c2 = 0;
        } else {
            RGBf c;
            if (PLAIN) {
                if (hasChroma) {
                    c.r = __fadd_rn(Y, __fmul_rn(J.c1, in.cr));
                    c.b = __fadd_rn(Y, __fmul_rn(J.c2, in.cb));
                    const float t = __fmul_rn(2.0f, __fadd_rn(__fmul_rn(J.c3, in.cr), __fmul_rn(J.c4, in.cb)));
                    c.g = __fsub_rn(Y, div_const(t, J.kg));
                } else {
                    c.r = c.g = c.b = Y;
                }
            } else {
                c = yuv_to_rgb_norm(Y, in.cb, in.cr, in.y, hasChroma, J);
            }
            float r = clamp01(c.r), g = clamp01(c.g), b = clamp01(c.b);
            if (!PLAIN && (J.alphaMul == AB_MUL_FLOAT_MUL || J.alphaMul == AB_MUL_FLOAT_UNMUL)) alpha_float(r, g, b, min(in.aRaw, J.yuvMax), J);
            o.c0 = quant_rgb(r, J.rgbMaxF);
            o.c1 = quant_rgb(g, J.rgbMaxF);
            o.c2 = quant_rgb(b, J.rgbMaxF);
        }
    }
    if (!PLAIN) {
        if (J.alphaMul == AB_MUL_INT_PREMUL || J.alphaMul == AB_MUL_INT_UNPREMUL) { /* ref: src/reformat.c:1575-1585 -> src/alpha.c */
            const bool un = (J.alphaMul == AB_MUL_INT_UNPREMUL);
            o.c0 = alpha_int(o.c0, o.a, J.rgbMax, J.rgbMaxDiv, un);
            o.c1 = alpha_int(o.c1, o.a, J.rgbMax, J.rgbMaxDiv, un);
            o.c2 = alpha_int(o.c2, o.a, J.rgbMax, J.rgbMaxDiv, un);
        }
        if (J.toF16) { /* ref: src/reformat.c:1419-1443: every channel, alpha included */
            o.c0 = to_f16_bits(o.c0, J.f16mul);
            o.c1 = to_f16_bits(o.c1, J.f16mul);
            o.c2 = to_f16_bits(o.c2, J.f16mul);
            o.a = to_f16_bits(o.a, J.f16mul);
        }
        if (J.is565) o.c0 = pack565(o.c0, o.c1, o.c2);
    }
    return o;
}

/* place one converted pixel at `px` (global or shared memory) */
template <typename RT, bool PLAIN>
__device__ __forceinline__ void store_pixel(uint8_t * px, const PixelOut & o, const AbY2RJob & J)
{
    if (!PLAIN && J.is565) {
        st_chan<uint16_t>(px, o.c0);
        return;
    }
    if (!PLAIN && J.isGray) {
        st_chan<RT>(px + J.offGray, o.c0);
    } else {
        st_chan<RT>(px + J.offR, o.c0);
        st_chan<RT>(px + J.offG, o.c1);
        st_chan<RT>(px + J.offB, o.c2);
    }
    if (J.hasAlphaChan) st_chan<RT>(px + J.offA, o.a);
}

/* blended chroma of pixel (i,j): 4 taps per plane with the reference's neighbour choice and edge replication
 * (ref: src/reformat.c:766-838) */
template <typename YT>
__device__ __forceinline__ void bilinear_chroma(uint32_t i, uint32_t j, const AbY2RJob & J, float & cb, float & cr)
{
    const uint32_t ui = i >> 1, uj = j >> J.shiftY;
    int dc, dr;
    if (i == 0 || (i == J.width - 1 && (i & 1))) dc = 0;
    else dc = (i & 1) ? 1 : -1;
    if (!J.vertBlend || j == 0 || (j == J.height - 1 && (j & 1))) dr = 0;
    else dr = (j & 1) ? 1 : -1;
    const uint8_t * u0 = J.u + (int64_t)uj * J.uRB;
    const uint8_t * v0 = J.v + (int64_t)uj * J.vRB;
    const uint8_t * u1 = u0 + (int64_t)dr * J.uRB;
    const uint8_t * v1 = v0 + (int64_t)dr * J.vRB;
    const uint32_t ub = (uint32_t)((int)ui + dc);
    const float u00 = norm_uv(ld_code<YT>(u0, ui, J.yuvMax), J), u10 = norm_uv(ld_code<YT>(u0, ub, J.yuvMax), J);
    const float u01 = norm_uv(ld_code<YT>(u1, ui, J.yuvMax), J), u11 = norm_uv(ld_code<YT>(u1, ub, J.yuvMax), J);
    const float v00 = norm_uv(ld_code<YT>(v0, ui, J.yuvMax), J), v10 = norm_uv(ld_code<YT>(v0, ub, J.yuvMax), J);
    const float v01 = norm_uv(ld_code<YT>(v1, ui, J.yuvMax), J), v11 = norm_uv(ld_code<YT>(v1, ub, J.yuvMax), J);
    cb = bilinear4(u00, u10, u01, u11);
    cr = bilinear4(v00, v10, v01, v11);
}

/* gather the inputs of pixel (i,j) */
template <typename YT, typename RT, int CHROMA>
__device__ __forceinline__ PixelIn load_pixel(uint32_t i, uint32_t j, const AbY2RJob & J)
{
    PixelIn in;
    in.cb = in.cr = 0.5f;
    in.u = in.v = 0;
    in.aRaw = 0;
    in.aDst = 0;
    in.y = ld_code<YT>(J.y + (int64_t)j * J.yRB, i, J.yuvMax);
    if (CHROMA == AB_CHROMA_FULL || CHROMA == AB_CHROMA_NEAREST) {
        const uint32_t ui = i >> J.shiftX, uj = j >> J.shiftY;
        in.u = ld_code<YT>(J.u + (int64_t)uj * J.uRB, ui, J.yuvMax);
        in.v = ld_code<YT>(J.v + (int64_t)uj * J.vRB, ui, J.yuvMax);
        in.cb = norm_uv(in.u, J);
        in.cr = norm_uv(in.v, J);
    } else if (CHROMA == AB_CHROMA_BILINEAR) {
        bilinear_chroma<YT>(i, j, J, in.cb, in.cr);
    }
    if (J.a) in.aRaw = ld_sample<YT>(J.a + (int64_t)j * J.aRB, i);
    if (J.alphaWrite == AB_ALPHA_KEEP && J.hasAlphaChan) {
        in.aDst = ld_sample<RT>(J.rgb + (int64_t)j * J.rgbRB + (int64_t)i * J.pixBytes + J.offA, 0);
    }
    return in;
}

/* ---------------------------------------------------------------------------------------------- */
/* generic kernel: one thread per pixel, direct byte-addressed stores (any alignment, any stride)   */

template <typename YT, typename RT>
__global__ void __launch_bounds__(256) y2r_generic_kernel(const __grid_constant__ AbY2RJob J)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = J.rowBegin + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= J.width || j >= J.rowEnd) return;
    PixelIn in;
    switch (J.chroma) {
        case AB_CHROMA_FULL: in = load_pixel<YT, RT, AB_CHROMA_FULL>(i, j, J); break;
        case AB_CHROMA_NEAREST: in = load_pixel<YT, RT, AB_CHROMA_NEAREST>(i, j, J); break;
        case AB_CHROMA_BILINEAR: in = load_pixel<YT, RT, AB_CHROMA_BILINEAR>(i, j, J); break;
        default: in = load_pixel<YT, RT, AB_CHROMA_NONE>(i, j, J); break;
    }
    const PixelOut o = convert_pixel<false>(in, J.chroma != AB_CHROMA_NONE, J);
    store_pixel<RT, false>(J.rgb + (int64_t)j * J.rgbRB + (int64_t)i * J.pixBytes, o, J);
}

/* ---------------------------------------------------------------------------------------------- */
/* tiled kernel: a warp converts 256 consecutive pixels of one row.  Lane l owns the pixel PAIRS          */
/* l, l+32, l+64, l+96 of the tile: a pair is one 32-bit (16-bit samples) or 16-bit (8-bit samples) load  */
/* per plane, one shared chroma sample for 4:2:x, and one 16/12/8/6/4/2-byte packed shared-memory store.  */
/* All loads of the tile are issued before any arithmetic; the packed tile then leaves the SM as 16-byte   */
/* coalesced vector stores whatever the pixel format.                                                      */

constexpr int kTilePx = 256;       /* pixels per warp tile */
constexpr int kPairsPerLane = 4;   /* kTilePx / 2 / 32 */
constexpr int kWarpsPerBlock = 8;

/* copy `n` bytes of a warp's staging tile to global memory with the widest granularity the address allows */
__device__ __forceinline__ void warp_copy_out(uint8_t * __restrict__ dst, const uint8_t * __restrict__ stage, uint32_t n, uint32_t lane)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(dst);
    if ((addr & 15) == 0) {
        const uint32_t nv = n >> 4;
        const uint4 * s = reinterpret_cast<const uint4 *>(stage);
        uint4 * d = reinterpret_cast<uint4 *>(dst);
        for (uint32_t k = lane; k < nv; k += 32) d[k] = s[k];
        for (uint32_t b = (nv << 4) + lane; b < n; b += 32) dst[b] = stage[b];
    } else if ((addr & 3) == 0) {
        const uint32_t nv = n >> 2;
        const uint32_t * s = reinterpret_cast<const uint32_t *>(stage);
        uint32_t * d = reinterpret_cast<uint32_t *>(dst);
        for (uint32_t k = lane; k < nv; k += 32) d[k] = s[k];
        for (uint32_t b = (nv << 2) + lane; b < n; b += 32) dst[b] = stage[b];
    } else {
        for (uint32_t b = lane; b < n; b += 32) dst[b] = stage[b];
    }
}

/* two horizontally adjacent samples of one plane row, x even; `both` = the second one exists */
template <typename T>
__device__ __forceinline__ void ld_pair(const uint8_t * row, uint32_t x, bool both, bool aligned, uint32_t & s0, uint32_t & s1)
{
    if (both && aligned) {
        if (sizeof(T) == 2) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(row) + (x >> 1));
            s0 = w & 0xFFFFu;
            s1 = w >> 16;
        } else {
            const uint32_t w = __ldg(reinterpret_cast<const uint16_t *>(row) + (x >> 1));
            s0 = w & 0xFFu;
            s1 = w >> 8;
        }
    } else {
        s0 = ld_sample<T>(row, x);
        s1 = both ? ld_sample<T>(row, x + 1) : 0;
    }
}

/* memory-order channel list of one converted pixel: n values of sizeof(RT) bytes each */
template <bool PLAIN>
__device__ __forceinline__ int channel_list(const PixelOut & o, const AbY2RJob & J, uint32_t ch[4])
{
    if (!PLAIN && (J.is565 || J.isGray)) {
        if (J.is565 || !J.hasAlphaChan) {
            ch[0] = o.c0;
            return 1;
        }
        ch[0] = J.alphaFirst ? o.a : o.c0;
        ch[1] = J.alphaFirst ? o.c0 : o.a;
        return 2;
    }
    const uint32_t first = J.swapRB ? o.c2 : o.c0, last = J.swapRB ? o.c0 : o.c2;
    if (!J.hasAlphaChan) {
        ch[0] = first;
        ch[1] = o.c1;
        ch[2] = last;
        return 3;
    }
    if (J.alphaFirst) {
        ch[0] = o.a;
        ch[1] = first;
        ch[2] = o.c1;
        ch[3] = last;
    } else {
        ch[0] = first;
        ch[1] = o.c1;
        ch[2] = last;
        ch[3] = o.a;
    }
    return 4;
}

/* pack one pixel pair (or the single last pixel of an odd row) into the staging tile at pixel index p0 */
template <typename RT, bool PLAIN>
__device__ __forceinline__ void stage_pair(uint8_t * stage, uint32_t p0, const PixelOut & o0, const PixelOut & o1, bool both, const AbY2RJob & J)
{
    uint32_t a[4], b[4];
    const int n = channel_list<PLAIN>(o0, J, a);
    channel_list<PLAIN>(o1, J, b);
    const bool wide = (sizeof(RT) == 2) || (!PLAIN && J.is565); /* 16-bit units */
    uint8_t * px = stage + p0 * J.pixBytes;
    if (!both) { /* odd tail: plain per-channel stores */
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < n) {
                if (wide) st_chan<uint16_t>(px + 2 * c, a[c]);
                else st_chan<uint8_t>(px + c, a[c]);
            }
        }
        return;
    }
    if (wide) {
        if (n == 4) {
            *reinterpret_cast<uint4 *>(px) = make_uint4(a[0] | (a[1] << 16), a[2] | (a[3] << 16), b[0] | (b[1] << 16), b[2] | (b[3] << 16));
        } else if (n == 3) {
            uint32_t * w = reinterpret_cast<uint32_t *>(px); /* 12 bytes per pair: 4-byte aligned */
            w[0] = a[0] | (a[1] << 16);
            w[1] = a[2] | (b[0] << 16);
            w[2] = b[1] | (b[2] << 16);
        } else if (n == 2) {
            *reinterpret_cast<uint2 *>(px) = make_uint2(a[0] | (a[1] << 16), b[0] | (b[1] << 16));
        } else {
            *reinterpret_cast<uint32_t *>(px) = a[0] | (b[0] << 16);
        }
    } else {
        if (n == 4) {
            *reinterpret_cast<uint2 *>(px) = make_uint2(a[0] | (a[1] << 8) | (a[2] << 16) | (a[3] << 24), b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24));
        } else if (n == 3) {
            uint16_t * h = reinterpret_cast<uint16_t *>(px); /* 6 bytes per pair: 2-byte aligned */
            h[0] = (uint16_t)(a[0] | (a[1] << 8));
            h[1] = (uint16_t)(a[2] | (b[0] << 8));
            h[2] = (uint16_t)(b[1] | (b[2] << 8));
        } else if (n == 2) {
            *reinterpret_cast<uint32_t *>(px) = a[0] | (a[1] << 8) | (b[0] << 16) | (b[1] << 24);
        } else {
            *reinterpret_cast<uint16_t *>(px) = (uint16_t)(a[0] | (b[0] << 8));
        }
    }
}

struct PairIn {
    uint32_t y0, y1;
    uint32_t u0, u1, v0, v1; /* raw chroma codes: one per pixel (4:4:4) or u0/v0 shared (nearest) */
    uint32_t a0, a1;
    float cb0, cr0, cb1, cr1; /* bilinear only: blended normalised chroma */
};

template <typename YT, typename RT, int CHROMA, bool PLAIN>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) y2r_tiled_kernel(const __grid_constant__ AbY2RJob J)
{
    __shared__ __align__(16) uint8_t stageAll[kWarpsPerBlock][kTilePx * 4 * sizeof(RT)];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t j = J.rowBegin + blockIdx.y * kWarpsPerBlock + warp;
    if (j >= J.rowEnd) return; /* warp-uniform */
    const uint32_t x0 = blockIdx.x * kTilePx;
    uint8_t * stage = stageAll[warp];
    const bool aligned = J.pairAligned != 0;

    const uint8_t * rowY = J.y + (int64_t)j * J.yRB;
    const uint8_t * rowA = J.a ? J.a + (int64_t)j * J.aRB : nullptr;
    const uint32_t uj = j >> J.shiftY;
    const uint8_t * rowU = (CHROMA != AB_CHROMA_NONE) ? J.u + (int64_t)uj * J.uRB : nullptr;
    const uint8_t * rowV = (CHROMA != AB_CHROMA_NONE) ? J.v + (int64_t)uj * J.vRB : nullptr;
    const bool keepAlpha = (J.alphaWrite == AB_ALPHA_KEEP) && J.hasAlphaChan;

    PairIn in[kPairsPerLane];
#pragma unroll
    for (int k = 0; k < kPairsPerLane; ++k) {
        const uint32_t i0 = x0 + 2 * (lane + 32 * k);
        PairIn & p = in[k];
        p.y0 = p.y1 = p.u0 = p.u1 = p.v0 = p.v1 = p.a0 = p.a1 = 0;
        p.cb0 = p.cr0 = p.cb1 = p.cr1 = 0.5f;
        if (i0 >= J.width) continue;
        const bool both = (i0 + 1 < J.width);
        ld_pair<YT>(rowY, i0, both, aligned, p.y0, p.y1);
        if (CHROMA == AB_CHROMA_FULL) {
            ld_pair<YT>(rowU, i0, both, aligned, p.u0, p.u1);
            ld_pair<YT>(rowV, i0, both, aligned, p.v0, p.v1);
        } else if (CHROMA == AB_CHROMA_NEAREST) {
            p.u0 = ld_sample<YT>(rowU, i0 >> 1);
            p.v0 = ld_sample<YT>(rowV, i0 >> 1);
        } else if (CHROMA == AB_CHROMA_BILINEAR) {
            bilinear_chroma<YT>(i0, j, J, p.cb0, p.cr0);
            if (both) bilinear_chroma<YT>(i0 + 1, j, J, p.cb1, p.cr1);
        }
        if (rowA) ld_pair<YT>(rowA, i0, both, aligned, p.a0, p.a1);
        if (keepAlpha) {
            const uint8_t * d = J.rgb + (int64_t)j * J.rgbRB + (int64_t)i0 * J.pixBytes + J.offA;
            p.a0 = ld_sample<RT>(d, 0);
            if (both) p.a1 = ld_sample<RT>(d + J.pixBytes, 0);
        }
    }
#pragma unroll
    for (int k = 0; k < kPairsPerLane; ++k) {
        const uint32_t pidx = 2 * (lane + 32 * k);
        const uint32_t i0 = x0 + pidx;
        if (i0 >= J.width) continue;
        const bool both = (i0 + 1 < J.width);
        const PairIn & p = in[k];
        PixelIn q0, q1;
        q0.y = (sizeof(YT) == 2) ? min(p.y0, J.yuvMax) : p.y0;
        q1.y = (sizeof(YT) == 2) ? min(p.y1, J.yuvMax) : p.y1;
        q0.cb = q0.cr = q1.cb = q1.cr = 0.5f;
        q0.u = q0.v = q1.u = q1.v = 0;
        if (CHROMA == AB_CHROMA_FULL) {
            q0.u = (sizeof(YT) == 2) ? min(p.u0, J.yuvMax) : p.u0;
            q0.v = (sizeof(YT) == 2) ? min(p.v0, J.yuvMax) : p.v0;
            q1.u = (sizeof(YT) == 2) ? min(p.u1, J.yuvMax) : p.u1;
            q1.v = (sizeof(YT) == 2) ? min(p.v1, J.yuvMax) : p.v1;
            q0.cb = norm_uv(q0.u, J);
            q0.cr = norm_uv(q0.v, J);
            q1.cb = norm_uv(q1.u, J);
            q1.cr = norm_uv(q1.v, J);
        } else if (CHROMA == AB_CHROMA_NEAREST) {
            q0.u = q1.u = (sizeof(YT) == 2) ? min(p.u0, J.yuvMax) : p.u0;
            q0.v = q1.v = (sizeof(YT) == 2) ? min(p.v0, J.yuvMax) : p.v0;
            q0.cb = q1.cb = norm_uv(q0.u, J);
            q0.cr = q1.cr = norm_uv(q0.v, J);
        } else if (CHROMA == AB_CHROMA_BILINEAR) {
            q0.cb = p.cb0;
            q0.cr = p.cr0;
            q1.cb = p.cb1;
            q1.cr = p.cr1;
        }
        q0.aRaw = q0.aDst = p.a0;
        q1.aRaw = q1.aDst = p.a1;
        const PixelOut o0 = convert_pixel<PLAIN>(q0, CHROMA != AB_CHROMA_NONE, J);
        const PixelOut o1 = convert_pixel<PLAIN>(q1, CHROMA != AB_CHROMA_NONE, J);
        stage_pair<RT, PLAIN>(stage, pidx, o0, o1, both, J);
    }
    __syncwarp();
    const uint32_t npx = min((uint32_t)kTilePx, J.width - x0);
    warp_copy_out(J.rgb + (int64_t)j * J.rgbRB + (int64_t)x0 * J.pixBytes, stage, npx * J.pixBytes, lane);
}

/* ---------------------------------------------------------------------------------------------- */
/* RGB -> YUV                                                                                      */

struct YUVf {
    float y, u, v;
};

struct RgbCodes {
    uint32_t r, g, b, a; /* gray input: the gray code is in r */
};

/* ref: src/reformat.c:311-387 (colour) and :479-511 (gray), one pixel, codes already in registers */
__device__ __forceinline__ YUVf rgb_codes_to_yuv(const RgbCodes & px, const AbR2YJob & J)
{
    YUVf o;
    const float m = J.rgbMaxF;
    float a = 1.0f;
    if (J.alphaMul != AB_MUL_NONE) a = div_const(u2f(px.a), J.rgbMaxDiv);
    if (J.isGray) {
        float g = div_const(u2f(px.r), J.rgbMaxDiv);
        if (J.alphaMul != AB_MUL_NONE) {
            if (a == 0.0f) g = 0.0f;
            else if (a < 1.0f) g = (J.alphaMul == AB_MUL_FLOAT_MUL) ? __fmul_rn(g, a) : min1(__fdiv_rn(g, a));
        }
        o.y = g;
        o.u = o.v = 0.0f;
        return o;
    }
    float r = div_const(u2f(px.r), J.rgbMaxDiv); /* c / rgbMaxF: true quotients, ref: :312-323 */
    float g = div_const(u2f(px.g), J.rgbMaxDiv);
    float b = div_const(u2f(px.b), J.rgbMaxDiv);
    if (J.alphaMul != AB_MUL_NONE) { /* ref: :325-358 */
        if (a == 0.0f) {
            r = g = b = 0.0f;
        } else if (a < 1.0f) {
            if (J.alphaMul == AB_MUL_FLOAT_MUL) {
                r = __fmul_rn(r, a);
                g = __fmul_rn(g, a);
                b = __fmul_rn(b, a);
            } else {
                r = min1(__fdiv_rn(r, a));
                g = min1(__fdiv_rn(g, a));
                b = min1(__fdiv_rn(b, a));
            }
        }
    }
    if (J.mode == AB_MODE_COEFF) { /* ref: :383-386 */
        const float Y = __fadd_rn(__fadd_rn(__fmul_rn(J.kr, r), __fmul_rn(J.kg, g)), __fmul_rn(J.kb, b));
        o.y = Y;
        o.u = div_const(__fsub_rn(b, Y), J.du);
        o.v = div_const(__fsub_rn(r, Y), J.dv);
    } else if (J.mode == AB_MODE_IDENTITY) { /* ref: :361-365 */
        o.y = g;
        o.u = b;
        o.v = r;
    } else if (J.mode == AB_MODE_YCGCO) { /* ref: :366-370 */
        const float s = __fmul_rn(0.25f, __fadd_rn(r, b));
        o.y = __fadd_rn(__fmul_rn(0.5f, g), s);
        o.u = __fsub_rn(__fmul_rn(0.5f, g), s);
        o.v = __fmul_rn(0.5f, __fsub_rn(r, b));
    } else { /* YCgCo-Re/Ro, ref: :371-381 */
        float t0 = __fmul_rn(r, m), t1 = __fmul_rn(g, m), t2 = __fmul_rn(b, m);
        t0 = (t0 < 0.0f) ? 0.0f : ((m < t0) ? m : t0);
        t1 = (t1 < 0.0f) ? 0.0f : ((m < t1) ? m : t1);
        t2 = (t2 < 0.0f) ? 0.0f : ((m < t2) ? m : t2);
        const int R = (int)round_half_up(t0), G = (int)round_half_up(t1), B = (int)round_half_up(t2);
        const int Co = R - B;
        const int t = B + (Co >> 1);
        const int Cg = G - t;
        o.y = __fdiv_rn((float)(t + (Cg >> 1)), J.rangeYdiv);
        o.u = __fdiv_rn((float)Cg, J.rangeUVdiv);
        o.v = __fdiv_rn((float)Co, J.rangeUVdiv);
    }
    return o;
}

/* ref: src/reformat.c:197-219: clamp((int)floorf(v*range + bias + 0.5f), 0, yuvMax) */
__device__ __forceinline__ uint32_t quant_yuv(float v, float range, float bias, uint32_t yuvMax)
{
    const int q = (int)round_half_up(__fadd_rn(__fmul_rn(v, range), bias));
    return (uint32_t)min(max(q, 0), (int)yuvMax);
}

template <typename RT>
__device__ __forceinline__ RgbCodes load_rgb_scalar(const uint8_t * px, const AbR2YJob & J)
{
    RgbCodes c;
    c.r = c.g = c.b = c.a = 0;
    if (J.isGray) {
        c.r = ld_sample<RT>(px + J.offGray, 0);
    } else {
        c.r = ld_sample<RT>(px + J.offR, 0);
        c.g = ld_sample<RT>(px + J.offG, 0);
        c.b = ld_sample<RT>(px + J.offB, 0);
    }
    if (J.hasAlphaChan) c.a = ld_sample<RT>(px + J.offA, 0);
    return c;
}

/* memory-order channel values -> named channels (inverse of channel_list) */
__device__ __forceinline__ RgbCodes name_channels(const uint32_t ch[4], const AbR2YJob & J)
{
    RgbCodes c;
    c.g = c.b = c.a = 0;
    if (J.isGray) {
        if (J.nChan == 1) {
            c.r = ch[0];
        } else {
            c.r = J.alphaFirst ? ch[1] : ch[0];
            c.a = J.alphaFirst ? ch[0] : ch[1];
        }
        return c;
    }
    uint32_t x0, x1, x2;
    if (J.nChan == 4 && J.alphaFirst) {
        c.a = ch[0];
        x0 = ch[1];
        x1 = ch[2];
        x2 = ch[3];
    } else {
        x0 = ch[0];
        x1 = ch[1];
        x2 = ch[2];
        if (J.nChan == 4) c.a = ch[3];
    }
    c.r = J.swapRB ? x2 : x0;
    c.g = x1;
    c.b = J.swapRB ? x0 : x2;
    return c;
}

/* two horizontally adjacent pixels with one vector load (rgb base/stride aligned to the pair size) */
template <typename RT>
__device__ __forceinline__ void load_rgb_pair_vec(const uint8_t * px, const AbR2YJob & J, RgbCodes & p0, RgbCodes & p1)
{
    uint32_t w[4] = { 0, 0, 0, 0 };
    const uint32_t pairBytes = 2u * J.pixBytes;
    if (pairBytes == 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(px));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if (pairBytes == 12) {
        const uint32_t * q = reinterpret_cast<const uint32_t *>(px);
        w[0] = __ldg(q); w[1] = __ldg(q + 1); w[2] = __ldg(q + 2);
    } else if (pairBytes == 8) {
        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(px));
        w[0] = v.x; w[1] = v.y;
    } else if (pairBytes == 6) {
        const uint16_t * q = reinterpret_cast<const uint16_t *>(px);
        w[0] = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 16);
        w[1] = __ldg(q + 2);
    } else if (pairBytes == 4) {
        w[0] = __ldg(reinterpret_cast<const uint32_t *>(px));
    } else {
        w[0] = __ldg(reinterpret_cast<const uint16_t *>(px));
    }
    uint32_t a[4] = { 0, 0, 0, 0 }, b[4] = { 0, 0, 0, 0 };
    const int n = J.nChan;
    if (sizeof(RT) == 2) { /* unit k of the pair = 16 bits at w[k/2] */
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < n) {
                a[k] = (w[k >> 1] >> ((k & 1) * 16)) & 0xFFFFu;
                const int kb = k + n;
                b[k] = (w[kb >> 1] >> ((kb & 1) * 16)) & 0xFFFFu;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < n) {
                a[k] = (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
                const int kb = k + n;
                b[k] = (w[kb >> 2] >> ((kb & 3) * 8)) & 0xFFu;
            }
        }
    }
    p0 = name_channels(a, J);
    p1 = name_channels(b, J);
}

/* store one or two horizontally adjacent samples of a plane row */
template <typename YT>
__device__ __forceinline__ void st_pair(uint8_t * row, uint32_t x, uint32_t s0, uint32_t s1, bool both, bool aligned)
{
    if (both && aligned) {
        if (sizeof(YT) == 2) reinterpret_cast<uint32_t *>(row)[x >> 1] = s0 | (s1 << 16);
        else reinterpret_cast<uint16_t *>(row)[x >> 1] = (uint16_t)(s0 | (s1 << 8));
    } else {
        reinterpret_cast<YT *>(row)[x] = (YT)s0;
        if (both) reinterpret_cast<YT *>(row)[x + 1] = (YT)s1;
    }
}

/* One thread = one 2x2 block, exactly the reference's block walk (ref: src/reformat.c:295-470).  Lanes own
 * horizontally consecutive blocks, so a warp reads two row segments of 64 pixels with one vector load per
 * pixel pair and writes Y/A (and 4:4:4 chroma) as packed pairs, subsampled chroma as one sample per lane.
 * VEC=false is the byte-addressed variant (any alignment; also the cross-check implementation). */
template <typename RT, typename YT, bool VEC>
__global__ void __launch_bounds__(256) r2y_block_kernel(const __grid_constant__ AbR2YJob J)
{
    const uint32_t oi = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const uint32_t oj = J.rowBegin + 2 * (blockIdx.y * blockDim.y + threadIdx.y);
    if (oi >= J.width || oj >= J.rowEnd) return;
    const bool both = (oi + 1 < J.width);
    const uint32_t bh = (oj + 1 >= J.height) ? 1 : 2;
    const bool srcVec = VEC && J.srcPairAligned && both;
    const bool dstVec = VEC && J.dstPairAligned;
    const bool full444 = J.hasChroma && !J.isGray && J.shiftX == 0;

    RgbCodes px[2][2];
#pragma unroll
    for (uint32_t bj = 0; bj < 2; ++bj) {
        if (bj < bh) {
            const uint8_t * p = J.rgb + (int64_t)(oj + bj) * J.rgbRB + (int64_t)oi * J.pixBytes;
            if (srcVec) {
                load_rgb_pair_vec<RT>(p, J, px[bj][0], px[bj][1]);
            } else {
                px[bj][0] = load_rgb_scalar<RT>(p, J);
                if (both) px[bj][1] = load_rgb_scalar<RT>(p + J.pixBytes, J);
            }
        }
    }

    float bu[2][2], bv[2][2]; /* [bj][bi] */
#pragma unroll
    for (uint32_t bj = 0; bj < 2; ++bj) {
        bu[bj][0] = bu[bj][1] = bv[bj][0] = bv[bj][1] = 0.0f;
        if (bj >= bh) continue;
        const uint32_t j = oj + bj;
        uint32_t yq[2] = { 0, 0 }, uq[2] = { 0, 0 }, vq[2] = { 0, 0 }, aq[2] = { 0, 0 };
#pragma unroll
        for (uint32_t bi = 0; bi < 2; ++bi) {
            if (bi == 1 && !both) continue;
            const YUVf c = rgb_codes_to_yuv(px[bj][bi], J);
            bu[bj][bi] = c.u;
            bv[bj][bi] = c.v;
            yq[bi] = quant_yuv(c.y, J.rangeY, J.biasY, J.yuvMax);
            if (full444) {
                uq[bi] = quant_yuv(c.u, J.rangeUV, J.biasUV, J.yuvMax);
                vq[bi] = quant_yuv(c.v, J.rangeUV, J.biasUV, J.yuvMax);
            }
            if (J.alphaWrite == AB_ALPHA_COPY) { /* ref: :545-569 -> src/alpha.c:37-149 */
                const uint32_t a = px[bj][bi].a;
                aq[bi] = J.alphaRescale ? alpha_rescale(a, J.rgbMaxDiv, J.yuvMaxF, J.yuvMax) : a;
            } else {
                aq[bi] = J.yuvMax;
            }
        }
        st_pair<YT>(J.y + (int64_t)j * J.yRB, oi, yq[0], yq[1], both, dstVec);
        if (full444) {
            st_pair<YT>(J.u + (int64_t)j * J.uRB, oi, uq[0], uq[1], both, dstVec);
            st_pair<YT>(J.v + (int64_t)j * J.vRB, oi, vq[0], vq[1], both, dstVec);
        }
        if (J.alphaWrite != AB_ALPHA_KEEP) st_pair<YT>(J.a + (int64_t)j * J.aRB, oi, aq[0], aq[1], both, dstVec);
    }
    if (!J.hasChroma || J.isGray || J.shiftX == 0) return;
    const uint32_t bw = both ? 2 : 1;
    if (J.shiftY) { /* 4:2:0: ordered sum TL,TR,BL,BR over the samples that exist (ref: :413-440) */
        float su = 0.0f, sv = 0.0f;
#pragma unroll
        for (uint32_t bj = 0; bj < 2; ++bj)
#pragma unroll
            for (uint32_t bi = 0; bi < 2; ++bi)
                if (bj < bh && bi < bw) {
                    su = __fadd_rn(su, bu[bj][bi]);
                    sv = __fadd_rn(sv, bv[bj][bi]);
                }
        const float n = (float)(bw * bh);
        const uint32_t ui = oi >> 1, uj = oj >> 1;
        st_chan<YT>(J.u + (int64_t)uj * J.uRB + (int64_t)ui * sizeof(YT), quant_yuv(__fdiv_rn(su, n), J.rangeUV, J.biasUV, J.yuvMax));
        st_chan<YT>(J.v + (int64_t)uj * J.vRB + (int64_t)ui * sizeof(YT), quant_yuv(__fdiv_rn(sv, n), J.rangeUV, J.biasUV, J.yuvMax));
    } else { /* 4:2:2: per-row pair (ref: :441-467) */
#pragma unroll
        for (uint32_t bj = 0; bj < 2; ++bj) {
            if (bj >= bh) break;
            float su = 0.0f, sv = 0.0f;
#pragma unroll
            for (uint32_t bi = 0; bi < 2; ++bi)
                if (bi < bw) {
                    su = __fadd_rn(su, bu[bj][bi]);
                    sv = __fadd_rn(sv, bv[bj][bi]);
                }
            const float n = (float)bw;
            const uint32_t ui = oi >> 1, uj = oj + bj;
            st_chan<YT>(J.u + (int64_t)uj * J.uRB + (int64_t)ui * sizeof(YT), quant_yuv(__fdiv_rn(su, n), J.rangeUV, J.biasUV, J.yuvMax));
            st_chan<YT>(J.v + (int64_t)uj * J.vRB + (int64_t)ui * sizeof(YT), quant_yuv(__fdiv_rn(sv, n), J.rangeUV, J.biasUV, J.yuvMax));
        }
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* standalone passes                                                                               */

/* ref: src/alpha.c:9-149 */
template <typename ST, typename DT>
__global__ void __launch_bounds__(256) alpha_kernel(const __grid_constant__ AbAlphaJob J)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= J.width || j >= J.height) return;
    uint8_t * d = J.dst + J.dstOff + (int64_t)j * J.dstRB + (int64_t)i * J.dstPix;
    if (!J.src) {
        st_chan<DT>(d, (J.dstDepth > 8) ? ((1u << J.dstDepth) - 1) : 255u);
        return;
    }
    const uint32_t a = ld_sample<ST>(J.src + J.srcOff + (int64_t)j * J.srcRB + (int64_t)i * J.srcPix, 0);
    if (J.srcDepth == J.dstDepth) {
        st_chan<DT>(d, a);
    } else {
        const uint32_t srcMax = (1u << J.srcDepth) - 1, dstMax = (1u << J.dstDepth) - 1;
        AbDiv sd;
        sd.d = (float)srcMax;
        sd.r = sd.rh = sd.rl = 0.0f;
        sd.exact = sd.intExact = 0;
        st_chan<DT>(d, alpha_rescale(a, sd, (float)dstMax, dstMax));
    }
}

/* ref: src/alpha.c:151-535, in place */
template <typename RT>
__global__ void __launch_bounds__(256) premultiply_kernel(const __grid_constant__ AbRgbPassJob J)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= J.width || j >= J.height) return;
    RT * px = reinterpret_cast<RT *>(J.pixels + (int64_t)j * J.rowBytes) + (size_t)i * J.nChan;
    const uint32_t max = (1u << J.depth) - 1;
    const uint32_t a = px[J.aIdx];
    if (a >= max) return;
    for (int c = 0; c < J.nColor; ++c) {
        RT * cp = px + J.cIdx[c];
        *cp = (RT)alpha_int(*cp, a, max, J.maxDiv, J.unmultiply != 0);
    }
}

/* ref: src/reformat.c:1419-1443, in place over width*channels samples per row */
__global__ void __launch_bounds__(256) f16_kernel(const __grid_constant__ AbRgbPassJob J)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= J.width * J.nChan || j >= J.height) return;
    uint16_t * p = reinterpret_cast<uint16_t *>(J.pixels + (int64_t)j * ((J.rowBytes >> 1) << 1)) + i;
    *p = (uint16_t)to_f16_bits(*p, J.f16mul);
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T * dst, size_t count, T value)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dst[k] = value;
}

/* ref: src/read.c:6724-6764 */
template <typename T>
__global__ void __launch_bounds__(256) limited_to_full_kernel(uint8_t * plane, int64_t rowBytes, uint32_t width, uint32_t height, uint32_t depth)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= width || j >= height) return;
    T * p = reinterpret_cast<T *>(plane + (int64_t)j * rowBytes) + i;
    *p = (T)limited_to_full_y(depth, (int)*p);
}

/* plane copy with replicate padding of the right/bottom borders (ref: src/write.c:1151-1205) */
template <typename T>
__global__ void __launch_bounds__(256) copy_pad_kernel(const uint8_t * src, int64_t srcRB, uint32_t srcW, uint32_t srcH, uint8_t * dst, int64_t dstRB,
                                                       uint32_t dstW, uint32_t dstH)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= dstW || j >= dstH) return;
    const uint32_t si = min(i, srcW - 1), sj = min(j, srcH - 1);
    reinterpret_cast<T *>(dst + (int64_t)j * dstRB)[i] = reinterpret_cast<const T *>(src + (int64_t)sj * srcRB)[si];
}

/* self-test of div_const's fast form: all 2^32 binary32 patterns x with x == +0 or 2^-60 <= |x| <= 2^60 */
__global__ void __launch_bounds__(256) div_selftest_kernel(float d, float r, unsigned long long * mismatches)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
        const uint32_t u = (uint32_t)b;
        const float x = __uint_as_float(u);
        const float ax = fabsf(x);
        if (!(u == 0u || (ax >= 8.6736173798840355e-19f && ax <= 1152921504606846976.0f))) continue;
        const float q = __fmul_rn(x, r);
        const float e = __fmaf_rn(-q, d, x);
        const float fast = __fmaf_rn(e, r, q);
        if (__float_as_uint(fast) != __float_as_uint(__fdiv_rn(x, d))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

/* self-test of the 2-op quotient on div_const's whole float domain */
__global__ void __launch_bounds__(256) div2_selftest_kernel(AbDiv dv, unsigned long long * mismatches)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
        const uint32_t u = (uint32_t)b;
        const float x = __uint_as_float(u);
        const float ax = fabsf(x);
        if (!(u == 0u || (ax >= 8.6736173798840355e-19f && ax <= 1152921504606846976.0f))) continue;
        if (__float_as_uint(div_int(x, dv)) != __float_as_uint(__fdiv_rn(x, dv.d))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

/* self-test of div_int: every integer numerator in [-65536, 65535] */
__global__ void __launch_bounds__(256) int_div_selftest_kernel(AbDiv dv, unsigned long long * mismatches)
{
    const int x = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 65536;
    if (x > 65535) return;
    const float xf = (float)x;
    if (__float_as_uint(div_int(xf, dv)) != __float_as_uint(__fdiv_rn(xf, dv.d))) atomicAdd(mismatches, 1ull);
}

/* self-test of the shared-reciprocal division (pixel_math.cuh: rcp_refined / div_by_rcp) against __fdiv_rn:
 * mode 0: x = c/max, a = k/max for every code pair (c, k), 0 < k < max (RGB->YUV unmultiply domain, exhaustive);
 * mode 1: a = k/max as above, x = 65536 pseudo-random floats in [0,1] per k (YUV->RGB unmultiply domain, sampled). */
__global__ void __launch_bounds__(256) alpha_div_selftest_kernel(uint32_t max, int mode, unsigned long long * mismatches)
{
    const uint64_t total = (mode == 0) ? (uint64_t)(max + 1) * (max - 1) : (uint64_t)65536 * (max - 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const float maxF = (float)max;
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t k = 1 + (uint32_t)(i % (max - 1));
        const uint32_t c = (uint32_t)(i / (max - 1));
        const float a = __fdiv_rn((float)k, maxF);
        float x;
        if (mode == 0) {
            x = __fdiv_rn((float)c, maxF);
        } else {
            uint32_t h = c * 2654435761u ^ (k * 40503u + 0x9E3779B9u);
            h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
            x = __fmul_rn((float)(h >> 8), 5.9604644775390625e-8f); /* 24 random bits in [0,1) */
            if ((h & 255u) == 0) x = 1.0f;
        }
        if (!(a < 1.0f) || a == 0.0f) continue;
        const float r1 = rcp_refined(a);
        if (__float_as_uint(div_by_rcp(x, a, r1)) != __float_as_uint(__fdiv_rn(x, a))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

/* premul_fma_magic against the reference form floorf(c * a / maxF + 0.5f) with the hardware's IEEE division, for every
 * code pair (c, a) in [0, max]^2. */
__global__ void __launch_bounds__(256) premul_fma_selftest_kernel(uint32_t max, float r, float k, unsigned long long * mismatches)
{
    const uint64_t total = (uint64_t)(max + 1) * (max + 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const float maxF = (float)max;
    unsigned long long bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t a = (uint32_t)(i % (max + 1)), c = (uint32_t)(i / (max + 1));
        const float ref = floorf(__fadd_rn(__fdiv_rn(__fmul_rn((float)c, (float)a), maxF), 0.5f));
        const float af = (float)a;
        const float got = premul_fma_magic(__uint_as_float(0x4B000000u | c), af, __fmul_rn(af, -8388608.0f), r, k);
        if ((__float_as_uint(got) & 0xFFFFu) != (uint32_t)ref || (__float_as_uint(got) >> 16) != 0x4B00u) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

/* ---------------------------------------------------------------------------------------------- */
/* launch helpers                                                                                  */

static inline int finish_launch()
{
    ++g_launches;
    return (int)cudaGetLastError();
}

static inline dim3 grid2d(uint32_t w, uint32_t h, dim3 block) { return dim3((w + block.x - 1) / block.x, (h + block.y - 1) / block.y); }

template <typename YT, typename RT>
static int launch_y2r_typed(const AbY2RJob & J, cudaStream_t st, int forceGeneric, const char ** name)
{
    const uint32_t rows = J.rowEnd - J.rowBegin;
    if (rows == 0 || J.width == 0) return 0;
    {
        /* gridDim.y <= 65535: very tall images are converted in row slices (the reference has no such limit; slices start on even rows) */
        const uint32_t maxRows = 65535u * 4u;
        if (rows > maxRows) {
            int err = 0;
            for (uint32_t r0 = J.rowBegin; r0 < J.rowEnd && !err; r0 += maxRows) {
                AbY2RJob S = J;
                S.rowBegin = r0;
                S.rowEnd = (J.rowEnd - r0 > maxRows) ? r0 + maxRows : J.rowEnd;
                err = launch_y2r_typed<YT, RT>(S, st, forceGeneric, name);
            }
            return err;
        }
    }
    if (forceGeneric) {
        const dim3 block(64, 4);
        y2r_generic_kernel<YT, RT><<<grid2d(J.width, rows, block), block, 0, st>>>(J);
        if (name) *name = "y2r_generic";
        return finish_launch();
    }
    const dim3 grid((J.width + kTilePx - 1) / kTilePx, (rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const dim3 block(kWarpsPerBlock * 32);
    const bool plain = J.mode == AB_MODE_COEFF && J.alphaMul == AB_MUL_NONE && !J.toF16 && !J.is565 && !J.isGray;
#define AB_LAUNCH(CH, PL, NAME)                                                \
    do {                                                                       \
        y2r_tiled_kernel<YT, RT, CH, PL><<<grid, block, 0, st>>>(J);           \
        if (name) *name = NAME;                                                \
    } while (0)
    switch (J.chroma) {
        case AB_CHROMA_FULL:
            if (plain) AB_LAUNCH(AB_CHROMA_FULL, true, "y2r_tiled_444_plain");
            else AB_LAUNCH(AB_CHROMA_FULL, false, "y2r_tiled_444_general");
            break;
        case AB_CHROMA_NEAREST:
            if (plain) AB_LAUNCH(AB_CHROMA_NEAREST, true, "y2r_tiled_nearest_plain");
            else AB_LAUNCH(AB_CHROMA_NEAREST, false, "y2r_tiled_nearest_general");
            break;
        case AB_CHROMA_BILINEAR:
            if (plain) AB_LAUNCH(AB_CHROMA_BILINEAR, true, "y2r_tiled_bilinear_plain");
            else AB_LAUNCH(AB_CHROMA_BILINEAR, false, "y2r_tiled_bilinear_general");
            break;
        default:
            if (plain) AB_LAUNCH(AB_CHROMA_NONE, true, "y2r_tiled_mono_plain");
            else AB_LAUNCH(AB_CHROMA_NONE, false, "y2r_tiled_mono_general");
            break;
    }
#undef AB_LAUNCH
    return finish_launch();
}

} // namespace ab

using namespace ab;

extern "C" {

int abY2RFast_8_8(const AbY2RJob * jobs, uint32_t count, void * stream);
int abY2RFast_8_16(const AbY2RJob * jobs, uint32_t count, void * stream);
int abY2RFast_16_8(const AbY2RJob * jobs, uint32_t count, void * stream);
int abY2RFast_16_16(const AbY2RJob * jobs, uint32_t count, void * stream);

static int y2r_fast_try(const AbY2RJob * jobs, uint32_t count, void * stream)
{
    if (jobs[0].yuvBytes == 2) return (jobs[0].rgbBytes == 2) ? abY2RFast_16_16(jobs, count, stream) : abY2RFast_16_8(jobs, count, stream);
    return (jobs[0].rgbBytes == 2) ? abY2RFast_8_16(jobs, count, stream) : abY2RFast_8_8(jobs, count, stream);
}

static const char * y2r_fast_name(const AbY2RJob * job)
{
    return (job->chroma == AB_CHROMA_BILINEAR) ? "y2r_fast_bilinear" : (job->chroma == AB_CHROMA_NEAREST) ? "y2r_fast_nearest" : "y2r_fast_444";
}

int abLaunchY2R(const AbY2RJob * job, void * stream, int forceGeneric, const char ** kernelName)
{
    cudaStream_t st = (cudaStream_t)stream;
    forceGeneric &= 3;
    if (forceGeneric == 0) { /* production: compile-time specialised kernels for the hot configurations (y2r_fast.cuh) */
        if (y2r_fast_try(job, 1, stream)) {
            if (kernelName) *kernelName = y2r_fast_name(job);
            return finish_launch();
        }
    }
    forceGeneric &= 1;
    if (job->yuvBytes == 2) {
        return job->rgbBytes == 2 ? launch_y2r_typed<uint16_t, uint16_t>(*job, st, forceGeneric, kernelName)
                                  : launch_y2r_typed<uint16_t, uint8_t>(*job, st, forceGeneric, kernelName);
    }
    return job->rgbBytes == 2 ? launch_y2r_typed<uint8_t, uint16_t>(*job, st, forceGeneric, kernelName)
                              : launch_y2r_typed<uint8_t, uint8_t>(*job, st, forceGeneric, kernelName);
}

int abLaunchY2RBatch(const AbY2RJob * jobs, uint32_t count, void * stream, int forceGeneric, const char ** kernelName)
{
    uint32_t k = 0;
    while (k < count) {
        /* longest run of frames with the same sample types, <= AB_MAX_BATCH: one specialised launch if they are compatible */
        uint32_t n = 1;
        while (k + n < count && n < AB_MAX_BATCH && jobs[k + n].yuvBytes == jobs[k].yuvBytes && jobs[k + n].rgbBytes == jobs[k].rgbBytes) ++n;
        /* longest prefix of the run that one specialised launch can take (frames must be eligible and share constants) */
        bool fused = false;
        if ((forceGeneric & 3) == 0) {
            for (; n > 1; --n) {
                if (y2r_fast_try(jobs + k, n, stream)) {
                    fused = true;
                    break;
                }
            }
        } else {
            n = 1;
        }
        if (fused) {
            if (kernelName) *kernelName = y2r_fast_name(jobs + k);
            const int err = finish_launch();
            if (err) return err;
        } else {
            n = 1;
            const int err = abLaunchY2R(jobs + k, stream, forceGeneric, kernelName);
            if (err) return err;
        }
        k += n;
    }
    return 0;
}

int abR2YFast_8_8(const AbR2YJob * jobs, uint32_t count, void * stream);
int abR2YFast_8_16(const AbR2YJob * jobs, uint32_t count, void * stream);
int abR2YFast_16_8(const AbR2YJob * jobs, uint32_t count, void * stream);
int abR2YFast_16_16(const AbR2YJob * jobs, uint32_t count, void * stream);

static int r2y_fast_try(const AbR2YJob * jobs, uint32_t count, void * stream)
{
    if (jobs[0].rgbBytes == 2) return (jobs[0].yuvBytes == 2) ? abR2YFast_16_16(jobs, count, stream) : abR2YFast_16_8(jobs, count, stream);
    return (jobs[0].yuvBytes == 2) ? abR2YFast_8_16(jobs, count, stream) : abR2YFast_8_8(jobs, count, stream);
}

int abLaunchR2YBatch(const AbR2YJob * jobs, uint32_t count, void * stream, int forceGeneric, const char ** kernelName)
{
    uint32_t k = 0;
    while (k < count) {
        uint32_t n = 1;
        while (k + n < count && n < AB_MAX_BATCH && jobs[k + n].yuvBytes == jobs[k].yuvBytes && jobs[k + n].rgbBytes == jobs[k].rgbBytes) ++n;
        bool fused = false;
        if ((forceGeneric & 3) == 0) {
            for (; n > 1; --n) {
                if (r2y_fast_try(jobs + k, n, stream)) {
                    fused = true;
                    break;
                }
            }
        } else {
            n = 1;
        }
        if (fused) {
            if (kernelName) *kernelName = "r2y_fast";
            const int err = finish_launch();
            if (err) return err;
        } else {
            n = 1;
            const int err = abLaunchR2Y(jobs + k, stream, forceGeneric, kernelName);
            if (err) return err;
        }
        k += n;
    }
    return 0;
}

int abLaunchR2Y(const AbR2YJob * job, void * stream, int forceGeneric, const char ** kernelName)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (job->width == 0 || job->rowEnd <= job->rowBegin) return 0;
    forceGeneric &= 3;
    if (forceGeneric == 0) { /* production: compile-time specialised kernels (r2y_fast.cuh) */
        if (r2y_fast_try(job, 1, stream)) {
            if (kernelName) *kernelName = "r2y_fast";
            return finish_launch();
        }
    }
    forceGeneric &= 1;
    {
        const uint32_t maxRows = 65535u * 16u; /* gridDim.y limit: 8 thread rows x 2 luma rows per block row */
        if (job->rowEnd - job->rowBegin > maxRows) {
            int err = 0;
            for (uint32_t r0 = job->rowBegin; r0 < job->rowEnd && !err; r0 += maxRows) {
                AbR2YJob S = *job;
                S.rowBegin = r0;
                S.rowEnd = (job->rowEnd - r0 > maxRows) ? r0 + maxRows : job->rowEnd;
                err = abLaunchR2Y(&S, stream, forceGeneric ? 1 : 4, kernelName); /* 4: skip the specialised kernels, keep the vector block kernel */
            }
            return err;
        }
    }
    const dim3 block(32, 8);
    const dim3 grid = grid2d((job->width + 1) / 2, (job->rowEnd - job->rowBegin + 1) / 2, block);
#define AB_R2Y(RT, YT)                                                                      \
    do {                                                                                    \
        if (forceGeneric) r2y_block_kernel<RT, YT, false><<<grid, block, 0, st>>>(*job);    \
        else r2y_block_kernel<RT, YT, true><<<grid, block, 0, st>>>(*job);                  \
    } while (0)
    if (job->rgbBytes == 2) {
        if (job->yuvBytes == 2) AB_R2Y(uint16_t, uint16_t);
        else AB_R2Y(uint16_t, uint8_t);
    } else {
        if (job->yuvBytes == 2) AB_R2Y(uint8_t, uint16_t);
        else AB_R2Y(uint8_t, uint8_t);
    }
#undef AB_R2Y
    if (kernelName) *kernelName = forceGeneric ? "r2y_block_generic" : "r2y_block_vec";
    return finish_launch();
}

static const uint32_t kMaxRows2D = 65535u * 4u; /* gridDim.y limit of the (64, 4)-thread 2-D launches below */

int abLaunchAlpha(const AbAlphaJob * job, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (job->width == 0 || job->height == 0) return 0;
    if (job->height > kMaxRows2D) { /* row slices */
        int err = 0;
        for (uint32_t r0 = 0; r0 < job->height && !err; r0 += kMaxRows2D) {
            AbAlphaJob S = *job;
            S.height = (job->height - r0 > kMaxRows2D) ? kMaxRows2D : job->height - r0;
            if (S.src) S.src += (int64_t)r0 * S.srcRB;
            S.dst += (int64_t)r0 * S.dstRB;
            err = abLaunchAlpha(&S, stream);
        }
        return err;
    }
    const dim3 block(64, 4);
    const dim3 grid = grid2d(job->width, job->height, block);
    const bool s16 = job->src && job->srcDepth > 8, d16 = job->dstDepth > 8;
    if (s16 && d16) alpha_kernel<uint16_t, uint16_t><<<grid, block, 0, st>>>(*job);
    else if (s16) alpha_kernel<uint16_t, uint8_t><<<grid, block, 0, st>>>(*job);
    else if (d16) alpha_kernel<uint8_t, uint16_t><<<grid, block, 0, st>>>(*job);
    else alpha_kernel<uint8_t, uint8_t><<<grid, block, 0, st>>>(*job);
    return finish_launch();
}

int abLaunchPremultiply(const AbRgbPassJob * job, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (job->width == 0 || job->height == 0) return 0;
    if (job->height > kMaxRows2D) { /* row slices */
        int err = 0;
        for (uint32_t r0 = 0; r0 < job->height && !err; r0 += kMaxRows2D) {
            AbRgbPassJob S = *job;
            S.height = (job->height - r0 > kMaxRows2D) ? kMaxRows2D : job->height - r0;
            S.pixels += (int64_t)r0 * S.rowBytes;
            err = abLaunchPremultiply(&S, stream);
        }
        return err;
    }
    const dim3 block(64, 4);
    const dim3 grid = grid2d(job->width, job->height, block);
    if (job->depth > 8) premultiply_kernel<uint16_t><<<grid, block, 0, st>>>(*job);
    else premultiply_kernel<uint8_t><<<grid, block, 0, st>>>(*job);
    return finish_launch();
}

int abLaunchToF16(const AbRgbPassJob * job, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (job->width == 0 || job->height == 0) return 0;
    if (job->height > kMaxRows2D) { /* row slices */
        int err = 0;
        for (uint32_t r0 = 0; r0 < job->height && !err; r0 += kMaxRows2D) {
            AbRgbPassJob S = *job;
            S.height = (job->height - r0 > kMaxRows2D) ? kMaxRows2D : job->height - r0;
            S.pixels += (int64_t)r0 * S.rowBytes;
            err = abLaunchToF16(&S, stream);
        }
        return err;
    }
    const dim3 block(64, 4);
    f16_kernel<<<grid2d(job->width * job->nChan, job->height, block), block, 0, st>>>(*job);
    return finish_launch();
}

int abLaunchFill16(uint8_t * dst, size_t count16, uint32_t value, void * stream)
{
    if (count16 == 0) return 0;
    fill_kernel<uint16_t><<<(unsigned)((count16 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<uint16_t *>(dst), count16, (uint16_t)value);
    return finish_launch();
}

int abLaunchFill8(uint8_t * dst, size_t count, uint32_t value, void * stream)
{
    if (count == 0) return 0;
    fill_kernel<uint8_t><<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dst, count, (uint8_t)value);
    return finish_launch();
}

int abLaunchLimitedToFull(uint8_t * plane, int64_t rowBytes, uint32_t width, uint32_t height, uint32_t depth, void * stream)
{
    if (width == 0 || height == 0) return 0;
    if (height > kMaxRows2D) {
        int err = 0;
        for (uint32_t r0 = 0; r0 < height && !err; r0 += kMaxRows2D)
            err = abLaunchLimitedToFull(plane + (int64_t)r0 * rowBytes, rowBytes, width, (height - r0 > kMaxRows2D) ? kMaxRows2D : height - r0, depth, stream);
        return err;
    }
    const dim3 block(64, 4);
    const dim3 grid = grid2d(width, height, block);
    if (depth > 8) limited_to_full_kernel<uint16_t><<<grid, block, 0, (cudaStream_t)stream>>>(plane, rowBytes, width, height, depth);
    else limited_to_full_kernel<uint8_t><<<grid, block, 0, (cudaStream_t)stream>>>(plane, rowBytes, width, height, depth);
    return finish_launch();
}

int abLaunchCopyPad(const uint8_t * src, int64_t srcRB, uint32_t srcW, uint32_t srcH, uint8_t * dst, int64_t dstRB, uint32_t dstW,
                    uint32_t dstH, uint32_t bytesPerSample, void * stream)
{
    if (dstW == 0 || dstH == 0 || srcW == 0 || srcH == 0) return 0;
    const dim3 block(64, 4);
    const dim3 grid = grid2d(dstW, dstH, block);
    if (bytesPerSample == 2) copy_pad_kernel<uint16_t><<<grid, block, 0, (cudaStream_t)stream>>>(src, srcRB, srcW, srcH, dst, dstRB, dstW, dstH);
    else copy_pad_kernel<uint8_t><<<grid, block, 0, (cudaStream_t)stream>>>(src, srcRB, srcW, srcH, dst, dstRB, dstW, dstH);
    return finish_launch();
}

int abDivSelfTest(float d, unsigned long long * mismatches, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long * dev = nullptr;
    int err = (int)cudaMalloc((void **)&dev, sizeof(*dev));
    if (err) return err;
    err = (int)cudaMemsetAsync(dev, 0, sizeof(*dev), st);
    if (!err) {
        div_selftest_kernel<<<148 * 16, 256, 0, st>>>(d, 1.0f / d, dev);
        err = finish_launch();
    }
    if (!err) err = (int)cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, st);
    if (!err) err = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    return err;
}

int abIntDivSelfTest(const AbDiv * dv, unsigned long long * mismatches, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long * dev = nullptr;
    int err = (int)cudaMalloc((void **)&dev, sizeof(*dev));
    if (err) return err;
    err = (int)cudaMemsetAsync(dev, 0, sizeof(*dev), st);
    if (!err) {
        int_div_selftest_kernel<<<131072 / 256, 256, 0, st>>>(*dv, dev);
        err = finish_launch();
    }
    if (!err) err = (int)cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, st);
    if (!err) err = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    return err;
}

int abDiv2SelfTest(const AbDiv * dv, unsigned long long * mismatches, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long * dev = nullptr;
    int err = (int)cudaMalloc((void **)&dev, sizeof(*dev));
    if (err) return err;
    err = (int)cudaMemsetAsync(dev, 0, sizeof(*dev), st);
    if (!err) {
        div2_selftest_kernel<<<148 * 16, 256, 0, st>>>(*dv, dev);
        err = finish_launch();
    }
    if (!err) err = (int)cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, st);
    if (!err) err = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    return err;
}

int abAlphaDivSelfTest(uint32_t max, int mode, unsigned long long * mismatches, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long * dev = nullptr;
    int err = (int)cudaMalloc((void **)&dev, sizeof(*dev));
    if (err) return err;
    err = (int)cudaMemsetAsync(dev, 0, sizeof(*dev), st);
    if (!err) {
        alpha_div_selftest_kernel<<<148 * 16, 256, 0, st>>>(max, mode, dev);
        err = finish_launch();
    }
    if (!err) err = (int)cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, st);
    if (!err) err = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    return err;
}

int abPremulFmaSelfTest(uint32_t max, float r, float k, unsigned long long * mismatches, void * stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long * dev = nullptr;
    int err = (int)cudaMalloc((void **)&dev, sizeof(*dev));
    if (err) return err;
    err = (int)cudaMemsetAsync(dev, 0, sizeof(*dev), st);
    if (!err) {
        premul_fma_selftest_kernel<<<148 * 16, 256, 0, st>>>(max, r, k, dev);
        err = finish_launch();
    }
    if (!err) err = (int)cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, st);
    if (!err) err = (int)cudaStreamSynchronize(st);
    cudaFree(dev);
    return err;
}

uint64_t abLaunchCount(void) { return g_launches; }
void abLaunchCountReset(void) { g_launches = 0; }

} // extern "C"
