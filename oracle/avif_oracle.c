/*
 * avif_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of libavif's built-in (libyuv OFF)
 * pixel-reformat + alpha path, used as the parity checker for the CUDA implementation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this.  The product library (libavif_b200/csrc) never links, calls or falls back to anything here.
 *
 * Parity status: PINNED.  This restatement is checked (tests/test_oracle_vs_reference.py, run where
 * /root/reference exists) against the reference itself compiled from its own sources
 * (oracle/Makefile -> oracle/_ref/libavif_ref.so), and (tests/test_oracle_golden.py, runs anywhere)
 * against the reference's single-pixel known answers and whole-frame FNV-1a-64 hashes recorded in
 * SURVEY.md Appendix C plus the fixtures under tests/golden/ generated from the compiled reference.
 *
 * Arithmetic contract: IEEE-754 binary32, one rounding per C operator, in
 * source order, no FMA contraction (compile with -ffp-contract=off, no -march=native/-ffast-math).
 *
 * All "ref:" citations are file:line in the reference tree (/root/reference).
 */
#include "avif_b200.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */

static float roundHalfUp(float v) { return floorf(v + 0.5f); } /* ref: src/utils.c:11-14 */

static int fmtHasAlpha(avifRGBFormat f) /* ref: src/avif.c:675-679 */
{
    return f != AVIF_RGB_FORMAT_RGB && f != AVIF_RGB_FORMAT_BGR && f != AVIF_RGB_FORMAT_RGB_565 && f != AVIF_RGB_FORMAT_GRAY;
}
static int fmtIsGray(avifRGBFormat f) /* ref: src/avif.c:670-673 */
{
    return f == AVIF_RGB_FORMAT_GRAY || f == AVIF_RGB_FORMAT_GRAYA || f == AVIF_RGB_FORMAT_AGRAY;
}
static uint32_t fmtChannels(avifRGBFormat f) /* ref: src/avif.c:681-690 */
{
    if (f == AVIF_RGB_FORMAT_GRAY) return 1;
    if (f == AVIF_RGB_FORMAT_GRAYA || f == AVIF_RGB_FORMAT_AGRAY) return 2;
    return fmtHasAlpha(f) ? 4 : 3;
}

/* Everything the loops need, derived once.  ref: avifReformatState, include/avif/internal.h:298-341 */
typedef struct OraclePlan {
    /* rgb side (ref: src/reformat.c:32-117) */
    uint32_t rgbChanBytes, rgbPixBytes, offR, offG, offB, offA, offGray;
    int rgbMax;
    float rgbMaxF;
    /* yuv side (ref: src/reformat.c:119-159) */
    float kr, kg, kb;
    uint32_t yuvChanBytes;
    int yuvMax;
    float biasY, biasUV, rangeY, rangeUV;
    int shiftX, shiftY;
    int mode; /* 0 coefficients, 1 identity, 2 YCgCo, 3 YCgCo-Re, 4 YCgCo-Ro (ref: internal.h:281-288) */
} OraclePlan;

/* ref: src/colr.c:517-542 (H.273 eq. 32-37), float arithmetic in this exact association */
static void chromaDerivedCoeffs(avifColorPrimaries cp, float * kr, float * kb)
{
    static const struct { uint16_t id; float p[8]; } T[] = {
        /* ref: src/colr.c:16-28, {rX,rY,gX,gY,bX,bY,wX,wY} */
        { 1, { 0.64f, 0.33f, 0.3f, 0.6f, 0.15f, 0.06f, 0.3127f, 0.329f } },
        { 4, { 0.67f, 0.33f, 0.21f, 0.71f, 0.14f, 0.08f, 0.310f, 0.316f } },
        { 5, { 0.64f, 0.33f, 0.29f, 0.60f, 0.15f, 0.06f, 0.3127f, 0.3290f } },
        { 6, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { 7, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { 8, { 0.681f, 0.319f, 0.243f, 0.692f, 0.145f, 0.049f, 0.310f, 0.316f } },
        { 9, { 0.708f, 0.292f, 0.170f, 0.797f, 0.131f, 0.046f, 0.3127f, 0.3290f } },
        { 10, { 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.3333f, 0.3333f } },
        { 11, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.314f, 0.351f } },
        { 12, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.3127f, 0.3290f } },
        { 22, { 0.630f, 0.340f, 0.295f, 0.605f, 0.155f, 0.077f, 0.3127f, 0.3290f } },
    };
    const float * p = T[0].p; /* unknown primaries fall back to BT.709's (ref: src/colr.c:41-42) */
    for (size_t i = 0; i < sizeof(T) / sizeof(T[0]); ++i) {
        if (T[i].id == cp) {
            p = T[i].p;
            break;
        }
    }
    const float rX = p[0], rY = p[1], gX = p[2], gY = p[3], bX = p[4], bY = p[5], wX = p[6], wY = p[7];
    const float rZ = 1.0f - (rX + rY);
    const float gZ = 1.0f - (gX + gY);
    const float bZ = 1.0f - (bX + bY);
    const float wZ = 1.0f - (wX + wY);
    const float den = wY * (rX * (gY * bZ - bY * gZ) + gX * (bY * rZ - rY * bZ) + bX * (rY * gZ - gY * rZ));
    *kr = (rY * (wX * (gY * bZ - bY * gZ) + wY * (bX * gZ - gX * bZ) + wZ * (gX * bY - bX * gY))) / den;
    *kb = (bY * (wX * (rY * gZ - gY * rZ) + wY * (gX * rZ - rX * gZ) + wZ * (rX * gY - gX * rY))) / den;
}

/* ref: src/colr.c:137-186 */
ORACLE_API void oracleCalcYUVCoefficients(const avifImage * image, float * outR, float * outG, float * outB)
{
    float kr = 0.299f, kb = 0.114f; /* default = BT.601 (ref: src/colr.c:173-175) */
    switch (image->matrixCoefficients) {
        case AVIF_MATRIX_COEFFICIENTS_BT709: kr = 0.2126f; kb = 0.0722f; break;
        case AVIF_MATRIX_COEFFICIENTS_FCC: kr = 0.30f; kb = 0.11f; break;
        case AVIF_MATRIX_COEFFICIENTS_BT470BG:
        case AVIF_MATRIX_COEFFICIENTS_BT601: kr = 0.299f; kb = 0.114f; break;
        case AVIF_MATRIX_COEFFICIENTS_SMPTE240: kr = 0.212f; kb = 0.087f; break;
        case AVIF_MATRIX_COEFFICIENTS_BT2020_NCL: kr = 0.2627f; kb = 0.0593f; break;
        case AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_NCL: chromaDerivedCoeffs(image->colorPrimaries, &kr, &kb); break;
        default: break;
    }
    *outR = kr;
    *outB = kb;
    *outG = 1.0f - kr - kb;
}

/* ref: src/reformat.c:161-194 (+ :32-159).  Returns 0 when the combination is rejected. */
static int makePlan(const avifImage * image, const avifRGBImage * rgb, OraclePlan * pl)
{
    const unsigned mc = image->matrixCoefficients;
    if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE && image->depth - 2 != rgb->depth) return 0; /* ref: :163-170 */
    if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO && image->depth - 1 != rgb->depth) return 0;

    /* rgb (ref: :34-42) */
    if (!(rgb->depth == 8 || rgb->depth == 10 || rgb->depth == 12 || rgb->depth == 16)) return 0;
    if (rgb->isFloat && rgb->depth != 16) return 0;
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565 && rgb->depth != 8) return 0;
    if ((int)rgb->format < AVIF_RGB_FORMAT_RGB || rgb->format >= AVIF_RGB_FORMAT_COUNT) return 0;
    const uint32_t cb = rgb->depth > 8 ? 2 : 1;
    pl->rgbChanBytes = cb;
    pl->rgbPixBytes = rgb->format == AVIF_RGB_FORMAT_RGB_565 ? 2 : fmtChannels(rgb->format) * cb;
    pl->offR = pl->offG = pl->offB = pl->offA = pl->offGray = 0;
    switch (rgb->format) { /* ref: :53-111 */
        case AVIF_RGB_FORMAT_RGB: pl->offG = cb; pl->offB = 2 * cb; break;
        case AVIF_RGB_FORMAT_RGBA: pl->offG = cb; pl->offB = 2 * cb; pl->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ARGB: pl->offR = cb; pl->offG = 2 * cb; pl->offB = 3 * cb; break;
        case AVIF_RGB_FORMAT_BGR: pl->offG = cb; pl->offR = 2 * cb; break;
        case AVIF_RGB_FORMAT_BGRA: pl->offG = cb; pl->offR = 2 * cb; pl->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ABGR: pl->offB = cb; pl->offG = 2 * cb; pl->offR = 3 * cb; break;
        case AVIF_RGB_FORMAT_GRAYA: pl->offA = cb; break;
        case AVIF_RGB_FORMAT_AGRAY: pl->offGray = cb; break;
        default: break;
    }
    pl->rgbMax = (1 << rgb->depth) - 1;
    pl->rgbMaxF = (float)pl->rgbMax;

    /* yuv (ref: :121-144) */
    if (!(image->depth == 8 || image->depth == 10 || image->depth == 12 || image->depth == 16)) return 0;
    if (image->yuvFormat < AVIF_PIXEL_FORMAT_YUV444 || image->yuvFormat >= AVIF_PIXEL_FORMAT_COUNT) return 0;
    if (image->yuvRange != AVIF_RANGE_LIMITED && image->yuvRange != AVIF_RANGE_FULL) return 0;
    const int isYCgCoFamily = mc == AVIF_MATRIX_COEFFICIENTS_YCGCO || mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE ||
                              mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO;
    if (mc == 3 || (isYCgCoFamily && image->yuvRange == AVIF_RANGE_LIMITED) || mc == AVIF_MATRIX_COEFFICIENTS_BT2020_CL ||
        mc == AVIF_MATRIX_COEFFICIENTS_SMPTE2085 || mc == AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_CL ||
        mc == AVIF_MATRIX_COEFFICIENTS_ICTCP || mc >= AVIF_MATRIX_COEFFICIENTS_LAST) {
        return 0;
    }
    if (mc == AVIF_MATRIX_COEFFICIENTS_IDENTITY && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV444 &&
        image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400) {
        return 0;
    }
    pl->shiftX = (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444) ? 0 : 1; /* ref: src/avif.c:39-73 */
    pl->shiftY = (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 || image->yuvFormat == AVIF_PIXEL_FORMAT_YUV422) ? 0 : 1;
    oracleCalcYUVCoefficients(image, &pl->kr, &pl->kg, &pl->kb);
    pl->yuvChanBytes = image->depth > 8 ? 2 : 1;
    pl->yuvMax = (1 << image->depth) - 1;
    const int lim = image->yuvRange == AVIF_RANGE_LIMITED;
    pl->biasY = lim ? (float)(16 << (image->depth - 8)) : 0.0f; /* ref: :153-156 */
    pl->biasUV = (float)(1 << (image->depth - 1));
    pl->rangeY = (float)(lim ? (219 << (image->depth - 8)) : pl->yuvMax);
    pl->rangeUV = (float)(lim ? (224 << (image->depth - 8)) : pl->yuvMax);

    pl->mode = 0; /* ref: :175-191 */
    if (mc == AVIF_MATRIX_COEFFICIENTS_IDENTITY) pl->mode = 1;
    else if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO) pl->mode = 2;
    else if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE) pl->mode = 3;
    else if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO) pl->mode = 4;
    if (pl->mode != 0) pl->kr = pl->kg = pl->kb = 0.0f;
    return 1;
}

static uint32_t loadSample(const uint8_t * p, uint32_t chanBytes)
{
    if (chanBytes > 1) {
        uint16_t v;
        memcpy(&v, p, 2);
        return v;
    }
    return *p;
}
static void storeSample(uint8_t * p, uint32_t chanBytes, uint32_t v)
{
    if (chanBytes > 1) {
        const uint16_t s = (uint16_t)v;
        memcpy(p, &s, 2);
    } else {
        *p = (uint8_t)v;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* alpha.c                                                                                     */

ORACLE_API void oracleFillAlpha(const avifAlphaParams * pr) /* ref: src/alpha.c:9-35 */
{
    const uint32_t cb = pr->dstDepth > 8 ? 2 : 1;
    const uint32_t opaque = pr->dstDepth > 8 ? (uint32_t)((1 << pr->dstDepth) - 1) : 255u;
    for (uint32_t j = 0; j < pr->height; ++j) {
        uint8_t * d = pr->dstPlane + pr->dstOffsetBytes + (size_t)j * pr->dstRowBytes;
        for (uint32_t i = 0; i < pr->width; ++i, d += pr->dstPixelBytes) storeSample(d, cb, opaque);
    }
}

ORACLE_API void oracleReformatAlpha(const avifAlphaParams * pr) /* ref: src/alpha.c:37-149 */
{
    const int srcMax = (1 << pr->srcDepth) - 1, dstMax = (1 << pr->dstDepth) - 1;
    const float srcMaxF = (float)srcMax, dstMaxF = (float)dstMax;
    const uint32_t scb = pr->srcDepth > 8 ? 2 : 1, dcb = pr->dstDepth > 8 ? 2 : 1;
    for (uint32_t j = 0; j < pr->height; ++j) {
        const uint8_t * s = pr->srcPlane + pr->srcOffsetBytes + (size_t)j * pr->srcRowBytes;
        uint8_t * d = pr->dstPlane + pr->dstOffsetBytes + (size_t)j * pr->dstRowBytes;
        for (uint32_t i = 0; i < pr->width; ++i, s += pr->srcPixelBytes, d += pr->dstPixelBytes) {
            const int a = (int)loadSample(s, scb);
            if (pr->srcDepth == pr->dstDepth) {
                storeSample(d, dcb, (uint32_t)a); /* plain copy, no clamp (ref: :44-79) */
            } else {
                const float alphaF = (float)a / srcMaxF; /* ref: :93-97 */
                int out = (int)(0.5f + (alphaF * dstMaxF));
                if (out < 0) out = 0;
                if (out > dstMax) out = dstMax;
                storeSample(d, dcb, (uint32_t)out);
            }
        }
    }
}

/* Position of the alpha channel and of the (up to three) colour channels inside one pixel, in channel units.
 * ref: the eight per-format loops of src/alpha.c:173-333 / :360-532 collapse to this table. */
static int premulLayout(avifRGBFormat f, int * aIdx, int cIdx[3], int * nColor, int * nChan)
{
    switch (f) {
        case AVIF_RGB_FORMAT_RGBA:
        case AVIF_RGB_FORMAT_BGRA: *aIdx = 3; cIdx[0] = 0; cIdx[1] = 1; cIdx[2] = 2; *nColor = 3; *nChan = 4; return 1;
        case AVIF_RGB_FORMAT_ARGB:
        case AVIF_RGB_FORMAT_ABGR: *aIdx = 0; cIdx[0] = 1; cIdx[1] = 2; cIdx[2] = 3; *nColor = 3; *nChan = 4; return 1;
        case AVIF_RGB_FORMAT_GRAYA: *aIdx = 1; cIdx[0] = 0; *nColor = 1; *nChan = 2; return 1;
        case AVIF_RGB_FORMAT_AGRAY: *aIdx = 0; cIdx[0] = 1; *nColor = 1; *nChan = 2; return 1;
        default: return 0;
    }
}

static avifResult premulImpl(avifRGBImage * rgb, int unmultiply)
{
    if (!rgb->pixels || !rgb->rowBytes) return AVIF_RESULT_REFORMAT_FAILED; /* ref: :154-156, :341-343 */
    if (!fmtHasAlpha(rgb->format)) return unmultiply ? AVIF_RESULT_REFORMAT_FAILED : AVIF_RESULT_INVALID_ARGUMENT; /* :159-161, :346-348 */
    int aIdx, cIdx[3], nColor, nChan;
    if (!premulLayout(rgb->format, &aIdx, cIdx, &nColor, &nChan)) return AVIF_RESULT_NOT_IMPLEMENTED;
    const uint32_t max = (1u << rgb->depth) - 1;
    const float maxF = (float)max;
    const uint32_t cb = rgb->depth > 8 ? 2 : 1;
    for (uint32_t j = 0; j < rgb->height; ++j) {
        uint8_t * px = rgb->pixels + (size_t)j * rgb->rowBytes;
        for (uint32_t i = 0; i < rgb->width; ++i, px += (size_t)nChan * cb) {
            const uint32_t a = loadSample(px + aIdx * cb, cb);
            if (a >= max) continue; /* opaque: untouched (ref: :180, :263, :367) */
            for (int c = 0; c < nColor; ++c) {
                uint8_t * cp = px + cIdx[c] * cb;
                if (a == 0) {
                    storeSample(cp, cb, 0);
                } else if (!unmultiply) {
                    const float v = roundHalfUp((float)loadSample(cp, cb) * (float)a / maxF); /* ref: :189 */
                    storeSample(cp, cb, (uint32_t)v);
                } else {
                    float v = roundHalfUp((float)loadSample(cp, cb) * maxF / (float)a); /* ref: :375-378 */
                    if (maxF < v) v = maxF;
                    storeSample(cp, cb, (uint32_t)v);
                }
            }
        }
    }
    return AVIF_RESULT_OK;
}

ORACLE_API avifResult oracleRGBImagePremultiplyAlpha(avifRGBImage * rgb) { return premulImpl(rgb, 0); }     /* ref: src/alpha.c:151 */
ORACLE_API avifResult oracleRGBImageUnpremultiplyAlpha(avifRGBImage * rgb) { return premulImpl(rgb, 1); }   /* ref: src/alpha.c:338 */

/* ------------------------------------------------------------------------------------------ */
/* limited <-> full integer helpers (ref: src/reformat.c:1750-1840)                            */

static int lim2full(int v, int lo, int hi, int full)
{
    v = (((v - lo) * full) + ((hi - lo) / 2)) / (hi - lo);
    return v < 0 ? 0 : (v > full ? full : v);
}
static int full2lim(int v, int lo, int hi, int full)
{
    v = (((v * (hi - lo)) + (full / 2)) / full) + lo;
    return v < lo ? lo : (v > hi ? hi : v);
}
static int depthIdx(uint32_t depth) { return depth == 8 ? 0 : depth == 10 ? 1 : depth == 12 ? 2 : -1; }
static const int kLo[3] = { 16, 64, 256 }, kHiY[3] = { 235, 940, 3760 }, kHiUV[3] = { 240, 960, 3840 }, kFull[3] = { 255, 1023, 4095 };

ORACLE_API int oracleLimitedToFullY(uint32_t depth, int v) { const int k = depthIdx(depth); return k < 0 ? v : lim2full(v, kLo[k], kHiY[k], kFull[k]); }
ORACLE_API int oracleLimitedToFullUV(uint32_t depth, int v) { const int k = depthIdx(depth); return k < 0 ? v : lim2full(v, kLo[k], kHiUV[k], kFull[k]); }
ORACLE_API int oracleFullToLimitedY(uint32_t depth, int v) { const int k = depthIdx(depth); return k < 0 ? v : full2lim(v, kLo[k], kHiY[k], kFull[k]); }
ORACLE_API int oracleFullToLimitedUV(uint32_t depth, int v) { const int k = depthIdx(depth); return k < 0 ? v : full2lim(v, kLo[k], kHiUV[k], kFull[k]); }

/* ------------------------------------------------------------------------------------------ */
/* YUV -> RGB                                                                                  */

static float clamp01(float v) { return v < 0.0f ? 0.0f : (1.0f < v ? 1.0f : v); }

/* code point -> normalised float; identical to the reference's look-up-table entries (ref: src/reformat.c:583, :598) */
static float normY(const OraclePlan * pl, uint32_t cp) { return ((float)cp - pl->biasY) / pl->rangeY; }
static float normUV(const OraclePlan * pl, uint32_t cp)
{
    if (pl->mode == 1) return ((float)cp - pl->biasY) / pl->rangeY; /* identity shares the luma table (ref: :587-589) */
    return ((float)cp - pl->biasUV) / pl->rangeUV;
}

/* sample fetch with the "clamp garbage high bits" rule for 16-bit containers (ref: :712, :727, :821, :1005-1007) */
static uint32_t fetchYUV(const uint8_t * row, uint32_t x, const OraclePlan * pl)
{
    if (pl->yuvChanBytes == 1) return row[x];
    uint32_t v = loadSample(row + 2 * (size_t)x, 2);
    return v > (uint32_t)pl->yuvMax ? (uint32_t)pl->yuvMax : v;
}

static void yuvToRgbMatrix(const OraclePlan * pl, float Y, float Cb, float Cr, float * R, float * G, float * B)
{
    const float kr = pl->kr, kg = pl->kg, kb = pl->kb;
    *R = Y + (2 * (1 - kr)) * Cr; /* ref: :874-876 == :1014-1016 */
    *B = Y + (2 * (1 - kb)) * Cb;
    *G = Y - ((2 * ((kr * (1 - kr) * Cr) + (kb * (1 - kb) * Cb))) / kg);
}

static void storeRgbPixel(const avifRGBImage * rgb, const OraclePlan * pl, uint8_t * px, float Rc, float Gc, float Bc)
{
    const float m = pl->rgbMaxF;
    if (rgb->depth == 8) { /* ref: :951-957, :621-633 */
        const uint8_t R = (uint8_t)(0.5f + (Rc * m)), G = (uint8_t)(0.5f + (Gc * m)), B = (uint8_t)(0.5f + (Bc * m));
        if (rgb->format == AVIF_RGB_FORMAT_RGB_565) {
            const uint16_t v = (uint16_t)((B >> 3) | ((G >> 2) << 5) | ((R >> 3) << 11)); /* ref: :619 */
            memcpy(px + pl->offR, &v, 2);
        } else {
            px[pl->offR] = R;
            px[pl->offG] = G;
            px[pl->offB] = B;
        }
    } else { /* ref: :959-961 */
        storeSample(px + pl->offR, 2, (uint16_t)(0.5f + (Rc * m)));
        storeSample(px + pl->offG, 2, (uint16_t)(0.5f + (Gc * m)));
        storeSample(px + pl->offB, 2, (uint16_t)(0.5f + (Bc * m)));
    }
}

/* the 9 "fast" loops (ref: src/reformat.c:980-1407): nearest chroma, no alpha math, colour or mono */
static void yuvToRgbFast(const avifImage * image, avifRGBImage * rgb, const OraclePlan * pl, int hasColor)
{
    for (uint32_t j = 0; j < image->height; ++j) {
        const uint32_t uvJ = j >> pl->shiftY;
        const uint8_t * rowY = image->yuvPlanes[0] + (size_t)j * image->yuvRowBytes[0];
        const uint8_t * rowU = hasColor ? image->yuvPlanes[1] + (size_t)uvJ * image->yuvRowBytes[1] : NULL;
        const uint8_t * rowV = hasColor ? image->yuvPlanes[2] + (size_t)uvJ * image->yuvRowBytes[2] : NULL;
        uint8_t * px = rgb->pixels + (size_t)j * rgb->rowBytes;
        for (uint32_t i = 0; i < image->width; ++i, px += pl->rgbPixBytes) {
            const float Y = normY(pl, fetchYUV(rowY, i, pl));
            float Cb = 0.0f, Cr = 0.0f; /* mono loops feed zeros through the same matrix (ref: :1057-1062) */
            if (hasColor) {
                const uint32_t uvI = i >> pl->shiftX;
                Cb = normUV(pl, fetchYUV(rowU, uvI, pl));
                Cr = normUV(pl, fetchYUV(rowV, uvI, pl));
            }
            float R, G, B;
            yuvToRgbMatrix(pl, Y, Cb, Cr, &R, &G, &B);
            storeRgbPixel(rgb, pl, px, clamp01(R), clamp01(G), clamp01(B));
        }
    }
}

/* ref: src/reformat.c:1278-1309 */
static void identity8Fast(const avifImage * image, avifRGBImage * rgb, const OraclePlan * pl)
{
    for (uint32_t j = 0; j < image->height; ++j) {
        const uint8_t * rowY = image->yuvPlanes[0] + (size_t)j * image->yuvRowBytes[0];
        const uint8_t * rowU = image->yuvPlanes[1] + (size_t)j * image->yuvRowBytes[1];
        const uint8_t * rowV = image->yuvPlanes[2] + (size_t)j * image->yuvRowBytes[2];
        uint8_t * px = rgb->pixels + (size_t)j * rgb->rowBytes;
        for (uint32_t i = 0; i < image->width; ++i, px += pl->rgbPixBytes) {
            const uint8_t R = rowV[i], G = rowY[i], B = rowU[i];
            if (rgb->format == AVIF_RGB_FORMAT_RGB_565) {
                const uint16_t v = (uint16_t)((B >> 3) | ((G >> 2) << 5) | ((R >> 3) << 11));
                memcpy(px + pl->offR, &v, 2);
            } else {
                px[pl->offR] = R;
                px[pl->offG] = G;
                px[pl->offB] = B;
            }
        }
    }
}

/* the general loop (ref: src/reformat.c:650-978); alphaMode: 0 no-op, 1 multiply, 2 unmultiply (in normalised float) */
static void yuvToRgbSlow(const avifImage * image, avifRGBImage * rgb, const OraclePlan * pl, int alphaMode)
{
    const uint8_t * uPlane = image->yuvPlanes[1];
    const uint8_t * vPlane = image->yuvPlanes[2];
    const int yuvHasColor = uPlane && vPlane && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400; /* ref: :676 */
    const int rgbHasColor = !fmtIsGray(rgb->format);
    const int nearest = rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_FASTEST || rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_NEAREST;
    const float m = pl->rgbMaxF;
    const ptrdiff_t cb = (ptrdiff_t)pl->yuvChanBytes;

    for (uint32_t j = 0; j < image->height; ++j) {
        const uint32_t uvJ = yuvHasColor ? (j >> pl->shiftY) : 0;
        const uint8_t * rowY = image->yuvPlanes[0] + (size_t)j * image->yuvRowBytes[0];
        const uint8_t * rowA = image->alphaPlane ? image->alphaPlane + (size_t)j * image->alphaRowBytes : NULL;
        uint8_t * px = rgb->pixels + (size_t)j * rgb->rowBytes;

        for (uint32_t i = 0; i < image->width; ++i, px += pl->rgbPixBytes) {
            const uint32_t unormY = fetchYUV(rowY, i, pl);
            const float Y = normY(pl, unormY);
            float Cb = 0.5f, Cr = 0.5f;
            if (yuvHasColor) {
                const uint32_t uvI = i >> pl->shiftX;
                const uint8_t * rowU = uPlane + (size_t)uvJ * image->yuvRowBytes[1];
                const uint8_t * rowV = vPlane + (size_t)uvJ * image->yuvRowBytes[2];
                if (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444) {
                    Cb = normUV(pl, fetchYUV(rowU, uvI, pl));
                    Cr = normUV(pl, fetchYUV(rowV, uvI, pl));
                } else {
                    /* neighbour selection with edge replication (ref: :766-795) */
                    ptrdiff_t adjCol, adjRowU, adjRowV;
                    if (i == 0 || (i == image->width - 1 && (i % 2) != 0)) adjCol = 0;
                    else adjCol = (i % 2) != 0 ? cb : -cb;
                    if (j == 0 || (j == image->height - 1 && (j % 2) != 0) || image->yuvFormat == AVIF_PIXEL_FORMAT_YUV422) {
                        adjRowU = adjRowV = 0;
                    } else if ((j % 2) != 0) {
                        adjRowU = (ptrdiff_t)image->yuvRowBytes[1];
                        adjRowV = (ptrdiff_t)image->yuvRowBytes[2];
                    } else {
                        adjRowU = -(ptrdiff_t)image->yuvRowBytes[1];
                        adjRowV = -(ptrdiff_t)image->yuvRowBytes[2];
                    }
                    const uint8_t * pU = rowU + (size_t)uvI * cb;
                    const uint8_t * pV = rowV + (size_t)uvI * cb;
                    if (nearest) { /* ref: :827-831 */
                        Cb = normUV(pl, fetchYUV(pU, 0, pl));
                        Cr = normUV(pl, fetchYUV(pV, 0, pl));
                    } else { /* ref: :833-838: ((a*9/16 + b*3/16) + c*3/16) + d*1/16, all four terms always present */
                        const float u00 = normUV(pl, fetchYUV(pU, 0, pl)), u10 = normUV(pl, fetchYUV(pU + adjCol, 0, pl));
                        const float u01 = normUV(pl, fetchYUV(pU + adjRowU, 0, pl)), u11 = normUV(pl, fetchYUV(pU + adjCol + adjRowU, 0, pl));
                        const float v00 = normUV(pl, fetchYUV(pV, 0, pl)), v10 = normUV(pl, fetchYUV(pV + adjCol, 0, pl));
                        const float v01 = normUV(pl, fetchYUV(pV + adjRowV, 0, pl)), v11 = normUV(pl, fetchYUV(pV + adjCol + adjRowV, 0, pl));
                        Cb = (u00 * (9.0f / 16.0f)) + (u10 * (3.0f / 16.0f)) + (u01 * (3.0f / 16.0f)) + (u11 * (1.0f / 16.0f));
                        Cr = (v00 * (9.0f / 16.0f)) + (v10 * (3.0f / 16.0f)) + (v01 * (3.0f / 16.0f)) + (v11 * (1.0f / 16.0f));
                    }
                }
            }

            float Rc = 0.0f, Gc = 0.0f, Bc = 0.0f, grayc = 0.0f;
            if (rgbHasColor) {
                float R, G, B;
                if (!yuvHasColor) {
                    R = G = B = Y; /* ref: :878-884 */
                } else if (pl->mode == 1) {
                    G = Y; B = Cb; R = Cr; /* ref: :846-851 */
                } else if (pl->mode == 2) {
                    const float t = Y - Cb; /* ref: :852-858 */
                    G = Y + Cb;
                    B = t - Cr;
                    R = t + Cr;
                } else if (pl->mode == 3 || pl->mode == 4) { /* ref: :859-871, integer lifting */
                    const int YY = (int)unormY;
                    const int Cg = (int)roundHalfUp(Cb * (float)pl->yuvMax);
                    const int Co = (int)roundHalfUp(Cr * (float)pl->yuvMax);
                    const int t = YY - (Cg >> 1);
                    int g = t + Cg, b = t - (Co >> 1);
                    g = g < 0 ? 0 : (g > pl->rgbMax ? pl->rgbMax : g);
                    b = b < 0 ? 0 : (b > pl->rgbMax ? pl->rgbMax : b);
                    G = (float)g;
                    B = (float)b;
                    R = B + (float)Co;
                    R = R < 0 ? 0 : ((float)pl->rgbMax < R ? (float)pl->rgbMax : R);
                    G /= m;
                    B /= m;
                    R /= m;
                } else {
                    yuvToRgbMatrix(pl, Y, Cb, Cr, &R, &G, &B);
                }
                Rc = clamp01(R);
                Gc = clamp01(G);
                Bc = clamp01(B);
            } else {
                grayc = clamp01(Y); /* ref: :888-892 */
            }

            if (alphaMode != 0) { /* ref: :894-947 */
                const float A = (float)fetchYUV(rowA, i, pl) / ((float)pl->yuvMax);
                const float Ac = clamp01(A);
                if (Ac == 0.0f) {
                    Rc = Gc = Bc = grayc = 0.0f;
                } else if (Ac < 1.0f) {
                    if (alphaMode == 1) {
                        Rc *= Ac; Gc *= Ac; Bc *= Ac; grayc *= Ac;
                    } else {
                        Rc /= Ac; Gc /= Ac; Bc /= Ac; grayc /= Ac;
                        if (1.0f < Rc) Rc = 1.0f;
                        if (1.0f < Gc) Gc = 1.0f;
                        if (1.0f < Bc) Bc = 1.0f;
                        if (1.0f < grayc) grayc = 1.0f;
                    }
                }
            }

            if (rgbHasColor) {
                storeRgbPixel(rgb, pl, px, Rc, Gc, Bc);
            } else if (rgb->depth == 8) { /* ref: :967-972 */
                px[pl->offGray] = (uint8_t)(0.5f + (grayc * m));
            } else {
                storeSample(px + pl->offGray, 2, (uint16_t)(0.5f + (grayc * m)));
            }
        }
    }
}

/* ref: src/reformat.c:1409-1443; truncating float -> half via the 2^-112 multiplier trick */
static void rgbToF16(avifRGBImage * rgb)
{
    const size_t n = (size_t)rgb->width * fmtChannels(rgb->format);
    const float scale = 1.0f / ((1 << rgb->depth) - 1);
    const float multiplier = 1.9259299444e-34f * scale;
    for (uint32_t j = 0; j < rgb->height; ++j) {
        uint8_t * row = rgb->pixels + (size_t)j * ((rgb->rowBytes >> 1) << 1);
        for (size_t i = 0; i < n; ++i) {
            union { float f; uint32_t u; } cv;
            cv.f = (float)loadSample(row + 2 * i, 2) * multiplier;
            storeSample(row + 2 * i, 2, (uint16_t)(cv.u >> 13));
        }
    }
}

ORACLE_API avifResult oracleImageYUVToRGB(const avifImage * image, avifRGBImage * rgb) /* ref: src/reformat.c:1649-1748 + :1445-1593 */
{
    if (!image->yuvPlanes[0] || rgb->maxThreads < 0) return AVIF_RESULT_REFORMAT_FAILED;
    OraclePlan pl;
    if (!makePlan(image, rgb, &pl)) return AVIF_RESULT_REFORMAT_FAILED;

    const int rgbHasAlpha = fmtHasAlpha(rgb->format);
    int alphaMode = 0; /* ref: :1662-1677 */
    if (image->alphaPlane) {
        if (!rgbHasAlpha || rgb->ignoreAlpha) {
            if (!image->alphaPremultiplied) alphaMode = 1;
        } else if (!image->alphaPremultiplied && rgb->alphaPremultiplied) {
            alphaMode = 1;
        } else if (image->alphaPremultiplied && !rgb->alphaPremultiplied) {
            alphaMode = 2;
        }
    }
    /* (row-slab threading, ref :1679-1747, never changes results: omitted) */

    if (rgbHasAlpha && (!rgb->ignoreAlpha || alphaMode != 0)) { /* ref: :1449-1486; alpha goes first */
        avifAlphaParams pr;
        pr.width = rgb->width;
        pr.height = rgb->height;
        pr.dstDepth = rgb->depth;
        pr.dstPlane = rgb->pixels;
        pr.dstRowBytes = rgb->rowBytes;
        pr.dstOffsetBytes = pl.offA;
        pr.dstPixelBytes = pl.rgbPixBytes;
        if (image->alphaPlane && image->alphaRowBytes) {
            pr.srcDepth = image->depth;
            pr.srcPlane = image->alphaPlane;
            pr.srcRowBytes = image->alphaRowBytes;
            pr.srcOffsetBytes = 0;
            pr.srcPixelBytes = pl.yuvChanBytes;
            oracleReformatAlpha(&pr);
        } else {
            oracleFillAlpha(&pr);
        }
    }

    const int hasColor = image->yuvRowBytes[1] && image->yuvRowBytes[2] && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400; /* ref: :1496 */
    const int nearest = rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_FASTEST || rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_NEAREST;
    int usedFast = 0;
    if (!fmtIsGray(rgb->format) && (!hasColor || image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 || nearest) &&
        (alphaMode == 0 || rgbHasAlpha)) { /* ref: :1499-1502 */
        if (pl.mode == 1) {
            if (image->depth == 8 && rgb->depth == 8 && image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 && image->yuvRange == AVIF_RANGE_FULL) {
                identity8Fast(image, rgb, &pl);
                usedFast = 1;
            }
        } else if (pl.mode == 0) {
            yuvToRgbFast(image, rgb, &pl, hasColor);
            usedFast = 1;
        }
    }
    if (!usedFast) {
        yuvToRgbSlow(image, rgb, &pl, alphaMode);
        alphaMode = 0; /* handled in-loop (ref: :1565-1566) */
    }
    if (alphaMode == 1) { /* integer post-pass on the quantised pixels (ref: :1575-1585) */
        const avifResult r = oracleRGBImagePremultiplyAlpha(rgb);
        if (r != AVIF_RESULT_OK) return r;
    } else if (alphaMode == 2) {
        const avifResult r = oracleRGBImageUnpremultiplyAlpha(rgb);
        if (r != AVIF_RESULT_OK) return r;
    }
    if (rgb->isFloat) rgbToF16(rgb); /* ref: :1588-1590 */
    return AVIF_RESULT_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* RGB -> YUV                                                                                  */

/* ref: src/avif.c:431-489 (pre-existing planes are kept) */
static avifResult allocPlanes(avifImage * image, int withAlpha)
{
    if (image->width == 0 || image->height == 0 || image->depth == 0 || image->depth > 16) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t cs = image->depth > 8 ? 2 : 1;
    if (image->width > UINT32_MAX / cs) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t fullRowBytes = cs * image->width;
    if (image->height > PTRDIFF_MAX / fullRowBytes) return AVIF_RESULT_INVALID_ARGUMENT;
    const size_t fullSize = (size_t)fullRowBytes * image->height;
    if (image->yuvFormat != AVIF_PIXEL_FORMAT_NONE) {
        image->imageOwnsYUVPlanes = AVIF_TRUE;
        if (!image->yuvPlanes[0]) {
            image->yuvPlanes[0] = (uint8_t *)malloc(fullSize);
            if (!image->yuvPlanes[0]) return AVIF_RESULT_OUT_OF_MEMORY;
            image->yuvRowBytes[0] = fullRowBytes;
        }
        if (image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400) {
            const int sx = image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 ? 0 : 1;
            const int sy = image->yuvFormat == AVIF_PIXEL_FORMAT_YUV420 ? 1 : 0;
            const uint32_t w = (uint32_t)(((uint64_t)image->width + sx) >> sx), h = (uint32_t)(((uint64_t)image->height + sy) >> sy);
            for (int p = 1; p <= 2; ++p) {
                if (!image->yuvPlanes[p]) {
                    image->yuvPlanes[p] = (uint8_t *)malloc((size_t)cs * w * h);
                    if (!image->yuvPlanes[p]) return AVIF_RESULT_OUT_OF_MEMORY;
                    image->yuvRowBytes[p] = cs * w;
                }
            }
        }
    }
    if (withAlpha) {
        image->imageOwnsAlphaPlane = AVIF_TRUE;
        if (!image->alphaPlane) {
            image->alphaPlane = (uint8_t *)malloc(fullSize);
            if (!image->alphaPlane) return AVIF_RESULT_OUT_OF_MEMORY;
            image->alphaRowBytes = fullRowBytes;
        }
    }
    return AVIF_RESULT_OK;
}

/* ref: src/reformat.c:197-219 */
static int quantY(const OraclePlan * pl, float v)
{
    const int q = (int)roundHalfUp(v * pl->rangeY + pl->biasY);
    return q < 0 ? 0 : (q > pl->yuvMax ? pl->yuvMax : q);
}
static int quantUV(const OraclePlan * pl, float v)
{
    const int q = (pl->mode == 1) ? (int)roundHalfUp(v * pl->rangeY + pl->biasY) : (int)roundHalfUp(v * pl->rangeUV + pl->biasUV);
    return q < 0 ? 0 : (q > pl->yuvMax ? pl->yuvMax : q);
}

/* in-float alpha handling shared by colour and gray inputs (ref: :325-358, :487-510) */
static void applyAlphaFloat(int alphaMode, float a, float * c, int n)
{
    if (alphaMode == 0) return;
    if (a == 0) {
        for (int k = 0; k < n; ++k) c[k] = 0;
    } else if (a < 1.0f) {
        for (int k = 0; k < n; ++k) {
            if (alphaMode == 1) {
                c[k] *= a;
            } else {
                c[k] /= a;
                if (1.0f < c[k]) c[k] = 1.0f;
            }
        }
    }
}

ORACLE_API avifResult oracleImageRGBToYUV(avifImage * image, const avifRGBImage * rgb) /* ref: src/reformat.c:221-571 */
{
    if (!rgb->pixels || rgb->format == AVIF_RGB_FORMAT_RGB_565) return AVIF_RESULT_REFORMAT_FAILED;
    OraclePlan pl;
    if (!makePlan(image, rgb, &pl)) return AVIF_RESULT_REFORMAT_FAILED;
    if (rgb->isFloat) return AVIF_RESULT_NOT_IMPLEMENTED;

    const int hasAlpha = fmtHasAlpha(rgb->format) && !rgb->ignoreAlpha;
    const avifResult ar = allocPlanes(image, hasAlpha);
    if (ar != AVIF_RESULT_OK) return ar;

    int alphaMode = 0; /* ref: :242-249 */
    if (hasAlpha) {
        if (!rgb->alphaPremultiplied && image->alphaPremultiplied) alphaMode = 1;
        else if (rgb->alphaPremultiplied && !image->alphaPremultiplied) alphaMode = 2;
    }
    const int isGray = fmtIsGray(rgb->format);
    if (!isGray && rgb->chromaDownsampling == AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV && image->yuvFormat == AVIF_PIXEL_FORMAT_YUV420) {
        return AVIF_RESULT_NOT_IMPLEMENTED; /* libsharpyuv absent (ref: :256-261, src/reformat_libsharpyuv.c:75-83) */
    }

    const uint32_t rcb = pl.rgbChanBytes, ycb = pl.yuvChanBytes;
    const float m = pl.rgbMaxF;
    if (!isGray) {
        for (uint32_t oj = 0; oj < image->height; oj += 2) {
            for (uint32_t oi = 0; oi < image->width; oi += 2) {
                const uint32_t bw = (oi + 1 >= image->width) ? 1 : 2, bh = (oj + 1 >= image->height) ? 1 : 2;
                float bu[2][2], bv[2][2]; /* [bI][bJ] */
                for (uint32_t bj = 0; bj < bh; ++bj) {
                    for (uint32_t bi = 0; bi < bw; ++bi) {
                        const uint32_t i = oi + bi, j = oj + bj;
                        const uint8_t * px = rgb->pixels + (size_t)i * pl.rgbPixBytes + (size_t)j * rgb->rowBytes;
                        float c[3];
                        c[0] = (float)loadSample(px + pl.offR, rcb) / m; /* ref: :312-323 (a true division) */
                        c[1] = (float)loadSample(px + pl.offG, rcb) / m;
                        c[2] = (float)loadSample(px + pl.offB, rcb) / m;
                        if (alphaMode != 0) applyAlphaFloat(alphaMode, (float)loadSample(px + pl.offA, rcb) / m, c, 3);

                        float y, u, v;
                        if (pl.mode == 1) { /* ref: :361-365 */
                            y = c[1]; u = c[2]; v = c[0];
                        } else if (pl.mode == 2) { /* ref: :366-370 */
                            y = 0.5f * c[1] + 0.25f * (c[0] + c[2]);
                            u = 0.5f * c[1] - 0.25f * (c[0] + c[2]);
                            v = 0.5f * (c[0] - c[2]);
                        } else if (pl.mode == 3 || pl.mode == 4) { /* ref: :371-381 */
                            float t0 = c[0] * m, t1 = c[1] * m, t2 = c[2] * m;
                            t0 = t0 < 0.0f ? 0.0f : (m < t0 ? m : t0);
                            t1 = t1 < 0.0f ? 0.0f : (m < t1 ? m : t1);
                            t2 = t2 < 0.0f ? 0.0f : (m < t2 ? m : t2);
                            const int R = (int)roundHalfUp(t0), G = (int)roundHalfUp(t1), B = (int)roundHalfUp(t2);
                            const int Co = R - B;
                            const int t = B + (Co >> 1);
                            const int Cg = G - t;
                            y = (t + (Cg >> 1)) / pl.rangeY;
                            u = Cg / pl.rangeUV;
                            v = Co / pl.rangeUV;
                        } else { /* ref: :383-386 */
                            const float Y = (pl.kr * c[0]) + (pl.kg * c[1]) + (pl.kb * c[2]);
                            y = Y;
                            u = (c[2] - Y) / (2 * (1 - pl.kb));
                            v = (c[0] - Y) / (2 * (1 - pl.kr));
                        }
                        bu[bi][bj] = u;
                        bv[bi][bj] = v;
                        storeSample(image->yuvPlanes[0] + (size_t)i * ycb + (size_t)j * image->yuvRowBytes[0], ycb, (uint32_t)quantY(&pl, y));
                        if (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444) {
                            storeSample(image->yuvPlanes[1] + (size_t)i * ycb + (size_t)j * image->yuvRowBytes[1], ycb, (uint32_t)quantUV(&pl, u));
                            storeSample(image->yuvPlanes[2] + (size_t)i * ycb + (size_t)j * image->yuvRowBytes[2], ycb, (uint32_t)quantUV(&pl, v));
                        }
                    }
                }
                if (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV420) { /* ref: :413-440; sum order TL,TR,BL,BR */
                    float su = 0.0f, sv = 0.0f;
                    for (uint32_t bj = 0; bj < bh; ++bj) {
                        for (uint32_t bi = 0; bi < bw; ++bi) {
                            su += bu[bi][bj];
                            sv += bv[bi][bj];
                        }
                    }
                    const float n = (float)(bw * bh);
                    const size_t ui = oi >> 1, uj = oj >> 1;
                    storeSample(image->yuvPlanes[1] + ui * ycb + uj * image->yuvRowBytes[1], ycb, (uint32_t)quantUV(&pl, su / n));
                    storeSample(image->yuvPlanes[2] + ui * ycb + uj * image->yuvRowBytes[2], ycb, (uint32_t)quantUV(&pl, sv / n));
                } else if (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV422) { /* ref: :441-467 */
                    for (uint32_t bj = 0; bj < bh; ++bj) {
                        float su = 0.0f, sv = 0.0f;
                        for (uint32_t bi = 0; bi < bw; ++bi) {
                            su += bu[bi][bj];
                            sv += bv[bi][bj];
                        }
                        const float n = (float)bw;
                        const size_t ui = oi >> 1, uj = oj + bj;
                        storeSample(image->yuvPlanes[1] + ui * ycb + uj * image->yuvRowBytes[1], ycb, (uint32_t)quantUV(&pl, su / n));
                        storeSample(image->yuvPlanes[2] + ui * ycb + uj * image->yuvRowBytes[2], ycb, (uint32_t)quantUV(&pl, sv / n));
                    }
                }
            }
        }
    } else { /* gray input (ref: :471-543) */
        for (uint32_t j = 0; j < image->height; ++j) {
            for (uint32_t i = 0; i < image->width; ++i) {
                const uint8_t * px = rgb->pixels + (size_t)i * pl.rgbPixBytes + (size_t)j * rgb->rowBytes;
                float g = (float)loadSample(px + pl.offGray, rcb) / m;
                if (alphaMode != 0) applyAlphaFloat(alphaMode, (float)loadSample(px + pl.offA, rcb) / m, &g, 1);
                storeSample(image->yuvPlanes[0] + (size_t)i * ycb + (size_t)j * image->yuvRowBytes[0], ycb, (uint32_t)quantY(&pl, g));
            }
        }
        /* chroma planes, if present, become constant "half" over rowBytes x chromaHeight, padding included (ref: :520-542) */
        const uint32_t shiftedH = (uint32_t)(((uint64_t)image->height + pl.shiftY) >> pl.shiftY);
        const int half = 1 << (image->depth - 1);
        for (int p = 1; p <= 2; ++p) {
            if (!image->yuvPlanes[p]) continue;
            const size_t bytes = (size_t)shiftedH * image->yuvRowBytes[p];
            if (ycb > 1) {
                for (size_t k = 0; k < bytes / 2; ++k) storeSample(image->yuvPlanes[p] + 2 * k, 2, (uint32_t)half);
            } else {
                memset(image->yuvPlanes[p], half, bytes);
            }
        }
    }

    if (image->alphaPlane && image->alphaRowBytes) { /* ref: :545-569 */
        avifAlphaParams pr;
        pr.width = image->width;
        pr.height = image->height;
        pr.dstDepth = image->depth;
        pr.dstPlane = image->alphaPlane;
        pr.dstRowBytes = image->alphaRowBytes;
        pr.dstOffsetBytes = 0;
        pr.dstPixelBytes = ycb;
        if (fmtHasAlpha(rgb->format) && !rgb->ignoreAlpha) {
            pr.srcDepth = rgb->depth;
            pr.srcPlane = rgb->pixels;
            pr.srcRowBytes = rgb->rowBytes;
            pr.srcOffsetBytes = pl.offA;
            pr.srcPixelBytes = pl.rgbPixBytes;
            oracleReformatAlpha(&pr);
        } else {
            oracleFillAlpha(&pr);
        }
    }
    return AVIF_RESULT_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* "next" rows of SURVEY.md section 8(f)                                                       */

/* ref: src/read.c:6724-6764 (avifImageLimitedToFullAlpha): integer range expansion of every alpha sample */
ORACLE_API avifResult oracleImageLimitedToFullAlpha(avifImage * image)
{
    if (!image->alphaPlane) return AVIF_RESULT_OK;
    const uint32_t cb = image->depth > 8 ? 2 : 1;
    for (uint32_t j = 0; j < image->height; ++j) {
        uint8_t * row = image->alphaPlane + (size_t)j * image->alphaRowBytes;
        for (uint32_t i = 0; i < image->width; ++i) {
            const int v = (int)loadSample(row + (size_t)i * cb, cb);
            storeSample(row + (size_t)i * cb, cb, (uint32_t)oracleLimitedToFullY(image->depth, v));
        }
    }
    return AVIF_RESULT_OK;
}

/* deterministic byte generator + hash shared with SURVEY.md Appendix D (normative for the golden hashes) */
ORACLE_API uint32_t oracleXorshiftFill(uint32_t state, void * dst, size_t count, uint32_t mask, int bytesPerSample)
{
    uint8_t * d8 = (uint8_t *)dst;
    for (size_t i = 0; i < count; ++i) {
        state ^= state << 13;
        state ^= state >> 17;
        state ^= state << 5;
        const uint32_t v = state & mask;
        if (bytesPerSample == 2) {
            const uint16_t s = (uint16_t)v;
            memcpy(d8 + 2 * i, &s, 2);
        } else {
            d8[i] = (uint8_t)v;
        }
    }
    return state;
}

ORACLE_API uint64_t oracleFnv1a64(uint64_t h, const void * data, size_t n)
{
    const uint8_t * p = (const uint8_t *)data;
    for (size_t i = 0; i < n; ++i) {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
