/*
 * avif_oracle_tonemap.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of gain-map tone mapping
 * (avifRGBImageApplyGainMap, ref: src/gainmap.c:73-315; gain maps of another size go through avif_oracle_scale.c first), with the
 * host libm like the reference: against the reference compiled from its own sources (oracle/_ref) it is BIT-EXACT
 * (tests/test_oracle_tonemap.py) -- parity status PINNED.  The CUDA kernel (libavif_b200/csrc/tonemap.cu) is compared with the
 * reference / this restatement to +-1 output code because CUDA's powf / exp2f are not glibc's.
 *
 * Only tests/ may load this.  Nothing under libavif_b200/ links or calls it.
 * Compile like avif_oracle.c: -ffp-contract=off, no -march=native / -ffast-math.
 */
#include "avif_b200.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

avifResult oracleImageYUVToRGB(const avifImage * image, avifRGBImage * rgb); /* avif_oracle.c */
avifResult oracleImageScale(avifImage * image, uint32_t dstWidth, uint32_t dstHeight, avifDiagnostics * diag); /* avif_oracle_scale.c */

/* the few host helpers of src/avif.c this file needs, restated locally so that liboracle.so stays self-contained
 * (ref: src/avif.c:670-741); the output pixels are malloc'ed like avifAlloc does */
static int fmtHasAlpha(avifRGBFormat f)
{
    return f == AVIF_RGB_FORMAT_RGBA || f == AVIF_RGB_FORMAT_ARGB || f == AVIF_RGB_FORMAT_BGRA || f == AVIF_RGB_FORMAT_ABGR ||
           f == AVIF_RGB_FORMAT_GRAYA || f == AVIF_RGB_FORMAT_AGRAY;
}
static uint32_t fmtChannels(avifRGBFormat f)
{
    switch (f) {
        case AVIF_RGB_FORMAT_GRAY: return 1;
        case AVIF_RGB_FORMAT_GRAYA:
        case AVIF_RGB_FORMAT_AGRAY: return 2;
        case AVIF_RGB_FORMAT_RGB:
        case AVIF_RGB_FORMAT_BGR: return 3;
        default: return 4;
    }
}
static uint32_t pixelSize(const avifRGBImage * rgb)
{
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565) return 2;
    return fmtChannels(rgb->format) * ((rgb->depth > 8) ? 2 : 1);
}
static avifResult allocPixels(avifRGBImage * rgb)
{
    free(rgb->pixels);
    rgb->pixels = NULL;
    rgb->rowBytes = 0;
    if (!rgb->width || !rgb->height) return AVIF_RESULT_INVALID_ARGUMENT;
    rgb->rowBytes = rgb->width * pixelSize(rgb);
    rgb->pixels = (uint8_t *)malloc((size_t)rgb->rowBytes * rgb->height);
    return rgb->pixels ? AVIF_RESULT_OK : AVIF_RESULT_OUT_OF_MEMORY;
}
static void setDefaults(avifRGBImage * rgb, const avifImage * image) /* ref: src/avif.c:700-717 */
{
    memset(rgb, 0, sizeof(*rgb));
    rgb->width = image->width;
    rgb->height = image->height;
    rgb->depth = image->depth;
    rgb->format = AVIF_RGB_FORMAT_RGBA;
    rgb->chromaUpsampling = AVIF_CHROMA_UPSAMPLING_AUTOMATIC;
    rgb->chromaDownsampling = AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC;
    rgb->maxThreads = 1;
}

static float clamp01(float v) { return (v < 0.0f) ? 0.0f : ((1.0f < v) ? 1.0f : v); } /* AVIF_CLAMP */
static float fmin2(float a, float b) { return (a < b) ? a : b; }                       /* AVIF_MIN */
static float fmax2(float a, float b) { return (a > b) ? a : b; }                       /* AVIF_MAX */

/* ref: src/colr.c:216-467, table :470-491, defaults :497,:508 */
static float toLinear(int tc, float g)
{
    switch (tc) {
        case 4: return powf(clamp01(g), 2.2f);
        case 5: return powf(clamp01(g), 2.8f);
        case 7:
            if (g < 0.0f) return 0.0f;
            if (g < 4.0f * 0.022821585529445f) return g / 4.0f;
            if (g < 1.0f) return powf((g + 0.111572195921731f) / 1.111572195921731f, 1.0f / 0.45f);
            return 1.0f;
        case 8: return clamp01(g);
        case 9: return (g <= 0.0f) ? (0.01f / 2.f) : powf(10.0f, 2.f * (fmin2(g, 1.f) - 1.0f));
        case 10: return (g <= 0.0f) ? (0.00316227766f / 2.f) : powf(10.0f, 2.5f * (fmin2(g, 1.f) - 1.0f));
        case 11:
            if (g < -4.5f * 0.018053968510807f) return -powf((g - 0.09929682680944f) / -1.09929682680944f, 1.0f / 0.45f);
            if (g < 4.5f * 0.018053968510807f) return g / 4.5f;
            return powf((g + 0.09929682680944f) / 1.09929682680944f, 1.0f / 0.45f);
        case 12:
            if (g < -0.25f) return -0.25f;
            if (g < 0.0f) return powf((g - 0.02482420670236f) / -0.27482420670236f, 1.0f / 0.45f) / -4.0f;
            if (g < 4.5f * 0.018053968510807f) return g / 4.5f;
            if (g < 1.0f) return powf((g + 0.09929682680944f) / 1.09929682680944f, 1.0f / 0.45f);
            return 1.0f;
        case 13:
            if (g < 0.0f) return 0.0f;
            if (g < 12.92f * 0.0030412825601275209f) return g / 12.92f;
            if (g < 1.0f) return powf((g + 0.0550107189475866f) / 1.0550107189475866f, 2.4f);
            return 1.0f;
        case 16:
            if (g > 0.0f) {
                const float pg = powf(g, 1.0f / 78.84375f);
                const float num = fmax2(pg - 0.8359375f, 0.0f);
                const float den = fmax2(18.8515625f - 18.6875f * pg, FLT_MIN);
                const float lin = powf(num / den, 1.0f / 0.1593017578125f);
                return lin * 10000.0f / 203.0f;
            }
            return 0.0f;
        case 17: return powf(fmax2(g, 0.0f), 2.6f) / 0.91655527974030934f;
        case 18: {
            if (g < 0.0f) return 0.0f;
            float lin;
            if (g <= 0.5f) lin = powf((g * g) * (1.0f / 3.0f), 1.2f);
            else lin = powf((expf((g - 0.55991073f) / 0.17883277f) + 0.28466892f) / 12.0f, 1.2f);
            return lin * 1000.0f / 203.0f;
        }
        default:
            if (g < 0.0f) return 0.0f;
            if (g < 4.5f * 0.018053968510807f) return g / 4.5f;
            if (g < 1.0f) return powf((g + 0.09929682680944f) / 1.09929682680944f, 1.0f / 0.45f);
            return 1.0f;
    }
}

static float toGamma(int tc, float l)
{
    switch (tc) {
        case 4: return powf(clamp01(l), 1.0f / 2.2f);
        case 5: return powf(clamp01(l), 1.0f / 2.8f);
        case 7:
            if (l < 0.0f) return 0.0f;
            if (l < 0.022821585529445f) return l * 4.0f;
            if (l < 1.0f) return 1.111572195921731f * powf(l, 0.45f) - 0.111572195921731f;
            return 1.0f;
        case 8: return clamp01(l);
        case 9: return (l <= 0.01f) ? 0.0f : 1.0f + log10f(fmin2(l, 1.0f)) / 2.0f;
        case 10: return (l <= 0.00316227766f) ? 0.0f : 1.0f + log10f(fmin2(l, 1.0f)) / 2.5f;
        case 11:
            if (l < -0.018053968510807f) return -1.09929682680944f * powf(-l, 0.45f) + 0.09929682680944f;
            if (l < 0.018053968510807f) return l * 4.5f;
            return 1.09929682680944f * powf(l, 0.45f) - 0.09929682680944f;
        case 12:
            if (l < -0.25f) return -0.25f;
            if (l < 0.0f) return -0.27482420670236f * powf(-4.0f * l, 0.45f) + 0.02482420670236f;
            if (l < 0.018053968510807f) return l * 4.5f;
            if (l < 1.0f) return 1.09929682680944f * powf(l, 0.45f) - 0.09929682680944f;
            return 1.0f;
        case 13:
            if (l < 0.0f) return 0.0f;
            if (l < 0.0030412825601275209f) return l * 12.92f;
            if (l < 1.0f) return 1.0550107189475866f * powf(l, 1.0f / 2.4f) - 0.0550107189475866f;
            return 1.0f;
        case 16:
            if (l > 0.0f) {
                const float s = clamp01(l * 203.0f / 10000.0f);
                const float pl = powf(s, 0.1593017578125f);
                const float num = 0.1640625f * pl - 0.1640625f;
                const float den = 1.0f + 18.6875f * pl;
                return powf(1.0f + num / den, 78.84375f);
            }
            return 0.0f;
        case 17: return powf(0.91655527974030934f * fmax2(l, 0.0f), 1.0f / 2.6f);
        case 18: {
            float s = clamp01(l * 203.0f / 1000.0f);
            s = powf(s, 1.0f / 1.2f);
            if (s < 0.0f) return 0.0f;
            if (s <= (1.0f / 12.0f)) return sqrtf(3.0f * s);
            return 0.17883277f * logf(12.0f * s - 0.28466892f) + 0.55991073f;
        }
        default:
            if (l < 0.0f) return 0.0f;
            if (l < 0.018053968510807f) return l * 4.5f;
            if (l < 1.0f) return 1.09929682680944f * powf(l, 0.45f) - 0.09929682680944f;
            return 1.0f;
    }
}

/* ---- pixel accessors, ref: src/reformat.c:32-117 (layout), :1842-1939 (get / set) ---- */
typedef struct Layout {
    uint32_t pixBytes, chanBytes, offR, offG, offB, offA;
    int hasAlpha, isFloat, is565;
    float maxF;
} Layout;

static int layoutOf(const avifRGBImage * rgb, Layout * L)
{
    if (!(rgb->depth == 8 || rgb->depth == 10 || rgb->depth == 12 || rgb->depth == 16)) return 0;
    if (rgb->isFloat && rgb->depth != 16) return 0;
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565 && rgb->depth != 8) return 0;
    if ((int)rgb->format < AVIF_RGB_FORMAT_RGB || rgb->format >= AVIF_RGB_FORMAT_COUNT) return 0;
    memset(L, 0, sizeof(*L));
    const uint32_t cb = (rgb->depth > 8) ? 2 : 1;
    L->chanBytes = cb;
    L->pixBytes = pixelSize(rgb);
    switch (rgb->format) {
        case AVIF_RGB_FORMAT_RGB: L->offG = cb; L->offB = 2 * cb; break;
        case AVIF_RGB_FORMAT_RGBA: L->offG = cb; L->offB = 2 * cb; L->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ARGB: L->offR = cb; L->offG = 2 * cb; L->offB = 3 * cb; break;
        case AVIF_RGB_FORMAT_BGR: L->offG = cb; L->offR = 2 * cb; break;
        case AVIF_RGB_FORMAT_BGRA: L->offG = cb; L->offR = 2 * cb; L->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ABGR: L->offB = cb; L->offG = 2 * cb; L->offR = 3 * cb; break;
        case AVIF_RGB_FORMAT_GRAYA: L->offA = cb; break;
        default: break;
    }
    L->hasAlpha = fmtHasAlpha(rgb->format);
    L->isFloat = rgb->isFloat ? 1 : 0;
    L->is565 = (rgb->format == AVIF_RGB_FORMAT_RGB_565);
    L->maxF = (float)((1 << rgb->depth) - 1);
    return 1;
}

#define F16_MULT 1.9259299444e-34f /* ref: src/reformat.c:1411 */
static float halfToFloat(uint16_t v)
{
    union { float f; uint32_t u; } x;
    x.u = (uint32_t)v << 13;
    return x.f / F16_MULT;
}
static uint16_t floatToHalf(float v)
{
    union { float f; uint32_t u; } x;
    x.f = v * F16_MULT;
    return (uint16_t)(x.u >> 13);
}

static void getPixel(const avifRGBImage * img, const Layout * L, uint32_t x, uint32_t y, float px[4])
{
    const uint8_t * p = img->pixels + (size_t)y * img->rowBytes + (size_t)x * L->pixBytes;
    if (L->chanBytes > 1) {
        uint16_t r, g, b, a = (uint16_t)L->maxF;
        memcpy(&r, p + L->offR, 2);
        memcpy(&g, p + L->offG, 2);
        memcpy(&b, p + L->offB, 2);
        if (L->hasAlpha) memcpy(&a, p + L->offA, 2);
        if (L->isFloat) {
            px[0] = halfToFloat(r); px[1] = halfToFloat(g); px[2] = halfToFloat(b);
            px[3] = L->hasAlpha ? halfToFloat(a) : 1.0f;
        } else {
            px[0] = r / L->maxF; px[1] = g / L->maxF; px[2] = b / L->maxF; px[3] = a / L->maxF;
        }
    } else if (L->is565) {
        uint16_t v;
        memcpy(&v, p + L->offR, 2);
        const uint16_t r5 = (v & 0xF800) >> 11, g6 = (v & 0x07E0) >> 5, b5 = v & 0x001F;
        px[0] = (uint8_t)((r5 << 3) | (r5 >> 2)) / L->maxF;
        px[1] = (uint8_t)((g6 << 2) | (g6 >> 4)) / L->maxF;
        px[2] = (uint8_t)((b5 << 3) | (b5 >> 2)) / L->maxF;
        px[3] = 1.0f;
    } else {
        px[0] = p[L->offR] / L->maxF; px[1] = p[L->offG] / L->maxF; px[2] = p[L->offB] / L->maxF;
        px[3] = L->hasAlpha ? (p[L->offA] / L->maxF) : 1.0f;
    }
}

static void setPixel(const avifRGBImage * img, const Layout * L, uint32_t x, uint32_t y, const float px[4])
{
    uint8_t * p = img->pixels + (size_t)y * img->rowBytes + (size_t)x * L->pixBytes;
    if (L->chanBytes > 1) {
        uint16_t v[4];
        for (int c = 0; c < 4; ++c) v[c] = L->isFloat ? floatToHalf(px[c]) : (uint16_t)(0.5f + (px[c] * L->maxF));
        memcpy(p + L->offR, &v[0], 2);
        memcpy(p + L->offG, &v[1], 2);
        memcpy(p + L->offB, &v[2], 2);
        if (L->hasAlpha) memcpy(p + L->offA, &v[3], 2);
    } else {
        const uint8_t R = (uint8_t)(0.5f + (px[0] * L->maxF)), G = (uint8_t)(0.5f + (px[1] * L->maxF)), B = (uint8_t)(0.5f + (px[2] * L->maxF));
        if (L->is565) {
            const uint16_t v = (uint16_t)((B >> 3) | ((G >> 2) << 5) | ((R >> 3) << 11)); /* ref: :619 */
            memcpy(p + L->offR, &v, 2);
            return;
        }
        p[L->offR] = R;
        p[L->offG] = G;
        p[L->offB] = B;
        if (L->hasAlpha) p[L->offA] = (uint8_t)(0.5f + (px[3] * L->maxF));
    }
}

/* ---- primaries, ref: src/colr.c:16-43, src/colrconvert.c ---- */
static const float * primariesOf(uint16_t cp)
{
    static const struct { uint16_t id; float v[8]; } t[] = {
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
    for (size_t k = 0; k < sizeof(t) / sizeof(t[0]); ++k) {
        if (t[k].id == cp) return t[k].v;
    }
    return t[0].v;
}
typedef double Mat[3][3];
static int matInv(Mat m, Mat o) /* ref: src/colrconvert.c:26-46 */
{
    double det = m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                 m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (fabs(det) < 1e-12) return 0;
    det = 1.0 / det;
    o[0][0] = (m[1][1] * m[2][2] - m[2][1] * m[1][2]) * det;
    o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * det;
    o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * det;
    o[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * det;
    o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * det;
    o[1][2] = (m[1][0] * m[0][2] - m[0][0] * m[1][2]) * det;
    o[2][0] = (m[1][0] * m[2][1] - m[2][0] * m[1][1]) * det;
    o[2][1] = (m[2][0] * m[0][1] - m[0][0] * m[2][1]) * det;
    o[2][2] = (m[0][0] * m[1][1] - m[1][0] * m[0][1]) * det;
    return 1;
}
static void matMul(Mat a, Mat b, Mat o)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) o[r][c] = a[r][0] * b[0][c] + a[r][1] * b[1][c] + a[r][2] * b[2][c];
}
static void matVec(Mat m, const double x[3], double y[3])
{
    for (int r = 0; r < 3; ++r) y[r] = m[r][0] * x[0] + m[r][1] * x[1] + m[r][2] * x[2];
}
static void matDiag(const double d[3], Mat o)
{
    memset(o, 0, sizeof(Mat));
    o[0][0] = d[0]; o[1][1] = d[1]; o[2][2] = d[2];
}
static int toXyzD50(uint16_t cp, Mat out) /* ref: src/colrconvert.c:88-141 */
{
    static Mat bradford = { { 0.8951, 0.2664, -0.1614 }, { -0.7502, 1.7135, 0.0367 }, { 0.0389, -0.0685, 1.0296 } };
    static const double lmsD50[3] = { 0.996284, 1.02043, 0.818644 };
    const float * p = primariesOf(cp);
    if (fabsf(p[7]) < 1e-12) return 0;
    const double f = 1.0 / p[7];
    const double white[3] = { p[6] * f, 1, (1 - p[6] - p[7]) * f };
    Mat prim = { { p[0], p[2], p[4] }, { p[1], p[3], p[5] }, { 1.0 - p[0] - p[1], 1.0 - p[2] - p[3], 1.0 - p[4] - p[5] } };
    Mat primInv, coefM, rgbXyz, adapt, tmp, bradInv;
    double coef[3], lms[3];
    if (!matInv(prim, primInv)) return 0;
    matVec(primInv, white, coef);
    matDiag(coef, coefM);
    matMul(prim, coefM, rgbXyz);
    matVec(bradford, white, lms);
    for (int k = 0; k < 3; ++k) {
        if (fabs(lms[k]) < 1e-12) return 0;
        lms[k] = lmsD50[k] / lms[k];
    }
    matDiag(lms, adapt);
    matMul(adapt, bradford, tmp);
    if (!matInv(bradford, bradInv)) return 0;
    matMul(bradInv, tmp, adapt);
    matMul(adapt, rgbXyz, out);
    return 1;
}
static int rgbToRgb(uint16_t src, uint16_t dst, Mat out) /* ref: src/colrconvert.c:157-171 */
{
    Mat a, b, bInv;
    if (!toXyzD50(src, a) || !toXyzD50(dst, b) || !matInv(b, bInv)) return 0;
    matMul(bInv, a, out);
    return 1;
}
static void convertPrimaries(float px[4], Mat m) /* ref: src/colrconvert.c:178-186 */
{
    const double x[3] = { px[0], px[1], px[2] };
    double y[3];
    matVec(m, x, y);
    px[0] = (float)y[0]; px[1] = (float)y[1]; px[2] = (float)y[2];
}

static float sfrac(avifSignedFraction f) { return f.d ? ((float)f.n / f.d) : 0.0f; }   /* ref: src/gainmap.c:30-36 */
static float ufrac(avifUnsignedFraction f) { return f.d ? ((float)f.n / f.d) : 0.0f; } /* ref: :38-44 */
static float clampf_u16(float v) { return (v < 0.0f) ? 0.0f : ((65535.0f < v) ? 65535.0f : v); }
static float nanSafeClamp(float v) { return fminf(1.0f, fmaxf(0.0f, v)); }             /* ref: :13-16 */

/* ref: src/gainmap.c:73-315 */
ORACLE_API avifResult oracleRGBImageApplyGainMap(const avifRGBImage * base, avifColorPrimaries baseCP, avifTransferCharacteristics baseTC,
                                                 const avifGainMap * gm, float hdrHeadroom, avifColorPrimaries outCP,
                                                 avifTransferCharacteristics outTC, avifRGBImage * out,
                                                 avifContentLightLevelInformationBox * clli, avifDiagnostics * diag)
{
    if (diag) diag->error[0] = '\0';
    if (hdrHeadroom < 0.0f || !base || !gm || !out) return AVIF_RESULT_INVALID_ARGUMENT;
    for (int c = 0; c < 3; ++c) { /* ref: :430-447 */
        if (!gm->gainMapMin[c].d || !gm->gainMapMax[c].d || !gm->gainMapGamma[c].d || !gm->baseOffset[c].d || !gm->alternateOffset[c].d) return AVIF_RESULT_INVALID_ARGUMENT;
        if ((int64_t)gm->gainMapMax[c].n * gm->gainMapMin[c].d < (int64_t)gm->gainMapMin[c].n * gm->gainMapMax[c].d) return AVIF_RESULT_INVALID_ARGUMENT;
        if (!gm->gainMapGamma[c].n) return AVIF_RESULT_INVALID_ARGUMENT;
    }
    if (!gm->baseHdrHeadroom.d || !gm->alternateHdrHeadroom.d) return AVIF_RESULT_INVALID_ARGUMENT;
    if (gm->useBaseColorSpace != 0 && gm->useBaseColorSpace != 1) return AVIF_RESULT_INVALID_ARGUMENT;

    const uint32_t W = base->width, H = base->height;
    const avifColorPrimaries mathCP = (gm->useBaseColorSpace || gm->altColorPrimaries == AVIF_COLOR_PRIMARIES_UNSPECIFIED) ? baseCP : gm->altColorPrimaries;
    const int needIn = (baseCP != mathCP), needOut = (mathCP != outCP);
    out->width = W;
    out->height = H;
    avifResult res = allocPixels(out);
    if (res != AVIF_RESULT_OK) return res;

    float weight; /* ref: :50-63 */
    {
        const float bh = ufrac(gm->baseHdrHeadroom), ah = ufrac(gm->alternateHdrHeadroom);
        if (bh == ah) weight = 0.0f;
        else {
            const float w = clamp01((hdrHeadroom - bh) / (ah - bh));
            weight = (ah < bh) ? -w : w;
        }
    }
    if (weight == 0.0f && outTC == baseTC && outCP == baseCP && base->format == out->format && base->depth == out->depth &&
        base->isFloat == out->isFloat && base->rowBytes == out->rowBytes) {
        memcpy(out->pixels, base->pixels, (size_t)base->rowBytes * base->height);
        return AVIF_RESULT_OK;
    }
    Layout LB, LO, LG;
    if (!layoutOf(base, &LB) || !layoutOf(out, &LO)) return AVIF_RESULT_NOT_IMPLEMENTED;

    if (weight == 0.0f) { /* ref: :141-169 */
        const int differ = (baseCP != outCP);
        Mat m;
        if (differ && !rgbToRgb(baseCP, outCP, m)) return AVIF_RESULT_NOT_IMPLEMENTED;
        for (uint32_t j = 0; j < H; ++j) {
            for (uint32_t i = 0; i < W; ++i) {
                float px[4];
                getPixel(base, &LB, i, j, px);
                if (outTC != baseTC || differ) {
                    for (int c = 0; c < 3; ++c) px[c] = toLinear(baseTC, px[c]);
                    if (differ) convertPrimaries(px, m);
                    for (int c = 0; c < 3; ++c) px[c] = nanSafeClamp(toGamma(outTC, px[c]));
                }
                setPixel(out, &LO, i, j, px);
            }
        }
        return AVIF_RESULT_OK;
    }

    Mat inM, outM;
    if (needIn && !rgbToRgb(baseCP, mathCP, inM)) return AVIF_RESULT_NOT_IMPLEMENTED;
    if (needOut && !rgbToRgb(mathCP, outCP, outM)) return AVIF_RESULT_NOT_IMPLEMENTED;
    if (!gm->image) return AVIF_RESULT_INVALID_ARGUMENT;
    const avifImage * gmImage = gm->image;
    avifImage scaled;
    int haveScaled = 0;
    if (gmImage->width != W || gmImage->height != H) { /* ref: :186-201, through avif_oracle_scale.c */
        scaled = *gmImage; /* a view: the scaler replaces the planes with its own and owns those */
        scaled.imageOwnsYUVPlanes = AVIF_FALSE;
        scaled.imageOwnsAlphaPlane = AVIF_FALSE;
        res = oracleImageScale(&scaled, W, H, NULL);
        if (res != AVIF_RESULT_OK) return res;
        haveScaled = 1;
        gmImage = &scaled;
    }

    avifRGBImage g; /* ref: :203-212 */
    setDefaults(&g, gmImage);
    res = allocPixels(&g);
    if (res == AVIF_RESULT_OK) res = oracleImageYUVToRGB(gmImage, &g);
    if (haveScaled) {
        for (int c = 0; c < 3; ++c) free(scaled.yuvPlanes[c]);
        free(scaled.alphaPlane);
    }
    if (res != AVIF_RESULT_OK || !layoutOf(&g, &LG)) {
        free(g.pixels);
        return (res != AVIF_RESULT_OK) ? res : AVIF_RESULT_NOT_IMPLEMENTED;
    }
    float gammaInv[3], gmin[3], gmax[3], boff[3], aoff[3];
    for (int c = 0; c < 3; ++c) {
        gammaInv[c] = 1.0f / ufrac(gm->gainMapGamma[c]);
        gmin[c] = sfrac(gm->gainMapMin[c]);
        gmax[c] = sfrac(gm->gainMapMax[c]);
        boff[c] = sfrac(gm->baseOffset[c]);
        aoff[c] = sfrac(gm->alternateOffset[c]);
    }
    float rgbMaxLinear = 0, rgbSumLinear = 0;
    for (uint32_t j = 0; j < H; ++j) {
        for (uint32_t i = 0; i < W; ++i) {
            float b[4], gg[4], t[4];
            getPixel(base, &LB, i, j, b);
            getPixel(&g, &LG, i, j, gg);
            float pixelMax = 0.0f;
            for (int c = 0; c < 3; ++c) b[c] = toLinear(baseTC, b[c]);
            if (needIn) convertPrimaries(b, inM);
            for (int c = 0; c < 3; ++c) {
                const float w = powf(gg[c], gammaInv[c]);
                const float log2g = (1.0f - w) * gmin[c] + w * gmax[c]; /* lerp, ref: :66-69 */
                const float v = (b[c] + boff[c]) * exp2f(log2g * weight) - aoff[c];
                if (v > rgbMaxLinear) rgbMaxLinear = v;
                if (v > pixelMax) pixelMax = v;
                t[c] = v;
            }
            if (needOut) convertPrimaries(t, outM);
            for (int c = 0; c < 3; ++c) {
                if (isnan(t[c])) {
                    free(g.pixels);
                    return AVIF_RESULT_INVALID_TONE_MAPPED_IMAGE;
                }
                t[c] = nanSafeClamp(toGamma(outTC, t[c]));
            }
            t[3] = b[3];
            rgbSumLinear += pixelMax;
            setPixel(out, &LO, i, j, t);
        }
    }
    if (clli) { /* ref: :297-306 */
        clli->maxCLL = (uint16_t)clampf_u16(floorf(rgbMaxLinear * 203.0f + 0.5f));
        const float avg = rgbSumLinear / ((size_t)W * H);
        clli->maxPALL = (uint16_t)clampf_u16(floorf(avg * 203.0f + 0.5f));
    }
    free(g.pixels);
    return AVIF_RESULT_OK;
}
