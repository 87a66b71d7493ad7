/*
 * gainmap.c -- host side of gain-map tone mapping: avifRGBImageApplyGainMap / avifImageApplyGainMap
 * with the reference's signatures, checks, result codes and early exits; the per-pixel work is tonemap.cu.
 *
 *   avifRGBImageApplyGainMap      src/gainmap.c:73-315
 *   avifImageApplyGainMap         src/gainmap.c:317-360
 *   avifGainMapValidateMetadata   src/gainmap.c:430-457
 *   weight / fractions            src/gainmap.c:30-63
 *   RGB -> RGB primaries matrix   src/colrconvert.c:10-176  (double precision, Bradford adaptation to D50)
 *   pixel layout                  src/reformat.c:32-117
 *
 * Uses only the library's own public entry points (the gain map's YUV -> RGB conversion is the product's
 * avifImageYUVToRGB writing into device memory) plus the CUDA runtime.  No CPU pixel arithmetic, no fallback.
 */
#define AVIF_B200_BUILDING
#include "avif_b200.h"

#include <cuda_runtime_api.h>
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plan.h"

static void diagClear(avifDiagnostics * diag)
{
    if (diag) diag->error[0] = '\0';
}
static void diagSet(avifDiagnostics * diag, const char * fmt, ...)
{
    if (!diag || diag->error[0]) return; /* first error wins, like src/diag.c */
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(diag->error, sizeof(diag->error), fmt, ap);
    va_end(ap);
}

static float sfrac(avifSignedFraction f) { return f.d ? ((float)f.n / (float)f.d) : 0.0f; }   /* src/gainmap.c:30-36 */
static float ufrac(avifUnsignedFraction f) { return f.d ? ((float)f.n / (float)f.d) : 0.0f; } /* :38-44 */

/* src/gainmap.c:430-457 */
static avifResult validateMetadata(const avifGainMap * gm, avifDiagnostics * diag)
{
    for (int c = 0; c < 3; ++c) {
        if (!gm->gainMapMin[c].d || !gm->gainMapMax[c].d || !gm->gainMapGamma[c].d || !gm->baseOffset[c].d || !gm->alternateOffset[c].d) {
            diagSet(diag, "gain map metadata: a per-channel denominator is 0");
            return AVIF_RESULT_INVALID_ARGUMENT;
        }
        if ((int64_t)gm->gainMapMax[c].n * gm->gainMapMin[c].d < (int64_t)gm->gainMapMin[c].n * gm->gainMapMax[c].d) {
            diagSet(diag, "gain map metadata: per-channel max below per-channel min");
            return AVIF_RESULT_INVALID_ARGUMENT;
        }
        if (!gm->gainMapGamma[c].n) {
            diagSet(diag, "gain map metadata: per-channel gamma is 0");
            return AVIF_RESULT_INVALID_ARGUMENT;
        }
    }
    if (!gm->baseHdrHeadroom.d || !gm->alternateHdrHeadroom.d) {
        diagSet(diag, "gain map metadata: headroom denominator is 0");
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    if (gm->useBaseColorSpace != 0 && gm->useBaseColorSpace != 1) {
        diagSet(diag, "gain map metadata: useBaseColorSpace is %d", gm->useBaseColorSpace);
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    return AVIF_RESULT_OK;
}

/* src/gainmap.c:50-63 */
static float gainMapWeight(float hdrHeadroom, const avifGainMap * gm)
{
    const float baseH = ufrac(gm->baseHdrHeadroom), altH = ufrac(gm->alternateHdrHeadroom);
    if (baseH == altH) return 0.0f;
    float w = (hdrHeadroom - baseH) / (altH - baseH);
    w = (w < 0.0f) ? 0.0f : ((1.0f < w) ? 1.0f : w);
    return (altH < baseH) ? -w : w;
}

/* ---- primaries: 3x3 matrices in double (src/colrconvert.c) ---- */

static const float * primariesOf(avifColorPrimaries cp)
{
    static const struct { uint16_t id; float v[8]; } table[] = { /* rX rY gX gY bX bY wX wY, src/colr.c:16-28 */
        { AVIF_COLOR_PRIMARIES_BT709, { 0.64f, 0.33f, 0.3f, 0.6f, 0.15f, 0.06f, 0.3127f, 0.329f } },
        { AVIF_COLOR_PRIMARIES_BT470M, { 0.67f, 0.33f, 0.21f, 0.71f, 0.14f, 0.08f, 0.310f, 0.316f } },
        { AVIF_COLOR_PRIMARIES_BT470BG, { 0.64f, 0.33f, 0.29f, 0.60f, 0.15f, 0.06f, 0.3127f, 0.3290f } },
        { AVIF_COLOR_PRIMARIES_BT601, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { AVIF_COLOR_PRIMARIES_SMPTE240, { 0.630f, 0.340f, 0.310f, 0.595f, 0.155f, 0.070f, 0.3127f, 0.3290f } },
        { AVIF_COLOR_PRIMARIES_GENERIC_FILM, { 0.681f, 0.319f, 0.243f, 0.692f, 0.145f, 0.049f, 0.310f, 0.316f } },
        { AVIF_COLOR_PRIMARIES_BT2020, { 0.708f, 0.292f, 0.170f, 0.797f, 0.131f, 0.046f, 0.3127f, 0.3290f } },
        { AVIF_COLOR_PRIMARIES_XYZ, { 1.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.3333f, 0.3333f } },
        { AVIF_COLOR_PRIMARIES_SMPTE431, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.314f, 0.351f } },
        { AVIF_COLOR_PRIMARIES_SMPTE432, { 0.680f, 0.320f, 0.265f, 0.690f, 0.150f, 0.060f, 0.3127f, 0.3290f } },
        { AVIF_COLOR_PRIMARIES_EBU3213, { 0.630f, 0.340f, 0.295f, 0.605f, 0.155f, 0.077f, 0.3127f, 0.3290f } },
    };
    for (size_t k = 0; k < sizeof(table) / sizeof(table[0]); ++k) {
        if (table[k].id == cp) return table[k].v;
    }
    return table[0].v; /* unknown -> BT.709 (src/colr.c:41-42) */
}

#define M(m, r, c) ((m)[3 * (r) + (c)])
static const double kTiny = 1e-12; /* src/colrconvert.c:8 */

/* cofactor inverse, same term order as src/colrconvert.c:26-46 (the products feed a float in the end, but keeping the
 * association keeps the doubles equal too) */
static int inv3(const double * m, double * o)
{
    double det = M(m, 0, 0) * (M(m, 1, 1) * M(m, 2, 2) - M(m, 2, 1) * M(m, 1, 2)) - M(m, 0, 1) * (M(m, 1, 0) * M(m, 2, 2) - M(m, 1, 2) * M(m, 2, 0)) +
                 M(m, 0, 2) * (M(m, 1, 0) * M(m, 2, 1) - M(m, 1, 1) * M(m, 2, 0));
    if (fabs(det) < kTiny) return 0;
    det = 1.0 / det;
    M(o, 0, 0) = (M(m, 1, 1) * M(m, 2, 2) - M(m, 2, 1) * M(m, 1, 2)) * det;
    M(o, 0, 1) = (M(m, 0, 2) * M(m, 2, 1) - M(m, 0, 1) * M(m, 2, 2)) * det;
    M(o, 0, 2) = (M(m, 0, 1) * M(m, 1, 2) - M(m, 0, 2) * M(m, 1, 1)) * det;
    M(o, 1, 0) = (M(m, 1, 2) * M(m, 2, 0) - M(m, 1, 0) * M(m, 2, 2)) * det;
    M(o, 1, 1) = (M(m, 0, 0) * M(m, 2, 2) - M(m, 0, 2) * M(m, 2, 0)) * det;
    M(o, 1, 2) = (M(m, 1, 0) * M(m, 0, 2) - M(m, 0, 0) * M(m, 1, 2)) * det;
    M(o, 2, 0) = (M(m, 1, 0) * M(m, 2, 1) - M(m, 2, 0) * M(m, 1, 1)) * det;
    M(o, 2, 1) = (M(m, 2, 0) * M(m, 0, 1) - M(m, 0, 0) * M(m, 2, 1)) * det;
    M(o, 2, 2) = (M(m, 0, 0) * M(m, 1, 1) - M(m, 1, 0) * M(m, 0, 1)) * det;
    return 1;
}
static void mul3(const double * a, const double * b, double * o)
{
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) M(o, r, c) = M(a, r, 0) * M(b, 0, c) + M(a, r, 1) * M(b, 1, c) + M(a, r, 2) * M(b, 2, c);
    }
}
static void mulv3(const double * m, const double * x, double * y)
{
    for (int r = 0; r < 3; ++r) y[r] = M(m, r, 0) * x[0] + M(m, r, 1) * x[1] + M(m, r, 2) * x[2];
}
static void diag3(const double * d, double * o)
{
    memset(o, 0, 9 * sizeof(double));
    M(o, 0, 0) = d[0];
    M(o, 1, 1) = d[1];
    M(o, 2, 2) = d[2];
}

/* src/colrconvert.c:88-141: primaries + white point -> RGB to XYZ, Bradford-adapted to D50 */
static int rgbToXyzD50(avifColorPrimaries cp, double * out)
{
    static const double bradford[9] = { 0.8951, 0.2664, -0.1614, -0.7502, 1.7135, 0.0367, 0.0389, -0.0685, 1.0296 };
    static const double lmsD50[3] = { 0.996284, 1.02043, 0.818644 };
    const float * p = primariesOf(cp);
    if (fabsf(p[7]) < kTiny) return 0;
    const double f = 1.0 / p[7];
    const double white[3] = { p[6] * f, 1, (1 - p[6] - p[7]) * f };
    const double prim[9] = { p[0], p[2], p[4], p[1], p[3], p[5], 1.0 - p[0] - p[1], 1.0 - p[2] - p[3], 1.0 - p[4] - p[5] };
    double primInv[9], coeff[3], coeffM[9], rgbXyz[9];
    if (!inv3(prim, primInv)) return 0;
    mulv3(primInv, white, coeff);
    diag3(coeff, coeffM);
    mul3(prim, coeffM, rgbXyz);
    double lms[3];
    mulv3(bradford, white, lms);
    for (int k = 0; k < 3; ++k) {
        if (fabs(lms[k]) < kTiny) return 0;
        lms[k] = lmsD50[k] / lms[k];
    }
    double adapt[9], tmp[9], bradfordInv[9];
    diag3(lms, adapt);
    mul3(adapt, bradford, tmp);
    if (!inv3(bradford, bradfordInv)) return 0;
    mul3(bradfordInv, tmp, adapt);
    mul3(adapt, rgbXyz, out);
    return 1;
}
/* src/colrconvert.c:157-171 */
static int rgbToRgb(avifColorPrimaries src, avifColorPrimaries dst, double * out)
{
    double a[9], b[9], bInv[9];
    if (!rgbToXyzD50(src, a) || !rgbToXyzD50(dst, b) || !inv3(b, bInv)) return 0;
    mul3(bInv, a, out);
    return 1;
}

/* Test hook (no device involved): the primaries conversion matrix the tone mapper would use, row-major.  Returns 0 when the
 * reference's avifColorPrimariesComputeRGBToRGBMatrix would fail. */
int avifB200DebugPrimariesMatrix(avifColorPrimaries src, avifColorPrimaries dst, double out[9]) { return rgbToRgb(src, dst, out); }

/* src/reformat.c:32-117 */
static int pixFmt(const avifRGBImage * rgb, AbPixFmt * f)
{
    if (!(rgb->depth == 8 || rgb->depth == 10 || rgb->depth == 12 || rgb->depth == 16)) return 0;
    if (rgb->isFloat && rgb->depth != 16) return 0;
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565 && rgb->depth != 8) return 0;
    if ((int)rgb->format < AVIF_RGB_FORMAT_RGB || rgb->format >= AVIF_RGB_FORMAT_COUNT) return 0;
    memset(f, 0, sizeof(*f));
    const uint32_t cb = (rgb->depth > 8) ? 2 : 1;
    f->chanBytes = cb;
    f->pixBytes = avifRGBImagePixelSize(rgb);
    switch (rgb->format) {
        case AVIF_RGB_FORMAT_RGB: f->offG = cb; f->offB = 2 * cb; break;
        case AVIF_RGB_FORMAT_RGBA: f->offG = cb; f->offB = 2 * cb; f->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ARGB: f->offR = cb; f->offG = 2 * cb; f->offB = 3 * cb; break;
        case AVIF_RGB_FORMAT_BGR: f->offG = cb; f->offR = 2 * cb; break;
        case AVIF_RGB_FORMAT_BGRA: f->offG = cb; f->offR = 2 * cb; f->offA = 3 * cb; break;
        case AVIF_RGB_FORMAT_ABGR: f->offB = cb; f->offG = 2 * cb; f->offR = 3 * cb; break;
        case AVIF_RGB_FORMAT_GRAYA: f->offA = cb; break; /* the colour offsets stay 0: the accessors read the gray sample three times */
        default: break; /* RGB_565, GRAY, AGRAY: every offset 0 */
    }
    f->hasAlpha = (uint8_t)avifRGBFormatHasAlpha(rgb->format);
    f->isFloat = (uint8_t)(rgb->isFloat ? 1 : 0);
    f->is565 = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_RGB_565);
    f->maxF = (float)((1 << rgb->depth) - 1);
    return 1;
}

static int onDevice(const void * p)
{
    struct cudaPointerAttributes at;
    if (!p) return 0;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

/* Scratch (converted gain map, staged base, output, tables) comes from a stream-ordered pool that keeps what it has been
 * given: the default pool hands freed memory back to the driver at every synchronisation (release threshold 0), which
 * turns each call into milliseconds of mapping work. */
int abScratchAlloc(void ** p, size_t bytes) { return abScratchAllocOn(p, bytes, AB_STREAM); }
int abScratchAllocOn(void ** p, size_t bytes, void * stream)
{
    enum { kMaxDevices = 64 };
    static cudaMemPool_t pools[kMaxDevices];
    static int made[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= kMaxDevices) return (int)cudaMallocAsync(p, bytes, (cudaStream_t)stream);
    if (!__atomic_load_n(&made[dev], __ATOMIC_ACQUIRE)) {
        static int lock = 0;
        while (__atomic_exchange_n(&lock, 1, __ATOMIC_ACQUIRE)) {
        }
        if (!made[dev]) {
            struct cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            e = cudaMemPoolCreate(&pools[dev], &props);
            if (e == cudaSuccess) {
                unsigned long long keep = ~0ull;
                e = cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
            }
            if (e == cudaSuccess) __atomic_store_n(&made[dev], 1, __ATOMIC_RELEASE);
        }
        __atomic_store_n(&lock, 0, __ATOMIC_RELEASE);
        if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaMallocFromPoolAsync(p, bytes, pools[dev], (cudaStream_t)stream);
}

#define GM_CUDA(call)                                                        \
    do {                                                                     \
        if ((call) != cudaSuccess) {                                         \
            (void)cudaGetLastError();                                        \
            diagSet(diag, "CUDA failure in %s", #call);                      \
            res = AVIF_RESULT_UNKNOWN_ERROR;                                 \
            goto cleanup;                                                    \
        }                                                                    \
    } while (0)

/* deviceOut: toneMappedImage->pixels / rowBytes are the caller's DEVICE buffer (nothing is allocated or copied back) */
static avifResult applyCommon(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                              avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap, float hdrHeadroom,
                              avifColorPrimaries outputColorPrimaries, avifTransferCharacteristics outputTransferCharacteristics,
                              avifRGBImage * toneMappedImage, avifContentLightLevelInformationBox * clli, avifDiagnostics * diag, int deviceOut)
{
    diagClear(diag);
    if (hdrHeadroom < 0.0f) {
        diagSet(diag, "hdrHeadroom should be >= 0, got %f", hdrHeadroom);
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    if (!baseImage || !gainMap || !toneMappedImage) {
        diagSet(diag, "NULL input image");
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    avifResult res = validateMetadata(gainMap, diag);
    if (res != AVIF_RESULT_OK) return res;

    const uint32_t width = baseImage->width, height = baseImage->height;
    const avifColorPrimaries mathPrimaries =
        (gainMap->useBaseColorSpace || gainMap->altColorPrimaries == AVIF_COLOR_PRIMARIES_UNSPECIFIED) ? baseColorPrimaries : gainMap->altColorPrimaries;
    const int needIn = (baseColorPrimaries != mathPrimaries), needOut = (mathPrimaries != outputColorPrimaries);

    if (deviceOut) {
        if (!toneMappedImage->pixels || toneMappedImage->width != width || toneMappedImage->height != height ||
            (size_t)toneMappedImage->rowBytes < (size_t)width * avifRGBImagePixelSize(toneMappedImage)) {
            diagSet(diag, "device output must be allocated by the caller with the base image's dimensions");
            return AVIF_RESULT_INVALID_ARGUMENT;
        }
    } else {
        toneMappedImage->width = width;
        toneMappedImage->height = height;
        res = avifRGBImageAllocatePixels(toneMappedImage);
        if (res != AVIF_RESULT_OK) return res;
    }

    uint8_t *dBase = NULL, *dGain = NULL, *dOut = NULL, *dStats = NULL, *dLut = NULL, *dScaled[3] = { NULL, NULL, NULL };
    avifImage scaled;
    const float weight = gainMapWeight(hdrHeadroom, gainMap);

    /* src/gainmap.c:120-128: nothing to apply and nothing to convert -> the base image, byte for byte */
    if (weight == 0.0f && outputTransferCharacteristics == baseTransferCharacteristics && outputColorPrimaries == baseColorPrimaries &&
        baseImage->format == toneMappedImage->format && baseImage->depth == toneMappedImage->depth && baseImage->isFloat == toneMappedImage->isFloat &&
        baseImage->rowBytes == toneMappedImage->rowBytes) {
        const size_t n = (size_t)baseImage->rowBytes * baseImage->height;
        GM_CUDA(cudaMemcpyAsync(toneMappedImage->pixels, baseImage->pixels, n, cudaMemcpyDefault, AB_STREAM)); /* either side may live on the device */
        GM_CUDA(cudaStreamSynchronize(AB_STREAM));
        return AVIF_RESULT_OK;
    }

    AbToneMapJob job;
    memset(&job, 0, sizeof(job));
    if (!pixFmt(baseImage, &job.baseFmt) || !pixFmt(toneMappedImage, &job.outFmt)) {
        diagSet(diag, "Unsupported RGB color space");
        res = AVIF_RESULT_NOT_IMPLEMENTED;
        goto cleanup;
    }
    if (!baseImage->pixels || !baseImage->rowBytes) {
        res = AVIF_RESULT_INVALID_ARGUMENT;
        goto cleanup;
    }
    job.width = width;
    job.height = height;
    job.tfIn = (uint8_t)((baseTransferCharacteristics < 256) ? baseTransferCharacteristics : 0);
    job.tfOut = (uint8_t)((outputTransferCharacteristics < 256) ? outputTransferCharacteristics : 0);
    job.weight = weight;

    if (weight == 0.0f) { /* src/gainmap.c:141-169 */
        const int differ = (baseColorPrimaries != outputColorPrimaries);
        if (differ && !rgbToRgb(baseColorPrimaries, outputColorPrimaries, job.inM)) {
            diagSet(diag, "Unsupported RGB color space conversion");
            res = AVIF_RESULT_NOT_IMPLEMENTED;
            goto cleanup;
        }
        job.applyGain = 0;
        job.transcode = (uint8_t)((outputTransferCharacteristics != baseTransferCharacteristics) || differ);
        job.inConv = (uint8_t)differ;
    } else {
        if (needIn && !rgbToRgb(baseColorPrimaries, mathPrimaries, job.inM)) {
            diagSet(diag, "Unsupported RGB color space conversion");
            res = AVIF_RESULT_NOT_IMPLEMENTED;
            goto cleanup;
        }
        if (needOut && !rgbToRgb(mathPrimaries, outputColorPrimaries, job.outM)) {
            diagSet(diag, "Unsupported RGB color space conversion");
            res = AVIF_RESULT_NOT_IMPLEMENTED;
            goto cleanup;
        }
        job.applyGain = 1;
        job.inConv = (uint8_t)needIn;
        job.outConv = (uint8_t)needOut;
        const avifImage * gmImage = gainMap->image;
        if (!gmImage) {
            res = AVIF_RESULT_INVALID_ARGUMENT;
            goto cleanup;
        }
        if (gmImage->width != width || gmImage->height != height) {
            /* src/gainmap.c:186-201: the gain map is brought to the base image's size with avifImageScale (libyuv's box /
             * bilinear integer scaler, csrc/scale.c + scale_kernels.cu) -- here straight into device planes */
            scaled = *gmImage;
            scaled.width = width;
            scaled.height = height;
            scaled.alphaPlane = NULL; /* a gain map's alpha never reaches the tone mapper */
            scaled.alphaRowBytes = 0;
            scaled.imageOwnsYUVPlanes = AVIF_FALSE;
            scaled.imageOwnsAlphaPlane = AVIF_FALSE;
            avifPixelFormatInfo fi;
            avifGetPixelFormatInfo(gmImage->yuvFormat, &fi);
            const uint32_t cs = (gmImage->depth > 8) ? 2 : 1;
            for (int c = 0; c < 3; ++c) {
                scaled.yuvPlanes[c] = NULL;
                scaled.yuvRowBytes[c] = 0;
                if (!gmImage->yuvPlanes[c] || (c > 0 && fi.monochrome)) continue;
                const uint32_t pw = c ? (uint32_t)(((uint64_t)width + fi.chromaShiftX) >> fi.chromaShiftX) : width;
                const uint32_t ph = c ? (uint32_t)(((uint64_t)height + fi.chromaShiftY) >> fi.chromaShiftY) : height;
                const size_t pitch = ((size_t)pw * cs + 255) & ~(size_t)255;
                GM_CUDA((cudaError_t)abScratchAlloc((void **)&dScaled[c], pitch * ph));
                scaled.yuvPlanes[c] = dScaled[c];
                scaled.yuvRowBytes[c] = (uint32_t)pitch;
            }
            res = avifB200ImageScaleToDevice(gmImage, &scaled);
            if (res != AVIF_RESULT_OK) {
                diagSet(diag, "scaling the gain map from %ux%u to %ux%u failed", gmImage->width, gmImage->height, width, height);
                goto cleanup;
            }
            gmImage = &scaled;
        }
        /* src/gainmap.c:203-212: the gain map as RGBA of its own depth, through the product's YUV -> RGB (device destination) */
        avifRGBImage rgbGain;
        avifRGBImageSetDefaults(&rgbGain, gmImage);
        if (!pixFmt(&rgbGain, &job.gainFmt)) {
            diagSet(diag, "Unsupported RGB color space");
            res = AVIF_RESULT_NOT_IMPLEMENTED;
            goto cleanup;
        }
        rgbGain.rowBytes = (uint32_t)(((size_t)width * job.gainFmt.pixBytes + 255) & ~(size_t)255);
        GM_CUDA((cudaError_t)abScratchAlloc((void **)&dGain, (size_t)rgbGain.rowBytes * height));
        GM_CUDA(cudaStreamSynchronize(AB_STREAM)); /* the conversion below runs on the library's own streams */
        rgbGain.pixels = dGain;
        res = avifImageYUVToRGB(gmImage, &rgbGain);
        if (res != AVIF_RESULT_OK) goto cleanup;
        job.gain = dGain;
        job.gainRB = rgbGain.rowBytes;
        for (int c = 0; c < 3; ++c) { /* src/gainmap.c:216-230 */
            job.gammaInv[c] = 1.0f / ufrac(gainMap->gainMapGamma[c]);
            job.gmMin[c] = sfrac(gainMap->gainMapMin[c]);
            job.gmMax[c] = sfrac(gainMap->gainMapMax[c]);
            job.baseOff[c] = sfrac(gainMap->baseOffset[c]);
            job.altOff[c] = sfrac(gainMap->alternateOffset[c]);
        }
    }

    /* per-code tables (tone_lut_kernel): one float per possible stored code of the base (unless half float) and three per
     * possible code of the gain map */
    job.baseLutN = (job.baseFmt.isFloat || !(job.applyGain || job.transcode)) ? 0u : ((job.baseFmt.chanBytes == 1) ? 256u : 65536u);
    job.gainLutN = job.applyGain ? ((job.gainFmt.chanBytes == 1) ? 256u : 65536u) : 0u;
    if (job.baseLutN + job.gainLutN) {
        GM_CUDA((cudaError_t)abScratchAlloc((void **)&dLut, ((size_t)job.baseLutN + 3u * (size_t)job.gainLutN) * sizeof(float)));
        job.baseLut = (float *)dLut;
        job.gainLut = (float *)dLut + job.baseLutN;
    }

    /* base pixels: used in place when already on the device */
    if (onDevice(baseImage->pixels)) {
        job.base = baseImage->pixels;
        job.baseRB = baseImage->rowBytes;
    } else {
        const size_t rowUsed = (size_t)width * job.baseFmt.pixBytes;
        const size_t pitch = (rowUsed + 255) & ~(size_t)255;
        GM_CUDA((cudaError_t)abScratchAlloc((void **)&dBase, pitch * height));
        GM_CUDA(cudaMemcpy2DAsync(dBase, pitch, baseImage->pixels, baseImage->rowBytes, rowUsed, height, cudaMemcpyHostToDevice, AB_STREAM));
        job.base = dBase;
        job.baseRB = (int64_t)pitch;
    }
    {
        const size_t rowUsed = (size_t)width * job.outFmt.pixBytes;
        const size_t pitch = (rowUsed + 255) & ~(size_t)255;
        if (deviceOut) {
            job.out = toneMappedImage->pixels;
            job.outRB = toneMappedImage->rowBytes;
        } else {
            GM_CUDA((cudaError_t)abScratchAlloc((void **)&dOut, pitch * height));
            job.out = dOut;
            job.outRB = (int64_t)pitch;
        }
        struct { double sum; uint32_t maxBits, nanIndex; } stats = { 0.0, 0u, 0xFFFFFFFFu };
        GM_CUDA((cudaError_t)abScratchAlloc((void **)&dStats, sizeof(stats)));
        GM_CUDA(cudaMemcpyAsync(dStats, &stats, sizeof(stats), cudaMemcpyHostToDevice, AB_STREAM));
        job.sum = (double *)dStats;
        job.maxBits = (uint32_t *)(dStats + offsetof(__typeof__(stats), maxBits));
        job.nanIndex = (uint32_t *)(dStats + offsetof(__typeof__(stats), nanIndex));
        GM_CUDA((cudaError_t)abLaunchToneMap(&job, AB_STREAM));
        GM_CUDA(cudaMemcpyAsync(&stats, dStats, sizeof(stats), cudaMemcpyDeviceToHost, AB_STREAM));
        GM_CUDA(cudaStreamSynchronize(AB_STREAM)); /* the launch and the read-back */
        if (job.applyGain && stats.nanIndex != 0xFFFFFFFFu) { /* src/gainmap.c:283-288 */
            diagSet(diag, "Degenerate gain map parameters produce NaN at pixel (%u, %u)", stats.nanIndex % width, stats.nanIndex / width);
            res = AVIF_RESULT_INVALID_TONE_MAPPED_IMAGE;
            goto cleanup;
        }
        if (!deviceOut) {
            GM_CUDA(cudaMemcpy2DAsync(toneMappedImage->pixels, toneMappedImage->rowBytes, dOut, pitch, rowUsed, height, cudaMemcpyDeviceToHost, AB_STREAM));
            GM_CUDA(cudaStreamSynchronize(AB_STREAM));
        }
        if (clli && job.applyGain) { /* src/gainmap.c:297-306 */
            float rgbMaxLinear;
            memcpy(&rgbMaxLinear, &stats.maxBits, sizeof(rgbMaxLinear));
            float v = floorf(rgbMaxLinear * 203.0f + 0.5f);
            v = (v < 0.0f) ? 0.0f : ((65535.0f < v) ? 65535.0f : v);
            clli->maxCLL = (uint16_t)v;
            /* The reference adds the per-pixel maxima one by one in float; here they are summed in double (every partial sum
             * exact to ~1e-16), so maxPALL can differ by the reference's own accumulated rounding on large images. */
            const float avg = (float)(stats.sum / (double)((size_t)width * height));
            v = floorf(avg * 203.0f + 0.5f);
            v = (v < 0.0f) ? 0.0f : ((65535.0f < v) ? 65535.0f : v);
            clli->maxPALL = (uint16_t)v;
        }
    }
    res = AVIF_RESULT_OK;

cleanup:
    if (dBase) (void)cudaFreeAsync(dBase, AB_STREAM);
    if (dGain) (void)cudaFreeAsync(dGain, AB_STREAM);
    if (dOut) (void)cudaFreeAsync(dOut, AB_STREAM);
    if (dStats) (void)cudaFreeAsync(dStats, AB_STREAM);
    if (dLut) (void)cudaFreeAsync(dLut, AB_STREAM);
    for (int c = 0; c < 3; ++c) {
        if (dScaled[c]) (void)cudaFreeAsync(dScaled[c], AB_STREAM);
    }
    return res;
}

/* applyCommon on the device the call belongs to (shim.c: device policy; resident buffers pin it) */
static avifResult applyScoped(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                              avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap, float hdrHeadroom,
                              avifColorPrimaries outputColorPrimaries, avifTransferCharacteristics outputTransferCharacteristics,
                              avifRGBImage * toneMappedImage, avifContentLightLevelInformationBox * clli, avifDiagnostics * diag, int deviceOut)
{
    const void * ptrs[3] = { (deviceOut && toneMappedImage) ? toneMappedImage->pixels : NULL, baseImage ? baseImage->pixels : NULL,
                             (gainMap && gainMap->image) ? gainMap->image->yuvPlanes[0] : NULL };
    int prev = -1;
    avifResult res = (avifResult)abDeviceEnter(ptrs, 3, &prev);
    if (res != AVIF_RESULT_OK) return res;
    res = applyCommon(baseImage, baseColorPrimaries, baseTransferCharacteristics, gainMap, hdrHeadroom, outputColorPrimaries,
                      outputTransferCharacteristics, toneMappedImage, clli, diag, deviceOut);
    abDeviceLeave(prev);
    return res;
}

avifResult avifRGBImageApplyGainMap(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                                    avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap, float hdrHeadroom,
                                    avifColorPrimaries outputColorPrimaries, avifTransferCharacteristics outputTransferCharacteristics,
                                    avifRGBImage * toneMappedImage, avifContentLightLevelInformationBox * clli, avifDiagnostics * diag)
{
    return applyScoped(baseImage, baseColorPrimaries, baseTransferCharacteristics, gainMap, hdrHeadroom, outputColorPrimaries,
                       outputTransferCharacteristics, toneMappedImage, clli, diag, 0);
}
/* extension: everything stays on the device -- base pixels, the gain map's planes and the caller-allocated output */
avifResult avifB200RGBImageApplyGainMapDevice(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                                              avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap,
                                              float hdrHeadroom, avifColorPrimaries outputColorPrimaries,
                                              avifTransferCharacteristics outputTransferCharacteristics, avifRGBImage * toneMappedImage,
                                              avifContentLightLevelInformationBox * clli, avifDiagnostics * diag)
{
    return applyScoped(baseImage, baseColorPrimaries, baseTransferCharacteristics, gainMap, hdrHeadroom, outputColorPrimaries,
                       outputTransferCharacteristics, toneMappedImage, clli, diag, 1);
}

/* src/gainmap.c:317-360 */
avifResult avifImageApplyGainMap(const avifImage * baseImage, const avifGainMap * gainMap, float hdrHeadroom, avifColorPrimaries outputColorPrimaries,
                                 avifTransferCharacteristics outputTransferCharacteristics, avifRGBImage * toneMappedImage,
                                 avifContentLightLevelInformationBox * clli, avifDiagnostics * diag)
{
    diagClear(diag);
    if (baseImage->icc.size > 0 || gainMap->altICC.size > 0) {
        diagSet(diag, "Tone mapping for images with ICC profiles is not supported");
        return AVIF_RESULT_NOT_IMPLEMENTED;
    }
    avifRGBImage baseRgb;
    avifRGBImageSetDefaults(&baseRgb, baseImage);
    /* the intermediate RGB rendition never leaves the device; rows packed like avifRGBImageAllocatePixels packs them, so the
     * "same layout" early exit of the RGB entry point triggers exactly when the reference's does */
    const uint32_t pixBytes = avifRGBImagePixelSize(&baseRgb);
    if (!baseImage->width || !baseImage->height) return AVIF_RESULT_INVALID_ARGUMENT;
    baseRgb.rowBytes = baseImage->width * pixBytes;
    uint8_t * dev = NULL;
    const void * ptrs[2] = { baseImage->yuvPlanes[0], (gainMap && gainMap->image) ? gainMap->image->yuvPlanes[0] : NULL };
    int prevDevice = -1;
    {
        const avifResult er = (avifResult)abDeviceEnter(ptrs, 2, &prevDevice);
        if (er != AVIF_RESULT_OK) return er;
    }
    if ((cudaError_t)abScratchAlloc((void **)&dev, (size_t)baseRgb.rowBytes * baseImage->height) != cudaSuccess || cudaStreamSynchronize(AB_STREAM) != cudaSuccess) {
        (void)cudaGetLastError();
        abDeviceLeave(prevDevice);
        return AVIF_RESULT_OUT_OF_MEMORY;
    }
    baseRgb.pixels = dev;
    avifResult res = avifImageYUVToRGB(baseImage, &baseRgb);
    if (res == AVIF_RESULT_OK) {
        res = avifRGBImageApplyGainMap(&baseRgb, baseImage->colorPrimaries, baseImage->transferCharacteristics, gainMap, hdrHeadroom,
                                       outputColorPrimaries, outputTransferCharacteristics, toneMappedImage, clli, diag);
    }
    (void)cudaFreeAsync(dev, AB_STREAM);
    abDeviceLeave(prevDevice);
    return res;
}
