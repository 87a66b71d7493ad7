/*
 * avif_b200.h -- C ABI of the B200-native pixel-reformat + alpha path.
 *
 * This library replaces ONE path of libavif (the reference): RGB<->YUV reformat,
 * alpha reformat and alpha (un)premultiply.  Every entry point below is the
 * symbol a libavif caller already binds; the struct layouts and enum values are
 * ABI-compatible with the reference's public header so that a program compiled
 * against the reference's <avif/avif.h> can link (or LD_PRELOAD) this library
 * for these calls without recompiling.  Layout compatibility is pinned by
 * tests/test_abi_layout.py (compiled against the reference header when present).
 *
 * Reference interface replaced (file:line relative to the reference tree):
 *   avifImageYUVToRGB               include/avif/avif.h:1040   src/reformat.c:1649
 *   avifImageRGBToYUV               include/avif/avif.h:1039   src/reformat.c:221
 *   avifRGBImagePremultiplyAlpha    include/avif/avif.h:1045   src/alpha.c:151
 *   avifRGBImageUnpremultiplyAlpha  include/avif/avif.h:1046   src/alpha.c:338
 *   avifReformatAlpha/avifFillAlpha include/avif/internal.h:277-278  src/alpha.c:9,37
 *   avif{FullToLimited,LimitedToFull}{Y,UV}  include/avif/avif.h:1051-1054  src/reformat.c:1778-1840
 *   avifRGBImage helpers            include/avif/avif.h:971-973,1031-1036  src/avif.c:670-760
 *   avifLibYUVVersion               include/avif/avif.h:152   (returns 0: no libyuv here)
 *
 * Pointers in avifImage / avifRGBImage may be HOST pointers (the reference's
 * contract; data is staged over PCIe) or DEVICE pointers (resident planes; no
 * copies) -- the library classifies every plane pointer through CUDA UVA.
 * There is no CPU fallback: without a usable CUDA device every conversion entry
 * point returns AVIF_RESULT_UNKNOWN_ERROR.
 *
 * If the reference's own <avif/avif.h> was included first, its definitions are
 * used and only the avifB200* extensions are declared here.
 */
#ifndef AVIF_B200_H
#define AVIF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(AVIF_B200_BUILDING)
#define AVIF_B200_API __attribute__((visibility("default")))
#else
#define AVIF_B200_API
#endif

#ifndef AVIF_AVIF_H /* reference header not included: provide ABI-identical definitions */

typedef int avifBool;
#define AVIF_TRUE 1
#define AVIF_FALSE 0

#define AVIF_PLANE_COUNT_YUV 3
enum { AVIF_CHAN_Y = 0, AVIF_CHAN_U = 1, AVIF_CHAN_V = 2, AVIF_CHAN_A = 3 };
enum { AVIF_PLANES_YUV = 1 << 0, AVIF_PLANES_A = 1 << 1, AVIF_PLANES_ALL = 0xff };
typedef uint32_t avifPlanesFlags;

/* Only the result codes this path can return are named; values match avif.h:163-203. */
typedef enum avifResult {
    AVIF_RESULT_OK = 0,
    AVIF_RESULT_UNKNOWN_ERROR = 1,
    AVIF_RESULT_REFORMAT_FAILED = 5,
    AVIF_RESULT_INVALID_ARGUMENT = 24,
    AVIF_RESULT_NOT_IMPLEMENTED = 25,
    AVIF_RESULT_OUT_OF_MEMORY = 26,
    AVIF_RESULT_INTERNAL_ERROR = 29,
    AVIF_RESULT_INVALID_TONE_MAPPED_IMAGE = 32
} avifResult;

typedef enum avifPixelFormat {
    AVIF_PIXEL_FORMAT_NONE = 0,
    AVIF_PIXEL_FORMAT_YUV444,
    AVIF_PIXEL_FORMAT_YUV422,
    AVIF_PIXEL_FORMAT_YUV420,
    AVIF_PIXEL_FORMAT_YUV400,
    AVIF_PIXEL_FORMAT_COUNT
} avifPixelFormat;

typedef struct avifPixelFormatInfo {
    avifBool monochrome;
    int chromaShiftX;
    int chromaShiftY;
} avifPixelFormatInfo;

typedef enum avifChromaSamplePosition {
    AVIF_CHROMA_SAMPLE_POSITION_UNKNOWN = 0,
    AVIF_CHROMA_SAMPLE_POSITION_VERTICAL = 1,
    AVIF_CHROMA_SAMPLE_POSITION_COLOCATED = 2,
    AVIF_CHROMA_SAMPLE_POSITION_RESERVED = 3
} avifChromaSamplePosition;

typedef enum avifRange { AVIF_RANGE_LIMITED = 0, AVIF_RANGE_FULL = 1 } avifRange;

/* CICP code points (ITU-T H.273) used by this path. */
enum {
    AVIF_COLOR_PRIMARIES_UNKNOWN = 0,
    AVIF_COLOR_PRIMARIES_BT709 = 1,
    AVIF_COLOR_PRIMARIES_UNSPECIFIED = 2,
    AVIF_COLOR_PRIMARIES_BT470M = 4,
    AVIF_COLOR_PRIMARIES_BT470BG = 5,
    AVIF_COLOR_PRIMARIES_BT601 = 6,
    AVIF_COLOR_PRIMARIES_SMPTE240 = 7,
    AVIF_COLOR_PRIMARIES_GENERIC_FILM = 8,
    AVIF_COLOR_PRIMARIES_BT2020 = 9,
    AVIF_COLOR_PRIMARIES_XYZ = 10,
    AVIF_COLOR_PRIMARIES_SMPTE431 = 11,
    AVIF_COLOR_PRIMARIES_SMPTE432 = 12,
    AVIF_COLOR_PRIMARIES_EBU3213 = 22
};
typedef uint16_t avifColorPrimaries;
typedef uint16_t avifTransferCharacteristics;
enum {
    AVIF_MATRIX_COEFFICIENTS_IDENTITY = 0,
    AVIF_MATRIX_COEFFICIENTS_BT709 = 1,
    AVIF_MATRIX_COEFFICIENTS_UNSPECIFIED = 2,
    AVIF_MATRIX_COEFFICIENTS_FCC = 4,
    AVIF_MATRIX_COEFFICIENTS_BT470BG = 5,
    AVIF_MATRIX_COEFFICIENTS_BT601 = 6,
    AVIF_MATRIX_COEFFICIENTS_SMPTE240 = 7,
    AVIF_MATRIX_COEFFICIENTS_YCGCO = 8,
    AVIF_MATRIX_COEFFICIENTS_BT2020_NCL = 9,
    AVIF_MATRIX_COEFFICIENTS_BT2020_CL = 10,
    AVIF_MATRIX_COEFFICIENTS_SMPTE2085 = 11,
    AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_NCL = 12,
    AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_CL = 13,
    AVIF_MATRIX_COEFFICIENTS_ICTCP = 14,
    AVIF_MATRIX_COEFFICIENTS_YCGCO_RE = 16,
    AVIF_MATRIX_COEFFICIENTS_YCGCO_RO = 17,
    AVIF_MATRIX_COEFFICIENTS_LAST
};
typedef uint16_t avifMatrixCoefficients;

/* Members this path never touches are kept only so that sizeof/offsetof match avif.h:777-851. */
typedef struct avifRWData { uint8_t * data; size_t size; } avifRWData;
typedef struct avifContentLightLevelInformationBox { uint16_t maxCLL, maxPALL; } avifContentLightLevelInformationBox;
typedef uint32_t avifTransformFlags;
typedef struct avifPixelAspectRatioBox { uint32_t hSpacing, vSpacing; } avifPixelAspectRatioBox;
typedef struct avifCleanApertureBox {
    uint32_t widthN, widthD, heightN, heightD, horizOffN, horizOffD, vertOffN, vertOffD;
} avifCleanApertureBox;
typedef struct avifImageRotation { uint8_t angle; } avifImageRotation;
typedef struct avifImageMirror { uint8_t axis; } avifImageMirror;
typedef struct avifCropRect { uint32_t x, y, width, height; } avifCropRect;
struct avifImageItemProperty;
struct avifGainMap;

typedef struct avifImage {
    uint32_t width;
    uint32_t height;
    uint32_t depth; /* 8, 10, 12 or 16; samples are uint16_t (low bits) when depth > 8 */

    avifPixelFormat yuvFormat;
    avifRange yuvRange;
    avifChromaSamplePosition yuvChromaSamplePosition;
    uint8_t * yuvPlanes[AVIF_PLANE_COUNT_YUV];
    uint32_t yuvRowBytes[AVIF_PLANE_COUNT_YUV];
    avifBool imageOwnsYUVPlanes;

    uint8_t * alphaPlane;
    uint32_t alphaRowBytes;
    avifBool imageOwnsAlphaPlane;
    avifBool alphaPremultiplied;

    avifRWData icc;

    avifColorPrimaries colorPrimaries;
    avifTransferCharacteristics transferCharacteristics;
    avifMatrixCoefficients matrixCoefficients;

    avifContentLightLevelInformationBox clli;
    avifTransformFlags transformFlags;
    avifPixelAspectRatioBox pasp;
    avifCleanApertureBox clap;
    avifImageRotation irot;
    avifImageMirror imir;
    avifRWData exif;
    avifRWData xmp;
    struct avifImageItemProperty * properties;
    size_t numProperties;
    struct avifGainMap * gainMap;
} avifImage;

typedef enum avifRGBFormat {
    AVIF_RGB_FORMAT_RGB = 0,
    AVIF_RGB_FORMAT_RGBA,
    AVIF_RGB_FORMAT_ARGB,
    AVIF_RGB_FORMAT_BGR,
    AVIF_RGB_FORMAT_BGRA,
    AVIF_RGB_FORMAT_ABGR,
    AVIF_RGB_FORMAT_RGB_565, /* YUV->RGB only, depth 8 */
    AVIF_RGB_FORMAT_GRAY,
    AVIF_RGB_FORMAT_GRAYA,
    AVIF_RGB_FORMAT_AGRAY,
    AVIF_RGB_FORMAT_COUNT
} avifRGBFormat;

typedef enum avifChromaUpsampling {
    AVIF_CHROMA_UPSAMPLING_AUTOMATIC = 0,    /* bilinear (no libyuv in this build) */
    AVIF_CHROMA_UPSAMPLING_FASTEST = 1,      /* nearest */
    AVIF_CHROMA_UPSAMPLING_BEST_QUALITY = 2, /* bilinear */
    AVIF_CHROMA_UPSAMPLING_NEAREST = 3,
    AVIF_CHROMA_UPSAMPLING_BILINEAR = 4
} avifChromaUpsampling;

typedef enum avifChromaDownsampling {
    AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC = 0,
    AVIF_CHROMA_DOWNSAMPLING_FASTEST = 1,
    AVIF_CHROMA_DOWNSAMPLING_BEST_QUALITY = 2,
    AVIF_CHROMA_DOWNSAMPLING_AVERAGE = 3,
    AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV = 4 /* 4:2:0: NOT_IMPLEMENTED (no libsharpyuv), as reformat.c:256-261 */
} avifChromaDownsampling;

typedef struct avifRGBImage {
    uint32_t width;
    uint32_t height;
    uint32_t depth; /* 8, 10, 12, 16 */
    avifRGBFormat format;
    avifChromaUpsampling chromaUpsampling;
    avifChromaDownsampling chromaDownsampling;
    avifBool avoidLibYUV; /* accepted, ignored: results are always the built-in (precise) path's */
    avifBool ignoreAlpha;
    avifBool alphaPremultiplied;
    avifBool isFloat; /* half-float output, depth must be 16 */
    int maxThreads;   /* hint only; negative values are rejected like the reference */

    uint8_t * pixels;
    uint32_t rowBytes;
} avifRGBImage;

/* ---- the reference's public entry points for this path (same names, same semantics) ---- */
AVIF_B200_API avifResult avifImageYUVToRGB(const avifImage * image, avifRGBImage * rgb);
AVIF_B200_API avifResult avifImageRGBToYUV(avifImage * image, const avifRGBImage * rgb);
AVIF_B200_API avifResult avifRGBImagePremultiplyAlpha(avifRGBImage * rgb);
AVIF_B200_API avifResult avifRGBImageUnpremultiplyAlpha(avifRGBImage * rgb);

AVIF_B200_API int avifFullToLimitedY(uint32_t depth, int v);
AVIF_B200_API int avifFullToLimitedUV(uint32_t depth, int v);
AVIF_B200_API int avifLimitedToFullY(uint32_t depth, int v);
AVIF_B200_API int avifLimitedToFullUV(uint32_t depth, int v);

AVIF_B200_API uint32_t avifRGBFormatChannelCount(avifRGBFormat format);
AVIF_B200_API avifBool avifRGBFormatHasAlpha(avifRGBFormat format);
AVIF_B200_API avifBool avifRGBFormatIsGray(avifRGBFormat format);
AVIF_B200_API void avifRGBImageSetDefaults(avifRGBImage * rgb, const avifImage * image);
AVIF_B200_API uint32_t avifRGBImagePixelSize(const avifRGBImage * rgb);
AVIF_B200_API avifResult avifRGBImageAllocatePixels(avifRGBImage * rgb);
AVIF_B200_API void avifRGBImageFreePixels(avifRGBImage * rgb);
AVIF_B200_API void avifGetPixelFormatInfo(avifPixelFormat format, avifPixelFormatInfo * info);
AVIF_B200_API unsigned int avifLibYUVVersion(void);

/* ---- gain-map tone mapping: ABI-identical subset of avif.h:366-387, :419-453, :630-712, :1727-1745 ---- */
enum {
    AVIF_TRANSFER_CHARACTERISTICS_UNKNOWN = 0,
    AVIF_TRANSFER_CHARACTERISTICS_BT709 = 1,
    AVIF_TRANSFER_CHARACTERISTICS_UNSPECIFIED = 2,
    AVIF_TRANSFER_CHARACTERISTICS_BT470M = 4,
    AVIF_TRANSFER_CHARACTERISTICS_BT470BG = 5,
    AVIF_TRANSFER_CHARACTERISTICS_BT601 = 6,
    AVIF_TRANSFER_CHARACTERISTICS_SMPTE240 = 7,
    AVIF_TRANSFER_CHARACTERISTICS_LINEAR = 8,
    AVIF_TRANSFER_CHARACTERISTICS_LOG100 = 9,
    AVIF_TRANSFER_CHARACTERISTICS_LOG100_SQRT10 = 10,
    AVIF_TRANSFER_CHARACTERISTICS_IEC61966 = 11,
    AVIF_TRANSFER_CHARACTERISTICS_BT1361 = 12,
    AVIF_TRANSFER_CHARACTERISTICS_SRGB = 13,
    AVIF_TRANSFER_CHARACTERISTICS_BT2020_10BIT = 14,
    AVIF_TRANSFER_CHARACTERISTICS_BT2020_12BIT = 15,
    AVIF_TRANSFER_CHARACTERISTICS_PQ = 16,
    AVIF_TRANSFER_CHARACTERISTICS_SMPTE2084 = 16,
    AVIF_TRANSFER_CHARACTERISTICS_SMPTE428 = 17,
    AVIF_TRANSFER_CHARACTERISTICS_HLG = 18
};
#define AVIF_DIAGNOSTICS_ERROR_BUFFER_SIZE 256
typedef struct avifDiagnostics { char error[AVIF_DIAGNOSTICS_ERROR_BUFFER_SIZE]; } avifDiagnostics;
typedef struct avifSignedFraction { int32_t n; uint32_t d; } avifSignedFraction;
typedef struct avifUnsignedFraction { uint32_t n; uint32_t d; } avifUnsignedFraction;
typedef struct avifGainMap {
    struct avifImage * image; /* gain-map pixels (YUV); metadata below interprets them */
    avifSignedFraction gainMapMin[3];
    avifSignedFraction gainMapMax[3];
    avifUnsignedFraction gainMapGamma[3];
    avifSignedFraction baseOffset[3];
    avifSignedFraction alternateOffset[3];
    avifUnsignedFraction baseHdrHeadroom;
    avifUnsignedFraction alternateHdrHeadroom;
    avifBool useBaseColorSpace;
    avifRWData altICC;
    avifColorPrimaries altColorPrimaries;
    avifTransferCharacteristics altTransferCharacteristics;
    avifMatrixCoefficients altMatrixCoefficients;
    avifRange altYUVRange;
    uint32_t altDepth;
    uint32_t altPlaneCount;
    avifContentLightLevelInformationBox altCLLI;
} avifGainMap;

/* src/gainmap.c:73-315 and :317-360: tone-maps baseImage with gainMap for a display of the given log2 HDR headroom.
 * Allocates toneMappedImage->pixels (width/height/rowBytes set here; format, depth, isFloat are the caller's request).
 * Arithmetic is binary32 with libm's powf/exp2f/logf in the reference and CUDA's on the device (evaluated in double and rounded
 * once wherever a 16-bit code depends on the last ulp): results agree to +-1 output code, not bit for bit.
 * clli->maxCLL follows the reference exactly up to that tolerance; clli->maxPALL is the average of the per-pixel maxima summed in
 * DOUBLE precision, where the reference accumulates them sequentially in a float (src/gainmap.c:213,290) -- on multi-megapixel
 * images the reference's running sum loses most of every addend, so its maxPALL can sit several nits below this (more accurate)
 * value; the two agree to +-1 nit on images up to ~1 megapixel.  A gain map of another size than the base image is rescaled
 * first with avifImageScale's arithmetic, like src/gainmap.c:186-201. */
AVIF_B200_API avifResult avifRGBImageApplyGainMap(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                                                  avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap,
                                                  float hdrHeadroom, avifColorPrimaries outputColorPrimaries,
                                                  avifTransferCharacteristics outputTransferCharacteristics, avifRGBImage * toneMappedImage,
                                                  avifContentLightLevelInformationBox * clli, avifDiagnostics * diag);
AVIF_B200_API avifResult avifImageApplyGainMap(const avifImage * baseImage, const avifGainMap * gainMap, float hdrHeadroom,
                                               avifColorPrimaries outputColorPrimaries,
                                               avifTransferCharacteristics outputTransferCharacteristics, avifRGBImage * toneMappedImage,
                                               avifContentLightLevelInformationBox * clli, avifDiagnostics * diag);
/* src/scale.c:203-207: scales the YUV / alpha planes in place (new planes, owned by the image) with the vendored libyuv's
 * kFilterBox arithmetic (third_party/libyuv/source/scale*.c), bit for bit.  Host planes. */
AVIF_B200_API avifResult avifImageScale(avifImage * image, uint32_t dstWidth, uint32_t dstHeight, avifDiagnostics * diag);
#endif /* AVIF_AVIF_H */

/* ---- internal seam of the reference (include/avif/internal.h:259-278), exported for drop-in at B-2 ---- */
#ifndef AVIF_INTERNAL_H
typedef struct avifAlphaParams {
    uint32_t width;
    uint32_t height;

    uint32_t srcDepth;
    const uint8_t * srcPlane;
    uint32_t srcRowBytes;
    uint32_t srcOffsetBytes;
    uint32_t srcPixelBytes;

    uint32_t dstDepth;
    uint8_t * dstPlane;
    uint32_t dstRowBytes;
    uint32_t dstOffsetBytes;
    uint32_t dstPixelBytes;
} avifAlphaParams;
AVIF_B200_API void avifFillAlpha(const avifAlphaParams * params);
AVIF_B200_API void avifReformatAlpha(const avifAlphaParams * params);

/* Colour-space descriptions and the per-pixel float accessors (include/avif/internal.h:281-348; src/reformat.c:32-159 and
 * :1856-1939).  Host-side scalar helpers of the reference's internal API: what its gain-map code (src/gainmap.c) calls around the
 * conversions.  Exported so that a reference build can drop src/reformat.c and src/alpha.c entirely and link this library instead
 * (seam 1, link flavour).  They touch HOST memory one pixel at a time by definition; the hot path never uses them. */
typedef enum avifReformatMode {
    AVIF_REFORMAT_MODE_YUV_COEFFICIENTS = 0,
    AVIF_REFORMAT_MODE_IDENTITY,
    AVIF_REFORMAT_MODE_YCGCO,
    AVIF_REFORMAT_MODE_YCGCO_RE,
    AVIF_REFORMAT_MODE_YCGCO_RO
} avifReformatMode;
typedef enum avifAlphaMultiplyMode {
    AVIF_ALPHA_MULTIPLY_MODE_NO_OP = 0,
    AVIF_ALPHA_MULTIPLY_MODE_MULTIPLY,
    AVIF_ALPHA_MULTIPLY_MODE_UNMULTIPLY
} avifAlphaMultiplyMode;
typedef struct avifRGBColorSpaceInfo {
    uint32_t channelBytes;
    uint32_t pixelBytes;
    uint32_t offsetBytesR;
    uint32_t offsetBytesG;
    uint32_t offsetBytesB;
    uint32_t offsetBytesA;
    uint32_t offsetBytesGray;
    int maxChannel;
    float maxChannelF;
} avifRGBColorSpaceInfo;
typedef struct avifYUVColorSpaceInfo {
    float kr;
    float kg;
    float kb;
    uint32_t channelBytes;
    uint32_t depth;
    avifRange range;
    int maxChannel;
    float biasY;
    float biasUV;
    float rangeY;
    float rangeUV;
    avifPixelFormatInfo formatInfo;
    avifReformatMode mode;
} avifYUVColorSpaceInfo;
AVIF_B200_API avifBool avifGetRGBColorSpaceInfo(const avifRGBImage * rgb, avifRGBColorSpaceInfo * info);
AVIF_B200_API avifBool avifGetYUVColorSpaceInfo(const avifImage * image, avifYUVColorSpaceInfo * info);
AVIF_B200_API void avifGetRGBAPixel(const avifRGBImage * src, uint32_t x, uint32_t y, const avifRGBColorSpaceInfo * info, float rgbaPixel[4]);
AVIF_B200_API void avifSetRGBAPixel(const avifRGBImage * dst, uint32_t x, uint32_t y, const avifRGBColorSpaceInfo * info, const float rgbaPixel[4]);
#endif

/* ---- B200 extensions (not in the reference) ----------------------------------------------- */

/* Stream-ordered variants: every plane / pixel pointer MUST be a device pointer on the current
 * device; nothing is copied, nothing is synchronised; the kernels are enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream).  Same validation and result
 * codes as the synchronous entry points.  avifB200ImageRGBToYUVAsync requires the destination
 * planes to exist already (it cannot malloc device memory behind the caller's back). */
AVIF_B200_API avifResult avifB200ImageYUVToRGBAsync(const avifImage * image, avifRGBImage * rgb, void * stream);
AVIF_B200_API avifResult avifB200ImageRGBToYUVAsync(avifImage * image, const avifRGBImage * rgb, void * stream);
AVIF_B200_API avifResult avifB200RGBImagePremultiplyAlphaAsync(avifRGBImage * rgb, void * stream);
AVIF_B200_API avifResult avifB200RGBImageUnpremultiplyAlphaAsync(avifRGBImage * rgb, void * stream);

/* Batches of device-resident frames (config C5 of BASELINE.json: batched decode-reformat).  Frames that share format,
 * depth, matrix, range and alpha handling are converted by ONE kernel launch per <= 24 frames: a launch costs ~4-5 us of
 * GPU ramp-up / drain, i.e. 15 % of a 4K frame and more than half of a 1080p one.  Same validation and result codes as
 * the per-image calls (the first failing image's code is returned; nothing is launched in that case). */
AVIF_B200_API avifResult avifB200ImageYUVToRGBBatchAsync(const avifImage * const * images, avifRGBImage * const * rgbs, uint32_t count, void * stream);
AVIF_B200_API avifResult avifB200ImageRGBToYUVBatchAsync(avifImage * const * images, const avifRGBImage * const * rgbs, uint32_t count, void * stream);

/* The synchronous entry points under names that do not clash with the reference's, for the internal-hook integration
 * (seam 2) where libavif keeps its own avifImageYUVToRGB and calls these from its accelerator hooks. */
AVIF_B200_API avifResult avifB200ImageYUVToRGBHost(const avifImage * image, avifRGBImage * rgb);
AVIF_B200_API avifResult avifB200ImageRGBToYUVHost(avifImage * image, const avifRGBImage * rgb);
AVIF_B200_API avifResult avifB200RGBImagePremultiplyAlphaHost(avifRGBImage * rgb);
AVIF_B200_API avifResult avifB200RGBImageUnpremultiplyAlphaHost(avifRGBImage * rgb);

/* Range conversion of a whole alpha plane (reference: avifImageLimitedToFullAlpha, src/read.c:6724-6764,
 * integer avifLimitedToFullY per sample).  Host or device pointers. */
AVIF_B200_API avifResult avifB200ImageLimitedToFullAlpha(avifImage * image);

/* Grid plane staging (reference: avifImageCopySamples src/avif.c:187-225 as used by
 * avifDecoderDataCopyTileToImage src/read.c:1823-1878, and avifImageCopyAndPad src/write.c:1151-1205).
 * Device pointers, stream-ordered. */
AVIF_B200_API avifResult avifB200ImageCopySamplesAsync(avifImage * dst, const avifImage * src, avifPlanesFlags planes, void * stream);
AVIF_B200_API avifResult avifB200ImageCopyAndPadAsync(avifImage * dst, const avifImage * src, void * stream);

/* Rows [rowBegin, rowEnd) of a resident frame (a tile / band pipeline converts rows as they arrive).  image and rgb describe the
 * WHOLE frame: edge replication of the bilinear filter happens at the real picture edges only (src/reformat.c:768-795).  Of the
 * planes only the rows the range touches need to exist behind the pointers -- luma / alpha rows rowBegin..rowEnd-1, chroma rows
 * (rowBegin >> shiftY) - 1 .. ((rowEnd - 1) >> shiftY) + 1 clipped to the plane -- so base pointers may be "virtual" (band buffer
 * minus rowBegin * rowBytes).  rowBegin must be even for 4:2:0.  Device pointers, stream-ordered. */
AVIF_B200_API avifResult avifB200ImageYUVToRGBRowsAsync(const avifImage * image, avifRGBImage * rgb, uint32_t rowBegin, uint32_t rowEnd, void * stream);

/* ---- tile grids across GPUs (BASELINE config 4; reference: avifImageSplitGrid apps/shared/avifutil.c:546-605,
 * avifEncoderAddImageGrid src/write.c:2147, tile copy + canvas conversion src/read.c:1823-1878, :6777-6866) --------------------
 * One process per GPU.  `comm` is an ncclComm_t (void *) over the participating ranks, or NULL for a single GPU; the library
 * reaches NCCL through dlopen("libnccl.so.2") and answers AVIF_RESULT_NOT_IMPLEMENTED where it is absent.  Every rank calls with
 * the same descriptors (dimensions, formats, flags); plane / pixel pointers are device memory on `root` and unused elsewhere
 * (for avifB200GridYUVToRGB the row strides tell the other ranks which planes exist).  All work is enqueued on `stream`.
 *   avifB200GridCellRect      cell `cellIndex` (row-major) of the reference's split rule; INVALID_ARGUMENT where the rule fails
 *   avifB200GridRGBToYUV      cells go to rank (root + cell) mod nranks, ONE batched launch per rank.  Peer-memory form (default): every
 *                             rank maps the root's canvas through CUDA IPC and converts views of it -- the kernel's loads and stores
 *                             cross NVLink themselves, nothing is staged; NCCL carries the canvas descriptor and one completion
 *                             token per rank.  Staged form (canvas memory that cannot be exported, AVIF_B200_GRID_PEER=0): grouped
 *                             ncclSend/ncclRecv of the cells' pixels, exchange of the planes back, placement into the canvas planes.
 *                             Either way bit-identical to converting the whole canvas (cell sizes are even when chroma is
 *                             subsampled).  phaseMs (may be NULL; synchronises when given), milliseconds on this rank: scatter,
 *                             convert, gather, assemble (staged) or descriptor, convert over NVLink, tokens, 0 (peer memory).
 *   avifB200GridLastMode      1 if the calling thread's last avifB200Grid* conversion ran in the peer-memory form, else 0.
 *   avifB200GridYUVToRGB      decode side: the canvas is cut into nranks even row bands, each converted against whole-canvas geometry
 *                             with one chroma row of halo per inner seam (bilinear upsampling is not tile-local); bands and results
 *                             travel by grouped ncclSend/ncclRecv.  AVIF_B200_GRID_PEER_DECODE=1 (every rank alike) converts straight
 *                             from and into the root's peer-mapped canvas instead -- measured slower (strided 16-byte pixel stores
 *                             over NVLink), kept as a switch.  phaseMs: scatter, convert, gather (or descriptor, convert, tokens).
 *   avifB200GridLastMode applies to both. */
AVIF_B200_API avifResult avifB200GridGetUniqueId(uint8_t id[128]);
AVIF_B200_API avifResult avifB200GridCommInit(void ** comm, int nranks, int rank, const uint8_t id[128]);
AVIF_B200_API avifResult avifB200GridCommDestroy(void * comm);
AVIF_B200_API avifResult avifB200GridCellRect(const avifImage * canvas, uint32_t gridCols, uint32_t gridRows, uint32_t cellIndex, avifCropRect * rect);
AVIF_B200_API avifResult avifB200GridRGBToYUV(avifImage * canvas, const avifRGBImage * canvasRgb, uint32_t gridCols, uint32_t gridRows, void * comm, int rank,
                                              int nranks, int root, void * stream, float phaseMs[4]);
AVIF_B200_API avifResult avifB200GridYUVToRGB(const avifImage * canvas, avifRGBImage * canvasRgb, void * comm, int rank, int nranks, int root, void * stream,
                                              float phaseMs[3]);
AVIF_B200_API int avifB200GridLastMode(void);

/* ---- which GPU serves a synchronous (host-pointer) call ----------------------------------------------------------------
 * The reference has no notion of a device; its own parallel unit is the per-call thread fan-out of
 * src/reformat.c:1679-1747.  Here several host threads may convert at once (avif.h:1012-1014), each on a GPU chosen as follows:
 *   1. a plane / pixel pointer that is device memory pins the call to that memory's device;
 *   2. else the device the calling thread selected with avifB200SetDevice(d)  (d < 0: back to the policy);
 *   3. else the device the process policy dealt to this thread the first time it entered the library: the list given to
 *      avifB200SetDevices(), else the environment variable AVIF_B200_DEVICES ("all" or a list like "0,2,5"), dealt round-robin
 *      in first-touch order; with neither, every visible device round-robin starting at the thread's current CUDA device.
 * The thread's current CUDA device is restored before the call returns.  Each (thread, device) pair owns its streams, its
 * device arena and its pinned ring; they are released when the thread exits.  The stream-ordered (*Async) entry points
 * never select a device: the caller's stream defines it.  avifB200GetDevice() = what rule 2/3 gives the calling thread. */
AVIF_B200_API avifResult avifB200SetDevice(int device);
AVIF_B200_API avifResult avifB200SetDevices(const int * devices, int count);
AVIF_B200_API int avifB200GetDevice(void);

/* Pinned host memory helpers for callers that want full PCIe rate on the host-pointer path.  Pageable (malloc) buffers work
 * as well: the calling thread stages them through a pinned ring while the copy engines run (shim.c: stagedUpload). */
AVIF_B200_API void * avifB200AllocPinned(size_t size);
AVIF_B200_API void avifB200FreePinned(void * p);

/* Selects which kernels serve subsequent calls made by THIS thread (bit field): bit 0 = use the byte-addressed
 * one-thread-per-pixel kernels instead of the tiled/vectorised ones; bit 1 = use IEEE division instructions instead
 * of the verified constant-division sequence.  0 = production.  Used by the parity tests to check every
 * implementation against the oracle. */
AVIF_B200_API void avifB200SetKernelPolicy(int policy);
/* Number of kernels this library has launched from the calling thread since the last reset. */
AVIF_B200_API uint64_t avifB200KernelLaunchCount(void);
AVIF_B200_API void avifB200ResetKernelLaunchCount(void);
/* Name of the kernel variant chosen by the most recent conversion on this thread (static string). */
AVIF_B200_API const char * avifB200LastKernelName(void);
/* The kernels divide by call constants (code ranges, kg, 2(1-kb), 2^depth-1) with a 3-operation sequence that is
 * proven equal to IEEE division for the divisors listed in csrc/verified_divisors.h; any other divisor uses the
 * hardware's IEEE division.  These expose that list and re-run the exhaustive proof on the device. */
AVIF_B200_API size_t avifB200VerifiedDivisorCount(void);
AVIF_B200_API uint32_t avifB200VerifiedDivisorBits(size_t index);
AVIF_B200_API avifResult avifB200SelfTestDivisor(uint32_t divisorBits, uint64_t * mismatches);
/* Sample and channel codes are normalised with a 2-operation quotient that is exact for integer numerators in
 * [-65536, 65535] (csrc/pixel_math.cuh: div_int).  The host proves that by exhaustion for every divisor it meets; this
 * re-runs the proof on the device.  hostVerdict receives the host's result (1 = proven). */
AVIF_B200_API avifResult avifB200SelfTestIntDivisor(uint32_t divisorBits, int * hostVerdict, uint64_t * mismatches);
/* The same 2-operation quotient serves arbitrary float numerators for every verified divisor but one (csrc/verified_divisors.h,
 * kAbNotTwoOpDivisors; proof: tools/verify_quant.c).  Re-runs that proof on the device: all 2^32 numerators against IEEE division.
 * hostVerdict receives whether the library uses the form for this divisor. */
AVIF_B200_API avifResult avifB200SelfTestDivisor2(uint32_t divisorBits, int * hostVerdict, uint64_t * mismatches);
/* Gain-map tone mapping with everything resident: base pixels, the gain map's planes and toneMappedImage->pixels (allocated by
 * the caller: width, height, rowBytes set) are device pointers; nothing crosses PCIe except the 16 bytes of statistics. */
AVIF_B200_API avifResult avifB200RGBImageApplyGainMapDevice(const avifRGBImage * baseImage, avifColorPrimaries baseColorPrimaries,
                                                            avifTransferCharacteristics baseTransferCharacteristics, const avifGainMap * gainMap,
                                                            float hdrHeadroom, avifColorPrimaries outputColorPrimaries,
                                                            avifTransferCharacteristics outputTransferCharacteristics, avifRGBImage * toneMappedImage,
                                                            avifContentLightLevelInformationBox * clli, avifDiagnostics * diag);

/* The same scaler between resident planes: src planes may be host or device pointers, dst's planes are DEVICE buffers the
 * caller allocated (width, height, row strides set; planes that are NULL on either side are skipped). */
AVIF_B200_API avifResult avifB200ImageScaleToDevice(const avifImage * src, avifImage * dst);

/* Host-logic test hooks (no device needed; tests/test_host_logic.py): the scaler's loop choice and position tables for one plane
 * (mode = AB_SCALE_* of csrc/plan.h, 0 = copy; rows / cols: 4 ints per destination row / column, may be NULL), and the tone
 * mapper's primaries matrix (row-major doubles; returns 0 where the reference's derivation fails). */
AVIF_B200_API avifResult avifB200DebugScalePlan(int srcWidth, int srcHeight, int dstWidth, int dstHeight, int sixteenBit, int * mode,
                                                int * linearOnly, int * minBoxWidth, int32_t * rows, int32_t * cols);
/* ... and whether the kernel computes that plane's column positions in closed form (x0 + i * dx, 16.16) instead of reading the table */
AVIF_B200_API avifResult avifB200DebugScalePlanColumns(int srcWidth, int srcHeight, int dstWidth, int dstHeight, int sixteenBit, int * closedForm,
                                                       int32_t * x0, int32_t * dx);
AVIF_B200_API int avifB200DebugPrimariesMatrix(avifColorPrimaries src, avifColorPrimaries dst, double out[9]);

/* ---- B-2: the reference's internal seam (include/avif/internal.h:350-387) ------------------------------------------
 * libavif_b200_hooks.so exports the six functions src/reformat.c and src/alpha.c call before their built-in loops
 * (avifImageYUVToRGBLibYUV, avifImageRGBToYUVLibYUV, avifRGBImage{Pre,Unpre}multiplyAlphaLibYUV, avifRGBImageToF16LibYUV,
 * avifImageRGBToYUVLibSharpYUV) and forwards them here: an unmodified reformat.c / alpha.c linked with that library instead
 * of reformat_libyuv.c runs its conversions on the GPU ("B-2").  Host pointers, synchronous.
 * YUV->RGB does the colour conversion (and the alpha channel when reformatAlpha is set) but never the alpha (un)multiply
 * or the F16 pass: the caller runs those through the other hooks.  It answers AVIF_RESULT_NOT_IMPLEMENTED for the
 * combinations where that split would not reproduce the built-in path (alpha multiply inside the general loop). */
AVIF_B200_API avifResult avifB200HookImageYUVToRGB(const avifImage * image, avifRGBImage * rgb, avifBool reformatAlpha, avifBool * alphaReformatted);
AVIF_B200_API avifResult avifB200HookImageRGBToYUV(avifImage * image, const avifRGBImage * rgb);
AVIF_B200_API avifResult avifB200HookRGBImageToF16(avifRGBImage * rgb);

/* The integer premultiply fused behind a conversion, floorf(c * a / maxF + 0.5f) (src/alpha.c:187-189), runs for channel
 * depths <= 12 as trunc(fma(c * a, r, k)) with the constants of csrc/verified_premul.h.  Both forms are compared on every
 * code pair (c, a): avifB200VerifyPremulFma on the host (no device needed), avifB200SelfTestPremulFma on the device.
 * AVIF_RESULT_NOT_IMPLEMENTED: no constants for that depth (the general form is used). */
AVIF_B200_API avifResult avifB200VerifyPremulFma(uint32_t depth, uint64_t * mismatches);
AVIF_B200_API avifResult avifB200SelfTestPremulFma(uint32_t depth, uint64_t * mismatches);
/* Alpha unmultiply divides three colour values by the same normalised alpha; the kernels share one reciprocal
 * between the three IEEE divisions (pixel_math.cuh: rcp_refined).  Re-checks that against the hardware's IEEE division:
 * sampledColour == 0: every (colour code, alpha code) pair of the given depth; != 0: random colour values in [0,1]. */
AVIF_B200_API avifResult avifB200SelfTestAlphaDivision(uint32_t depth, int sampledColour, uint64_t * mismatches);
/* 1 if a CUDA device is usable, else 0 (conversions then fail with AVIF_RESULT_UNKNOWN_ERROR). */
AVIF_B200_API int avifB200DeviceAvailable(void);
AVIF_B200_API const char * avifB200Version(void);

#ifdef __cplusplus
}
#endif
#endif /* AVIF_B200_H */
