/*
 * plan.h -- plain-data job descriptors handed from the host shim (shim.c, C) to the CUDA launchers
 * (kernels.cu).  A descriptor holds DEVICE pointers only and every constant the per-pixel arithmetic
 * needs, derived on the host exactly as the reference derives them (float, one rounding per operator).
 *
 * Which reference variant a call must mirror (fast loop + integer post-pass vs. the general loop with
 * in-float alpha) is decided on the host and encoded in `alphaMul`/`mode`/`chroma`;
 * the kernels never re-derive policy.
 */
#ifndef AVIF_B200_PLAN_H
#define AVIF_B200_PLAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AB_MODE_COEFF = 0, AB_MODE_IDENTITY = 1, AB_MODE_YCGCO = 2, AB_MODE_YCGCO_R = 3, AB_MODE_IDENTITY8_SHUFFLE = 4 };
enum { AB_CHROMA_NONE = 0, AB_CHROMA_FULL = 1, AB_CHROMA_NEAREST = 2, AB_CHROMA_BILINEAR = 3 };
enum { AB_ALPHA_KEEP = 0, AB_ALPHA_COPY = 1, AB_ALPHA_FILL = 2 };
/* alpha (un)multiply flavours: the reference does it either in normalised float inside the general loop
 * (reformat.c:894-947) or as an integer-domain post-pass on the quantised pixels (alpha.c:151-535). */
enum { AB_MUL_NONE = 0, AB_MUL_FLOAT_MUL = 1, AB_MUL_FLOAT_UNMUL = 2, AB_MUL_INT_PREMUL = 3, AB_MUL_INT_UNPREMUL = 4,
       AB_MUL_INT_PREMUL_FMA = 5 /* kernel-template value only: INT_PREMUL with the verified 2-fma form (pmR != 0) */ };

/* Exact division by a call-constant divisor d:  q = x*r; e = fma(-q,d,x); q' = fma(e,r,q) with r = RN(1/d).
 * `exact` is set by the host only for divisors whose (x,d) domain has been verified exhaustively to give the
 * IEEE-754 correctly rounded quotient (tests/test_division.py); otherwise the kernels use __fdiv_rn. */
typedef struct AbDiv {
    float d;      /* divisor */
    float r;      /* RN(1/d) */
    float rh, rl; /* 1/d in double precision, split into two floats: rh = RN(1/d), rl = RN(1/d - rh) */
    int exact;    /* 1: q=x*r; e=fma(-q,d,x); q'=fma(e,r,q) is proven == x/d for every float x (verified_divisors.h) */
    int intExact; /* 1: fma(x, rh, RN(x*rl)) is proven == x/d for every INTEGER x in [-65536, 65535] (checked by exhaustion
                     on the host when the divisor is first seen: shim.c divisorIntVerified) -- the 2-op form used to
                     normalise sample codes */
    int exact2;   /* 1: the same 2-op form fma(x, rh, RN(x*rl)) is proven == x/d for EVERY float x of div_const's domain
                     (tools/verify_quant.c `div`: all 2^32 numerators; verified_divisors.h lists the one divisor that fails) */
} AbDiv;

typedef struct AbY2RJob {
    const uint8_t * y;
    const uint8_t * u;
    const uint8_t * v;
    const uint8_t * a; /* source alpha plane (may be NULL) */
    uint8_t * rgb;
    int64_t yRB, uRB, vRB, aRB, rgbRB; /* row strides in bytes */
    uint32_t width, height;            /* extent of the colour loop (image dims) */
    uint32_t rowBegin, rowEnd;         /* rows [rowBegin,rowEnd) are converted by this launch (band pipelining) */
    uint32_t yuvMax, rgbMax;
    float yuvMaxF, rgbMaxF;
    float biasY, biasUV;
    AbDiv rangeY, rangeUV; /* identity mode: UV uses Y's bias/range (reformat.c:587-589) */
    AbDiv kg;
    AbDiv yuvMaxDiv, rgbMaxDiv; /* (2^depth - 1) as divisors: alpha normalisation / rescale / premultiply */
    float c1, c2, c3, c4; /* 2(1-kr), 2(1-kb), kr(1-kr), kb(1-kb) as float constants (reformat.c:1014-1016) */
    float f16mul;         /* 2^-112 * (1/65535)  (reformat.c:1429-1430) */
    float pmR, pmK;       /* integer premultiply as floor(fma(c*a, pmR, pmK)); 0 when no verified constants exist for rgbDepth */
    uint32_t magicK;               /* 0x4B000000, passed as data so that PRMT takes it from a uniform register and keeps its selector immediate */
    float magicBiasY, magicBiasUV; /* 2^23 + bias: (float)cp - bias == as_float(cp | 0x4B000000) - magicBias, exactly */
    uint8_t yuvBytes, rgbBytes; /* bytes per sample / per channel */
    uint8_t chroma;             /* AB_CHROMA_* */
    uint8_t vertBlend;          /* bilinear also blends rows (4:2:0); 0 for 4:2:2 */
    uint8_t shiftX, shiftY;
    uint8_t mode;       /* AB_MODE_* */
    uint8_t rgbFormat;  /* avifRGBFormat */
    uint8_t pixBytes, offR, offG, offB, offA, offGray;
    uint8_t nChan;      /* channels per pixel (565 counts as 1 x u16) */
    uint8_t isGray, is565, hasAlphaChan;
    uint8_t alphaWrite; /* AB_ALPHA_* : what lands in the RGB alpha channel */
    uint8_t alphaRescale; /* 1: source alpha depth != rgb depth -> float rescale (alpha.c:93-97) */
    uint8_t alphaMul;   /* AB_MUL_* */
    uint8_t toF16;
    uint8_t swapRB;      /* B before R in memory (BGR, BGRA, ABGR) */
    uint8_t alphaFirst;  /* alpha precedes colour / gray (ARGB, ABGR, AGRAY) */
    uint8_t pairAligned; /* every source plane base and stride is a multiple of 2*yuvBytes: pixel pairs load as one word */
} AbY2RJob;

typedef struct AbR2YJob {
    const uint8_t * rgb;
    uint8_t * y;
    uint8_t * u;
    uint8_t * v;
    uint8_t * a; /* destination alpha plane (may be NULL) */
    int64_t rgbRB, yRB, uRB, vRB, aRB;
    uint32_t width, height;
    uint32_t rowBegin, rowEnd; /* luma rows [rowBegin,rowEnd) are converted by this launch; rowBegin is even */
    uint32_t yuvMax, rgbMax;
    float yuvMaxF, rgbMaxF;
    float biasY, biasUV, rangeY, rangeUV; /* identity: UV quantised with Y's (reformat.c:212-213) */
    float rangeYdiv, rangeUVdiv;          /* divisors of the YCgCo-R branch: always the true Y / UV ranges (reformat.c:379-381) */
    float kr, kg, kb;
    AbDiv du, dv;     /* 2(1-kb), 2(1-kr) (reformat.c:385-386) */
    AbDiv rgbMaxDiv;  /* c / rgbMaxF (reformat.c:312-323) */
    uint32_t magicK;  /* 0x4B000000 (see AbY2RJob.magicK) */
    uint8_t rgbBytes, yuvBytes;
    uint8_t shiftX, shiftY; /* chroma subsampling of the destination; both 0 for 4:4:4 */
    uint8_t hasChroma;      /* 0 for 4:0:0 */
    uint8_t mode;
    uint8_t pixBytes, offR, offG, offB, offA, offGray;
    uint8_t isGray;
    uint8_t alphaMul;   /* AB_MUL_NONE / FLOAT_MUL / FLOAT_UNMUL */
    uint8_t alphaWrite; /* AB_ALPHA_KEEP (no plane) / COPY (from RGB alpha) / FILL */
    uint8_t alphaRescale;
    uint8_t nChan, swapRB, alphaFirst, hasAlphaChan;
    uint8_t srcPairAligned; /* rgb base and stride allow one vector load per pixel pair */
    uint8_t dstPairAligned; /* every destination plane base and stride is a multiple of 2*yuvBytes */
} AbR2YJob;

/* ---- batches: one launch converts several frames that share every conversion constant ------------------ */
/* A kernel launch costs ~4-5 us of ramp-up / drain during which the SMs idle (ncu: sm__cycles_elapsed - active), i.e.
 * 15 % of a 4K frame and half of a 1080p one.  The specialised kernels therefore take a small table of frames: work
 * is cut into GROUPS of 8 consecutive 256-pixel tiles of one frame (one tile per warp of a block) and every block is
 * bound to ONE frame for its whole life (blocks are shared out in proportion to the frames' group counts), so the
 * frame -- pointers, strides, extents -- is a loop invariant held in (uniform) registers. */
#define AB_MAX_BATCH 24

typedef struct AbY2RFrame {
    const uint8_t * y;
    const uint8_t * u;
    const uint8_t * v;
    const uint8_t * a;
    uint8_t * rgb;
    int64_t yRB, uRB, vRB, aRB, rgbRB;
    uint32_t width, height;
    uint32_t rowBegin, rowEnd;
    uint32_t tilesX;     /* 256-pixel tiles per row */
    uint32_t tiles;      /* tilesX * row units */
    uint32_t groups;     /* ceil(tiles / 8) */
    uint32_t blockStart; /* blocks [blockStart, blockStart + blocks) of the launch work on this frame only, */
    uint32_t blocks;     /* striding over its groups: the frame is invariant for the whole life of a block */
    uint32_t vecWidth;   /* == width when every plane row and the pixel rows are vector-aligned; 0 when they are only sample-aligned
                          * (views into larger images, odd strides): every tile then takes the ragged code with element-wise loads
                          * and stores */
} AbY2RFrame;

typedef struct AbY2RBatch {
    AbY2RJob common; /* constants (per-frame members unused) */
    uint32_t nFrames, totalBlocks;
    AbY2RFrame frame[AB_MAX_BATCH];
} AbY2RBatch;

typedef struct AbR2YFrame {
    const uint8_t * rgb;
    uint8_t * y;
    uint8_t * u;
    uint8_t * v;
    uint8_t * a;
    int64_t rgbRB, yRB, uRB, vRB, aRB;
    uint32_t width, height;
    uint32_t rowBegin, rowEnd;
    uint32_t tilesX, tiles, groups, blockStart, blocks;
    uint32_t vecWidth; /* as in AbY2RFrame */
} AbR2YFrame;

typedef struct AbR2YBatch {
    AbR2YJob common;
    uint32_t nFrames, totalBlocks;
    AbR2YFrame frame[AB_MAX_BATCH];
} AbR2YBatch;

/* avifAlphaParams with device pointers (alpha.c:9-149) */
typedef struct AbAlphaJob {
    const uint8_t * src; /* NULL: fill opaque */
    uint8_t * dst;
    int64_t srcRB, dstRB;
    uint32_t width, height;
    uint32_t srcDepth, dstDepth;
    uint32_t srcOff, srcPix, dstOff, dstPix;
} AbAlphaJob;

/* in-place passes over an interleaved RGB(A) buffer */
typedef struct AbRgbPassJob {
    uint8_t * pixels;
    int64_t rowBytes;
    uint32_t width, height;
    uint32_t depth;
    AbDiv maxDiv; /* 2^depth - 1 */
    uint8_t aIdx, nColor, nChan; /* channel indices, alpha.c:173-333 */
    uint8_t cIdx[3];
    uint8_t unmultiply;
    float f16mul;
} AbRgbPassJob;

/* ---- gain-map tone mapping (tonemap.cu; src/gainmap.c:73-315) ---- */
typedef struct AbPixFmt { /* what avifGetRGBAPixel / avifSetRGBAPixel need (src/reformat.c:32-117, :1856-1939) */
    uint32_t pixBytes, chanBytes;
    uint32_t offR, offG, offB, offA; /* byte offsets exactly as avifGetRGBColorSpaceInfo leaves them (0 for the gray formats) */
    uint8_t hasAlpha, isFloat, is565;
    float maxF;
} AbPixFmt;
typedef struct AbToneMapJob {
    const uint8_t * base; /* base rendition, RGB */
    const uint8_t * gain; /* gain map converted to RGB(A) (src/gainmap.c:203-212); unused when !applyGain */
    uint8_t * out;
    int64_t baseRB, gainRB, outRB;
    uint32_t width, height;
    AbPixFmt baseFmt, gainFmt, outFmt;
    uint8_t tfIn, tfOut;    /* avifTransferCharacteristics of the base / of the output (table: src/colr.c:470-491) */
    uint8_t applyGain;      /* weight != 0 */
    uint8_t transcode;      /* weight == 0: run gamma -> linear (-> primaries) -> gamma at all (src/gainmap.c:156) */
    uint8_t inConv, outConv; /* primaries conversions (3x3 in double, src/colrconvert.c:178-186) */
    double inM[9], outM[9];
    float gammaInv[3], gmMin[3], gmMax[3], baseOff[3], altOff[3];
    float weight;
    /* per-call tables indexed by a stored code (tonemap.cu: tone_lut_kernel); N = 256 or 65536 by container, 0 = not tabulated.
     * gainLut must be present whenever applyGain is set. */
    float * baseLut; /* [baseLutN]      to_linear(code / max) */
    float * gainLut; /* [3][gainLutN]   exp2f(lerp(min, max, powf(code / max, 1/gamma)) * weight) per channel */
    uint32_t baseLutN, gainLutN;
    /* results: [0] = max over every tone-mapped linear channel value (as float bits; values <= 0 never raise it),
     * nanIndex = smallest y*width+x whose output is NaN (0xFFFFFFFF if none), sum = sum over pixels of max(r,g,b,0) */
    uint32_t * maxBits;
    uint32_t * nanIndex;
    double * sum;
} AbToneMapJob;
int abLaunchToneMap(const AbToneMapJob * job, void * stream);
/* Every host-synchronous helper outside shim.c (tone mapping, scaling) works on the calling thread's own default stream:
 * unlike the legacy NULL stream it does not serialise with the streams of other threads or of the host application. */
#define AB_STREAM ((cudaStream_t)0x2) /* == cudaStreamPerThread */
/* Device selection of a synchronous call (shim.c: policy of avifB200SetDevice / avifB200SetDevices; a pointer that is device
 * memory pins the call to its device): makes that device current and returns the previous one in *prevDevice (-1: unchanged).
 * Returns an avifResult value; succeeds without doing anything when no CUDA device exists (the CUDA calls that follow report it). */
int abDeviceEnter(const void * const * ptrs, int n, int * prevDevice);
void abDeviceLeave(int prevDevice);
/* Stream-ordered (AB_STREAM) device scratch from a per-device pool that keeps freed memory instead of returning it
 * to the driver at every synchronisation (gainmap.c); release with cudaFreeAsync(p, AB_STREAM).  Returns a cudaError_t value. */
int abScratchAlloc(void ** p, size_t bytes);
int abScratchAllocOn(void ** p, size_t bytes, void * stream); /* the same pool, ordered on `stream`; release with cudaFreeAsync(p, stream) */

/* ---- plane scaling (scale_kernels.cu; src/scale.c over third_party/libyuv/source/scale*.c) ----
 * The host walks libyuv's 16.16 stepping state machines once per plane (scale.c here: O(width + height) integer work, no
 * pixels) and hands the kernel two tables; every sample operation happens on the device. */
enum { AB_SCALE_VERTICAL = 1, AB_SCALE_SIMPLE, AB_SCALE_BILINEAR_DOWN, AB_SCALE_BILINEAR_UP, AB_SCALE_UP2_LINEAR, AB_SCALE_UP2_BILINEAR, AB_SCALE_BOX };
typedef struct AbScaleJob {
    const uint8_t * src;
    uint8_t * dst;
    int64_t srcRB, dstRB; /* bytes */
    int32_t srcW, srcH, dstW, dstH;
    uint8_t bps;        /* bytes per sample: 1 or 2 */
    uint8_t mode;       /* AB_SCALE_* */
    uint8_t linearOnly; /* kFilterLinear: horizontal filter only */
    uint8_t colsArith;  /* column i sits at x0 + i * dx (16.16): the kernel computes positions instead of reading cols[] */
    int32_t minBoxWidth; /* box: dx >> 16 (or the constant box width) */
    int32_t segLen;      /* destination samples per unit of work (a multiple of 8 bytes); filled by the launcher */
    int32_t rowsPerUnit; /* destination rows per unit of work; filled by the launcher */
    int32_t x0, dx;      /* see colsArith */
    const int32_t * rows; /* 4 ints per destination row, meaning depends on mode (scale_kernels.cu) */
    const int32_t * cols; /* 4 ints per destination column */
} AbScaleJob;
enum { AB_SCALE_MAX_PLANES = 4 };
typedef struct AbScaleBatch { /* the planes of one image (same sample size), one launch */
    AbScaleJob job[AB_SCALE_MAX_PLANES];
    uint32_t count;
    uint32_t blockStart[AB_SCALE_MAX_PLANES + 1]; /* filled by the launcher */
} AbScaleBatch;
int abLaunchScaleBatch(AbScaleBatch * batch, void * stream);

/* Launchers (kernels.cu).  `stream` is a cudaStream_t.  Return 0 on success, else a cudaError_t value.
 * `forceGeneric` selects the byte-addressed reference kernel instead of the tiled/vectorised one. */
int abLaunchY2R(const AbY2RJob * job, void * stream, int forceGeneric, const char ** kernelName);
int abLaunchR2Y(const AbR2YJob * job, void * stream, int forceGeneric, const char ** kernelName);
/* `count` jobs in as few launches as possible (one per <= AB_MAX_BATCH compatible frames when the specialised kernels apply) */
int abLaunchY2RBatch(const AbY2RJob * jobs, uint32_t count, void * stream, int forceGeneric, const char ** kernelName);
int abLaunchR2YBatch(const AbR2YJob * jobs, uint32_t count, void * stream, int forceGeneric, const char ** kernelName);
int abLaunchAlpha(const AbAlphaJob * job, void * stream);
int abLaunchPremultiply(const AbRgbPassJob * job, void * stream);
int abLaunchToF16(const AbRgbPassJob * job, void * stream);
int abLaunchFill16(uint8_t * dst, size_t count16, uint32_t value, void * stream); /* avifMemset16, reformat.c:17-23 */
int abLaunchFill8(uint8_t * dst, size_t count, uint32_t value, void * stream);
int abLaunchLimitedToFull(uint8_t * plane, int64_t rowBytes, uint32_t width, uint32_t height, uint32_t depth, void * stream);
int abLaunchCopyPad(const uint8_t * src, int64_t srcRB, uint32_t srcW, uint32_t srcH, uint8_t * dst, int64_t dstRB, uint32_t dstW,
                    uint32_t dstH, uint32_t bytesPerSample, void * stream);
/* exhaustive device check of div_const's fast form for divisor d (counts x whose result differs from x/d) */
int abDivSelfTest(float d, unsigned long long * mismatches, void * stream);
int abIntDivSelfTest(const AbDiv * dv, unsigned long long * mismatches, void * stream);
int abDiv2SelfTest(const AbDiv * dv, unsigned long long * mismatches, void * stream); /* 2-op form, every float numerator of the domain */
int abPremulFmaSelfTest(uint32_t max, float r, float k, unsigned long long * mismatches, void * stream);
int abAlphaDivSelfTest(uint32_t max, int mode, unsigned long long * mismatches, void * stream);
/* number of kernel launches issued by the calling thread through the functions above */
uint64_t abLaunchCount(void);
void abLaunchCountReset(void);

#ifdef __cplusplus
}
#endif
#endif
