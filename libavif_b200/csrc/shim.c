/*
 * shim.c -- host side of the B200 reformat path: the reference's public entry points (same names, argument
 * meaning, result codes) implemented as validation + constant derivation + variant classification on the
 * CPU, and fused CUDA kernels (kernels.cu) for every byte of pixel work.  Plain C, like the reference.
 *
 * Host logic mirrored here (file:line in the reference tree):
 *   avifPrepareReformatState / avifGet{RGB,YUV}ColorSpaceInfo   src/reformat.c:32-194
 *   avifCalcYUVCoefficients, avifColorPrimariesComputeYCoeffs   src/colr.c:137-186, :517-542
 *   avifImageYUVToRGB / avifImageYUVToRGBImpl control flow      src/reformat.c:1445-1593, :1649-1748
 *   avifImageRGBToYUV control flow                              src/reformat.c:221-273, :545-569
 *   avifImageAllocatePlanes, avifRGBImage helpers               src/avif.c:431-489, :670-741
 *
 * Plane / pixel pointers may be host or device memory (classified through UVA).  Host data is staged into a
 * per-thread device arena with 256-byte aligned pitches; there is NO CPU fallback for the pixel work.
 */
#define AVIF_B200_BUILDING
#include "avif_b200.h"

#include <cuda_runtime_api.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plan.h"
#include "verified_divisors.h"
#include "verified_premul.h"

#define AB_VERSION "libavif_b200 0.1 (sm_100a)"

/* ------------------------------------------------------------------------------------------ */
/* small host helpers (src/avif.c)                                                             */

avifBool avifRGBFormatIsGray(avifRGBFormat f)
{
    return f == AVIF_RGB_FORMAT_GRAY || f == AVIF_RGB_FORMAT_GRAYA || f == AVIF_RGB_FORMAT_AGRAY;
}
avifBool avifRGBFormatHasAlpha(avifRGBFormat f)
{
    return f != AVIF_RGB_FORMAT_RGB && f != AVIF_RGB_FORMAT_BGR && f != AVIF_RGB_FORMAT_RGB_565 && f != AVIF_RGB_FORMAT_GRAY;
}
uint32_t avifRGBFormatChannelCount(avifRGBFormat f)
{
    if (f == AVIF_RGB_FORMAT_GRAY) return 1;
    if (f == AVIF_RGB_FORMAT_GRAYA || f == AVIF_RGB_FORMAT_AGRAY) return 2;
    return avifRGBFormatHasAlpha(f) ? 4 : 3;
}
uint32_t avifRGBImagePixelSize(const avifRGBImage * rgb)
{
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565) return 2;
    return avifRGBFormatChannelCount(rgb->format) * ((rgb->depth > 8) ? 2 : 1);
}
void avifRGBImageSetDefaults(avifRGBImage * rgb, const avifImage * image)
{
    rgb->width = image->width;
    rgb->height = image->height;
    rgb->depth = image->depth;
    rgb->format = AVIF_RGB_FORMAT_RGBA;
    rgb->chromaUpsampling = AVIF_CHROMA_UPSAMPLING_AUTOMATIC;
    rgb->chromaDownsampling = AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC;
    rgb->avoidLibYUV = AVIF_FALSE;
    rgb->ignoreAlpha = AVIF_FALSE;
    rgb->alphaPremultiplied = AVIF_FALSE;
    rgb->isFloat = AVIF_FALSE;
    rgb->maxThreads = 1;
    rgb->pixels = NULL;
    rgb->rowBytes = 0;
}
void avifRGBImageFreePixels(avifRGBImage * rgb)
{
    free(rgb->pixels); /* avifFree == free (src/mem.c) */
    rgb->pixels = NULL;
    rgb->rowBytes = 0;
}
avifResult avifRGBImageAllocatePixels(avifRGBImage * rgb)
{
    avifRGBImageFreePixels(rgb);
    if (rgb->width == 0 || rgb->height == 0) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t pixelSize = avifRGBImagePixelSize(rgb);
    if (rgb->width > UINT32_MAX / pixelSize) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t rowBytes = rgb->width * pixelSize;
    if (rgb->height > PTRDIFF_MAX / rowBytes) return AVIF_RESULT_INVALID_ARGUMENT;
    rgb->pixels = (uint8_t *)malloc((size_t)rowBytes * rgb->height);
    if (!rgb->pixels) return AVIF_RESULT_OUT_OF_MEMORY;
    rgb->rowBytes = rowBytes;
    return AVIF_RESULT_OK;
}
void avifGetPixelFormatInfo(avifPixelFormat format, avifPixelFormatInfo * info)
{
    memset(info, 0, sizeof(*info));
    switch (format) {
        case AVIF_PIXEL_FORMAT_YUV422: info->chromaShiftX = 1; break;
        case AVIF_PIXEL_FORMAT_YUV420: info->chromaShiftX = 1; info->chromaShiftY = 1; break;
        case AVIF_PIXEL_FORMAT_YUV400: info->monochrome = AVIF_TRUE; info->chromaShiftX = 1; info->chromaShiftY = 1; break;
        default: break;
    }
}
unsigned int avifLibYUVVersion(void) { return 0; }

/* limited <-> full integer helpers (src/reformat.c:1750-1840): scalar host API, also run per-sample on the
 * device by avifB200ImageLimitedToFullAlpha */
static int rangeRow(uint32_t depth, int uv, int * lo, int * hi, int * full)
{
    switch (depth) {
        case 8: *lo = 16; *hi = uv ? 240 : 235; *full = 255; return 1;
        case 10: *lo = 64; *hi = uv ? 960 : 940; *full = 1023; return 1;
        case 12: *lo = 256; *hi = uv ? 3840 : 3760; *full = 4095; return 1;
        default: return 0;
    }
}
static int limToFull(uint32_t depth, int uv, int v)
{
    int lo, hi, full;
    if (!rangeRow(depth, uv, &lo, &hi, &full)) return v;
    v = (((v - lo) * full) + ((hi - lo) / 2)) / (hi - lo);
    return v < 0 ? 0 : (v > full ? full : v);
}
static int fullToLim(uint32_t depth, int uv, int v)
{
    int lo, hi, full;
    if (!rangeRow(depth, uv, &lo, &hi, &full)) return v;
    v = (((v * (hi - lo)) + (full / 2)) / full) + lo;
    return v < lo ? lo : (v > hi ? hi : v);
}
int avifLimitedToFullY(uint32_t depth, int v) { return limToFull(depth, 0, v); }
int avifLimitedToFullUV(uint32_t depth, int v) { return limToFull(depth, 1, v); }
int avifFullToLimitedY(uint32_t depth, int v) { return fullToLim(depth, 0, v); }
int avifFullToLimitedUV(uint32_t depth, int v) { return fullToLim(depth, 1, v); }

/* ------------------------------------------------------------------------------------------ */
/* per-thread, per-device CUDA contexts: streams + staging arena + pinned ring                 */

#define AB_MAX_BANDS 16
#define AB_MAX_DEVICES 16
#define AB_RING_SLOTS 4

/* One context = everything a synchronous host-pointer call needs on ONE device.  A host thread owns up to one context
 * per device (created lazily); which one a call uses is decided per call (ctxEnter).  Contexts are destroyed when the
 * thread exits (pthread key destructor). */
typedef struct AbCtx {
    int ready;
    int device;
    cudaStream_t stream;           /* kernels (and everything, on the unpipelined path) */
    cudaStream_t up, up2, down;    /* host->device (two, alternating bands) / device->host copy streams of the band pipeline */
    cudaEvent_t evUp[AB_MAX_BANDS], evKernel[AB_MAX_BANDS];
    uint8_t * arena;
    size_t arenaCap;
    /* pinned staging ring for pageable host memory (stagedUpload / stagedDownload): AB_RING_SLOTS chunks that the host
     * thread fills / drains with memcpy while the copy engines move the previous ones */
    uint8_t * ring;
    size_t ringSlotBytes;
    cudaEvent_t evRing[AB_RING_SLOTS];
    struct AbHelperPool * helpers; /* host threads that share the staging copies of this context (stagingHelpers) */
} AbCtx;

typedef struct AbThread {
    AbCtx ctx[AB_MAX_DEVICES];
    int assigned;    /* device handed to this thread by the process policy at first touch (-1: none yet) */
    int explicitDev; /* avifB200SetDevice (-1: follow the policy) */
} AbThread;

typedef struct AbScope { /* one synchronous call: device selection + context binding (ctxEnter / ctxLeave) */
    int prevDevice;
    AbCtx * prevCx;
} AbScope;

static _Thread_local AbThread * tlsThread = NULL;
static _Thread_local AbCtx * cx = NULL; /* context of the synchronous call in progress */
static _Thread_local struct {
    int policy; /* 0 auto, 1 force generic kernels */
    const char * lastKernel;
} tls;
static _Thread_local int tlsZeroCopy = 0;      /* experiment: kernels read / write pinned host memory directly over PCIe */
static _Thread_local int tlsExactDivision = 0; /* 1: always use IEEE division instructions (tests cross-check both forms) */
/* Set while serving the reference's internal libyuv seam (avifB200HookImageYUVToRGB): colour conversion only -- the caller
 * (the reference's avifImageYUVToRGBImpl, src/reformat.c:1445-1593) runs the integer (un)premultiply and F16 passes itself. */
static _Thread_local struct { int active, reformatAlpha; } tlsHook = { 0, 0 };

static avifResult cudaToResult(int err)
{
    if (err == cudaSuccess) return AVIF_RESULT_OK;
    (void)cudaGetLastError();
    return (err == cudaErrorMemoryAllocation) ? AVIF_RESULT_OUT_OF_MEMORY : AVIF_RESULT_UNKNOWN_ERROR;
}
#define AB_CUDA(call)                                      \
    do {                                                   \
        const int abErr_ = (int)(call);                    \
        if (abErr_ != cudaSuccess) return cudaToResult(abErr_); \
    } while (0)

static AbDiv makeIntDiv(float d);
static AbDiv makeDiv(float d);
static avifResult ctxEnter(AbScope * sc, int pointerDevice);
static avifResult ctxLeave(AbScope * sc, avifResult res);

int avifB200DeviceAvailable(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n > 0;
}
const char * avifB200Version(void) { return AB_VERSION; }
void avifB200SetKernelPolicy(int policy)
{
    /* bit 0: generic byte-addressed kernels; bit 1: IEEE division instructions instead of the verified 3-op form;
     * bit 2: skip the compile-time specialised kernels (run the runtime-configured tiled kernel) */
    tls.policy = (policy & 1) | ((policy & 4) >> 1) | (policy & 0x18);
    tlsZeroCopy = (policy >> 5) & 1;
    tlsExactDivision = (policy >> 1) & 1;
}
uint64_t avifB200KernelLaunchCount(void) { return abLaunchCount(); }
/* AVIF_B200_TRACE_CALLS=1: report at process exit how much work went through the GPU; this is how an interposed
 * (LD_PRELOAD) deployment shows that it is live (tests/test_gpu_dropin.py). */
__attribute__((destructor)) static void abReportAtExit(void)
{
    const char * e = getenv("AVIF_B200_TRACE_CALLS");
    if (e && e[0] == '1') fprintf(stderr, "avif_b200: kernel launches %llu\n", (unsigned long long)abLaunchCount());
}
void avifB200ResetKernelLaunchCount(void) { abLaunchCountReset(); }
const char * avifB200LastKernelName(void) { return tls.lastKernel ? tls.lastKernel : ""; }

/* verified_divisors.h as data + the on-device exhaustive re-check used by tests/test_gpu_division.py */
size_t avifB200VerifiedDivisorCount(void) { return sizeof(kAbVerifiedDivisors) / sizeof(kAbVerifiedDivisors[0]); }
uint32_t avifB200VerifiedDivisorBits(size_t index) { return index < avifB200VerifiedDivisorCount() ? kAbVerifiedDivisors[index] : 0; }
avifResult avifB200SelfTestDivisor(uint32_t divisorBits, uint64_t * mismatches)
{
    float d;
    memcpy(&d, &divisorBits, sizeof(d));
    AbScope sc;
    avifResult res = ctxEnter(&sc, -1);
    if (res != AVIF_RESULT_OK) return res;
    unsigned long long bad = 0;
    res = cudaToResult(abDivSelfTest(d, &bad, cx->stream));
    *mismatches = bad;
    return ctxLeave(&sc, res);
}

/* the 2-operation quotient (AbDiv.exact2) against IEEE division for every float numerator of the domain; hostVerdict = whether
 * the library would use it for this divisor */
avifResult avifB200SelfTestDivisor2(uint32_t divisorBits, int * hostVerdict, uint64_t * mismatches)
{
    float d;
    memcpy(&d, &divisorBits, sizeof(d));
    AbScope sc;
    avifResult res = ctxEnter(&sc, -1);
    if (res != AVIF_RESULT_OK) return res;
    const int saved = tlsExactDivision;
    tlsExactDivision = 0;
    AbDiv dv = makeDiv(d);
    tlsExactDivision = saved;
    if (hostVerdict) *hostVerdict = dv.exact2;
    unsigned long long bad = 0;
    res = cudaToResult(abDiv2SelfTest(&dv, &bad, cx->stream));
    *mismatches = bad;
    return ctxLeave(&sc, res);
}

avifResult avifB200SelfTestIntDivisor(uint32_t divisorBits, int * hostVerdict, uint64_t * mismatches)
{
    float d;
    memcpy(&d, &divisorBits, sizeof(d));
    AbScope sc;
    avifResult res = ctxEnter(&sc, -1);
    if (res != AVIF_RESULT_OK) return res;
    const int saved = tlsExactDivision;
    tlsExactDivision = 0;
    AbDiv dv = makeIntDiv(d);
    tlsExactDivision = saved;
    if (hostVerdict) *hostVerdict = dv.intExact;
    unsigned long long bad = 0;
    res = cudaToResult(abIntDivSelfTest(&dv, &bad, cx->stream));
    *mismatches = bad;
    return ctxLeave(&sc, res);
}

/* verified_premul.h lookup: constants of the 2-fma integer premultiply for this channel depth (0 = none) */
static int premulFmaConstants(uint32_t depth, float * r, float * k)
{
    for (size_t i = 0; i < sizeof(kAbVerifiedPremul) / sizeof(kAbVerifiedPremul[0]); ++i) {
        if (kAbVerifiedPremul[i].depth == depth) {
            memcpy(r, &kAbVerifiedPremul[i].rBits, sizeof(*r));
            memcpy(k, &kAbVerifiedPremul[i].kBits, sizeof(*k));
            return 1;
        }
    }
    return 0;
}

/* Host re-run of tools/verify_premul.c for one depth (no device needed): every code pair, reference form vs fma form. */
avifResult avifB200VerifyPremulFma(uint32_t depth, uint64_t * mismatches)
{
    float r, k;
    if (!mismatches) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!premulFmaConstants(depth, &r, &k)) return AVIF_RESULT_NOT_IMPLEMENTED;
    const uint32_t max = (1u << depth) - 1;
    const float maxF = (float)max;
    uint64_t bad = 0;
    for (uint32_t a = 0; a <= max; ++a) {
        for (uint32_t c = 0; c <= max; ++c) {
            const float x = (float)c * (float)a;
            bad += ((uint32_t)floorf(x / maxF + 0.5f) != (uint32_t)fmaf(x, r, k));
        }
    }
    *mismatches = bad;
    return AVIF_RESULT_OK;
}

avifResult avifB200SelfTestPremulFma(uint32_t depth, uint64_t * mismatches)
{
    float r, k;
    if (!mismatches) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!premulFmaConstants(depth, &r, &k)) return AVIF_RESULT_NOT_IMPLEMENTED;
    AbScope sc;
    avifResult res = ctxEnter(&sc, -1);
    if (res != AVIF_RESULT_OK) return res;
    unsigned long long bad = 0;
    res = cudaToResult(abPremulFmaSelfTest((1u << depth) - 1, r, k, &bad, cx->stream));
    *mismatches = bad;
    return ctxLeave(&sc, res);
}

avifResult avifB200SelfTestAlphaDivision(uint32_t depth, int sampledColour, uint64_t * mismatches)
{
    if (depth < 2 || depth > 16) return AVIF_RESULT_INVALID_ARGUMENT;
    AbScope sc;
    avifResult res = ctxEnter(&sc, -1);
    if (res != AVIF_RESULT_OK) return res;
    unsigned long long bad = 0;
    res = cudaToResult(abAlphaDivSelfTest((1u << depth) - 1, sampledColour ? 1 : 0, &bad, cx->stream));
    *mismatches = bad;
    return ctxLeave(&sc, res);
}

void * avifB200AllocPinned(size_t size)
{
    void * p = NULL;
    if (cudaHostAlloc(&p, size, cudaHostAllocDefault) != cudaSuccess) {
        (void)cudaGetLastError();
        return NULL;
    }
    return p;
}
void avifB200FreePinned(void * p)
{
    if (p) (void)cudaFreeHost(p);
}

/* ---- device policy ---------------------------------------------------------------------------
 * Which GPU serves a synchronous host-pointer call (the reference has no notion of a device; its own parallel unit is the
 * per-call thread fan-out of src/reformat.c:1679-1747, here: host threads x GPUs):
 *   1. a plane / pixel pointer that is device memory pins the call to that memory's device;
 *   2. else the device the calling thread chose with avifB200SetDevice();
 *   3. else the device the process policy handed this thread when it first entered the library: the list given to
 *      avifB200SetDevices() / the environment variable AVIF_B200_DEVICES ("all" or "0,2,5"), dealt round-robin in
 *      first-touch order; without either, all visible devices round-robin STARTING AT the thread's current CUDA device
 *      (so a single-threaded CUDA application stays on the device it selected, and a pool of N worker threads of a
 *      CUDA-unaware application is spread over the GPUs instead of piling up on device 0).
 * The caller's current device is restored when the call returns.  Stream-ordered (Async) entry points never touch the
 * current device: the caller's stream defines it. */
static struct {
    int count;                /* 0: not initialised */
    int list[AB_MAX_DEVICES];
    int fromUser;             /* list came from avifB200SetDevices / AVIF_B200_DEVICES */
    unsigned next;            /* round-robin cursor */
} gPolicy;
static pthread_mutex_t gPolicyLock = PTHREAD_MUTEX_INITIALIZER;
static pthread_key_t gThreadKey;
static pthread_once_t gThreadKeyOnce = PTHREAD_ONCE_INIT;

struct AbHelperPool;
static void helpersDestroy(struct AbHelperPool * pool);
static void ctxDestroy(AbCtx * c)
{
    if (!c->ready) return;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (cudaSetDevice(c->device) == cudaSuccess) {
        (void)cudaStreamSynchronize(c->stream);
        (void)cudaStreamSynchronize(c->up);
        (void)cudaStreamSynchronize(c->up2);
        (void)cudaStreamSynchronize(c->down);
        for (int k = 0; k < AB_MAX_BANDS; ++k) {
            if (c->evUp[k]) (void)cudaEventDestroy(c->evUp[k]);
            if (c->evKernel[k]) (void)cudaEventDestroy(c->evKernel[k]);
        }
        for (int k = 0; k < AB_RING_SLOTS; ++k)
            if (c->evRing[k]) (void)cudaEventDestroy(c->evRing[k]);
        if (c->stream) (void)cudaStreamDestroy(c->stream);
        if (c->up) (void)cudaStreamDestroy(c->up);
        if (c->up2) (void)cudaStreamDestroy(c->up2);
        if (c->down) (void)cudaStreamDestroy(c->down);
        if (c->arena) (void)cudaFree(c->arena);
        if (c->ring) (void)cudaFreeHost(c->ring);
    }
    helpersDestroy(c->helpers);
    (void)cudaGetLastError(); /* e.g. cudaErrorCudartUnloading while the process is going down */
    memset(c, 0, sizeof(*c));
    if (prev >= 0) (void)cudaSetDevice(prev);
}

/* thread exit: streams, events, device arena and pinned ring of every context of the thread */
static void threadDestroy(void * p)
{
    AbThread * t = (AbThread *)p;
    if (!t) return;
    for (int d = 0; d < AB_MAX_DEVICES; ++d) ctxDestroy(&t->ctx[d]);
    free(t);
}
static void makeThreadKey(void) { (void)pthread_key_create(&gThreadKey, threadDestroy); }

static AbThread * threadState(void)
{
    if (tlsThread) return tlsThread;
    AbThread * t = (AbThread *)calloc(1, sizeof(AbThread));
    if (!t) return NULL;
    t->assigned = -1;
    t->explicitDev = -1;
    (void)pthread_once(&gThreadKeyOnce, makeThreadKey);
    (void)pthread_setspecific(gThreadKey, t);
    tlsThread = t;
    return t;
}

static int visibleDevices(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n > AB_MAX_DEVICES ? AB_MAX_DEVICES : n;
}

/* gPolicyLock held */
static void policyInitLocked(int startAt)
{
    if (gPolicy.count) return;
    const int n = visibleDevices();
    if (n <= 0) return;
    const char * e = getenv("AVIF_B200_DEVICES");
    if (e && *e && strcmp(e, "all") != 0) {
        int count = 0;
        const char * q = e;
        while (*q && count < AB_MAX_DEVICES) {
            char * end = NULL;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            if (v >= 0 && v < n) gPolicy.list[count++] = (int)v;
            q = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') break;
        }
        if (count) {
            gPolicy.count = count;
            gPolicy.fromUser = 1;
            return;
        }
    }
    for (int k = 0; k < n; ++k) gPolicy.list[k] = (startAt + k) % n;
    gPolicy.count = n;
    gPolicy.fromUser = (e && strcmp(e, "all") == 0);
}

avifResult avifB200SetDevices(const int * devices, int count)
{
    const int n = visibleDevices();
    if (count < 0 || count > AB_MAX_DEVICES || (count > 0 && !devices)) return AVIF_RESULT_INVALID_ARGUMENT;
    for (int k = 0; k < count; ++k)
        if (devices[k] < 0 || devices[k] >= n) return AVIF_RESULT_INVALID_ARGUMENT;
    pthread_mutex_lock(&gPolicyLock);
    gPolicy.count = 0; /* count == 0: back to the default policy (re-read lazily) */
    gPolicy.fromUser = 0;
    gPolicy.next = 0;
    if (count) {
        memcpy(gPolicy.list, devices, (size_t)count * sizeof(int));
        gPolicy.count = count;
        gPolicy.fromUser = 1;
    }
    pthread_mutex_unlock(&gPolicyLock);
    return AVIF_RESULT_OK;
}

avifResult avifB200SetDevice(int device)
{
    AbThread * t = threadState();
    if (!t) return AVIF_RESULT_OUT_OF_MEMORY;
    if (device < 0) {
        t->explicitDev = -1;
        return AVIF_RESULT_OK;
    }
    if (device >= visibleDevices()) return AVIF_RESULT_INVALID_ARGUMENT;
    t->explicitDev = device;
    return AVIF_RESULT_OK;
}

/* device that the next synchronous host-pointer call of this thread would use (-1: no device / error) */
int avifB200GetDevice(void)
{
    AbThread * t = threadState();
    if (!t) return -1;
    if (t->explicitDev >= 0) return t->explicitDev;
    if (t->assigned < 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) {
            (void)cudaGetLastError();
            return -1;
        }
        pthread_mutex_lock(&gPolicyLock);
        policyInitLocked(cur);
        if (gPolicy.count) t->assigned = gPolicy.list[gPolicy.next++ % (unsigned)gPolicy.count];
        pthread_mutex_unlock(&gPolicyLock);
    }
    return t->assigned;
}

static avifResult ctxCreate(AbCtx * c, int device)
{
    /* the device is current */
    memset(c, 0, sizeof(*c));
    c->device = device;
    int err = (int)cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (!err) err = (int)cudaStreamCreateWithFlags(&c->up, cudaStreamNonBlocking);
    if (!err) err = (int)cudaStreamCreateWithFlags(&c->up2, cudaStreamNonBlocking);
    if (!err) err = (int)cudaStreamCreateWithFlags(&c->down, cudaStreamNonBlocking);
    for (int k = 0; k < AB_MAX_BANDS && !err; ++k) {
        err = (int)cudaEventCreateWithFlags(&c->evUp[k], cudaEventDisableTiming);
        if (!err) err = (int)cudaEventCreateWithFlags(&c->evKernel[k], cudaEventDisableTiming);
    }
    for (int k = 0; k < AB_RING_SLOTS && !err; ++k) err = (int)cudaEventCreateWithFlags(&c->evRing[k], cudaEventDisableTiming);
    if (err) { /* unwind what exists */
        c->ready = 1;
        ctxDestroy(c);
        return cudaToResult(err);
    }
    c->ready = 1;
    return AVIF_RESULT_OK;
}

/* Scope of one synchronous call: selects the device (see the policy above), makes it current, binds `cx`. */
static avifResult ctxEnter(AbScope * sc, int pointerDevice)
{
    sc->prevDevice = -1;
    sc->prevCx = cx;
    if (!avifB200DeviceAvailable()) return AVIF_RESULT_UNKNOWN_ERROR; /* no CPU fallback by design */
    AbThread * t = threadState();
    if (!t) return AVIF_RESULT_OUT_OF_MEMORY;
    int device = pointerDevice;
    if (device < 0) device = avifB200GetDevice();
    if (device < 0 || device >= AB_MAX_DEVICES) return AVIF_RESULT_UNKNOWN_ERROR;
    int cur = -1;
    AB_CUDA(cudaGetDevice(&cur));
    if (cur != device) {
        AB_CUDA(cudaSetDevice(device));
        sc->prevDevice = cur;
    }
    AbCtx * c = &t->ctx[device];
    if (!c->ready) {
        const avifResult res = ctxCreate(c, device);
        if (res != AVIF_RESULT_OK) {
            if (sc->prevDevice >= 0) (void)cudaSetDevice(sc->prevDevice);
            sc->prevDevice = -1;
            return res;
        }
    }
    cx = c;
    return AVIF_RESULT_OK;
}

/* Leaves the scope.  A failed call first waits for everything it enqueued: copies must not keep running on the caller's
 * host buffers (or on the arena the thread's next call reuses) after the error has been reported. */
static avifResult ctxLeave(AbScope * sc, avifResult res)
{
    if (cx && res != AVIF_RESULT_OK) {
        (void)cudaStreamSynchronize(cx->up);
        (void)cudaStreamSynchronize(cx->up2);
        (void)cudaStreamSynchronize(cx->stream);
        (void)cudaStreamSynchronize(cx->down);
        (void)cudaGetLastError();
    }
    cx = sc->prevCx;
    if (sc->prevDevice >= 0) (void)cudaSetDevice(sc->prevDevice);
    return res;
}

/* device ordinal of the first pointer that is device memory, -1 if all are host (or NULL) */
static int deviceOfPointers(const void * const * ptrs, int n)
{
    for (int k = 0; k < n; ++k) {
        struct cudaPointerAttributes at;
        if (!ptrs[k]) continue;
        if (cudaPointerGetAttributes(&at, ptrs[k]) != cudaSuccess) {
            (void)cudaGetLastError();
            continue;
        }
        if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return at.device;
    }
    return -1;
}
static int deviceOfImages(const avifImage * image, const avifRGBImage * rgb)
{
    const void * p[5] = { NULL, NULL, NULL, NULL, NULL };
    if (image) {
        p[0] = image->yuvPlanes[0];
        p[1] = image->yuvPlanes[1];
        p[2] = image->yuvPlanes[2];
        p[3] = image->alphaPlane;
    }
    if (rgb) p[4] = rgb->pixels;
    return deviceOfPointers(p, 5);
}

/* device selection for the host-synchronous helpers outside this file (plan.h) */
int abDeviceEnter(const void * const * ptrs, int n, int * prevDevice)
{
    *prevDevice = -1;
    if (!avifB200DeviceAvailable()) return AVIF_RESULT_OK; /* the caller's first CUDA call reports the missing device */
    int device = deviceOfPointers(ptrs, n);
    if (device < 0) device = avifB200GetDevice();
    if (device < 0) return AVIF_RESULT_UNKNOWN_ERROR;
    int cur = -1;
    AB_CUDA(cudaGetDevice(&cur));
    if (cur != device) {
        AB_CUDA(cudaSetDevice(device));
        *prevDevice = cur;
    }
    return AVIF_RESULT_OK;
}
void abDeviceLeave(int prevDevice)
{
    if (prevDevice >= 0) (void)cudaSetDevice(prevDevice);
}

/* grow-only device arena of the current context; contents are dead between calls */
static avifResult arenaReserve(size_t bytes)
{
    if (bytes <= cx->arenaCap) return AVIF_RESULT_OK;
    if (cx->arena) {
        AB_CUDA(cudaStreamSynchronize(cx->stream));
        AB_CUDA(cudaStreamSynchronize(cx->down));
        AB_CUDA(cudaFree(cx->arena));
        cx->arena = NULL;
        cx->arenaCap = 0;
    }
    const size_t cap = bytes + (bytes >> 3) + (1u << 20);
    AB_CUDA(cudaMalloc((void **)&cx->arena, cap));
    cx->arenaCap = cap;
    return AVIF_RESULT_OK;
}

/* cudaMemoryType of a pointer: Unregistered = pageable host memory, Host = pinned / registered, Device, Managed */
static int ptrMemoryType(const void * p, struct cudaPointerAttributes * out)
{
    struct cudaPointerAttributes at;
    memset(&at, 0, sizeof(at));
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        at.type = cudaMemoryTypeUnregistered;
    }
    if (out) *out = at;
    return (int)at.type;
}
static int ptrOnDevice(const void * p)
{
    if (!p) return 0;
    const int t = ptrMemoryType(p, NULL);
    return t == cudaMemoryTypeDevice || t == cudaMemoryTypeManaged;
}

static size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* One 2-D byte region that must be visible to a kernel.  Host regions get a slot in the arena. */
typedef struct AbRegion {
    uint8_t * user;     /* caller's pointer (host or device) */
    int64_t userRB;     /* caller's row stride */
    size_t rowUsed;     /* bytes per row the path touches */
    uint32_t rows;
    int onDevice;
    int pageable;       /* host memory the copy engines cannot reach directly: staged through the context's pinned ring */
    uint8_t * dev;      /* pointer the kernel uses */
    int64_t devRB;
    int present;
    int exactPitch;     /* keep the caller's pitch and copy rows*pitch as one block (gray chroma fill) */
} AbRegion;

static void regionInit(AbRegion * r, const void * user, int64_t rb, size_t rowUsed, uint32_t rows, int assumeDevice)
{
    memset(r, 0, sizeof(*r));
    if (!user) return;
    r->present = 1;
    r->user = (uint8_t *)user;
    r->userRB = rb;
    r->rowUsed = rowUsed;
    r->rows = rows;
    if (assumeDevice) {
        r->onDevice = 1;
        return;
    }
    struct cudaPointerAttributes at;
    const int type = ptrMemoryType(user, &at);
    r->onDevice = (type == cudaMemoryTypeDevice || type == cudaMemoryTypeManaged);
    r->pageable = (type == cudaMemoryTypeUnregistered);
    if (!r->onDevice && tlsZeroCopy && type == cudaMemoryTypeHost && at.devicePointer) {
        r->user = (uint8_t *)at.devicePointer; /* mapped pinned memory: usable by kernels as is */
        r->onDevice = 1;
    }
}

/* rows the arena holds: a stride-0 plane (every row aliases row 0, like in the reference's pointer arithmetic) is one row
 * that the kernels re-read with a zero device stride */
static uint32_t regionArenaRows(const AbRegion * r) { return (r->userRB == 0 && !r->exactPitch) ? 1u : r->rows; }

static size_t regionArenaBytes(const AbRegion * r)
{
    if (!r->present || r->onDevice) return 0;
    if (r->exactPitch) return alignUp((size_t)r->userRB * r->rows, 256);
    return alignUp(r->rowUsed, 256) * regionArenaRows(r);
}

static void regionPlace(AbRegion * r, uint8_t ** cursor)
{
    if (!r->present) return;
    if (r->onDevice) {
        r->dev = r->user;
        r->devRB = r->userRB;
        return;
    }
    r->dev = *cursor;
    r->devRB = r->exactPitch ? r->userRB : ((r->userRB == 0) ? 0 : (int64_t)alignUp(r->rowUsed, 256));
    *cursor += regionArenaBytes(r);
}

/* ---- pageable host memory: pinned ring ---------------------------------------------------------
 * cudaMemcpy*Async on pageable memory makes the driver stage every byte through its own bounce buffer, one blocking piece
 * after the other.  Here the calling thread moves rows between the caller's buffer and AB_RING_SLOTS pinned chunks with
 * memcpy while the copy engine transfers the chunks filled before (uploads) / fills the next ones (downloads), so the host
 * copy and the DMA overlap and a call costs about one memcpy of its bytes.  AVIF_B200_STAGING=0 disables it (driver path),
 * AVIF_B200_RING_MB sets the chunk size. */
#define AB_STAGE_MIN_BYTES ((size_t)256 << 10)
static int stagingEnabled(void)
{
    static int v = -1;
    if (v < 0) {
        const char * e = getenv("AVIF_B200_STAGING");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v;
}
static avifResult ringReserve(size_t rowBytes)
{
    size_t want = (size_t)2 << 20;
    const char * e = getenv("AVIF_B200_RING_MB");
    if (e && atol(e) > 0) want = (size_t)atol(e) << 20;
    if (want < rowBytes) want = alignUp(rowBytes, 4096); /* a chunk holds whole rows */
    if (cx->ring && cx->ringSlotBytes >= want) return AVIF_RESULT_OK;
    if (cx->ring) {
        for (int k = 0; k < AB_RING_SLOTS; ++k) AB_CUDA(cudaEventSynchronize(cx->evRing[k]));
        AB_CUDA(cudaFreeHost(cx->ring));
        cx->ring = NULL;
        cx->ringSlotBytes = 0;
    }
    AB_CUDA(cudaHostAlloc((void **)&cx->ring, want * AB_RING_SLOTS, cudaHostAllocDefault));
    cx->ringSlotBytes = want;
    return AVIF_RESULT_OK;
}
static int regionStaged(const AbRegion * r, uint32_t rows)
{
    return r->pageable && stagingEnabled() && !r->exactPitch && r->userRB != 0 && (size_t)r->rowUsed * rows >= AB_STAGE_MIN_BYTES;
}
static unsigned gRingCursor(void)
{
    static _Thread_local unsigned cursor = 0;
    return cursor++;
}

/* Host copy between a caller's pageable buffer and a pinned chunk.  With several host threads converting at once the staging
 * copies are bound by the socket's memory bandwidth, and an ordinary store stream costs a read-for-ownership of every destination
 * line on top of the read and the write: non-temporal stores drop that third of the traffic (glibc's memcpy only switches to them
 * above a threshold far larger than a chunk).  AVIF_B200_NT_COPY=0 falls back to memcpy. */
#if defined(__x86_64__)
#include <emmintrin.h>
static void copyStreaming(uint8_t * dst, const uint8_t * src, size_t n)
{
    static int useNt = -1;
    if (useNt < 0) {
        const char * e = getenv("AVIF_B200_NT_COPY");
        useNt = (e && e[0] == '0') ? 0 : 1;
    }
    if (!useNt || n < 4096) {
        memcpy(dst, src, n);
        return;
    }
    const size_t head = (size_t)(-(uintptr_t)dst) & 63;
    if (head) {
        memcpy(dst, src, head);
        dst += head;
        src += head;
        n -= head;
    }
    const size_t blocks = n / 64;
    const __m128i * s = (const __m128i *)src;
    __m128i * d = (__m128i *)dst;
    for (size_t k = 0; k < blocks; ++k) {
        const __m128i a = _mm_loadu_si128(s + 0), b = _mm_loadu_si128(s + 1), c = _mm_loadu_si128(s + 2), e = _mm_loadu_si128(s + 3);
        _mm_stream_si128(d + 0, a);
        _mm_stream_si128(d + 1, b);
        _mm_stream_si128(d + 2, c);
        _mm_stream_si128(d + 3, e);
        s += 4;
        d += 4;
    }
    _mm_sfence(); /* the chunk is handed to the copy engine / the caller right after */
    if (n & 63) memcpy(dst + blocks * 64, src + blocks * 64, n & 63);
}
#else
static void copyStreaming(uint8_t * dst, const uint8_t * src, size_t n) { memcpy(dst, src, n); }
#endif

/* ---- helper threads of the staging copies ------------------------------------------------------------------------
 * One host thread copies ~6 GB/s between pageable memory and the pinned ring, an eighth of what the copy engines move.  The
 * reference lets a caller spend rgb->maxThreads threads on a conversion (src/reformat.c:1679-1747); here the extra threads
 * (maxThreads - 1, at most AB_MAX_HELPERS; AVIF_B200_STAGE_THREADS overrides) share every chunk copy of the call: rows are cut
 * into equal slices, the calling thread takes the first.  The pool belongs to the (thread, device) context and dies with it. */
#define AB_MAX_HELPERS 7
typedef struct AbHelperPool {
    int n;
    pthread_t th[AB_MAX_HELPERS];
    pthread_mutex_t mu;
    pthread_cond_t cv;
    unsigned gen; /* job generation, written under mu, also polled */
    int stop;
    /* the job: a 2-D copy cut into `parts` row slices; helper k takes slice k + 1 */
    uint8_t * dst;
    const uint8_t * src;
    size_t dstPitch, srcPitch, rowBytes;
    uint32_t rows;
    int parts;
    int done; /* helpers that finished the current job */
} AbHelperPool;
typedef struct AbHelperArg {
    AbHelperPool * pool;
    int index;
} AbHelperArg;

static void copyRows(uint8_t * dst, size_t dstPitch, const uint8_t * src, size_t srcPitch, size_t rowBytes, uint32_t r0, uint32_t r1)
{
    if (r1 <= r0) return;
    if (dstPitch == srcPitch && rowBytes == dstPitch) { /* dense rows: one run (row padding, where there is any, is never written) */
        copyStreaming(dst + (size_t)r0 * dstPitch, src + (size_t)r0 * srcPitch, dstPitch * (r1 - r0));
        return;
    }
    for (uint32_t k = r0; k < r1; ++k) copyStreaming(dst + (size_t)k * dstPitch, src + (size_t)k * srcPitch, rowBytes);
}

static void * helperMain(void * p)
{
    AbHelperArg * arg = (AbHelperArg *)p;
    AbHelperPool * pool = arg->pool;
    const int index = arg->index;
    free(arg);
    unsigned seen = 0;
    for (;;) {
        /* jobs of one call follow each other within microseconds: poll briefly before sleeping */
        int spins = 0;
        while (__atomic_load_n(&pool->gen, __ATOMIC_ACQUIRE) == seen && !__atomic_load_n(&pool->stop, __ATOMIC_ACQUIRE) && spins < 20000) {
            ++spins;
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        pthread_mutex_lock(&pool->mu);
        while (pool->gen == seen && !pool->stop) pthread_cond_wait(&pool->cv, &pool->mu);
        if (pool->stop) {
            pthread_mutex_unlock(&pool->mu);
            return NULL;
        }
        seen = pool->gen;
        uint8_t * dst = pool->dst;
        const uint8_t * src = pool->src;
        const size_t dp = pool->dstPitch, sp = pool->srcPitch, rb = pool->rowBytes;
        const uint32_t rows = pool->rows;
        const int parts = pool->parts;
        pthread_mutex_unlock(&pool->mu);
        if (index + 1 < parts) {
            const uint32_t r0 = (uint32_t)((uint64_t)rows * (uint32_t)(index + 1) / (uint32_t)parts);
            const uint32_t r1 = (uint32_t)((uint64_t)rows * (uint32_t)(index + 2) / (uint32_t)parts);
            copyRows(dst, dp, src, sp, rb, r0, r1);
        }
        __atomic_add_fetch(&pool->done, 1, __ATOMIC_RELEASE);
    }
}

static void helpersDestroy(AbHelperPool * pool)
{
    if (!pool) return;
    pthread_mutex_lock(&pool->mu);
    pool->stop = 1;
    pthread_cond_broadcast(&pool->cv);
    pthread_mutex_unlock(&pool->mu);
    for (int k = 0; k < pool->n; ++k) pthread_join(pool->th[k], NULL);
    pthread_mutex_destroy(&pool->mu);
    pthread_cond_destroy(&pool->cv);
    free(pool);
}

/* how many helper threads the call in progress may use (set by the conversion entry points from rgb->maxThreads) */
static _Thread_local int tlsStageHelpers = 0;
static int stageHelperCount(int maxThreads)
{
    static int envCount = -2;
    if (envCount == -2) {
        const char * e = getenv("AVIF_B200_STAGE_THREADS");
        envCount = (e && *e) ? atoi(e) : -1;
    }
    int n = ((envCount >= 0) ? envCount : maxThreads) - 1;
    if (n < 0) n = 0;
    return n > AB_MAX_HELPERS ? AB_MAX_HELPERS : n;
}

static void helpersEnsure(int want)
{
    if (want <= 0) return;
    AbHelperPool * pool = cx->helpers;
    if (!pool) {
        pool = (AbHelperPool *)calloc(1, sizeof(AbHelperPool));
        if (!pool) return;
        pthread_mutex_init(&pool->mu, NULL);
        pthread_cond_init(&pool->cv, NULL);
        cx->helpers = pool;
    }
    while (pool->n < want) {
        AbHelperArg * arg = (AbHelperArg *)malloc(sizeof(AbHelperArg));
        if (!arg) return;
        arg->pool = pool;
        arg->index = pool->n;
        if (pthread_create(&pool->th[pool->n], NULL, helperMain, arg) != 0) {
            free(arg);
            return;
        }
        ++pool->n;
    }
}

/* 2-D host copy shared with the context's helper threads */
static void copyRowsShared(uint8_t * dst, size_t dstPitch, const uint8_t * src, size_t srcPitch, size_t rowBytes, uint32_t rows)
{
    AbHelperPool * pool = cx->helpers;
    int helpers = tlsStageHelpers;
    if (pool && helpers > pool->n) helpers = pool->n;
    if (!pool || helpers <= 0 || rows < 8 || (size_t)rows * rowBytes < ((size_t)256 << 10)) {
        copyRows(dst, dstPitch, src, srcPitch, rowBytes, 0, rows);
        return;
    }
    pthread_mutex_lock(&pool->mu);
    pool->dst = dst;
    pool->src = src;
    pool->dstPitch = dstPitch;
    pool->srcPitch = srcPitch;
    pool->rowBytes = rowBytes;
    pool->rows = rows;
    pool->parts = helpers + 1;
    __atomic_store_n(&pool->done, 0, __ATOMIC_RELAXED);
    __atomic_add_fetch(&pool->gen, 1, __ATOMIC_RELEASE);
    pthread_cond_broadcast(&pool->cv);
    pthread_mutex_unlock(&pool->mu);
    copyRows(dst, dstPitch, src, srcPitch, rowBytes, 0, (uint32_t)((uint64_t)rows / (uint32_t)(helpers + 1)));
    while (__atomic_load_n(&pool->done, __ATOMIC_ACQUIRE) < pool->n) { /* every helper acknowledges every job */
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

static avifResult stagedUpload(const AbRegion * r, uint32_t row0, uint32_t row1, cudaStream_t st)
{
    const size_t pitch = (size_t)r->devRB; /* the chunk mirrors the arena layout: one 1-D DMA per chunk */
    avifResult res = ringReserve(pitch);
    if (res != AVIF_RESULT_OK) return res;
    helpersEnsure(tlsStageHelpers);
    const uint32_t perChunk = (uint32_t)(cx->ringSlotBytes / pitch);
    for (uint32_t row = row0; row < row1; row += perChunk) {
        const uint32_t n = (row1 - row < perChunk) ? row1 - row : perChunk;
        const unsigned slot = gRingCursor() % AB_RING_SLOTS;
        uint8_t * chunk = cx->ring + (size_t)slot * cx->ringSlotBytes;
        AB_CUDA(cudaEventSynchronize(cx->evRing[slot])); /* the DMA that last used this chunk */
        const uint8_t * src = r->user + (int64_t)row * r->userRB;
        copyRowsShared(chunk, pitch, src, (size_t)r->userRB, r->rowUsed, n);
        AB_CUDA(cudaMemcpyAsync(r->dev + (size_t)row * pitch, chunk, pitch * (n - 1) + r->rowUsed, cudaMemcpyHostToDevice, st));
        AB_CUDA(cudaEventRecord(cx->evRing[slot], st));
    }
    return AVIF_RESULT_OK;
}

/* returns when the rows are in the caller's buffer */
static avifResult stagedDownload(const AbRegion * r, uint32_t row0, uint32_t row1, cudaStream_t st)
{
    const size_t pitch = (size_t)r->devRB;
    avifResult res = ringReserve(pitch);
    if (res != AVIF_RESULT_OK) return res;
    helpersEnsure(tlsStageHelpers);
    const uint32_t perChunk = (uint32_t)(cx->ringSlotBytes / pitch);
    struct { uint32_t row, n; unsigned slot; } pending[AB_RING_SLOTS];
    int head = 0, count = 0;
    uint32_t row = row0;
    while (row < row1 || count) {
        if (row < row1 && count < AB_RING_SLOTS) { /* keep the copy engine AB_RING_SLOTS chunks ahead of the host copy */
            const uint32_t n = (row1 - row < perChunk) ? row1 - row : perChunk;
            const unsigned slot = gRingCursor() % AB_RING_SLOTS;
            /* slots are handed out round-robin and drained in order: `slot` is free unless an upload of this call still uses it */
            AB_CUDA(cudaEventSynchronize(cx->evRing[slot]));
            AB_CUDA(cudaMemcpyAsync(cx->ring + (size_t)slot * cx->ringSlotBytes, r->dev + (size_t)row * pitch, pitch * (n - 1) + r->rowUsed,
                                    cudaMemcpyDeviceToHost, st));
            AB_CUDA(cudaEventRecord(cx->evRing[slot], st));
            const int tail = (head + count) % AB_RING_SLOTS;
            pending[tail].row = row;
            pending[tail].n = n;
            pending[tail].slot = slot;
            ++count;
            row += n;
            continue;
        }
        const uint32_t prow = pending[head].row, pn = pending[head].n;
        const unsigned slot = pending[head].slot;
        head = (head + 1) % AB_RING_SLOTS;
        --count;
        AB_CUDA(cudaEventSynchronize(cx->evRing[slot]));
        const uint8_t * chunk = cx->ring + (size_t)slot * cx->ringSlotBytes;
        uint8_t * dst = r->user + (int64_t)prow * r->userRB;
        copyRowsShared(dst, (size_t)r->userRB, chunk, pitch, r->rowUsed, pn);
    }
    return AVIF_RESULT_OK;
}

static int regionUpload(const AbRegion * r, uint32_t row0, uint32_t row1, cudaStream_t st)
{
    if (!r->present || r->onDevice || row1 <= row0) return cudaSuccess;
    if (r->userRB == 0 || r->rowUsed == 0) { /* stride-0 plane (every row aliases row 0): upload the single row */
        return (int)cudaMemcpyAsync(r->dev, r->user, r->rowUsed, cudaMemcpyHostToDevice, st);
    }
    if (regionStaged(r, row1 - row0)) return (stagedUpload(r, row0, row1, st) == AVIF_RESULT_OK) ? cudaSuccess : cudaErrorUnknown;
    return (int)cudaMemcpy2DAsync(r->dev + (int64_t)row0 * r->devRB, (size_t)r->devRB, r->user + (int64_t)row0 * r->userRB, (size_t)r->userRB,
                                  r->rowUsed, row1 - row0, cudaMemcpyHostToDevice, st);
}

static int regionDownload(const AbRegion * r, uint32_t row0, uint32_t row1, cudaStream_t st)
{
    if (!r->present || r->onDevice || row1 <= row0) return cudaSuccess;
    if (r->exactPitch) {
        return (int)cudaMemcpyAsync(r->user + (int64_t)row0 * r->userRB, r->dev + (int64_t)row0 * r->devRB, (size_t)r->userRB * (row1 - row0),
                                    cudaMemcpyDeviceToHost, st);
    }
    if (regionStaged(r, row1 - row0)) return (stagedDownload(r, row0, row1, st) == AVIF_RESULT_OK) ? cudaSuccess : cudaErrorUnknown;
    return (int)cudaMemcpy2DAsync(r->user + (int64_t)row0 * r->userRB, (size_t)r->userRB, r->dev + (int64_t)row0 * r->devRB, (size_t)r->devRB,
                                  r->rowUsed, row1 - row0, cudaMemcpyDeviceToHost, st);
}

/* Band pipeline of the host-pointer path: the frame is cut into row bands; band b's planes go up on `up`, its
 * kernel runs on `stream` once they (and, for vertical chroma blending, the next band's) have landed, its result
 * goes down on `down` -- PCIe is full duplex, so a frame costs max(upload, download) instead of their sum. */
/* upload stream of band b.  One stream by default: alternating bands between two upload streams (AVIF_B200_UP_STREAMS=2)
 * measured 10 % slower end to end on this host (profiles/r01_variants.md) */
static cudaStream_t upStream(int band)
{
    static _Thread_local int nUp = 0;
    if (!nUp) {
        const char * e = getenv("AVIF_B200_UP_STREAMS");
        nUp = (e && atoi(e) == 2) ? 2 : 1;
    }
    return (nUp == 2 && (band & 1)) ? cx->up2 : cx->up;
}

/* synchronous host-pointer conversions currently inside the library, over all threads */
static int gHostCallsInFlight = 0;

static int planBands(uint32_t height, size_t frameBytes, uint32_t bandRows[AB_MAX_BANDS + 1])
{
    static _Thread_local size_t bandBytes = 0;
    static _Thread_local int bandFixed = 0;
    if (!bandBytes) {
        const char * e = getenv("AVIF_B200_BAND_MB"); /* tuning knob; default from profiles/r01_variants.md */
        const long mb = e ? atol(e) : 0;
        bandFixed = (mb > 0);
        bandBytes = (size_t)((mb > 0) ? mb : 24) << 20;
    }
    int n = (int)(frameBytes / bandBytes);
    /* A lone caller overlaps its own upload, kernel and download band by band.  When several host threads convert at once
     * their calls already overlap each other, and whole-plane copies use PCIe better than 5 MB pieces (uploads below 16 MB run
     * at half rate on this host): one band then (measured: 4.7 -> 5.2-5.4 Gpixel/s with 4-8 threads, profiles/r01_variants.md). */
    if (!bandFixed && __atomic_load_n(&gHostCallsInFlight, __ATOMIC_RELAXED) > 1) n = 1;
    if (n < 1) n = 1;
    if (n > AB_MAX_BANDS) n = AB_MAX_BANDS;
    uint32_t per = (height + (uint32_t)n - 1) / (uint32_t)n;
    per = (per + 1) & ~1u; /* even: chroma rows and 2x2 blocks never straddle a band */
    if (per < 2) per = 2;
    int count = 0;
    uint32_t r = 0;
    while (r < height && count < AB_MAX_BANDS) {
        bandRows[count++] = r;
        r += per;
    }
    bandRows[count] = height;
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* reformat state (src/reformat.c:32-194)                                                      */

typedef struct AbState {
    /* rgb */
    uint32_t rgbChanBytes, rgbPixBytes, offR, offG, offB, offA, offGray;
    int rgbMax;
    float rgbMaxF;
    /* yuv */
    float kr, kg, kb;
    uint32_t yuvChanBytes;
    int yuvMax;
    float biasY, biasUV, rangeY, rangeUV;
    int shiftX, shiftY;
    int mode; /* AB_MODE_COEFF / IDENTITY / YCGCO / YCGCO_R */
} AbState;

/* H.273 eq. 32-37 in float, same association as src/colr.c:517-542 */
static void chromaDerivedKrKb(avifColorPrimaries cp, float * outKr, float * outKb)
{
    static const struct {
        uint16_t id;
        float v[8]; /* rX rY gX gY bX bY wX wY (src/colr.c:16-28) */
    } table[] = {
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
    const float * v = table[0].v; /* unknown -> BT.709 primaries (src/colr.c:41-42) */
    for (size_t k = 0; k < sizeof(table) / sizeof(table[0]); ++k) {
        if (table[k].id == cp) {
            v = table[k].v;
            break;
        }
    }
    const float rX = v[0], rY = v[1], gX = v[2], gY = v[3], bX = v[4], bY = v[5], wX = v[6], wY = v[7];
    const float rZ = 1.0f - (rX + rY), gZ = 1.0f - (gX + gY), bZ = 1.0f - (bX + bY), wZ = 1.0f - (wX + wY);
    const float denom = wY * (rX * (gY * bZ - bY * gZ) + gX * (bY * rZ - rY * bZ) + bX * (rY * gZ - gY * rZ));
    *outKr = (rY * (wX * (gY * bZ - bY * gZ) + wY * (bX * gZ - gX * bZ) + wZ * (gX * bY - bX * gY))) / denom;
    *outKb = (bY * (wX * (rY * gZ - gY * rZ) + wY * (gX * rZ - rX * gZ) + wZ * (rX * gY - gX * rY))) / denom;
}

static void calcCoefficients(const avifImage * image, float * kr, float * kg, float * kb)
{
    float r = 0.299f, b = 0.114f; /* unspecified / unknown MC -> BT.601 (src/colr.c:173-175) */
    switch (image->matrixCoefficients) {
        case AVIF_MATRIX_COEFFICIENTS_BT709: r = 0.2126f; b = 0.0722f; break;
        case AVIF_MATRIX_COEFFICIENTS_FCC: r = 0.30f; b = 0.11f; break;
        case AVIF_MATRIX_COEFFICIENTS_SMPTE240: r = 0.212f; b = 0.087f; break;
        case AVIF_MATRIX_COEFFICIENTS_BT2020_NCL: r = 0.2627f; b = 0.0593f; break;
        case AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_NCL: chromaDerivedKrKb(image->colorPrimaries, &r, &b); break;
        default: break; /* BT.470BG, BT.601 and everything else: BT.601 numbers */
    }
    *kr = r;
    *kb = b;
    *kg = 1.0f - r - b;
}

/* the checks of avifGetYUVColorSpaceInfo, src/reformat.c:119-141 */
static int yuvSideValid(const avifImage * image)
{
    const unsigned mc = image->matrixCoefficients;
    if (image->depth != 8 && image->depth != 10 && image->depth != 12 && image->depth != 16) return 0;
    if (image->yuvFormat < AVIF_PIXEL_FORMAT_YUV444 || image->yuvFormat >= AVIF_PIXEL_FORMAT_COUNT) return 0;
    if (image->yuvRange != AVIF_RANGE_LIMITED && image->yuvRange != AVIF_RANGE_FULL) return 0;
    const int ycgcoFamily = (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO || mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE || mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO);
    if (mc == 3 || (ycgcoFamily && image->yuvRange == AVIF_RANGE_LIMITED) || mc == AVIF_MATRIX_COEFFICIENTS_BT2020_CL ||
        mc == AVIF_MATRIX_COEFFICIENTS_SMPTE2085 || mc == AVIF_MATRIX_COEFFICIENTS_CHROMA_DERIVED_CL || mc == AVIF_MATRIX_COEFFICIENTS_ICTCP ||
        mc >= AVIF_MATRIX_COEFFICIENTS_LAST) {
        return 0;
    }
    if (mc == AVIF_MATRIX_COEFFICIENTS_IDENTITY && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV444 && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400) {
        return 0;
    }
    return 1;
}

static int prepareState(const avifImage * image, const avifRGBImage * rgb, AbState * s)
{
    const unsigned mc = image->matrixCoefficients;
    if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE && image->depth - 2 != rgb->depth) return 0;
    if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO && image->depth - 1 != rgb->depth) return 0;

    if (rgb->depth != 8 && rgb->depth != 10 && rgb->depth != 12 && rgb->depth != 16) return 0;
    if (rgb->isFloat && rgb->depth != 16) return 0;
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565 && rgb->depth != 8) return 0;
    if ((int)rgb->format < AVIF_RGB_FORMAT_RGB || rgb->format >= AVIF_RGB_FORMAT_COUNT) return 0;

    const uint32_t cb = (rgb->depth > 8) ? 2 : 1;
    s->rgbChanBytes = cb;
    s->rgbPixBytes = avifRGBImagePixelSize(rgb);
    s->offR = s->offG = s->offB = s->offA = s->offGray = 0;
    /* channel order as byte offsets (src/reformat.c:53-111) */
    static const int8_t order[AVIF_RGB_FORMAT_COUNT][5] = {
        /*            R   G   B   A  Gray */
        /* RGB   */ { 0, 1, 2, -1, -1 },
        /* RGBA  */ { 0, 1, 2, 3, -1 },
        /* ARGB  */ { 1, 2, 3, 0, -1 },
        /* BGR   */ { 2, 1, 0, -1, -1 },
        /* BGRA  */ { 2, 1, 0, 3, -1 },
        /* ABGR  */ { 3, 2, 1, 0, -1 },
        /* 565   */ { 0, 0, 0, -1, -1 },
        /* GRAY  */ { -1, -1, -1, -1, 0 },
        /* GRAYA */ { -1, -1, -1, 1, 0 },
        /* AGRAY */ { -1, -1, -1, 0, 1 },
    };
    const int8_t * o = order[rgb->format];
    if (o[0] > 0) s->offR = (uint32_t)o[0] * cb;
    if (o[1] > 0) s->offG = (uint32_t)o[1] * cb;
    if (o[2] > 0) s->offB = (uint32_t)o[2] * cb;
    if (o[3] > 0) s->offA = (uint32_t)o[3] * cb;
    if (o[4] > 0) s->offGray = (uint32_t)o[4] * cb;
    s->rgbMax = (1 << rgb->depth) - 1;
    s->rgbMaxF = (float)s->rgbMax;

    if (!yuvSideValid(image)) return 0;
    avifPixelFormatInfo fi;
    avifGetPixelFormatInfo(image->yuvFormat, &fi);
    s->shiftX = fi.chromaShiftX;
    s->shiftY = fi.chromaShiftY;
    calcCoefficients(image, &s->kr, &s->kg, &s->kb);
    s->yuvChanBytes = (image->depth > 8) ? 2 : 1;
    s->yuvMax = (1 << image->depth) - 1;
    const int limited = (image->yuvRange == AVIF_RANGE_LIMITED);
    s->biasY = limited ? (float)(16 << (image->depth - 8)) : 0.0f;
    s->biasUV = (float)(1 << (image->depth - 1));
    s->rangeY = (float)(limited ? (219 << (image->depth - 8)) : s->yuvMax);
    s->rangeUV = (float)(limited ? (224 << (image->depth - 8)) : s->yuvMax);

    s->mode = AB_MODE_COEFF;
    if (mc == AVIF_MATRIX_COEFFICIENTS_IDENTITY) s->mode = AB_MODE_IDENTITY;
    else if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO) s->mode = AB_MODE_YCGCO;
    else if (mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RE || mc == AVIF_MATRIX_COEFFICIENTS_YCGCO_RO) s->mode = AB_MODE_YCGCO_R;
    if (s->mode != AB_MODE_COEFF) s->kr = s->kg = s->kb = 0.0f;
    return 1;
}

/* ---- the reference's internal colour-space descriptions and per-pixel accessors ----------------
 * src/reformat.c:32-159 and :1842-1939.  Host-side scalar API (one pixel of HOST memory per call): what the reference's own
 * gain-map code calls; exported so that src/reformat.c + src/alpha.c can be left out of a reference build altogether. */
avifBool avifGetRGBColorSpaceInfo(const avifRGBImage * rgb, avifRGBColorSpaceInfo * info)
{
    /* same checks, in the same order, as prepareState's RGB half */
    if (rgb->depth != 8 && rgb->depth != 10 && rgb->depth != 12 && rgb->depth != 16) return AVIF_FALSE;
    if (rgb->isFloat && rgb->depth != 16) return AVIF_FALSE;
    if (rgb->format == AVIF_RGB_FORMAT_RGB_565 && rgb->depth != 8) return AVIF_FALSE;
    if ((int)rgb->format < AVIF_RGB_FORMAT_RGB || rgb->format >= AVIF_RGB_FORMAT_COUNT) return AVIF_FALSE;
    avifImage probe;
    memset(&probe, 0, sizeof(probe));
    probe.depth = 8;
    probe.yuvFormat = AVIF_PIXEL_FORMAT_YUV444;
    probe.yuvRange = AVIF_RANGE_FULL;
    probe.matrixCoefficients = AVIF_MATRIX_COEFFICIENTS_BT601;
    AbState s;
    if (!prepareState(&probe, rgb, &s)) return AVIF_FALSE;
    info->channelBytes = s.rgbChanBytes;
    info->pixelBytes = s.rgbPixBytes;
    info->offsetBytesR = s.offR;
    info->offsetBytesG = s.offG;
    info->offsetBytesB = s.offB;
    info->offsetBytesA = s.offA;
    info->offsetBytesGray = s.offGray;
    info->maxChannel = s.rgbMax;
    info->maxChannelF = s.rgbMaxF;
    return AVIF_TRUE;
}

avifBool avifGetYUVColorSpaceInfo(const avifImage * image, avifYUVColorSpaceInfo * info)
{
    if (!yuvSideValid(image)) return AVIF_FALSE;
    avifGetPixelFormatInfo(image->yuvFormat, &info->formatInfo);
    calcCoefficients(image, &info->kr, &info->kg, &info->kb);
    const int limited = (image->yuvRange == AVIF_RANGE_LIMITED);
    info->channelBytes = (image->depth > 8) ? 2 : 1;
    info->depth = image->depth;
    info->range = image->yuvRange;
    info->maxChannel = (1 << image->depth) - 1;
    info->biasY = limited ? (float)(16 << (image->depth - 8)) : 0.0f;
    info->biasUV = (float)(1 << (image->depth - 1));
    info->rangeY = (float)(limited ? (219 << (image->depth - 8)) : info->maxChannel);
    info->rangeUV = (float)(limited ? (224 << (image->depth - 8)) : info->maxChannel);
    /* info->mode is not assigned here: the reference's function leaves it to avifPrepareReformatState as well */
    return AVIF_TRUE;
}

/* src/reformat.c:1411, :1842-1854 */
static const float kF16Multiplier = 1.9259299444e-34f;
static inline uint16_t floatToF16(float v)
{
    union { float f; uint32_t u; } c;
    c.f = v * kF16Multiplier;
    return (uint16_t)(c.u >> 13);
}
static inline float f16ToFloat(uint16_t v)
{
    union { float f; uint32_t u; } c;
    c.u = (uint32_t)v << 13;
    return c.f / kF16Multiplier;
}

void avifGetRGBAPixel(const avifRGBImage * src, uint32_t x, uint32_t y, const avifRGBColorSpaceInfo * info, float rgbaPixel[4])
{
    const uint8_t * const px = &src->pixels[(size_t)y * src->rowBytes + (size_t)x * info->pixelBytes];
    const int hasAlpha = avifRGBFormatHasAlpha(src->format);
    if (info->channelBytes > 1) {
        uint16_t r, g, b, a = (uint16_t)info->maxChannel;
        memcpy(&r, px + info->offsetBytesR, 2);
        memcpy(&g, px + info->offsetBytesG, 2);
        memcpy(&b, px + info->offsetBytesB, 2);
        if (hasAlpha) memcpy(&a, px + info->offsetBytesA, 2);
        if (src->isFloat) {
            rgbaPixel[0] = f16ToFloat(r);
            rgbaPixel[1] = f16ToFloat(g);
            rgbaPixel[2] = f16ToFloat(b);
            rgbaPixel[3] = hasAlpha ? f16ToFloat(a) : 1.0f;
        } else {
            rgbaPixel[0] = r / info->maxChannelF;
            rgbaPixel[1] = g / info->maxChannelF;
            rgbaPixel[2] = b / info->maxChannelF;
            rgbaPixel[3] = a / info->maxChannelF;
        }
    } else if (src->format == AVIF_RGB_FORMAT_RGB_565) { /* src/reformat.c:621-632: 5/6/5 fields widened by bit replication */
        uint16_t v;
        memcpy(&v, px + info->offsetBytesR, 2);
        const uint16_t r5 = (v >> 11) & 0x1F, g6 = (v >> 5) & 0x3F, b5 = v & 0x1F;
        rgbaPixel[0] = (uint8_t)((r5 << 3) | (r5 >> 2)) / info->maxChannelF;
        rgbaPixel[1] = (uint8_t)((g6 << 2) | (g6 >> 4)) / info->maxChannelF;
        rgbaPixel[2] = (uint8_t)((b5 << 3) | (b5 >> 2)) / info->maxChannelF;
        rgbaPixel[3] = 1.0f;
    } else {
        rgbaPixel[0] = px[info->offsetBytesR] / info->maxChannelF;
        rgbaPixel[1] = px[info->offsetBytesG] / info->maxChannelF;
        rgbaPixel[2] = px[info->offsetBytesB] / info->maxChannelF;
        rgbaPixel[3] = hasAlpha ? (px[info->offsetBytesA] / info->maxChannelF) : 1.0f;
    }
}

void avifSetRGBAPixel(const avifRGBImage * dst, uint32_t x, uint32_t y, const avifRGBColorSpaceInfo * info, const float rgbaPixel[4])
{
    uint8_t * const px = &dst->pixels[(size_t)y * dst->rowBytes + (size_t)x * info->pixelBytes];
    const int hasAlpha = avifRGBFormatHasAlpha(dst->format);
    if (dst->depth > 8) {
        uint16_t v[4];
        for (int c = 0; c < 4; ++c) v[c] = dst->isFloat ? floatToF16(rgbaPixel[c]) : (uint16_t)(0.5f + (rgbaPixel[c] * info->maxChannelF));
        memcpy(px + info->offsetBytesR, &v[0], 2);
        memcpy(px + info->offsetBytesG, &v[1], 2);
        memcpy(px + info->offsetBytesB, &v[2], 2);
        if (hasAlpha) memcpy(px + info->offsetBytesA, &v[3], 2);
        return;
    }
    const uint8_t r = (uint8_t)(0.5f + (rgbaPixel[0] * info->maxChannelF));
    const uint8_t g = (uint8_t)(0.5f + (rgbaPixel[1] * info->maxChannelF));
    const uint8_t b = (uint8_t)(0.5f + (rgbaPixel[2] * info->maxChannelF));
    if (dst->format == AVIF_RGB_FORMAT_RGB_565) { /* avifStoreRGB8Pixel, src/reformat.c:605-617 */
        const uint16_t v = (uint16_t)((b >> 3) | ((g >> 2) << 5) | ((r >> 3) << 11));
        memcpy(px + info->offsetBytesR, &v, 2);
    } else {
        px[info->offsetBytesR] = r;
        px[info->offsetBytesG] = g;
        px[info->offsetBytesB] = b;
    }
    if (hasAlpha) px[info->offsetBytesA] = (uint8_t)(0.5f + (rgbaPixel[3] * info->maxChannelF));
}

/* binary search of the exhaustively verified divisors (verified_divisors.h) */
static int divisorVerified(float d)
{
    uint32_t bits;
    memcpy(&bits, &d, sizeof(bits));
    size_t lo = 0, hi = sizeof(kAbVerifiedDivisors) / sizeof(kAbVerifiedDivisors[0]);
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (kAbVerifiedDivisors[mid] == bits) return 1;
        if (kAbVerifiedDivisors[mid] < bits) lo = mid + 1;
        else hi = mid;
    }
    return 0;
}

/* Proof by exhaustion, on the host, that fma(x, rh, RN(x*rl)) == x/d for every integer x in [-65536, 65535]: sample codes
 * minus their bias, and channel codes, are such integers.  131072 fmaf() calls (< 1 ms), cached per thread. */
static int divisorIntVerified(float d, float rh, float rl)
{
    enum { kCache = 16 };
    static _Thread_local struct { uint32_t bits; int ok; int used; } cache[kCache];
    uint32_t bits;
    memcpy(&bits, &d, sizeof(bits));
    for (int k = 0; k < kCache; ++k) {
        if (cache[k].used && cache[k].bits == bits) return cache[k].ok;
    }
    int ok = isfinite(d) && d != 0.0f;
    for (int x = -65536; ok && x <= 65535; ++x) {
        const float xf = (float)x;
        const float fast = fmaf(xf, rh, xf * rl);
        const float ref = xf / d;
        uint32_t a, b;
        memcpy(&a, &fast, 4);
        memcpy(&b, &ref, 4);
        if (a != b) ok = 0;
    }
    for (int k = 0; k < kCache; ++k) {
        if (!cache[k].used) {
            cache[k].used = 1;
            cache[k].bits = bits;
            cache[k].ok = ok;
            break;
        }
    }
    return ok;
}

static AbDiv makeDiv(float d)
{
    AbDiv v;
    v.d = d;
    v.r = 1.0f / d;
    const double inv = 1.0 / (double)d;
    v.rh = (float)inv;
    v.rl = (float)(inv - (double)v.rh);
    v.exact = !tlsExactDivision && divisorVerified(d);
    v.intExact = 0;
    v.exact2 = v.exact;
    {
        uint32_t bits;
        memcpy(&bits, &d, sizeof(bits));
        for (size_t k = 0; k < sizeof(kAbNotTwoOpDivisors) / sizeof(kAbNotTwoOpDivisors[0]); ++k)
            if (kAbNotTwoOpDivisors[k] == bits) v.exact2 = 0;
    }
    return v;
}

/* divisors that only ever see integer numerators (code ranges, 2^depth - 1) */
static AbDiv makeIntDiv(float d)
{
    AbDiv v = makeDiv(d);
    v.intExact = !tlsExactDivision && divisorIntVerified(d, v.rh, v.rl);
    return v;
}

/* ------------------------------------------------------------------------------------------ */
/* in-place RGB passes (premultiply / unpremultiply / F16)                                      */

static int premulLayout(avifRGBFormat f, AbRgbPassJob * job)
{
    switch (f) {
        case AVIF_RGB_FORMAT_RGBA:
        case AVIF_RGB_FORMAT_BGRA: job->aIdx = 3; job->cIdx[0] = 0; job->cIdx[1] = 1; job->cIdx[2] = 2; job->nColor = 3; job->nChan = 4; return 1;
        case AVIF_RGB_FORMAT_ARGB:
        case AVIF_RGB_FORMAT_ABGR: job->aIdx = 0; job->cIdx[0] = 1; job->cIdx[1] = 2; job->cIdx[2] = 3; job->nColor = 3; job->nChan = 4; return 1;
        case AVIF_RGB_FORMAT_GRAYA: job->aIdx = 1; job->cIdx[0] = 0; job->nColor = 1; job->nChan = 2; return 1;
        case AVIF_RGB_FORMAT_AGRAY: job->aIdx = 0; job->cIdx[0] = 1; job->nColor = 1; job->nChan = 2; return 1;
        default: return 0;
    }
}

static float f16Multiplier(uint32_t depth)
{
    const float scale = 1.0f / ((1 << depth) - 1); /* src/reformat.c:1429-1430 */
    return 1.9259299444e-34f * scale;
}

/* src/alpha.c:151-336 / :338-535 */
static avifResult premultiplyBody(avifRGBImage * rgb, int unmultiply, void * userStream, int async)
{
    if (!rgb->pixels || !rgb->rowBytes) return AVIF_RESULT_REFORMAT_FAILED;
    if (!avifRGBFormatHasAlpha(rgb->format)) return unmultiply ? AVIF_RESULT_REFORMAT_FAILED : AVIF_RESULT_INVALID_ARGUMENT;
    AbRgbPassJob job;
    memset(&job, 0, sizeof(job));
    if (!premulLayout(rgb->format, &job)) return AVIF_RESULT_NOT_IMPLEMENTED;
    if (rgb->depth < 8 || rgb->depth > 16) return AVIF_RESULT_REFORMAT_FAILED; /* the reference asserts */
    avifResult res;
    cudaStream_t st = async ? (cudaStream_t)userStream : cx->stream;

    AbRegion px;
    regionInit(&px, rgb->pixels, rgb->rowBytes, (size_t)rgb->width * avifRGBImagePixelSize(rgb), rgb->height, async);
    if (!async) {
        res = arenaReserve(regionArenaBytes(&px));
        if (res != AVIF_RESULT_OK) return res;
    }
    uint8_t * cursor = async ? NULL : cx->arena;
    regionPlace(&px, &cursor);
    AB_CUDA(regionUpload(&px, 0, px.rows, st));
    job.pixels = px.dev;
    job.rowBytes = px.devRB;
    job.width = rgb->width;
    job.height = rgb->height;
    job.depth = rgb->depth;
    job.maxDiv = makeDiv((float)((1u << rgb->depth) - 1));
    job.unmultiply = (uint8_t)unmultiply;
    AB_CUDA(abLaunchPremultiply(&job, st));
    tls.lastKernel = unmultiply ? "unpremultiply" : "premultiply";
    AB_CUDA(regionDownload(&px, 0, px.rows, st));
    if (!async) AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

static avifResult premultiplyCommon(avifRGBImage * rgb, int unmultiply, void * userStream, int async)
{
    if (async) return premultiplyBody(rgb, unmultiply, userStream, 1);
    /* argument errors are reported without needing a device, like the reference */
    if (!rgb->pixels || !rgb->rowBytes) return AVIF_RESULT_REFORMAT_FAILED;
    if (!avifRGBFormatHasAlpha(rgb->format)) return unmultiply ? AVIF_RESULT_REFORMAT_FAILED : AVIF_RESULT_INVALID_ARGUMENT;
    AbScope sc;
    avifResult res = ctxEnter(&sc, deviceOfImages(NULL, rgb));
    if (res != AVIF_RESULT_OK) return res;
    return ctxLeave(&sc, premultiplyBody(rgb, unmultiply, NULL, 0));
}
avifResult avifRGBImagePremultiplyAlpha(avifRGBImage * rgb) { return premultiplyCommon(rgb, 0, NULL, 0); }
avifResult avifRGBImageUnpremultiplyAlpha(avifRGBImage * rgb) { return premultiplyCommon(rgb, 1, NULL, 0); }
avifResult avifB200RGBImagePremultiplyAlphaHost(avifRGBImage * rgb) { return premultiplyCommon(rgb, 0, NULL, 0); }
avifResult avifB200RGBImageUnpremultiplyAlphaHost(avifRGBImage * rgb) { return premultiplyCommon(rgb, 1, NULL, 0); }
avifResult avifB200RGBImagePremultiplyAlphaAsync(avifRGBImage * rgb, void * stream) { return premultiplyCommon(rgb, 0, stream, 1); }
avifResult avifB200RGBImageUnpremultiplyAlphaAsync(avifRGBImage * rgb, void * stream) { return premultiplyCommon(rgb, 1, stream, 1); }

/* src/reformat.c:1419-1443 as a pass of its own (the conversions fuse it; the internal seam needs it separately) */
static avifResult toF16Body(avifRGBImage * rgb)
{
    if (!rgb->pixels || !rgb->rowBytes) return AVIF_RESULT_INVALID_ARGUMENT;
    if (rgb->depth <= 8 || rgb->depth > 16) return AVIF_RESULT_INVALID_ARGUMENT; /* half floats live in 16-bit channels */
    avifResult res;
    cudaStream_t st = cx->stream;
    AbRgbPassJob job;
    memset(&job, 0, sizeof(job));
    AbRegion px;
    regionInit(&px, rgb->pixels, rgb->rowBytes, (size_t)rgb->width * avifRGBImagePixelSize(rgb), rgb->height, 0);
    res = arenaReserve(regionArenaBytes(&px));
    if (res != AVIF_RESULT_OK) return res;
    uint8_t * cursor = cx->arena;
    regionPlace(&px, &cursor);
    AB_CUDA(regionUpload(&px, 0, px.rows, st));
    job.pixels = px.dev;
    job.rowBytes = px.devRB;
    job.width = rgb->width;
    job.height = rgb->height;
    job.depth = rgb->depth;
    job.nChan = (uint8_t)avifRGBFormatChannelCount(rgb->format);
    job.f16mul = f16Multiplier(rgb->depth);
    AB_CUDA(abLaunchToF16(&job, st));
    tls.lastKernel = "to_f16";
    AB_CUDA(regionDownload(&px, 0, px.rows, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

static avifResult toF16Common(avifRGBImage * rgb)
{
    if (!rgb->pixels || !rgb->rowBytes) return AVIF_RESULT_INVALID_ARGUMENT;
    AbScope sc;
    avifResult res = ctxEnter(&sc, deviceOfImages(NULL, rgb));
    if (res != AVIF_RESULT_OK) return res;
    return ctxLeave(&sc, toF16Body(rgb));
}

/* ------------------------------------------------------------------------------------------ */
/* alpha plane <-> channel (src/alpha.c:9-149); void like the reference: failures cannot be reported */

static avifResult alphaBody(const avifAlphaParams * p, int fill)
{
    cudaStream_t st = cx->stream;
    const uint32_t dcb = (p->dstDepth > 8) ? 2 : 1, scb = (p->srcDepth > 8) ? 2 : 1;
    AbRegion src, dst;
    regionInit(&dst, p->dstPlane, p->dstRowBytes, (size_t)p->dstOffsetBytes + (size_t)(p->width - 1) * p->dstPixelBytes + dcb, p->height, 0);
    regionInit(&src, fill ? NULL : p->srcPlane, p->srcRowBytes, (size_t)p->srcOffsetBytes + (size_t)(p->width - 1) * p->srcPixelBytes + scb, p->height, 0);
    if (arenaReserve(regionArenaBytes(&src) + regionArenaBytes(&dst)) != AVIF_RESULT_OK) return AVIF_RESULT_OUT_OF_MEMORY;
    uint8_t * cursor = cx->arena;
    regionPlace(&src, &cursor);
    regionPlace(&dst, &cursor);
    AB_CUDA(regionUpload(&src, 0, src.rows, st));
    AB_CUDA(regionUpload(&dst, 0, dst.rows, st)); /* untouched channels must survive the round trip */
    AbAlphaJob job;
    memset(&job, 0, sizeof(job));
    job.src = fill ? NULL : src.dev;
    job.dst = dst.dev;
    job.srcRB = src.devRB;
    job.dstRB = dst.devRB;
    job.width = p->width;
    job.height = p->height;
    job.srcDepth = fill ? p->dstDepth : p->srcDepth;
    job.dstDepth = p->dstDepth;
    job.srcOff = p->srcOffsetBytes;
    job.srcPix = p->srcPixelBytes;
    job.dstOff = p->dstOffsetBytes;
    job.dstPix = p->dstPixelBytes;
    AB_CUDA(abLaunchAlpha(&job, st));
    tls.lastKernel = fill ? "alpha_fill" : "alpha_reformat";
    AB_CUDA(regionDownload(&dst, 0, dst.rows, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

static void alphaCommon(const avifAlphaParams * p, int fill)
{
    if (p->width == 0 || p->height == 0) return;
    const void * ptrs[2] = { p->dstPlane, fill ? NULL : p->srcPlane };
    AbScope sc;
    if (ctxEnter(&sc, deviceOfPointers(ptrs, 2)) != AVIF_RESULT_OK) return;
    (void)ctxLeave(&sc, alphaBody(p, fill));
}


void avifFillAlpha(const avifAlphaParams * params) { alphaCommon(params, 1); }
void avifReformatAlpha(const avifAlphaParams * params) { alphaCommon(params, 0); }

/* ------------------------------------------------------------------------------------------ */
/* YUV -> RGB                                                                                  */

/* everything the launch needs besides device pointers: validation, constants, variant classification */
typedef struct AbY2RPlan {
    AbState s;
    int hasChroma, rgbHasAlpha, reformatAlpha;
} AbY2RPlan;

static avifResult y2rBuildJob(const avifImage * image, const avifRGBImage * rgb, AbY2RJob * jobOut, AbY2RPlan * plan)
{
    /* src/reformat.c:1649-1677 */
    if (!image->yuvPlanes[AVIF_CHAN_Y] || rgb->maxThreads < 0) return AVIF_RESULT_REFORMAT_FAILED;
    AbState s;
    if (!prepareState(image, rgb, &s)) return AVIF_RESULT_REFORMAT_FAILED;
    if (!rgb->pixels) return AVIF_RESULT_REFORMAT_FAILED; /* the reference would dereference NULL */

    const int rgbHasAlpha = avifRGBFormatHasAlpha(rgb->format);
    const int rgbIsGray = avifRGBFormatIsGray(rgb->format);
    int alphaMode = 0; /* 0 no-op, 1 multiply, 2 unmultiply */
    if (image->alphaPlane) {
        if (!rgbHasAlpha || rgb->ignoreAlpha) {
            if (!image->alphaPremultiplied) alphaMode = 1;
        } else if (!image->alphaPremultiplied && rgb->alphaPremultiplied) {
            alphaMode = 1;
        } else if (image->alphaPremultiplied && !rgb->alphaPremultiplied) {
            alphaMode = 2;
        }
    }
    /* rgb->maxThreads only splits rows in the reference and never changes results (:1679-1747): a hint here. */

    /* src/reformat.c:1449-1450, :1475 */
    int reformatAlpha = rgbHasAlpha && (!rgb->ignoreAlpha || alphaMode != 0);
    const int srcAlphaUsable = image->alphaPlane && image->alphaRowBytes;

    /* variant classification, src/reformat.c:1496-1567 */
    const int nearest = (rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_FASTEST) || (rgb->chromaUpsampling == AVIF_CHROMA_UPSAMPLING_NEAREST);
    const int hasColorByStride = image->yuvRowBytes[AVIF_CHAN_U] && image->yuvRowBytes[AVIF_CHAN_V] && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400;
    int fast = 0, shuffle = 0;
    if (!rgbIsGray && (!hasColorByStride || image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 || nearest) && (alphaMode == 0 || rgbHasAlpha)) {
        if (s.mode == AB_MODE_IDENTITY) {
            if (image->depth == 8 && rgb->depth == 8 && image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444 && image->yuvRange == AVIF_RANGE_FULL) {
                fast = 1;
                shuffle = 1;
            }
        } else if (s.mode == AB_MODE_COEFF) {
            fast = 1;
        }
    }
    if (tlsHook.active) {
        /* The caller applies alpha (un)multiply as an integer pass over the quantised pixels after this returns.  That equals
         * the built-in path only where the built-in path does the same (the "fast" class); the general loop multiplies in
         * normalised float before quantising, so those combinations are declined (the caller then runs its own loop). */
        if (alphaMode != 0 && !fast) return AVIF_RESULT_NOT_IMPLEMENTED;
        alphaMode = 0;
        reformatAlpha = tlsHook.reformatAlpha;
    }
    int hasChroma;
    if (fast) {
        hasChroma = hasColorByStride; /* :1496-1497 picks the Color or Mono loop */
        if (hasChroma && (!image->yuvPlanes[AVIF_CHAN_U] || !image->yuvPlanes[AVIF_CHAN_V])) return AVIF_RESULT_REFORMAT_FAILED; /* reference: NULL deref */
    } else {
        hasChroma = image->yuvPlanes[AVIF_CHAN_U] && image->yuvPlanes[AVIF_CHAN_V] && image->yuvFormat != AVIF_PIXEL_FORMAT_YUV400; /* :676 */
    }

    AbY2RJob job;
    memset(&job, 0, sizeof(job));
    job.width = image->width;
    job.height = image->height;
    job.yuvMax = (uint32_t)s.yuvMax;
    job.rgbMax = (uint32_t)s.rgbMax;
    job.yuvMaxF = (float)s.yuvMax;
    job.rgbMaxF = s.rgbMaxF;
    job.biasY = s.biasY;
    job.rangeY = makeIntDiv(s.rangeY);
    if (s.mode == AB_MODE_IDENTITY) { /* chroma shares the luma table, :587-589 */
        job.biasUV = s.biasY;
        job.rangeUV = makeIntDiv(s.rangeY);
    } else {
        job.biasUV = s.biasUV;
        job.rangeUV = makeIntDiv(s.rangeUV);
    }
    job.kg = makeDiv(s.kg);
    job.yuvMaxDiv = makeIntDiv((float)s.yuvMax);
    job.rgbMaxDiv = makeIntDiv(s.rgbMaxF);
    job.c1 = 2 * (1 - s.kr);
    job.c2 = 2 * (1 - s.kb);
    job.c3 = s.kr * (1 - s.kr);
    job.c4 = s.kb * (1 - s.kb);
    job.f16mul = f16Multiplier(rgb->depth);
    job.magicK = 0x4B000000u;
    job.magicBiasY = 8388608.0f + job.biasY;
    job.magicBiasUV = 8388608.0f + job.biasUV;
    job.yuvBytes = (uint8_t)s.yuvChanBytes;
    job.rgbBytes = (uint8_t)s.rgbChanBytes;
    job.shiftX = (uint8_t)s.shiftX;
    job.shiftY = (uint8_t)s.shiftY;
    if (!hasChroma) {
        job.chroma = AB_CHROMA_NONE;
    } else if (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV444) {
        job.chroma = AB_CHROMA_FULL;
        job.shiftX = job.shiftY = 0;
    } else if (nearest) {
        job.chroma = AB_CHROMA_NEAREST;
    } else {
        job.chroma = AB_CHROMA_BILINEAR;
        job.vertBlend = (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV420);
    }
    job.mode = (uint8_t)(shuffle ? AB_MODE_IDENTITY8_SHUFFLE : s.mode);
    job.rgbFormat = (uint8_t)rgb->format;
    job.pixBytes = (uint8_t)s.rgbPixBytes;
    job.offR = (uint8_t)s.offR;
    job.offG = (uint8_t)s.offG;
    job.offB = (uint8_t)s.offB;
    job.offA = (uint8_t)s.offA;
    job.offGray = (uint8_t)s.offGray;
    job.nChan = (uint8_t)((rgb->format == AVIF_RGB_FORMAT_RGB_565) ? 1 : avifRGBFormatChannelCount(rgb->format));
    job.isGray = (uint8_t)rgbIsGray;
    job.is565 = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_RGB_565);
    job.hasAlphaChan = (uint8_t)rgbHasAlpha;
    job.alphaWrite = (uint8_t)(!reformatAlpha ? AB_ALPHA_KEEP : (srcAlphaUsable ? AB_ALPHA_COPY : AB_ALPHA_FILL));
    job.alphaRescale = (uint8_t)(image->depth != rgb->depth);
    if (alphaMode == 0) job.alphaMul = AB_MUL_NONE;
    else if (fast) job.alphaMul = (uint8_t)((alphaMode == 1) ? AB_MUL_INT_PREMUL : AB_MUL_INT_UNPREMUL);
    else job.alphaMul = (uint8_t)((alphaMode == 1) ? AB_MUL_FLOAT_MUL : AB_MUL_FLOAT_UNMUL);
    job.pmR = job.pmK = 0.0f;
    if (job.alphaMul == AB_MUL_INT_PREMUL && !tlsExactDivision) (void)premulFmaConstants(rgb->depth, &job.pmR, &job.pmK);
    job.toF16 = (uint8_t)((rgb->isFloat && !tlsHook.active) ? 1 : 0);
    job.swapRB = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_BGR || rgb->format == AVIF_RGB_FORMAT_BGRA || rgb->format == AVIF_RGB_FORMAT_ABGR);
    job.alphaFirst = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_ARGB || rgb->format == AVIF_RGB_FORMAT_ABGR || rgb->format == AVIF_RGB_FORMAT_AGRAY);

    *jobOut = job;
    plan->s = s;
    plan->hasChroma = hasChroma;
    plan->rgbHasAlpha = rgbHasAlpha;
    plan->reformatAlpha = reformatAlpha;
    return AVIF_RESULT_OK;
}

static avifResult yuvToRgbBody(const avifImage * image, avifRGBImage * rgb, void * userStream, int async, AbY2RJob job, const AbY2RPlan * planIn)
{
    const AbY2RPlan plan = *planIn;
    avifResult res;
    const AbState s = plan.s;
    const int hasChroma = plan.hasChroma, rgbHasAlpha = plan.rgbHasAlpha, reformatAlpha = plan.reformatAlpha;
    cudaStream_t st = async ? (cudaStream_t)userStream : cx->stream;

    /* the alpha channel / post-passes iterate over rgb->width x rgb->height, the colour loop over the image's
     * (src/reformat.c:1467-1468 vs :686,:703).  When they agree (always, for sane callers) everything is one kernel. */
    const int sameDims = (rgb->width == image->width) && (rgb->height == image->height);
    const uint32_t rgbRows = (rgb->height > image->height) ? rgb->height : image->height;
    const uint32_t rgbCols = (rgb->width > image->width) ? rgb->width : image->width;

    const uint32_t cw = (uint32_t)(((uint64_t)image->width + s.shiftX) >> s.shiftX);
    const uint32_t ch = (uint32_t)(((uint64_t)image->height + s.shiftY) >> s.shiftY);
    AbRegion ry, ru, rv, ra, rrgb;
    regionInit(&ry, image->yuvPlanes[0], image->yuvRowBytes[0], (size_t)image->width * s.yuvChanBytes, image->height, async);
    regionInit(&ru, hasChroma ? image->yuvPlanes[1] : NULL, image->yuvRowBytes[1], (size_t)cw * s.yuvChanBytes, ch, async);
    regionInit(&rv, hasChroma ? image->yuvPlanes[2] : NULL, image->yuvRowBytes[2], (size_t)cw * s.yuvChanBytes, ch, async);
    const int needSrcAlpha = (job.alphaWrite == AB_ALPHA_COPY) || job.alphaMul == AB_MUL_FLOAT_MUL || job.alphaMul == AB_MUL_FLOAT_UNMUL;
    const uint32_t aCols = sameDims ? image->width : ((job.alphaWrite == AB_ALPHA_COPY) ? rgbCols : image->width);
    const uint32_t aRows = sameDims ? image->height : ((job.alphaWrite == AB_ALPHA_COPY) ? rgbRows : image->height);
    regionInit(&ra, needSrcAlpha ? image->alphaPlane : NULL, image->alphaRowBytes, (size_t)aCols * s.yuvChanBytes, aRows, async);
    regionInit(&rrgb, rgb->pixels, rgb->rowBytes, (size_t)rgbCols * s.rgbPixBytes, rgbRows, async);

    if (!async) {
        res = arenaReserve(regionArenaBytes(&ry) + regionArenaBytes(&ru) + regionArenaBytes(&rv) + regionArenaBytes(&ra) + regionArenaBytes(&rrgb));
        if (res != AVIF_RESULT_OK) return res;
    }
    uint8_t * cursor = async ? NULL : cx->arena;
    regionPlace(&ry, &cursor);
    regionPlace(&ru, &cursor);
    regionPlace(&rv, &cursor);
    regionPlace(&ra, &cursor);
    regionPlace(&rrgb, &cursor);

    const int keepsDstBytes = (job.alphaWrite == AB_ALPHA_KEEP && rgbHasAlpha) || !sameDims;
    const int anyHost = !ry.onDevice || (ru.present && !ru.onDevice) || (rv.present && !rv.onDevice) || (ra.present && !ra.onDevice) || !rrgb.onDevice;
    const size_t frameBytes = (size_t)ry.rowUsed * ry.rows + (ru.present ? 2 * (size_t)ru.rowUsed * ru.rows : 0) +
                              (ra.present ? (size_t)ra.rowUsed * ra.rows : 0) + (size_t)rrgb.rowUsed * rrgb.rows;
    /* pageable buffers are staged chunk by chunk by this thread (stagedUpload / stagedDownload): the host copy is the
     * critical path there and already overlaps the DMA, so bands would only add launches */
    const int anyStaged = regionStaged(&ry, ry.rows) || regionStaged(&ra, ra.rows) || regionStaged(&rrgb, rrgb.rows);
    const int pipelined = !async && anyHost && !anyStaged && sameDims && !keepsDstBytes && frameBytes >= ((size_t)8 << 20) &&
                          (!ra.present || image->alphaRowBytes != 0) && image->yuvRowBytes[0] != 0;

    job.y = ry.dev;
    job.yRB = ry.devRB;
    job.u = ru.dev;
    job.uRB = ru.devRB;
    job.v = rv.dev;
    job.vRB = rv.devRB;
    job.a = ra.dev;
    job.aRB = ra.present ? ((image->alphaRowBytes == 0) ? 0 : ra.devRB) : 0;
    job.rgb = rrgb.dev;
    job.rgbRB = rrgb.devRB;
    job.rowBegin = 0;
    job.rowEnd = image->height;
    {
        const uintptr_t m = 2u * s.yuvChanBytes - 1;
        uintptr_t bits = (uintptr_t)job.y | (uintptr_t)job.yRB;
        if (job.u) bits |= (uintptr_t)job.u | (uintptr_t)job.uRB | (uintptr_t)job.v | (uintptr_t)job.vRB;
        if (job.a) bits |= (uintptr_t)job.a | (uintptr_t)job.aRB;
        job.pairAligned = (uint8_t)((bits & m) == 0);
    }

    if (pipelined) {
        uint32_t bands[AB_MAX_BANDS + 1];
        const int nb = planBands(image->height, frameBytes, bands);
        const int lookAhead = (job.chroma == AB_CHROMA_BILINEAR && job.vertBlend) ? 1 : 0;
        const char * kname = "";
        for (int b = 0; b < nb; ++b) {
            const uint32_t r0 = bands[b], r1 = bands[b + 1];
            const uint32_t c0 = r0 >> s.shiftY, c1 = (uint32_t)(((uint64_t)r1 + s.shiftY) >> s.shiftY);
            cudaStream_t us = upStream(b);
            AB_CUDA(regionUpload(&ry, r0, r1, us));
            AB_CUDA(regionUpload(&ru, c0, c1, us));
            AB_CUDA(regionUpload(&rv, c0, c1, us));
            AB_CUDA(regionUpload(&ra, r0, r1, us));
            AB_CUDA(cudaEventRecord(cx->evUp[b], us));
        }
        for (int b = 0; b < nb; ++b) {
            const int need = (b + lookAhead < nb) ? b + lookAhead : nb - 1;
            AB_CUDA(cudaStreamWaitEvent(st, cx->evUp[need], 0));
            job.rowBegin = bands[b];
            job.rowEnd = bands[b + 1];
            AB_CUDA(abLaunchY2R(&job, st, tls.policy, &kname));
            AB_CUDA(cudaEventRecord(cx->evKernel[b], st));
            AB_CUDA(cudaStreamWaitEvent(cx->down, cx->evKernel[b], 0));
            AB_CUDA(regionDownload(&rrgb, bands[b], bands[b + 1], cx->down));
        }
        tls.lastKernel = kname;
        AB_CUDA(cudaStreamSynchronize(cx->down));
        return AVIF_RESULT_OK;
    }

    AB_CUDA(regionUpload(&ry, 0, ry.rows, st));
    AB_CUDA(regionUpload(&ru, 0, ru.rows, st));
    AB_CUDA(regionUpload(&rv, 0, rv.rows, st));
    AB_CUDA(regionUpload(&ra, 0, ra.rows, st));
    if (keepsDstBytes) AB_CUDA(regionUpload(&rrgb, 0, rrgb.rows, st)); /* bytes the path must leave untouched */

    const char * kname = "";
    if (sameDims) {
        AB_CUDA(abLaunchY2R(&job, st, tls.policy, &kname));
    } else {
        /* faithful multi-pass composition for mismatched extents */
        const uint8_t intMul = (job.alphaMul == AB_MUL_INT_PREMUL || job.alphaMul == AB_MUL_INT_UNPREMUL) ? job.alphaMul : 0;
        if (reformatAlpha) {
            AbAlphaJob aj;
            memset(&aj, 0, sizeof(aj));
            aj.src = (job.alphaWrite == AB_ALPHA_COPY) ? ra.dev : NULL;
            aj.dst = rrgb.dev;
            aj.srcRB = job.aRB;
            aj.dstRB = rrgb.devRB;
            aj.width = rgb->width;
            aj.height = rgb->height;
            aj.srcDepth = image->depth;
            aj.dstDepth = rgb->depth;
            aj.srcOff = 0;
            aj.srcPix = s.yuvChanBytes;
            aj.dstOff = s.offA;
            aj.dstPix = s.rgbPixBytes;
            AB_CUDA(abLaunchAlpha(&aj, st));
        }
        AbY2RJob cj = job;
        cj.alphaWrite = AB_ALPHA_KEEP;
        if (intMul) cj.alphaMul = AB_MUL_NONE;
        cj.toF16 = 0;
        AB_CUDA(abLaunchY2R(&cj, st, tls.policy, &kname));
        AbRgbPassJob pj;
        memset(&pj, 0, sizeof(pj));
        pj.pixels = rrgb.dev;
        pj.rowBytes = rrgb.devRB;
        pj.width = rgb->width;
        pj.height = rgb->height;
        pj.depth = rgb->depth;
        pj.maxDiv = makeDiv(s.rgbMaxF);
        if (intMul && premulLayout(rgb->format, &pj)) {
            pj.unmultiply = (uint8_t)(intMul == AB_MUL_INT_UNPREMUL);
            AB_CUDA(abLaunchPremultiply(&pj, st));
        }
        if (job.toF16) {
            pj.nChan = (uint8_t)avifRGBFormatChannelCount(rgb->format);
            pj.f16mul = job.f16mul;
            AB_CUDA(abLaunchToF16(&pj, st));
        }
        kname = "y2r_multipass";
    }
    tls.lastKernel = kname;

    AB_CUDA(regionDownload(&rrgb, 0, rrgb.rows, st));
    if (!async) AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

static avifResult yuvToRgbCommon(const avifImage * image, avifRGBImage * rgb, void * userStream, int async)
{
    AbY2RJob job;
    AbY2RPlan plan;
    avifResult res = y2rBuildJob(image, rgb, &job, &plan); /* validation and constants: no device involved */
    if (res != AVIF_RESULT_OK) return res;
    if (async) return yuvToRgbBody(image, rgb, userStream, 1, job, &plan);
    AbScope sc;
    res = ctxEnter(&sc, deviceOfImages(image, rgb));
    if (res != AVIF_RESULT_OK) return res;
    __atomic_add_fetch(&gHostCallsInFlight, 1, __ATOMIC_RELAXED);
    tlsStageHelpers = stageHelperCount(rgb->maxThreads);
    res = yuvToRgbBody(image, rgb, NULL, 0, job, &plan);
    __atomic_sub_fetch(&gHostCallsInFlight, 1, __ATOMIC_RELAXED);
    return ctxLeave(&sc, res);
}
avifResult avifImageYUVToRGB(const avifImage * image, avifRGBImage * rgb) { return yuvToRgbCommon(image, rgb, NULL, 0); }
avifResult avifB200ImageYUVToRGBHost(const avifImage * image, avifRGBImage * rgb) { return yuvToRgbCommon(image, rgb, NULL, 0); }
avifResult avifB200ImageYUVToRGBAsync(const avifImage * image, avifRGBImage * rgb, void * stream) { return yuvToRgbCommon(image, rgb, stream, 1); }

/* The reference's internal seam, include/avif/internal.h:350-387 (its libyuv hooks): colour conversion only, the
 * caller keeps the surrounding passes.  libavif_b200_hooks.so exports these under the reference's hook names. */
avifResult avifB200HookImageYUVToRGB(const avifImage * image, avifRGBImage * rgb, avifBool reformatAlpha, avifBool * alphaReformatted)
{
    if (alphaReformatted) *alphaReformatted = AVIF_FALSE;
    tlsHook.active = 1;
    tlsHook.reformatAlpha = reformatAlpha ? 1 : 0;
    const avifResult res = yuvToRgbCommon(image, rgb, NULL, 0);
    tlsHook.active = 0;
    if (res == AVIF_RESULT_OK && reformatAlpha && alphaReformatted) *alphaReformatted = AVIF_TRUE; /* the kernel wrote the channel */
    return res;
}
avifResult avifB200HookRGBImageToF16(avifRGBImage * rgb) { return toF16Common(rgb); }

/* ------------------------------------------------------------------------------------------ */
/* RGB -> YUV                                                                                  */

/* src/avif.c:431-489: missing planes are malloc'ed and owned by the image; existing ones are kept */
static avifResult allocatePlanes(avifImage * image, int withAlpha, int validateOnly)
{
    if (image->width == 0 || image->height == 0 || image->depth == 0 || image->depth > 16) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t cs = (image->depth > 8) ? 2 : 1;
    if (image->width > UINT32_MAX / cs) return AVIF_RESULT_INVALID_ARGUMENT;
    const uint32_t fullRowBytes = cs * image->width;
    if (image->height > PTRDIFF_MAX / fullRowBytes) return AVIF_RESULT_INVALID_ARGUMENT;
    const size_t fullSize = (size_t)fullRowBytes * image->height;
    if (validateOnly) return AVIF_RESULT_OK; /* device-resident planes stay the caller's: no allocation, ownership flags untouched */
    if (image->yuvFormat != AVIF_PIXEL_FORMAT_NONE) {
        avifPixelFormatInfo fi;
        avifGetPixelFormatInfo(image->yuvFormat, &fi);
        image->imageOwnsYUVPlanes = AVIF_TRUE;
        if (!image->yuvPlanes[AVIF_CHAN_Y]) {
            image->yuvPlanes[AVIF_CHAN_Y] = (uint8_t *)malloc(fullSize);
            if (!image->yuvPlanes[AVIF_CHAN_Y]) return AVIF_RESULT_OUT_OF_MEMORY;
            image->yuvRowBytes[AVIF_CHAN_Y] = fullRowBytes;
        }
        if (!fi.monochrome) {
            const uint32_t w = (uint32_t)(((uint64_t)image->width + fi.chromaShiftX) >> fi.chromaShiftX);
            const uint32_t h = (uint32_t)(((uint64_t)image->height + fi.chromaShiftY) >> fi.chromaShiftY);
            for (int p = AVIF_CHAN_U; p <= AVIF_CHAN_V; ++p) {
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

typedef struct AbR2YPlan {
    AbState s;
    int monochrome, isGray, hasAlpha, writeAlphaPlane;
} AbR2YPlan;

/* validation, plane allocation, constants (src/reformat.c:221-273); no CUDA */
static avifResult r2yBuildJob(avifImage * image, const avifRGBImage * rgb, int async, AbR2YJob * jobOut, AbR2YPlan * plan)
{
    /* src/reformat.c:221-273 */
    if (!rgb->pixels || rgb->format == AVIF_RGB_FORMAT_RGB_565) return AVIF_RESULT_REFORMAT_FAILED;
    AbState s;
    if (!prepareState(image, rgb, &s)) return AVIF_RESULT_REFORMAT_FAILED;
    if (rgb->isFloat) return AVIF_RESULT_NOT_IMPLEMENTED;

    const int hasAlpha = avifRGBFormatHasAlpha(rgb->format) && !rgb->ignoreAlpha;
    const int wantAlphaPlane = hasAlpha && !tlsHook.active; /* seam: the caller owns allocation and the alpha plane */
    const int monochrome = (image->yuvFormat == AVIF_PIXEL_FORMAT_YUV400);
    const int planesMissing = !image->yuvPlanes[0] || (!monochrome && (!image->yuvPlanes[1] || !image->yuvPlanes[2])) || (wantAlphaPlane && !image->alphaPlane);
    if (planesMissing) {
        /* Missing planes are malloc'ed host memory owned by the image, like the reference.  That is only
         * meaningful when the call may stage through the host. */
        if (async || ptrOnDevice(rgb->pixels)) return AVIF_RESULT_INVALID_ARGUMENT;
    }
    /* The reference (src/avif.c:431-489) marks the image as owner of its planes whenever it passes through here, which is
     * only ever true of host memory it can free().  Planes that live in device memory (stream-ordered entry points, or the
     * synchronous ones called with resident planes) belong to the caller: imageOwns* stay as they are. */
    const int residentPlanes = async || (!planesMissing && ptrOnDevice(image->yuvPlanes[0]));
    avifResult res = allocatePlanes(image, wantAlphaPlane, residentPlanes);
    if (res != AVIF_RESULT_OK) return res;

    int alphaMode = 0;
    if (hasAlpha) {
        if (!rgb->alphaPremultiplied && image->alphaPremultiplied) alphaMode = 1;
        else if (rgb->alphaPremultiplied && !image->alphaPremultiplied) alphaMode = 2;
    }
    if (tlsHook.active && alphaMode != 0) return AVIF_RESULT_NOT_IMPLEMENTED; /* the caller only asks for NO_OP (src/reformat.c:265) */
    const int isGray = avifRGBFormatIsGray(rgb->format);
    if (!isGray && rgb->chromaDownsampling == AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV && image->yuvFormat == AVIF_PIXEL_FORMAT_YUV420) {
        return AVIF_RESULT_NOT_IMPLEMENTED; /* no libsharpyuv: src/reformat.c:256-261, src/reformat_libsharpyuv.c:75-83 */
    }

    AbR2YJob job;
    memset(&job, 0, sizeof(job));
    job.width = image->width;
    job.height = image->height;
    job.yuvMax = (uint32_t)s.yuvMax;
    job.rgbMax = (uint32_t)s.rgbMax;
    job.yuvMaxF = (float)s.yuvMax;
    job.rgbMaxF = s.rgbMaxF;
    job.biasY = s.biasY;
    job.rangeY = s.rangeY;
    if (s.mode == AB_MODE_IDENTITY) { /* :212-213 */
        job.biasUV = s.biasY;
        job.rangeUV = s.rangeY;
    } else {
        job.biasUV = s.biasUV;
        job.rangeUV = s.rangeUV;
    }
    job.rangeYdiv = s.rangeY;
    job.rangeUVdiv = s.rangeUV;
    job.kr = s.kr;
    job.kg = s.kg;
    job.kb = s.kb;
    job.du = makeDiv(2 * (1 - s.kb));
    job.dv = makeDiv(2 * (1 - s.kr));
    job.rgbMaxDiv = makeIntDiv(s.rgbMaxF);
    job.magicK = 0x4B000000u;
    job.nChan = (uint8_t)avifRGBFormatChannelCount(rgb->format);
    job.swapRB = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_BGR || rgb->format == AVIF_RGB_FORMAT_BGRA || rgb->format == AVIF_RGB_FORMAT_ABGR);
    job.alphaFirst = (uint8_t)(rgb->format == AVIF_RGB_FORMAT_ARGB || rgb->format == AVIF_RGB_FORMAT_ABGR || rgb->format == AVIF_RGB_FORMAT_AGRAY);
    job.hasAlphaChan = (uint8_t)avifRGBFormatHasAlpha(rgb->format);
    job.rgbBytes = (uint8_t)s.rgbChanBytes;
    job.yuvBytes = (uint8_t)s.yuvChanBytes;
    job.hasChroma = (uint8_t)!monochrome;
    job.shiftX = (uint8_t)(monochrome ? 0 : s.shiftX);
    job.shiftY = (uint8_t)(monochrome ? 0 : s.shiftY);
    job.mode = (uint8_t)s.mode;
    job.pixBytes = (uint8_t)s.rgbPixBytes;
    job.offR = (uint8_t)s.offR;
    job.offG = (uint8_t)s.offG;
    job.offB = (uint8_t)s.offB;
    job.offA = (uint8_t)s.offA;
    job.offGray = (uint8_t)s.offGray;
    job.isGray = (uint8_t)isGray;
    job.alphaMul = (uint8_t)((alphaMode == 0) ? AB_MUL_NONE : ((alphaMode == 1) ? AB_MUL_FLOAT_MUL : AB_MUL_FLOAT_UNMUL));
    const int writeAlphaPlane = !tlsHook.active && image->alphaPlane && image->alphaRowBytes; /* :545 */
    job.alphaWrite = (uint8_t)(!writeAlphaPlane ? AB_ALPHA_KEEP : (hasAlpha ? AB_ALPHA_COPY : AB_ALPHA_FILL));
    job.alphaRescale = (uint8_t)(image->depth != rgb->depth);

    *jobOut = job;
    plan->s = s;
    plan->monochrome = monochrome;
    plan->isGray = isGray;
    plan->hasAlpha = hasAlpha;
    plan->writeAlphaPlane = writeAlphaPlane;
    return AVIF_RESULT_OK;
}

static avifResult rgbToYuvBody(avifImage * image, const avifRGBImage * rgb, void * userStream, int async, AbR2YJob job, const AbR2YPlan * planIn)
{
    const AbR2YPlan plan = *planIn;
    avifResult res;
    const AbState s = plan.s;
    const int monochrome = plan.monochrome, isGray = plan.isGray, writeAlphaPlane = plan.writeAlphaPlane;
    cudaStream_t st = async ? (cudaStream_t)userStream : cx->stream;

    const uint32_t cw = (uint32_t)(((uint64_t)image->width + s.shiftX) >> s.shiftX);
    const uint32_t ch = (uint32_t)(((uint64_t)image->height + s.shiftY) >> s.shiftY);
    /* gray input: existing chroma planes become constant "half", padding included (:520-542) */
    const int grayChromaFill = isGray;
    AbRegion rrgb, ry, ru, rv, ra;
    regionInit(&rrgb, rgb->pixels, rgb->rowBytes, (size_t)image->width * s.rgbPixBytes, image->height, async);
    regionInit(&ry, image->yuvPlanes[0], image->yuvRowBytes[0], (size_t)image->width * s.yuvChanBytes, image->height, async);
    const int useU = grayChromaFill ? (image->yuvPlanes[1] != NULL) : !monochrome;
    const int useV = grayChromaFill ? (image->yuvPlanes[2] != NULL) : !monochrome;
    regionInit(&ru, useU ? image->yuvPlanes[1] : NULL, image->yuvRowBytes[1], (size_t)cw * s.yuvChanBytes, ch, async);
    regionInit(&rv, useV ? image->yuvPlanes[2] : NULL, image->yuvRowBytes[2], (size_t)cw * s.yuvChanBytes, ch, async);
    if (grayChromaFill) {
        ru.exactPitch = 1;
        rv.exactPitch = 1;
    }
    regionInit(&ra, writeAlphaPlane ? image->alphaPlane : NULL, image->alphaRowBytes, (size_t)image->width * s.yuvChanBytes, image->height, async);

    if (!async) {
        res = arenaReserve(regionArenaBytes(&rrgb) + regionArenaBytes(&ry) + regionArenaBytes(&ru) + regionArenaBytes(&rv) + regionArenaBytes(&ra));
        if (res != AVIF_RESULT_OK) return res;
    }
    uint8_t * cursor = async ? NULL : cx->arena;
    regionPlace(&rrgb, &cursor);
    regionPlace(&ry, &cursor);
    regionPlace(&ru, &cursor);
    regionPlace(&rv, &cursor);
    regionPlace(&ra, &cursor);
    job.rgb = rrgb.dev;
    job.rgbRB = rrgb.devRB;
    job.y = ry.dev;
    job.yRB = ry.devRB;
    job.u = ru.dev;
    job.uRB = ru.devRB;
    job.v = rv.dev;
    job.vRB = rv.devRB;
    job.a = ra.dev;
    job.aRB = ra.devRB;
    job.rowBegin = 0;
    job.rowEnd = image->height;
    {
        /* natural alignment of one pixel pair: 16 / 8 / 4 / 2 bytes for pairs of 16 / 8 / (12|4) / (6|2) bytes */
        const uint32_t pairBytes = 2u * s.rgbPixBytes;
        const uintptr_t srcAlign = (pairBytes % 16 == 0) ? 16 : (pairBytes % 8 == 0) ? 8 : (pairBytes % 4 == 0) ? 4 : 2;
        job.srcPairAligned = (uint8_t)((((uintptr_t)job.rgb | (uintptr_t)job.rgbRB) & (srcAlign - 1)) == 0);
        const uintptr_t m = 2u * s.yuvChanBytes - 1;
        uintptr_t bits = (uintptr_t)job.y | (uintptr_t)job.yRB;
        if (job.u) bits |= (uintptr_t)job.u | (uintptr_t)job.uRB;
        if (job.v) bits |= (uintptr_t)job.v | (uintptr_t)job.vRB;
        if (job.a) bits |= (uintptr_t)job.a | (uintptr_t)job.aRB;
        job.dstPairAligned = (uint8_t)((bits & m) == 0);
    }

    const int anyHost = !rrgb.onDevice || !ry.onDevice || (ru.present && !ru.onDevice) || (rv.present && !rv.onDevice) || (ra.present && !ra.onDevice);
    const size_t frameBytes = (size_t)rrgb.rowUsed * rrgb.rows + (size_t)ry.rowUsed * ry.rows + (ru.present ? 2 * (size_t)ru.rowUsed * ru.rows : 0) +
                              (ra.present ? (size_t)ra.rowUsed * ra.rows : 0);
    const int anyStaged = regionStaged(&rrgb, rrgb.rows) || regionStaged(&ry, ry.rows) || regionStaged(&ra, ra.rows);
    if (!async && anyHost && !anyStaged && !grayChromaFill && frameBytes >= ((size_t)8 << 20) && rgb->rowBytes != 0) {
        /* band pipeline (see planBands): pixels up, kernel, planes down, on three streams */
        uint32_t bands[AB_MAX_BANDS + 1];
        const int nb = planBands(image->height, frameBytes, bands);
        const char * pname = "";
        for (int b = 0; b < nb; ++b) {
            cudaStream_t us = upStream(b);
            AB_CUDA(regionUpload(&rrgb, bands[b], bands[b + 1], us));
            AB_CUDA(cudaEventRecord(cx->evUp[b], us));
        }
        for (int b = 0; b < nb; ++b) {
            const uint32_t r0 = bands[b], r1 = bands[b + 1];
            const uint32_t c0 = r0 >> job.shiftY, c1 = (uint32_t)(((uint64_t)r1 + job.shiftY) >> job.shiftY);
            AB_CUDA(cudaStreamWaitEvent(st, cx->evUp[b], 0));
            job.rowBegin = r0;
            job.rowEnd = r1;
            AB_CUDA(abLaunchR2Y(&job, st, tls.policy, &pname));
            AB_CUDA(cudaEventRecord(cx->evKernel[b], st));
            AB_CUDA(cudaStreamWaitEvent(cx->down, cx->evKernel[b], 0));
            AB_CUDA(regionDownload(&ry, r0, r1, cx->down));
            AB_CUDA(regionDownload(&ru, c0, c1, cx->down));
            AB_CUDA(regionDownload(&rv, c0, c1, cx->down));
            AB_CUDA(regionDownload(&ra, r0, r1, cx->down));
        }
        tls.lastKernel = pname;
        AB_CUDA(cudaStreamSynchronize(cx->down));
        return AVIF_RESULT_OK;
    }

    AB_CUDA(regionUpload(&rrgb, 0, rrgb.rows, st));

    const char * kname = "";
    AB_CUDA(abLaunchR2Y(&job, st, tls.policy, &kname));
    tls.lastKernel = kname;
    if (grayChromaFill) {
        const int half = 1 << (image->depth - 1);
        AbRegion * pr[2] = { &ru, &rv };
        for (int k = 0; k < 2; ++k) {
            if (!pr[k]->present) continue;
            const size_t bytes = (size_t)ch * (size_t)pr[k]->userRB;
            if (s.yuvChanBytes > 1) AB_CUDA(abLaunchFill16(pr[k]->dev, bytes / 2, (uint32_t)half, st));
            else AB_CUDA(abLaunchFill8(pr[k]->dev, bytes, (uint32_t)half, st));
        }
    }

    AB_CUDA(regionDownload(&ry, 0, ry.rows, st));
    AB_CUDA(regionDownload(&ru, 0, ru.rows, st));
    AB_CUDA(regionDownload(&rv, 0, rv.rows, st));
    AB_CUDA(regionDownload(&ra, 0, ra.rows, st));
    if (!async) AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

static avifResult rgbToYuvCommon(avifImage * image, const avifRGBImage * rgb, void * userStream, int async)
{
    AbR2YJob job;
    AbR2YPlan plan;
    avifResult res = r2yBuildJob(image, rgb, async, &job, &plan); /* validation, plane allocation, constants: no device involved */
    if (res != AVIF_RESULT_OK) return res;
    if (async) return rgbToYuvBody(image, rgb, userStream, 1, job, &plan);
    AbScope sc;
    res = ctxEnter(&sc, deviceOfImages(image, rgb));
    if (res != AVIF_RESULT_OK) return res;
    __atomic_add_fetch(&gHostCallsInFlight, 1, __ATOMIC_RELAXED);
    tlsStageHelpers = stageHelperCount(rgb->maxThreads);
    res = rgbToYuvBody(image, rgb, NULL, 0, job, &plan);
    __atomic_sub_fetch(&gHostCallsInFlight, 1, __ATOMIC_RELAXED);
    return ctxLeave(&sc, res);
}
avifResult avifB200HookImageRGBToYUV(avifImage * image, const avifRGBImage * rgb)
{
    /* src/reformat.c:265-273: planes are allocated by the caller, alpha mode is NO_OP, the alpha plane is written by the
     * caller afterwards (:545-569) */
    tlsHook.active = 1;
    const avifResult res = rgbToYuvCommon(image, rgb, NULL, 0);
    tlsHook.active = 0;
    return res;
}
avifResult avifImageRGBToYUV(avifImage * image, const avifRGBImage * rgb) { return rgbToYuvCommon(image, rgb, NULL, 0); }
avifResult avifB200ImageRGBToYUVHost(avifImage * image, const avifRGBImage * rgb) { return rgbToYuvCommon(image, rgb, NULL, 0); }
avifResult avifB200ImageRGBToYUVAsync(avifImage * image, const avifRGBImage * rgb, void * stream) { return rgbToYuvCommon(image, rgb, stream, 1); }

/* ------------------------------------------------------------------------------------------ */
/* "next" rows: alpha range expansion and grid plane staging                    */

/* src/read.c:6724-6764 */
static avifResult limitedToFullAlphaBody(avifImage * image)
{
    avifResult res;
    cudaStream_t st = cx->stream;
    const uint32_t cb = (image->depth > 8) ? 2 : 1;
    AbRegion ra;
    regionInit(&ra, image->alphaPlane, image->alphaRowBytes, (size_t)image->width * cb, image->height, 0);
    res = arenaReserve(regionArenaBytes(&ra));
    if (res != AVIF_RESULT_OK) return res;
    uint8_t * cursor = cx->arena;
    regionPlace(&ra, &cursor);
    AB_CUDA(regionUpload(&ra, 0, ra.rows, st));
    AB_CUDA(abLaunchLimitedToFull(ra.dev, ra.devRB, image->width, image->height, image->depth, st));
    tls.lastKernel = "limited_to_full";
    AB_CUDA(regionDownload(&ra, 0, ra.rows, st));
    AB_CUDA(cudaStreamSynchronize(st));
    return AVIF_RESULT_OK;
}

avifResult avifB200ImageLimitedToFullAlpha(avifImage * image)
{
    if (!image->alphaPlane) return AVIF_RESULT_OK;
    const void * ptrs[1] = { image->alphaPlane };
    AbScope sc;
    avifResult res = ctxEnter(&sc, deviceOfPointers(ptrs, 1));
    if (res != AVIF_RESULT_OK) return res;
    return ctxLeave(&sc, limitedToFullAlphaBody(image));
}

static void planeGeometry(const avifImage * im, int plane, uint32_t * w, uint32_t * h)
{
    avifPixelFormatInfo fi;
    avifGetPixelFormatInfo(im->yuvFormat, &fi);
    if (plane == AVIF_CHAN_Y || plane == AVIF_CHAN_A) {
        *w = im->width;
        *h = im->height;
    } else if (fi.monochrome) {
        *w = *h = 0;
    } else {
        *w = (uint32_t)(((uint64_t)im->width + fi.chromaShiftX) >> fi.chromaShiftX);
        *h = (uint32_t)(((uint64_t)im->height + fi.chromaShiftY) >> fi.chromaShiftY);
    }
}

/* src/avif.c:187-225: row-wise copy of the src-sized area of every selected plane (device, stream-ordered) */
avifResult avifB200ImageCopySamplesAsync(avifImage * dst, const avifImage * src, avifPlanesFlags planes, void * stream)
{
    if (dst->depth != src->depth || dst->yuvFormat != src->yuvFormat) return AVIF_RESULT_INVALID_ARGUMENT;
    const size_t bps = (src->depth > 8) ? 2 : 1;
    for (int c = AVIF_CHAN_Y; c <= AVIF_CHAN_A; ++c) {
        const int isAlpha = (c == AVIF_CHAN_A);
        if (isAlpha ? !(planes & AVIF_PLANES_A) : !(planes & AVIF_PLANES_YUV)) continue;
        const uint8_t * sp = isAlpha ? src->alphaPlane : src->yuvPlanes[c];
        uint8_t * dp = isAlpha ? dst->alphaPlane : dst->yuvPlanes[c];
        const size_t srb = isAlpha ? src->alphaRowBytes : src->yuvRowBytes[c];
        const size_t drb = isAlpha ? dst->alphaRowBytes : dst->yuvRowBytes[c];
        if (!sp || !dp) continue;
        uint32_t w, h;
        planeGeometry(src, c, &w, &h);
        if (!w || !h) continue;
        AB_CUDA(cudaMemcpy2DAsync(dp, drb, sp, srb, (size_t)w * bps, h, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return AVIF_RESULT_OK;
}

/* src/write.c:1151-1205: copy into a larger cell, replicating the last column / row */
avifResult avifB200ImageCopyAndPadAsync(avifImage * dst, const avifImage * src, void * stream)
{
    if (dst->depth != src->depth || dst->yuvFormat != src->yuvFormat || dst->width < src->width || dst->height < src->height) {
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    const uint32_t bps = (src->depth > 8) ? 2 : 1;
    for (int c = AVIF_CHAN_Y; c <= AVIF_CHAN_A; ++c) {
        const int isAlpha = (c == AVIF_CHAN_A);
        const uint8_t * sp = isAlpha ? src->alphaPlane : src->yuvPlanes[c];
        uint8_t * dp = isAlpha ? dst->alphaPlane : dst->yuvPlanes[c];
        if (!sp || !dp) continue;
        uint32_t sw, sh, dw, dh;
        planeGeometry(src, c, &sw, &sh);
        planeGeometry(dst, c, &dw, &dh);
        if (!sw || !sh) continue;
        AB_CUDA(abLaunchCopyPad(sp, isAlpha ? src->alphaRowBytes : src->yuvRowBytes[c], sw, sh, dp, isAlpha ? dst->alphaRowBytes : dst->yuvRowBytes[c], dw,
                                dh, bps, stream));
    }
    tls.lastKernel = "copy_pad";
    return AVIF_RESULT_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* batches of device-resident frames: as few launches as possible (see plan.h, AbY2RBatch)     */

avifResult avifB200ImageYUVToRGBBatchAsync(const avifImage * const * images, avifRGBImage * const * rgbs, uint32_t count, void * stream)
{
    if (count == 0) return AVIF_RESULT_OK;
    if (!images || !rgbs) return AVIF_RESULT_INVALID_ARGUMENT;
    AbY2RJob * jobs = (AbY2RJob *)malloc((size_t)count * sizeof(AbY2RJob));
    if (!jobs) return AVIF_RESULT_OUT_OF_MEMORY;
    avifResult res = AVIF_RESULT_OK;
    int fusable = 1;
    for (uint32_t k = 0; k < count && res == AVIF_RESULT_OK; ++k) {
        AbY2RPlan plan;
        res = y2rBuildJob(images[k], rgbs[k], &jobs[k], &plan);
        if (res != AVIF_RESULT_OK) break;
        const avifImage * image = images[k];
        const avifRGBImage * rgb = rgbs[k];
        if (rgb->width != image->width || rgb->height != image->height) fusable = 0; /* multi-pass composition: per image */
        AbY2RJob * job = &jobs[k];
        const int needSrcAlpha = (job->alphaWrite == AB_ALPHA_COPY) || job->alphaMul == AB_MUL_FLOAT_MUL || job->alphaMul == AB_MUL_FLOAT_UNMUL;
        job->y = image->yuvPlanes[0];
        job->yRB = image->yuvRowBytes[0];
        job->u = plan.hasChroma ? image->yuvPlanes[1] : NULL;
        job->uRB = image->yuvRowBytes[1];
        job->v = plan.hasChroma ? image->yuvPlanes[2] : NULL;
        job->vRB = image->yuvRowBytes[2];
        job->a = needSrcAlpha ? image->alphaPlane : NULL;
        job->aRB = job->a ? image->alphaRowBytes : 0;
        job->rgb = rgb->pixels;
        job->rgbRB = rgb->rowBytes;
        job->rowBegin = 0;
        job->rowEnd = image->height;
        const uintptr_t m = 2u * plan.s.yuvChanBytes - 1;
        uintptr_t bits = (uintptr_t)job->y | (uintptr_t)job->yRB;
        if (job->u) bits |= (uintptr_t)job->u | (uintptr_t)job->uRB | (uintptr_t)job->v | (uintptr_t)job->vRB;
        if (job->a) bits |= (uintptr_t)job->a | (uintptr_t)job->aRB;
        job->pairAligned = (uint8_t)((bits & m) == 0);
    }
    if (res == AVIF_RESULT_OK && !avifB200DeviceAvailable()) res = AVIF_RESULT_UNKNOWN_ERROR;
    if (res == AVIF_RESULT_OK) {
        if (fusable) {
            const char * kname = "";
            res = cudaToResult(abLaunchY2RBatch(jobs, count, stream, tls.policy, &kname));
            tls.lastKernel = kname;
        } else {
            for (uint32_t k = 0; k < count && res == AVIF_RESULT_OK; ++k) res = yuvToRgbCommon(images[k], rgbs[k], stream, 1);
        }
    }
    free(jobs);
    return res;
}

/* Rows [rowBegin, rowEnd) of a resident frame: what a tile / band pipeline calls as decoded rows arrive.  The image and rgb
 * structs describe the WHOLE frame (bilinear upsampling replicates chroma only at the real picture edges, src/reformat.c:768-795,
 * and reads one chroma row beyond the band inside the picture); of the planes only the rows the band touches need to exist behind
 * the pointers -- luma / alpha rows rowBegin..rowEnd-1, chroma rows (rowBegin >> shiftY) - 1 .. ((rowEnd - 1) >> shiftY) + 1 clipped
 * to the plane -- so a caller may pass "virtual" base pointers (band buffer minus rowBegin * rowBytes).  rowBegin must be even for
 * vertically subsampled formats.  Device pointers, stream-ordered. */
avifResult avifB200ImageYUVToRGBRowsAsync(const avifImage * image, avifRGBImage * rgb, uint32_t rowBegin, uint32_t rowEnd, void * stream)
{
    AbY2RJob job;
    AbY2RPlan plan;
    avifResult res = y2rBuildJob(image, rgb, &job, &plan);
    if (res != AVIF_RESULT_OK) return res;
    if (rgb->width != image->width || rgb->height != image->height) return AVIF_RESULT_INVALID_ARGUMENT;
    if (rowBegin > rowEnd || rowEnd > image->height || (plan.s.shiftY && (rowBegin & 1))) return AVIF_RESULT_INVALID_ARGUMENT;
    if (job.alphaWrite == AB_ALPHA_KEEP && plan.rgbHasAlpha) {
        /* untouched alpha bytes: fine for resident buffers (nothing is staged), nothing to do */
    }
    if (!avifB200DeviceAvailable()) return AVIF_RESULT_UNKNOWN_ERROR;
    if (rowBegin == rowEnd) return AVIF_RESULT_OK;
    const int needSrcAlpha = (job.alphaWrite == AB_ALPHA_COPY) || job.alphaMul == AB_MUL_FLOAT_MUL || job.alphaMul == AB_MUL_FLOAT_UNMUL;
    job.y = image->yuvPlanes[0];
    job.yRB = image->yuvRowBytes[0];
    job.u = plan.hasChroma ? image->yuvPlanes[1] : NULL;
    job.uRB = image->yuvRowBytes[1];
    job.v = plan.hasChroma ? image->yuvPlanes[2] : NULL;
    job.vRB = image->yuvRowBytes[2];
    job.a = needSrcAlpha ? image->alphaPlane : NULL;
    job.aRB = job.a ? image->alphaRowBytes : 0;
    job.rgb = rgb->pixels;
    job.rgbRB = rgb->rowBytes;
    job.rowBegin = rowBegin;
    job.rowEnd = rowEnd;
    const uintptr_t m = 2u * plan.s.yuvChanBytes - 1;
    uintptr_t bits = (uintptr_t)job.y | (uintptr_t)job.yRB;
    if (job.u) bits |= (uintptr_t)job.u | (uintptr_t)job.uRB | (uintptr_t)job.v | (uintptr_t)job.vRB;
    if (job.a) bits |= (uintptr_t)job.a | (uintptr_t)job.aRB;
    job.pairAligned = (uint8_t)((bits & m) == 0);
    const char * kname = "";
    res = cudaToResult(abLaunchY2R(&job, stream, tls.policy, &kname));
    tls.lastKernel = kname;
    return res;
}

avifResult avifB200ImageRGBToYUVBatchAsync(avifImage * const * images, const avifRGBImage * const * rgbs, uint32_t count, void * stream)
{
    if (count == 0) return AVIF_RESULT_OK;
    if (!images || !rgbs) return AVIF_RESULT_INVALID_ARGUMENT;
    AbR2YJob * jobs = (AbR2YJob *)malloc((size_t)count * sizeof(AbR2YJob));
    if (!jobs) return AVIF_RESULT_OUT_OF_MEMORY;
    avifResult res = AVIF_RESULT_OK;
    int fusable = 1;
    for (uint32_t k = 0; k < count && res == AVIF_RESULT_OK; ++k) {
        AbR2YPlan plan;
        res = r2yBuildJob(images[k], rgbs[k], 1, &jobs[k], &plan);
        if (res != AVIF_RESULT_OK) break;
        if (plan.isGray) fusable = 0; /* gray input also fills the chroma planes: per image */
        avifImage * image = images[k];
        AbR2YJob * job = &jobs[k];
        job->rgb = rgbs[k]->pixels;
        job->rgbRB = rgbs[k]->rowBytes;
        job->y = image->yuvPlanes[0];
        job->yRB = image->yuvRowBytes[0];
        job->u = plan.monochrome ? NULL : image->yuvPlanes[1];
        job->uRB = image->yuvRowBytes[1];
        job->v = plan.monochrome ? NULL : image->yuvPlanes[2];
        job->vRB = image->yuvRowBytes[2];
        job->a = plan.writeAlphaPlane ? image->alphaPlane : NULL;
        job->aRB = image->alphaRowBytes;
        job->rowBegin = 0;
        job->rowEnd = image->height;
        const uint32_t pairBytes = 2u * plan.s.rgbPixBytes;
        const uintptr_t srcAlign = (pairBytes % 16 == 0) ? 16 : (pairBytes % 8 == 0) ? 8 : (pairBytes % 4 == 0) ? 4 : 2;
        job->srcPairAligned = (uint8_t)((((uintptr_t)job->rgb | (uintptr_t)job->rgbRB) & (srcAlign - 1)) == 0);
        const uintptr_t m = 2u * plan.s.yuvChanBytes - 1;
        uintptr_t bits = (uintptr_t)job->y | (uintptr_t)job->yRB;
        if (job->u) bits |= (uintptr_t)job->u | (uintptr_t)job->uRB;
        if (job->v) bits |= (uintptr_t)job->v | (uintptr_t)job->vRB;
        if (job->a) bits |= (uintptr_t)job->a | (uintptr_t)job->aRB;
        job->dstPairAligned = (uint8_t)((bits & m) == 0);
    }
    if (res == AVIF_RESULT_OK && !avifB200DeviceAvailable()) res = AVIF_RESULT_UNKNOWN_ERROR;
    if (res == AVIF_RESULT_OK) {
        if (fusable) {
            const char * kname = "";
            res = cudaToResult(abLaunchR2YBatch(jobs, count, stream, tls.policy, &kname));
            tls.lastKernel = kname;
        } else {
            for (uint32_t k = 0; k < count && res == AVIF_RESULT_OK; ++k) res = rgbToYuvCommon(images[k], rgbs[k], stream, 1);
        }
    }
    free(jobs);
    return res;
}
