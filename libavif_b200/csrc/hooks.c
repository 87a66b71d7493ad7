/*
 * hooks.c -- libavif_b200_hooks.so: the reference's INTERNAL seam (B-2).
 *
 * src/reformat.c and src/alpha.c of the reference call six functions before their built-in loops and fall through when
 * one answers AVIF_RESULT_NOT_IMPLEMENTED (include/avif/internal.h:350-387; call sites src/reformat.c:257,266,1423,1454,
 * src/alpha.c:163,350).  In the reference they live in src/reformat_libyuv.c / src/reformat_libsharpyuv.c.  This file is
 * the replacement for those two: link an UNMODIFIED libavif with it instead and every conversion of the reformat path runs
 * on the B200 while the rest of the library (and its whole test-suite) stays as it is.
 *
 * The functions forward to libavif_b200.so, which is opened with RTLD_LOCAL: that library also exports the PUBLIC entry
 * points (avifImageYUVToRGB ...; seam B-1) and those must not interpose the reference's own definitions here.
 * Location: $AVIF_B200_LIB, else libavif_b200.so next to this library.
 *
 * No CPU arithmetic happens here.  If the GPU library cannot be loaded the hooks answer AVIF_RESULT_UNKNOWN_ERROR --
 * loudly wrong rather than silently falling back to the caller's CPU loops.
 */
#define _GNU_SOURCE
#include "avif_b200.h"

#include <dlfcn.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define HOOK_API __attribute__((visibility("default")))

typedef struct avifReformatState avifReformatState; /* opaque here: only passed through (include/avif/internal.h:311) */

static struct {
    avifResult (*y2r)(const avifImage *, avifRGBImage *, avifBool, avifBool *);
    avifResult (*r2y)(avifImage *, const avifRGBImage *);
    avifResult (*premul)(avifRGBImage *);
    avifResult (*unpremul)(avifRGBImage *);
    avifResult (*f16)(avifRGBImage *);
    int ok;
} gpu;
static pthread_once_t once = PTHREAD_ONCE_INIT;

static void loadGpuLibrary(void)
{
    char path[4096];
    const char * env = getenv("AVIF_B200_LIB");
    if (env && env[0]) {
        snprintf(path, sizeof(path), "%s", env);
    } else {
        Dl_info self;
        if (!dladdr((void *)&loadGpuLibrary, &self) || !self.dli_fname) return;
        snprintf(path, sizeof(path), "%s", self.dli_fname);
        char * slash = strrchr(path, '/');
        if (!slash) return;
        snprintf(slash + 1, sizeof(path) - (size_t)(slash + 1 - path), "libavif_b200.so");
    }
    void * h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "libavif_b200_hooks: cannot load %s: %s\n", path, dlerror());
        return;
    }
    *(void **)&gpu.y2r = dlsym(h, "avifB200HookImageYUVToRGB");
    *(void **)&gpu.r2y = dlsym(h, "avifB200HookImageRGBToYUV");
    *(void **)&gpu.premul = dlsym(h, "avifB200RGBImagePremultiplyAlphaHost");
    *(void **)&gpu.unpremul = dlsym(h, "avifB200RGBImageUnpremultiplyAlphaHost");
    *(void **)&gpu.f16 = dlsym(h, "avifB200HookRGBImageToF16");
    gpu.ok = gpu.y2r && gpu.r2y && gpu.premul && gpu.unpremul && gpu.f16;
}

static int ready(void)
{
    pthread_once(&once, loadGpuLibrary);
    return gpu.ok;
}

/* include/avif/internal.h:354 ; called from src/reformat.c:266 when the alpha mode is NO_OP */
HOOK_API avifResult avifImageRGBToYUVLibYUV(avifImage * image, const avifRGBImage * rgb)
{
    return ready() ? gpu.r2y(image, rgb) : AVIF_RESULT_UNKNOWN_ERROR;
}

/* include/avif/internal.h:369 ; called from src/reformat.c:1454 */
HOOK_API avifResult avifImageYUVToRGBLibYUV(const avifImage * image, avifRGBImage * rgb, avifBool reformatAlpha, avifBool * alphaReformattedWithLibYUV)
{
    if (alphaReformattedWithLibYUV) *alphaReformattedWithLibYUV = AVIF_FALSE;
    return ready() ? gpu.y2r(image, rgb, reformatAlpha, alphaReformattedWithLibYUV) : AVIF_RESULT_UNKNOWN_ERROR;
}

/* include/avif/internal.h:375 ; src/reformat.c:257.  The sharp-YUV downsampler is not part of the path:
 * same answer as a reference built without libsharpyuv (src/reformat_libsharpyuv.c:75-83). */
HOOK_API avifResult avifImageRGBToYUVLibSharpYUV(avifImage * image, const avifRGBImage * rgb, const avifReformatState * state)
{
    (void)image;
    (void)rgb;
    (void)state;
    return AVIF_RESULT_NOT_IMPLEMENTED;
}

/* include/avif/internal.h:381 ; src/reformat.c:1423 */
HOOK_API avifResult avifRGBImageToF16LibYUV(avifRGBImage * rgb) { return ready() ? gpu.f16(rgb) : AVIF_RESULT_UNKNOWN_ERROR; }

/* include/avif/internal.h:386-387 ; src/alpha.c:163, :350 (the caller has already rejected formats without alpha) */
HOOK_API avifResult avifRGBImagePremultiplyAlphaLibYUV(avifRGBImage * rgb) { return ready() ? gpu.premul(rgb) : AVIF_RESULT_UNKNOWN_ERROR; }
HOOK_API avifResult avifRGBImageUnpremultiplyAlphaLibYUV(avifRGBImage * rgb) { return ready() ? gpu.unpremul(rgb) : AVIF_RESULT_UNKNOWN_ERROR; }

/* src/reformat_libyuv.c:20-25 reports the libyuv version; 0 = "no libyuv" keeps callers that gate on it (avifenc's
 * --libyuv options, tests that skip libyuv-only expectations) on the built-in-path expectations this seam reproduces. */
HOOK_API unsigned int avifLibYUVVersion(void) { return 0; }
