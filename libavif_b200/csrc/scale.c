/*
 * scale.c -- host side of avifImageScale: the reference's checks and plane bookkeeping (src/scale.c:24-201)
 * and libyuv's stepping state machines (third_party/libyuv/source/scale.c, scale_common.c in the reference tree), walked once
 * per plane to tabulate, for every destination row and column, the source positions the reference's loops would visit.
 * Only integers about POSITIONS are computed here (O(width + height)); every sample is read, filtered and written by
 * scale_kernels.cu on the device.  No CPU pixel arithmetic, no fallback.
 */
#define AVIF_B200_BUILDING
#include "avif_b200.h"

#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "plan.h"

enum { F_NONE = 0, F_LINEAR = 1, F_BILINEAR = 2, F_BOX = 3 }; /* libyuv FilterMode, include/libyuv/scale.h:17-22 */

/* scale_common.c:428-469 */
static int filterReduce(int sw, int sh, int dw, int dh, int f)
{
    if (f == F_BOX && (dw * 2 >= sw || dh * 2 >= sh)) f = F_BILINEAR;
    if (f == F_BILINEAR) {
        if (sh == 1) f = F_LINEAR;
        if (dh == sh || dh * 3 == sh) f = F_LINEAR;
        if (sw == 1) f = F_NONE;
    }
    if (f == F_LINEAR) {
        if (sw == 1) f = F_NONE;
        if (dw == sw || dw * 3 == sw) f = F_NONE;
    }
    return f;
}
static int fixedDiv(int num, int div) { return (int)(((int64_t)num << 16) / div); }                        /* :472-474 */
static int fixedDiv1(int num, int div) { return (int)((((int64_t)num << 16) - 0x00010001) / (div - 1)); } /* :477-479 */
static int centerStart(int d, int s) { return (d < 0) ? -((-d >> 1) + s) : ((d >> 1) + s); }              /* CENTERSTART */
static int min1(int v) { return v < 1 ? 1 : v; }

/* scale_common.c:484-553 for positive widths; outputs keep their previous value where the reference leaves them untouched */
static void slope(int sw, int sh, int dw, int dh, int f, int * x, int * y, int * dx, int * dy)
{
    if (dw == 1 && sw >= 32768) dw = sw;
    if (dh == 1 && sh >= 32768) dh = sh;
    if (f == F_BOX) {
        *dx = fixedDiv(sw, dw);
        *dy = fixedDiv(sh, dh);
        *x = 0;
        *y = 0;
    } else if (f == F_BILINEAR) {
        if (dw <= sw) {
            *dx = fixedDiv(sw, dw);
            *x = centerStart(*dx, -32768);
        } else if (sw > 1 && dw > 1) {
            *dx = fixedDiv1(sw, dw);
            *x = 0;
        }
        if (dh <= sh) {
            *dy = fixedDiv(sh, dh);
            *y = centerStart(*dy, -32768);
        } else if (sh > 1 && dh > 1) {
            *dy = fixedDiv1(sh, dh);
            *y = 0;
        }
    } else if (f == F_LINEAR) {
        if (dw <= sw) {
            *dx = fixedDiv(sw, dw);
            *x = centerStart(*dx, -32768);
        } else if (sw > 1 && dw > 1) {
            *dx = fixedDiv1(sw, dw);
            *x = 0;
        }
        *dy = fixedDiv(sh, dh);
        *y = *dy >> 1;
    } else {
        *dx = fixedDiv(sw, dw);
        *dy = fixedDiv(sh, dh);
        *x = centerStart(*dx, 0);
        *y = centerStart(*dy, 0);
    }
}

typedef struct ScalePlan {
    int mode; /* 0 = straight copy */
    int linearOnly;
    int minBoxWidth;
    int colsArith; /* 1: column positions are x0 + i * dx (16.16) -- the kernel may compute them instead of reading cols[] */
    int32_t x0, dx;
    int32_t * rows; /* 4 per destination row */
    int32_t * cols; /* 4 per destination column */
} ScalePlan;

static void planFree(ScalePlan * p)
{
    free(p->rows);
    free(p->cols);
    p->rows = p->cols = NULL;
}

/* positions of ScaleFilterCols / ScaleCols: x, x+dx, ... (scale_common.c:114-130, :197-222) */
static void colsStepped(ScalePlan * p, int dw, int x, int dx)
{
    if (x >= 0 && dx > 0 && (int64_t)x + (int64_t)dw * dx < 0x7fffffffll) { /* no 32-bit wrap on the way: closed form == the walk */
        p->colsArith = 1;
        p->x0 = x;
        p->dx = dx;
    }
    for (int i = 0; i < dw; ++i) {
        p->cols[4 * i + 0] = x >> 16;
        p->cols[4 * i + 1] = x & 0xffff;
        x += dx;
    }
}

/* The dispatch of ScalePlane (scale.c:829-897), ScalePlane_16 (:899-971) and ScalePlane_12 (:973-1007) for kFilterBox, and
 * the row / column walk of the loop it lands in. */
static int planScale(int sw, int sh, int dw, int dh, int sixteen, ScalePlan * p)
{
    memset(p, 0, sizeof(*p));
    const int f = filterReduce(sw, sh, dw, dh, F_BOX);
    int mode = -1;
    if (sixteen) { /* ScalePlane_12 tries its own 2x paths first */
        if ((dw + 1) / 2 == sw && f == F_LINEAR) mode = AB_SCALE_UP2_LINEAR;
        else if ((dh + 1) / 2 == sh && (dw + 1) / 2 == sw && (f == F_BILINEAR || f == F_BOX)) mode = AB_SCALE_UP2_BILINEAR;
    }
    if (mode < 0) {
        if (dw == sw && dh == sh) mode = 0;
        else if (dw == sw && f != F_BOX) mode = AB_SCALE_VERTICAL;
        else if (f == F_BOX && dh * 2 < sh) mode = AB_SCALE_BOX;
        else if ((dw + 1) / 2 == sw && f == F_LINEAR) mode = AB_SCALE_UP2_LINEAR;
        else if ((dh + 1) / 2 == sh && (dw + 1) / 2 == sw && (f == F_BILINEAR || f == F_BOX)) mode = AB_SCALE_UP2_BILINEAR;
        else if (f && dh > sh) mode = AB_SCALE_BILINEAR_UP;
        else if (f) mode = AB_SCALE_BILINEAR_DOWN;
        else mode = AB_SCALE_SIMPLE;
    }
    p->mode = mode;
    if (mode == 0) return 1;
    p->rows = (int32_t *)calloc((size_t)dh * 4, sizeof(int32_t));
    p->cols = (int32_t *)calloc((size_t)dw * 4, sizeof(int32_t));
    if (!p->rows || !p->cols) {
        planFree(p);
        return 0;
    }
    int x = 0, y = 0, dx = 0, dy = 0;
    switch (mode) {
        case AB_SCALE_VERTICAL: { /* scale.c:855-871 + scale_common.c:348-386 */
            if (dh <= sh) {
                dy = fixedDiv(sh, dh);
                y = centerStart(dy, -32768);
            } else if (sh > 1 && dh > 1) {
                dy = fixedDiv1(sh, dh);
            }
            const int maxY = (sh > 1) ? ((sh - 1) << 16) - 1 : 0;
            for (int j = 0; j < dh; ++j) {
                if (y > maxY) y = maxY;
                p->rows[4 * j + 0] = y >> 16;
                p->rows[4 * j + 1] = f ? ((y >> 8) & 255) : 0;
                y += dy;
            }
            break;
        }
        case AB_SCALE_BOX: { /* scale.c:148-198 / :200-256 */
            slope(sw, sh, dw, dh, F_BOX, &x, &y, &dx, &dy);
            const int maxY = sh << 16;
            for (int j = 0; j < dh; ++j) {
                const int iy = y >> 16;
                y += dy;
                if (y > maxY) y = maxY;
                p->rows[4 * j + 0] = iy;
                p->rows[4 * j + 1] = min1((y >> 16) - iy);
            }
            if (x == 0 && dx > 0 && (int64_t)(dw + 1) * dx < 0x7fffffffll) { /* box i spans columns (i*dx) >> 16 .. ((i+1)*dx) >> 16 in both forms below */
                p->colsArith = 1;
                p->x0 = 0;
                p->dx = dx;
            }
            if (dx & 0xffff) { /* ScaleAddCols2: two box widths */
                p->minBoxWidth = dx >> 16;
                for (int i = 0; i < dw; ++i) {
                    const int ix = x >> 16;
                    x += dx;
                    const int bw = min1((x >> 16) - ix);
                    p->cols[4 * i + 0] = ix;
                    p->cols[4 * i + 1] = bw;
                    p->cols[4 * i + 2] = bw - p->minBoxWidth;
                }
            } else { /* ScaleAddCols1 (and ScaleAddCols0 when dx == 1.0): one box width */
                const int bw = min1(dx >> 16);
                p->minBoxWidth = bw;
                int xs = x >> 16;
                for (int i = 0; i < dw; ++i) {
                    p->cols[4 * i + 0] = xs;
                    p->cols[4 * i + 1] = bw;
                    xs += bw;
                }
            }
            break;
        }
        case AB_SCALE_UP2_LINEAR: { /* scale.c:467-501 */
            if (dh == 1) {
                p->rows[0] = (sh - 1) / 2;
            } else {
                dy = fixedDiv(sh - 1, dh - 1);
                y = (1 << 15) - 1;
                for (int j = 0; j < dh; ++j) {
                    p->rows[4 * j + 0] = y >> 16;
                    y += dy;
                }
            }
            break;
        }
        case AB_SCALE_UP2_BILINEAR: break; /* positions follow from (i, j) alone */
        case AB_SCALE_BILINEAR_UP: { /* scale.c:371-466: two cached filtered rows; which source rows they hold is state */
            slope(sw, sh, dw, dh, f, &x, &y, &dx, &dy);
            const int maxY = (sh - 1) << 16;
            if (y > maxY) y = maxY;
            int yi = y >> 16;
            int src = yi;          /* row the reference's `src` pointer is at */
            int bufA = src;        /* row held at rowptr */
            if (sh > 1) ++src;
            int bufB = src;        /* row held at rowptr + rowstride */
            if (sh > 2) ++src;
            int lasty = yi;
            for (int j = 0; j < dh; ++j) {
                yi = y >> 16;
                if (yi != lasty) {
                    if (y > maxY) {
                        y = maxY;
                        yi = y >> 16;
                        src = yi;
                    }
                    if (yi != lasty) {
                        bufA = src; /* filtered into rowptr ... */
                        const int t = bufA;
                        bufA = bufB; /* ... then the buffers swap roles */
                        bufB = t;
                        lasty = yi;
                        if ((y + 65536) < maxY) ++src;
                    }
                }
                p->rows[4 * j + 0] = bufA;
                p->rows[4 * j + 1] = bufB;
                p->rows[4 * j + 2] = (y >> 8) & 255;
                y += dy;
            }
            p->linearOnly = (f == F_LINEAR);
            colsStepped(p, dw, x, dx);
            break;
        }
        case AB_SCALE_BILINEAR_DOWN: { /* scale.c:258-312 */
            slope(sw, sh, dw, dh, f, &x, &y, &dx, &dy);
            const int maxY = (sh - 1) << 16;
            if (y > maxY) y = maxY;
            for (int j = 0; j < dh; ++j) {
                p->rows[4 * j + 0] = y >> 16;
                p->rows[4 * j + 1] = (y >> 8) & 255;
                y += dy;
                if (y > maxY) y = maxY;
            }
            p->linearOnly = (f == F_LINEAR);
            colsStepped(p, dw, x, dx);
            break;
        }
        default: { /* AB_SCALE_SIMPLE, scale.c:763-793 */
            slope(sw, sh, dw, dh, F_NONE, &x, &y, &dx, &dy);
            for (int j = 0; j < dh; ++j) {
                p->rows[4 * j + 0] = y >> 16;
                y += dy;
            }
            if (sw * 2 == dw && x < 0x8000) { /* ScaleColsUp2 */
                for (int i = 0; i < dw; ++i) p->cols[4 * i + 0] = i >> 1;
                p->colsArith = 1; /* i >> 1 == (i * 0x8000) >> 16 */
                p->x0 = 0;
                p->dx = 0x8000;
            } else {
                colsStepped(p, dw, x, dx);
            }
            break;
        }
    }
    if (p->colsArith) { /* the closed form has to reproduce the walk: checked once per plan, O(width) */
        const int box = (mode == AB_SCALE_BOX), frac = (mode == AB_SCALE_BILINEAR_UP || mode == AB_SCALE_BILINEAR_DOWN);
        for (int i = 0; i < dw && p->colsArith; ++i) {
            const int64_t t = (int64_t)p->x0 + (int64_t)i * p->dx, t1 = t + p->dx;
            if (p->cols[4 * i + 0] != (int32_t)(t >> 16)) p->colsArith = 0;
            if (box && p->cols[4 * i + 1] != (int32_t)((t1 >> 16) - (t >> 16))) p->colsArith = 0;
            if (frac && p->cols[4 * i + 1] != (int32_t)(t & 0xffff)) p->colsArith = 0;
        }
    }
    return 1;
}

/* Test hook (no device involved): the loop libyuv's dispatcher picks for one plane and the position tables this file would
 * hand to the kernel.  rows / cols receive 4 ints per destination row / column when non-NULL (caller-allocated). */
avifResult avifB200DebugScalePlan(int srcWidth, int srcHeight, int dstWidth, int dstHeight, int sixteenBit, int * mode, int * linearOnly,
                                  int * minBoxWidth, int32_t * rows, int32_t * cols)
{
    if (srcWidth <= 0 || srcHeight <= 0 || dstWidth <= 0 || dstHeight <= 0) return AVIF_RESULT_INVALID_ARGUMENT;
    ScalePlan plan;
    if (!planScale(srcWidth, srcHeight, dstWidth, dstHeight, sixteenBit, &plan)) return AVIF_RESULT_OUT_OF_MEMORY;
    if (mode) *mode = plan.mode;
    if (linearOnly) *linearOnly = plan.linearOnly;
    if (minBoxWidth) *minBoxWidth = plan.minBoxWidth;
    if (plan.mode) {
        if (rows) memcpy(rows, plan.rows, (size_t)dstHeight * 4 * sizeof(int32_t));
        if (cols) memcpy(cols, plan.cols, (size_t)dstWidth * 4 * sizeof(int32_t));
    }
    planFree(&plan);
    return AVIF_RESULT_OK;
}

typedef struct PlaneIO { /* one plane, both ends on the device; strides in bytes */
    const uint8_t * src;
    uint8_t * dst;
    int64_t srcRB, dstRB;
    int sw, sh, dw, dh;
} PlaneIO;

/* Test hook, second half: whether the kernel may compute this plane's column positions in closed form (x0 + i * dx, 16.16) instead of
 * reading cols[] -- what planScale decided after checking the closed form against its own walk. */
avifResult avifB200DebugScalePlanColumns(int srcWidth, int srcHeight, int dstWidth, int dstHeight, int sixteenBit, int * closedForm, int32_t * x0, int32_t * dx)
{
    if (srcWidth <= 0 || srcHeight <= 0 || dstWidth <= 0 || dstHeight <= 0) return AVIF_RESULT_INVALID_ARGUMENT;
    ScalePlan plan;
    if (!planScale(srcWidth, srcHeight, dstWidth, dstHeight, sixteenBit, &plan)) return AVIF_RESULT_OUT_OF_MEMORY;
    if (closedForm) *closedForm = plan.colsArith;
    if (x0) *x0 = plan.x0;
    if (dx) *dx = plan.dx;
    planFree(&plan);
    return AVIF_RESULT_OK;
}

/* Position tables depend on the four dimensions (and the sample size) only, and a scaling service sees the same few geometries
 * over and over (tiles of one size, gain maps of one size): the first call for a geometry replays libyuv's state machines on the
 * host and uploads the tables, later calls find them resident.  Process-wide, per device, at most kPlanCacheSize geometries
 * (a few megabytes); entries are immutable and live until the process exits; when the cache is full further geometries are
 * planned and uploaded per call as before. */
enum { kPlanCacheSize = 32 };
typedef struct PlanCacheEntry {
    int used, device, sw, sh, dw, dh, sixteen;
    int mode, linearOnly, minBoxWidth, colsArith;
    int32_t x0, dx;
    int32_t * dRows;
    int32_t * dCols;
} PlanCacheEntry;
static PlanCacheEntry gPlanCache[kPlanCacheSize];
static pthread_mutex_t gPlanLock = PTHREAD_MUTEX_INITIALIZER;

/* 1: *out is valid (found or inserted); 0: not cacheable now (cache full / no memory) -- the caller plans per call */
static int planCacheGet(int sw, int sh, int dw, int dh, int sixteen, PlanCacheEntry * out)
{
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int found = 0;
    pthread_mutex_lock(&gPlanLock);
    int freeSlot = -1;
    for (int k = 0; k < kPlanCacheSize; ++k) {
        PlanCacheEntry * c = &gPlanCache[k];
        if (!c->used) {
            if (freeSlot < 0) freeSlot = k;
            continue;
        }
        if (c->device == device && c->sw == sw && c->sh == sh && c->dw == dw && c->dh == dh && c->sixteen == sixteen) {
            *out = *c;
            found = 1;
            break;
        }
    }
    if (!found && freeSlot >= 0) {
        ScalePlan plan;
        memset(&plan, 0, sizeof(plan));
        if (planScale(sw, sh, dw, dh, sixteen, &plan)) {
            PlanCacheEntry c;
            memset(&c, 0, sizeof(c));
            c.used = 1;
            c.device = device;
            c.sw = sw; c.sh = sh; c.dw = dw; c.dh = dh; c.sixteen = sixteen;
            c.mode = plan.mode;
            c.linearOnly = plan.linearOnly;
            c.minBoxWidth = plan.minBoxWidth;
            c.colsArith = plan.colsArith;
            c.x0 = plan.x0;
            c.dx = plan.dx;
            int ok = 1;
            if (plan.mode) {
                const size_t rb = (size_t)dh * 4 * sizeof(int32_t), cb = (size_t)dw * 4 * sizeof(int32_t);
                ok = cudaMalloc((void **)&c.dRows, rb + cb) == cudaSuccess;
                if (ok) {
                    c.dCols = c.dRows + (size_t)dh * 4;
                    ok = cudaMemcpy(c.dRows, plan.rows, rb, cudaMemcpyHostToDevice) == cudaSuccess &&
                         cudaMemcpy(c.dCols, plan.cols, cb, cudaMemcpyHostToDevice) == cudaSuccess;
                    if (!ok) (void)cudaFree(c.dRows);
                }
                if (!ok) (void)cudaGetLastError();
            }
            if (ok) {
                gPlanCache[freeSlot] = c;
                *out = c;
                found = 1;
            }
        }
        planFree(&plan);
    }
    pthread_mutex_unlock(&gPlanLock);
    return found;
}

/* Up to four planes of one image: their position tables go up in ONE copy, the samples are produced by ONE launch. */
static cudaError_t scalePlanesOnDevice(const PlaneIO * io, int count, int bps, int * oom)
{
    ScalePlan plans[AB_SCALE_MAX_PLANES];
    PlanCacheEntry cached[AB_SCALE_MAX_PLANES];
    int isCached[AB_SCALE_MAX_PLANES] = { 0, 0, 0, 0 };
    memset(plans, 0, sizeof(plans));
    size_t tableInts = 0;
    cudaError_t e = cudaSuccess;
    int32_t *hTab = NULL, *dTab = NULL;
    for (int k = 0; k < count; ++k) {
        isCached[k] = planCacheGet(io[k].sw, io[k].sh, io[k].dw, io[k].dh, bps == 2, &cached[k]);
        if (isCached[k]) {
            plans[k].mode = cached[k].mode;
            plans[k].linearOnly = cached[k].linearOnly;
            plans[k].minBoxWidth = cached[k].minBoxWidth;
            plans[k].colsArith = cached[k].colsArith;
            plans[k].x0 = cached[k].x0;
            plans[k].dx = cached[k].dx;
            continue;
        }
        if (!planScale(io[k].sw, io[k].sh, io[k].dw, io[k].dh, bps == 2, &plans[k])) {
            *oom = 1;
            goto done;
        }
        if (plans[k].mode) tableInts += 4 * ((size_t)io[k].dw + (size_t)io[k].dh);
    }
    AbScaleBatch batch;
    memset(&batch, 0, sizeof(batch));
    if (tableInts) {
        hTab = (int32_t *)malloc(tableInts * sizeof(int32_t));
        if (!hTab) {
            *oom = 1;
            goto done;
        }
        e = (cudaError_t)abScratchAlloc((void **)&dTab, tableInts * sizeof(int32_t));
        if (e != cudaSuccess) goto done;
    }
    size_t at = 0;
    for (int k = 0; k < count; ++k) {
        if (plans[k].mode == 0) { /* same size: CopyPlane (scale.c:849-853) */
            e = cudaMemcpy2DAsync(io[k].dst, (size_t)io[k].dstRB, io[k].src, (size_t)io[k].srcRB, (size_t)io[k].dw * bps, (size_t)io[k].dh,
                                  cudaMemcpyDeviceToDevice, AB_STREAM);
            if (e != cudaSuccess) goto done;
            continue;
        }
        AbScaleJob * job = &batch.job[batch.count++];
        job->src = io[k].src;
        job->dst = io[k].dst;
        job->srcRB = io[k].srcRB;
        job->dstRB = io[k].dstRB;
        job->srcW = io[k].sw;
        job->srcH = io[k].sh;
        job->dstW = io[k].dw;
        job->dstH = io[k].dh;
        job->bps = (uint8_t)bps;
        job->mode = (uint8_t)plans[k].mode;
        job->linearOnly = (uint8_t)plans[k].linearOnly;
        job->minBoxWidth = plans[k].minBoxWidth;
        job->colsArith = (uint8_t)plans[k].colsArith;
        job->x0 = plans[k].x0;
        job->dx = plans[k].dx;
        if (isCached[k]) {
            job->rows = cached[k].dRows;
            job->cols = cached[k].dCols;
            continue;
        }
        memcpy(hTab + at, plans[k].rows, (size_t)io[k].dh * 4 * sizeof(int32_t));
        job->rows = dTab + at;
        at += (size_t)io[k].dh * 4;
        memcpy(hTab + at, plans[k].cols, (size_t)io[k].dw * 4 * sizeof(int32_t));
        job->cols = dTab + at;
        at += (size_t)io[k].dw * 4;
    }
    if (batch.count) {
        if (tableInts) e = cudaMemcpyAsync(dTab, hTab, tableInts * sizeof(int32_t), cudaMemcpyHostToDevice, AB_STREAM);
        if (e == cudaSuccess) e = (cudaError_t)abLaunchScaleBatch(&batch, AB_STREAM);
    }
    /* hTab is pageable: the copy has been staged when cudaMemcpyAsync returns, but one synchronisation per image keeps the
     * lifetime rules simple (and callers of the synchronous API expect finished planes) */
    if (e == cudaSuccess) e = cudaStreamSynchronize(AB_STREAM);

done:
    if (dTab) (void)cudaFreeAsync(dTab, AB_STREAM);
    free(hTab);
    for (int k = 0; k < count; ++k) planFree(&plans[k]);
    return e;
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

static void diagf(avifDiagnostics * diag, const char * msg, unsigned a, unsigned b)
{
    if (diag && !diag->error[0]) snprintf(diag->error, sizeof(diag->error), msg, a, b);
}

/* Extension used by the tone mapper: scale the planes of `src` (host or device pointers) into DEVICE planes described by
 * `dst` (planes allocated by the caller with cudaMalloc-family calls, dimensions and row strides set). */
static avifResult scaleToDeviceBody(const avifImage * src, avifImage * dst);
avifResult avifB200ImageScaleToDevice(const avifImage * src, avifImage * dst)
{
    if (!src || !dst || !dst->width || !dst->height) return AVIF_RESULT_INVALID_ARGUMENT;
    const void * ptrs[8] = { dst->yuvPlanes[0], dst->yuvPlanes[1], dst->yuvPlanes[2], dst->alphaPlane,
                             src->yuvPlanes[0], src->yuvPlanes[1], src->yuvPlanes[2], src->alphaPlane };
    int prev = -1;
    avifResult res = (avifResult)abDeviceEnter(ptrs, 8, &prev);
    if (res != AVIF_RESULT_OK) return res;
    res = scaleToDeviceBody(src, dst);
    abDeviceLeave(prev);
    return res;
}
static avifResult scaleToDeviceBody(const avifImage * src, avifImage * dst)
{
    if (src->width > 16384 || src->height > 16384) return AVIF_RESULT_NOT_IMPLEMENTED; /* src/scale.c:65-76 */
    const int bps = (src->depth > 8) ? 2 : 1;
    avifPixelFormatInfo info;
    avifGetPixelFormatInfo(src->yuvFormat, &info);
    PlaneIO io[AB_SCALE_MAX_PLANES];
    uint8_t * staged[AB_SCALE_MAX_PLANES] = { NULL, NULL, NULL, NULL };
    int count = 0, oom = 0;
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < 4 && e == cudaSuccess; ++c) {
        const uint8_t * sp = (c < 3) ? src->yuvPlanes[c] : src->alphaPlane;
        uint8_t * dp = (c < 3) ? dst->yuvPlanes[c] : dst->alphaPlane;
        if (!sp || !dp) continue;
        const uint32_t srb = (c < 3) ? src->yuvRowBytes[c] : src->alphaRowBytes, drb = (c < 3) ? dst->yuvRowBytes[c] : dst->alphaRowBytes;
        const int chroma = (c == 1 || c == 2);
        if (chroma && info.monochrome) continue;
        PlaneIO * p = &io[count];
        p->sw = chroma ? (int)((src->width + info.chromaShiftX) >> info.chromaShiftX) : (int)src->width;
        p->sh = chroma ? (int)((src->height + info.chromaShiftY) >> info.chromaShiftY) : (int)src->height;
        p->dw = chroma ? (int)((dst->width + info.chromaShiftX) >> info.chromaShiftX) : (int)dst->width;
        p->dh = chroma ? (int)((dst->height + info.chromaShiftY) >> info.chromaShiftY) : (int)dst->height;
        p->src = sp;
        p->srcRB = srb;
        p->dst = dp;
        p->dstRB = drb;
        if (!onDevice(sp)) {
            const size_t pitch = ((size_t)p->sw * bps + 255) & ~(size_t)255;
            e = (cudaError_t)abScratchAlloc((void **)&staged[count], pitch * (size_t)p->sh);
            if (e == cudaSuccess) e = cudaMemcpy2DAsync(staged[count], pitch, sp, srb, (size_t)p->sw * bps, (size_t)p->sh, cudaMemcpyHostToDevice, AB_STREAM);
            p->src = staged[count];
            p->srcRB = (int64_t)pitch;
        }
        ++count;
    }
    if (e == cudaSuccess && count) e = scalePlanesOnDevice(io, count, bps, &oom);
    for (int k = 0; k < AB_SCALE_MAX_PLANES; ++k) {
        if (staged[k]) (void)cudaFreeAsync(staged[k], AB_STREAM);
    }
    if (oom) return AVIF_RESULT_OUT_OF_MEMORY;
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return (e == cudaErrorMemoryAllocation) ? AVIF_RESULT_OUT_OF_MEMORY : AVIF_RESULT_UNKNOWN_ERROR;
    }
    return AVIF_RESULT_OK;
}

/* src/scale.c:24-201 with the default limits of :203-207 */
avifResult avifImageScale(avifImage * image, uint32_t dstWidth, uint32_t dstHeight, avifDiagnostics * diag)
{
    if (diag) diag->error[0] = '\0';
    if (image->width == dstWidth && image->height == dstHeight) return AVIF_RESULT_OK;
    if (!dstWidth || !dstHeight) {
        diagf(diag, "avifImageScale requested invalid dst dimensions [%ux%u]", dstWidth, dstHeight);
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    if (dstWidth > (16384u * 16384u) / dstHeight || dstWidth > 32768u || dstHeight > 32768u) { /* avifDimensionsTooLarge, src/avif.c:646-655 */
        diagf(diag, "avifImageScale requested dst dimensions that are too large [%ux%u]", dstWidth, dstHeight);
        return AVIF_RESULT_NOT_IMPLEMENTED;
    }
    const void * devPtrs[4] = { image->yuvPlanes[0], image->yuvPlanes[1], image->yuvPlanes[2], image->alphaPlane };
    int prevDevice = -1;
    {
        const avifResult er = (avifResult)abDeviceEnter(devPtrs, 4, &prevDevice);
        if (er != AVIF_RESULT_OK) return er;
    }
    /* the source planes step aside, the image takes the new size (src/scale.c:40-63) */
    avifImage old = *image;
    for (int c = 0; c < 3; ++c) {
        image->yuvPlanes[c] = NULL;
        image->yuvRowBytes[c] = 0;
    }
    image->imageOwnsYUVPlanes = AVIF_FALSE;
    image->alphaPlane = NULL;
    image->alphaRowBytes = 0;
    image->imageOwnsAlphaPlane = AVIF_FALSE;
    image->width = dstWidth;
    image->height = dstHeight;

    avifResult res = AVIF_RESULT_OK;
    uint8_t * dPlanes[4] = { NULL, NULL, NULL, NULL };
    if ((old.yuvPlanes[0] || old.alphaPlane) && (old.width > 16384 || old.height > 16384)) {
        diagf(diag, "avifImageScale requested invalid scale for libyuv [%u -> %u]", old.width > 16384 ? old.width : old.height,
              old.width > 16384 ? dstWidth : dstHeight);
        res = AVIF_RESULT_NOT_IMPLEMENTED;
        goto cleanup;
    }
    {
        const uint32_t cs = (image->depth > 8) ? 2 : 1;
        avifPixelFormatInfo info;
        avifGetPixelFormatInfo(image->yuvFormat, &info);
        /* new host planes, like avifImageAllocatePlanes (src/avif.c:431-489): tightly packed rows, owned by the image */
        avifImage dev = *image; /* the same geometry with device planes: where the kernels write */
        for (int c = 0; c < 4; ++c) {
            const int chroma = (c == 1 || c == 2);
            const uint8_t * sp = (c < 3) ? old.yuvPlanes[c] : old.alphaPlane;
            const int wanted = (c < 3) ? (old.yuvPlanes[0] != NULL && image->yuvFormat != AVIF_PIXEL_FORMAT_NONE && !(chroma && info.monochrome))
                                       : (old.alphaPlane != NULL);
            if (!wanted) continue;
            const uint32_t w = chroma ? (uint32_t)(((uint64_t)dstWidth + info.chromaShiftX) >> info.chromaShiftX) : dstWidth;
            const uint32_t h = chroma ? (uint32_t)(((uint64_t)dstHeight + info.chromaShiftY) >> info.chromaShiftY) : dstHeight;
            uint8_t * host = (uint8_t *)malloc((size_t)cs * w * h);
            if (!host) {
                res = AVIF_RESULT_OUT_OF_MEMORY;
                goto cleanup;
            }
            if (c < 3) {
                image->yuvPlanes[c] = host;
                image->yuvRowBytes[c] = cs * w;
                image->imageOwnsYUVPlanes = AVIF_TRUE;
            } else {
                image->alphaPlane = host;
                image->alphaRowBytes = cs * w;
                image->imageOwnsAlphaPlane = AVIF_TRUE;
            }
            if (!sp) continue; /* allocated but not scaled, like src/scale.c:88-91 */
            const size_t pitch = ((size_t)cs * w + 255) & ~(size_t)255;
            if ((cudaError_t)abScratchAlloc((void **)&dPlanes[c], pitch * h) != cudaSuccess) {
                (void)cudaGetLastError();
                res = AVIF_RESULT_OUT_OF_MEMORY;
                goto cleanup;
            }
            if (c < 3) {
                dev.yuvPlanes[c] = dPlanes[c];
                dev.yuvRowBytes[c] = (uint32_t)pitch;
            } else {
                dev.alphaPlane = dPlanes[c];
                dev.alphaRowBytes = (uint32_t)pitch;
            }
        }
        for (int c = 0; c < 3; ++c) {
            if (!dPlanes[c]) dev.yuvPlanes[c] = NULL;
        }
        if (!dPlanes[3]) dev.alphaPlane = NULL;
        res = avifB200ImageScaleToDevice(&old, &dev);
        if (res != AVIF_RESULT_OK) goto cleanup;
        for (int c = 0; c < 4; ++c) {
            if (!dPlanes[c]) continue;
            const int chroma = (c == 1 || c == 2);
            const uint32_t w = chroma ? (uint32_t)(((uint64_t)dstWidth + info.chromaShiftX) >> info.chromaShiftX) : dstWidth;
            const uint32_t h = chroma ? (uint32_t)(((uint64_t)dstHeight + info.chromaShiftY) >> info.chromaShiftY) : dstHeight;
            uint8_t * host = (c < 3) ? image->yuvPlanes[c] : image->alphaPlane;
            const uint32_t hostRB = (c < 3) ? image->yuvRowBytes[c] : image->alphaRowBytes;
            const uint32_t devRB = (c < 3) ? dev.yuvRowBytes[c] : dev.alphaRowBytes;
            if (cudaMemcpy2DAsync(host, hostRB, dPlanes[c], devRB, (size_t)cs * w, h, cudaMemcpyDeviceToHost, AB_STREAM) != cudaSuccess ||
                cudaStreamSynchronize(AB_STREAM) != cudaSuccess) {
                (void)cudaGetLastError();
                res = AVIF_RESULT_UNKNOWN_ERROR;
                goto cleanup;
            }
        }
    }

cleanup:
    for (int c = 0; c < 4; ++c) {
        if (dPlanes[c]) (void)cudaFreeAsync(dPlanes[c], AB_STREAM);
    }
    /* the old planes are released if the image owned them (src/scale.c:189-198); on a device box that may be cudaFree'd memory
     * only if the caller allocated it that way AND handed ownership over, which the reference's avifFree == free forbids too */
    if (old.yuvPlanes[0] && old.imageOwnsYUVPlanes) {
        for (int c = 0; c < 3; ++c) free(old.yuvPlanes[c]);
    }
    if (old.alphaPlane && old.imageOwnsAlphaPlane) free(old.alphaPlane);
    abDeviceLeave(prevDevice);
    return res;
}
