/*
 * grid.c -- the tile-grid paths across GPUs (BASELINE.json config 4).  Host C, like the reference.
 *
 * Encode side (reference flow: apps/avifenc.c converts the whole canvas with avifImageRGBToYUV, avifImageSplitGrid --
 * apps/shared/avifutil.c:546-605 -- cuts it into cell views, avifEncoderAddImageGrid -- src/write.c:2147 -- hands the cells to the
 * CPU AV1 encoder): avifB200GridRGBToYUV converts the canvas cell by cell.  Cell sizes follow the reference's split rule, which
 * makes them even whenever chroma is subsampled, so no 2x2 averaging block straddles a cell and the result is bit-identical to
 * converting the whole canvas.  One process per GPU: the rank that holds the canvas (root) sends every other rank its cells'
 * pixels (ONE grouped NCCL point-to-point exchange, NVLink), every rank converts its cells in ONE batched launch of the fused
 * kernel, the planes travel back in a second grouped exchange and land in the canvas planes.  Root's own cells are views of the
 * canvas: no copy, no allocation.  With one rank (comm == NULL) the whole call is one batched launch over 16 views.
 *
 * Decode side (reference flow: tiles are copied into the canvas, src/read.c:1823-1878, then the canvas is converted,
 * src/read.c:6777-6866 + src/reformat.c:764-838): bilinear chroma upsampling is NOT tile-local -- the filter reads across cell
 * seams and replicates only at the canvas borders -- so avifB200GridYUVToRGB shards the canvas by even ROW BANDS with a one-row
 * chroma halo on each inner seam and converts each band with avifB200ImageYUVToRGBRowsAsync against whole-canvas geometry.
 *
 * NCCL is reached through dlopen("libnccl.so.2") so that the library keeps working (single GPU) where NCCL is absent; nothing of
 * NCCL's ABI is needed beyond ncclUniqueId being 128 bytes, ncclComm_t a pointer and ncclUint8 == 1.
 */
#define AVIF_B200_BUILDING
#include "avif_b200.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "plan.h"

typedef struct AbNcclId { char internal[128]; } AbNcclId;
enum { kNcclUint8 = 1 };
static struct {
    int tried, ok;
    int (*GetUniqueId)(AbNcclId *);
    int (*CommInitRank)(void **, int, AbNcclId, int);
    int (*CommDestroy)(void *);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t);
} nccl;
static pthread_mutex_t ncclLock = PTHREAD_MUTEX_INITIALIZER;

static int ncclLoad(void)
{
    pthread_mutex_lock(&ncclLock);
    if (!nccl.tried) {
        nccl.tried = 1;
        void * h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            *(void **)&nccl.GetUniqueId = dlsym(h, "ncclGetUniqueId");
            *(void **)&nccl.CommInitRank = dlsym(h, "ncclCommInitRank");
            *(void **)&nccl.CommDestroy = dlsym(h, "ncclCommDestroy");
            *(void **)&nccl.GroupStart = dlsym(h, "ncclGroupStart");
            *(void **)&nccl.GroupEnd = dlsym(h, "ncclGroupEnd");
            *(void **)&nccl.Send = dlsym(h, "ncclSend");
            *(void **)&nccl.Recv = dlsym(h, "ncclRecv");
            nccl.ok = nccl.GetUniqueId && nccl.CommInitRank && nccl.CommDestroy && nccl.GroupStart && nccl.GroupEnd && nccl.Send && nccl.Recv;
        }
    }
    pthread_mutex_unlock(&ncclLock);
    return nccl.ok;
}

avifResult avifB200GridGetUniqueId(uint8_t id[128])
{
    if (!id) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!ncclLoad()) return AVIF_RESULT_NOT_IMPLEMENTED;
    AbNcclId v;
    if (nccl.GetUniqueId(&v) != 0) return AVIF_RESULT_UNKNOWN_ERROR;
    memcpy(id, &v, sizeof(v));
    return AVIF_RESULT_OK;
}

/* collective over the ranks; the calling thread's current CUDA device becomes the communicator's device */
avifResult avifB200GridCommInit(void ** comm, int nranks, int rank, const uint8_t id[128])
{
    if (!comm || !id || nranks < 1 || rank < 0 || rank >= nranks) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!ncclLoad()) return AVIF_RESULT_NOT_IMPLEMENTED;
    AbNcclId v;
    memcpy(&v, id, sizeof(v));
    *comm = NULL;
    return (nccl.CommInitRank(comm, nranks, v, rank) == 0) ? AVIF_RESULT_OK : AVIF_RESULT_UNKNOWN_ERROR;
}

avifResult avifB200GridCommDestroy(void * comm)
{
    if (!comm) return AVIF_RESULT_OK;
    if (!ncclLoad()) return AVIF_RESULT_NOT_IMPLEMENTED;
    return (nccl.CommDestroy(comm) == 0) ? AVIF_RESULT_OK : AVIF_RESULT_UNKNOWN_ERROR;
}

/* ---- the reference's split rule (apps/shared/avifutil.c:493-546) ---- */
static int bestCellSize(uint32_t numPixels, uint32_t numCells, int subsampled, uint32_t * cell)
{
    if (!numPixels || !numCells) return 0;
    uint32_t c = (uint32_t)(((uint64_t)numPixels + numCells - 1) / numCells);
    if (c < 64) {
        c = 64;
        if ((uint64_t)(numCells - 1) * c >= (uint64_t)numPixels) return 0; /* cells entirely off-canvas */
    }
    if (c > 65536) return 0;
    if (subsampled && (c & 1)) {
        ++c;
        if ((uint64_t)(numCells - 1) * c >= (uint64_t)numPixels) return 0;
    }
    *cell = c;
    return 1;
}

avifResult avifB200GridCellRect(const avifImage * canvas, uint32_t gridCols, uint32_t gridRows, uint32_t cellIndex, avifCropRect * rect)
{
    if (!canvas || !rect || !gridCols || !gridRows || cellIndex >= gridCols * gridRows) return AVIF_RESULT_INVALID_ARGUMENT;
    avifPixelFormatInfo fi;
    avifGetPixelFormatInfo(canvas->yuvFormat, &fi);
    uint32_t cw, ch;
    if (!bestCellSize(canvas->width, gridCols, !fi.monochrome && fi.chromaShiftX, &cw) ||
        !bestCellSize(canvas->height, gridRows, !fi.monochrome && fi.chromaShiftY, &ch)) {
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    const uint32_t gx = cellIndex % gridCols, gy = cellIndex / gridCols;
    rect->x = gx * cw;
    rect->y = gy * ch;
    rect->width = (rect->x + cw > canvas->width) ? canvas->width - rect->x : cw;
    rect->height = (rect->y + ch > canvas->height) ? canvas->height - rect->y : ch;
    return AVIF_RESULT_OK;
}

/* ---- helpers ---- */
#define GR_CUDA(call)                               \
    do {                                            \
        if ((call) != cudaSuccess) {                \
            (void)cudaGetLastError();               \
            res = AVIF_RESULT_UNKNOWN_ERROR;        \
            goto cleanup;                           \
        }                                           \
    } while (0)
#define GR_NCCL(call)                               \
    do {                                            \
        if ((call) != 0) {                          \
            res = AVIF_RESULT_UNKNOWN_ERROR;        \
            goto cleanup;                           \
        }                                           \
    } while (0)

typedef struct CellGeom {
    avifCropRect r;
    uint32_t cw, ch;                    /* chroma plane extent of the cell */
    size_t rgbBytes, yBytes, cBytes, aBytes, yuvBytes; /* packed (contiguous) sizes */
    int owner;
} CellGeom;

static void viewOfCanvas(const avifImage * canvas, const CellGeom * g, uint32_t bps, int sx, int sy, avifImage * v)
{
    *v = *canvas;
    v->width = g->r.width;
    v->height = g->r.height;
    v->imageOwnsYUVPlanes = AVIF_FALSE;
    v->imageOwnsAlphaPlane = AVIF_FALSE;
    if (canvas->yuvPlanes[0]) v->yuvPlanes[0] = canvas->yuvPlanes[0] + (size_t)g->r.y * canvas->yuvRowBytes[0] + (size_t)g->r.x * bps;
    for (int p = 1; p < 3; ++p)
        if (canvas->yuvPlanes[p]) v->yuvPlanes[p] = canvas->yuvPlanes[p] + (size_t)(g->r.y >> sy) * canvas->yuvRowBytes[p] + (size_t)(g->r.x >> sx) * bps;
    if (canvas->alphaPlane) v->alphaPlane = canvas->alphaPlane + (size_t)g->r.y * canvas->alphaRowBytes + (size_t)g->r.x * bps;
}

/* a cell image over one packed buffer [Y | U | V | A] */
static void imageOverBuffer(const avifImage * canvas, const CellGeom * g, uint32_t bps, int mono, int alpha, uint8_t * buf, avifImage * v)
{
    *v = *canvas;
    v->width = g->r.width;
    v->height = g->r.height;
    v->imageOwnsYUVPlanes = AVIF_FALSE;
    v->imageOwnsAlphaPlane = AVIF_FALSE;
    v->yuvPlanes[0] = buf;
    v->yuvRowBytes[0] = g->r.width * bps;
    v->yuvPlanes[1] = v->yuvPlanes[2] = NULL;
    v->yuvRowBytes[1] = v->yuvRowBytes[2] = 0;
    if (!mono) {
        v->yuvPlanes[1] = buf + g->yBytes;
        v->yuvPlanes[2] = buf + g->yBytes + g->cBytes;
        v->yuvRowBytes[1] = v->yuvRowBytes[2] = g->cw * bps;
    }
    v->alphaPlane = alpha ? buf + g->yBytes + 2 * g->cBytes : NULL;
    v->alphaRowBytes = alpha ? g->r.width * bps : 0;
}

/* ---- peer-memory path of the encode grid ----------------------------------------------------------------------------------
 * One process per GPU, NVSwitch between them: instead of packing cells, shipping them with NCCL, converting and shipping the planes
 * back, every rank maps the ROOT's canvas (CUDA IPC: the pixel buffer and the four planes) and runs the fused conversion kernel on
 * views of it -- its loads fetch the cell's pixels over NVLink and its stores put the planes where they belong, tile by tile, so the
 * transfer overlaps the arithmetic and nothing is staged, packed or assembled.  NCCL carries two ordering points only: the
 * descriptor of the canvas (IPC handles + offsets; stream-ordered behind whatever produced the canvas on the root), and a one-byte
 * token per rank that tells the root its planes are complete.  Memory that cannot be exported (not from cudaMalloc), a driver
 * without IPC or AVIF_B200_GRID_PEER=0 fall back to the staged exchange below, decided by the root and followed by every rank. */
enum { kPeerBufs = 5 }; /* pixels, Y, U, V, A */
typedef struct PeerCanvas {
    uint8_t handle[kPeerBufs][64]; /* cudaIpcMemHandle_t of the allocation each buffer lives in */
    uint64_t offset[kPeerBufs];    /* of the buffer inside its allocation */
    uint8_t present[kPeerBufs];
    uint8_t ok;                    /* root: every present buffer could be exported */
    uint8_t pad[2];
} PeerCanvas;

typedef int (*AbMemGetAddressRange)(unsigned long long * base, size_t * size, unsigned long long ptr);
static AbMemGetAddressRange peerRangeFn(void)
{
    static AbMemGetAddressRange fn = NULL;
    static int tried = 0;
    if (!tried) {
        tried = 1;
        void * f = NULL;
        enum cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (AbMemGetAddressRange)f;
        else (void)cudaGetLastError();
    }
    return fn;
}

static int peerExport(const void * ptr, uint8_t handle[64], uint64_t * offset)
{
    AbMemGetAddressRange range = peerRangeFn();
    unsigned long long base = 0;
    size_t size = 0;
    if (!range || range(&base, &size, (unsigned long long)(uintptr_t)ptr) != 0 || !base) return 0;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, (void *)(uintptr_t)base) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    memcpy(handle, &h, 64);
    *offset = (uint64_t)((uintptr_t)ptr - (uintptr_t)base);
    return 1;
}

/* allocations of other processes this process has mapped: kept open for the life of the process (a canvas is reused frame after
 * frame; opening costs a few hundred microseconds) */
static struct { uint8_t handle[64]; void * base; int device; } gPeerOpen[64];
static int gPeerOpenCount = 0;
static pthread_mutex_t gPeerLock = PTHREAD_MUTEX_INITIALIZER;

static void * peerOpen(const uint8_t handle[64])
{
    int device = 0;
    (void)cudaGetDevice(&device);
    void * base = NULL;
    pthread_mutex_lock(&gPeerLock);
    for (int k = 0; k < gPeerOpenCount; ++k)
        if (gPeerOpen[k].base && gPeerOpen[k].device == device && memcmp(gPeerOpen[k].handle, handle, 64) == 0) base = gPeerOpen[k].base;
    if (!base) {
        int slot = gPeerOpenCount;
        if (slot >= 64) { /* table full: unmap the oldest entry of this device and take its place (its allocation is long gone or idle) */
            static int next = 0;
            slot = -1;
            for (int tries = 0; tries < 64 && slot < 0; ++tries, next = (next + 1) % 64)
                if (gPeerOpen[next].device == device) slot = next;
            if (slot >= 0) {
                next = (slot + 1) % 64;
                (void)cudaIpcCloseMemHandle(gPeerOpen[slot].base);
                (void)cudaGetLastError();
                memset(gPeerOpen[slot].handle, 0, 64); /* nothing matches the slot until it is filled again */
                gPeerOpen[slot].base = NULL;
            }
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, 64);
        if (slot >= 0 && cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
            memcpy(gPeerOpen[slot].handle, handle, 64);
            gPeerOpen[slot].base = base;
            gPeerOpen[slot].device = device;
            if (slot == gPeerOpenCount) ++gPeerOpenCount;
        } else {
            (void)cudaGetLastError();
            base = NULL;
        }
    }
    pthread_mutex_unlock(&gPeerLock);
    return base;
}

static __thread int gGridLastMode = 0; /* 0: one rank / staged NCCL exchange, 1: peer memory */
int avifB200GridLastMode(void) { return gGridLastMode; }

static int peerWanted(void)
{
    const char * e = getenv("AVIF_B200_GRID_PEER");
    return !(e && e[0] == '0');
}
static int peerDecodeWanted(void)
{
    const char * e = getenv("AVIF_B200_GRID_PEER_DECODE");
    return e && e[0] == '1';
}

avifResult avifB200GridRGBToYUV(avifImage * canvas, const avifRGBImage * canvasRgb, uint32_t gridCols, uint32_t gridRows, void * comm, int rank, int nranks,
                                int root, void * streamPtr, float phaseMs[4])
{
    if (!canvas || !canvasRgb || !gridCols || !gridRows || gridCols * gridRows > 1024) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!comm) {
        nranks = 1;
        rank = root = 0;
    }
    if (nranks < 1 || rank < 0 || rank >= nranks || root < 0 || root >= nranks) return AVIF_RESULT_INVALID_ARGUMENT;
    if (nranks > 1 && !ncclLoad()) return AVIF_RESULT_NOT_IMPLEMENTED;
    if (canvasRgb->width != canvas->width || canvasRgb->height != canvas->height) return AVIF_RESULT_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)streamPtr;
    const uint32_t n = gridCols * gridRows;
    avifPixelFormatInfo fi;
    avifGetPixelFormatInfo(canvas->yuvFormat, &fi);
    const int mono = fi.monochrome, sx = mono ? 0 : fi.chromaShiftX, sy = mono ? 0 : fi.chromaShiftY;
    const uint32_t bps = (canvas->depth > 8) ? 2 : 1, pix = avifRGBImagePixelSize(canvasRgb);
    const int alpha = avifRGBFormatHasAlpha(canvasRgb->format) && !canvasRgb->ignoreAlpha; /* src/reformat.c:236-237: AVIF_PLANES_ALL */
    const int isRoot = (rank == root);
    if (isRoot && (!canvasRgb->pixels || !canvas->yuvPlanes[0] || (!mono && (!canvas->yuvPlanes[1] || !canvas->yuvPlanes[2])) || (alpha && !canvas->alphaPlane))) {
        return AVIF_RESULT_INVALID_ARGUMENT; /* resident canvases are allocated by the caller */
    }

    avifResult res = AVIF_RESULT_OK;
    CellGeom * cells = (CellGeom *)calloc(n, sizeof(CellGeom));
    avifImage * imgs = (avifImage *)calloc(n, sizeof(avifImage));
    avifRGBImage * rgbs = (avifRGBImage *)calloc(n, sizeof(avifRGBImage));
    avifImage ** pImgs = (avifImage **)calloc(n, sizeof(void *));
    const avifRGBImage ** pRgbs = (const avifRGBImage **)calloc(n, sizeof(void *));
    uint8_t ** stageRgb = (uint8_t **)calloc(n, sizeof(void *)); /* root: packed pixels of a remote cell; owner: its received pixels */
    uint8_t ** stageYuv = (uint8_t **)calloc(n, sizeof(void *)); /* owner: converted planes; root: received planes of a remote cell */
    cudaEvent_t ev[5] = { NULL, NULL, NULL, NULL, NULL };
    uint32_t mine = 0;
    uint8_t *dDesc = NULL, *dTok = NULL; /* peer path: canvas descriptor, completion tokens */
    if (!cells || !imgs || !rgbs || !pImgs || !pRgbs || !stageRgb || !stageYuv) {
        res = AVIF_RESULT_OUT_OF_MEMORY;
        goto cleanup;
    }
    for (uint32_t u = 0; u < n; ++u) {
        CellGeom * g = &cells[u];
        res = avifB200GridCellRect(canvas, gridCols, gridRows, u, &g->r);
        if (res != AVIF_RESULT_OK) goto cleanup;
        g->cw = mono ? 0 : (uint32_t)(((uint64_t)g->r.width + sx) >> sx);
        g->ch = mono ? 0 : (uint32_t)(((uint64_t)g->r.height + sy) >> sy);
        g->rgbBytes = (size_t)g->r.width * pix * g->r.height;
        g->yBytes = (size_t)g->r.width * bps * g->r.height;
        g->cBytes = (size_t)g->cw * bps * g->ch;
        g->aBytes = alpha ? g->yBytes : 0;
        g->yuvBytes = g->yBytes + 2 * g->cBytes + g->aBytes;
        g->owner = (int)((root + u) % (uint32_t)nranks); /* cell 0 stays on the root, the rest round-robin */
    }
    if (phaseMs) {
        for (int k = 0; k < 5; ++k) GR_CUDA(cudaEventCreate(&ev[k]));
        GR_CUDA(cudaEventRecord(ev[0], st));
    }

    gGridLastMode = 0;
    if (nranks > 1) {
        /* ---- peer-memory path: descriptor -> kernels on views of the root's canvas -> tokens ---- */
        PeerCanvas pc;
        memset(&pc, 0, sizeof(pc));
        const void * bufs[kPeerBufs] = { NULL, NULL, NULL, NULL, NULL };
        if (isRoot) {
            bufs[0] = canvasRgb->pixels;
            bufs[1] = canvas->yuvPlanes[0];
            bufs[2] = mono ? NULL : canvas->yuvPlanes[1];
            bufs[3] = mono ? NULL : canvas->yuvPlanes[2];
            bufs[4] = alpha ? canvas->alphaPlane : NULL;
            pc.ok = (uint8_t)peerWanted();
            for (int b = 0; b < kPeerBufs && pc.ok; ++b) {
                pc.present[b] = (bufs[b] != NULL);
                if (bufs[b] && !peerExport(bufs[b], pc.handle[b], &pc.offset[b])) pc.ok = 0;
            }
        }
        GR_CUDA((cudaError_t)abScratchAllocOn((void **)&dDesc, sizeof(pc) + 64, st));
        if (isRoot) GR_CUDA(cudaMemcpyAsync(dDesc, &pc, sizeof(pc), cudaMemcpyHostToDevice, st));
        GR_NCCL(nccl.GroupStart());
        for (int r = 0; r < nranks; ++r) {
            if (r == root) continue;
            if (isRoot) GR_NCCL(nccl.Send(dDesc, sizeof(pc), kNcclUint8, r, comm, st));
            else if (r == rank) GR_NCCL(nccl.Recv(dDesc, sizeof(pc), kNcclUint8, root, comm, st));
        }
        GR_NCCL(nccl.GroupEnd());
        if (!isRoot) { /* the descriptor is behind everything the root had queued before the call: the canvas is ready when it arrives */
            GR_CUDA(cudaMemcpyAsync(&pc, dDesc, sizeof(pc), cudaMemcpyDeviceToHost, st));
            GR_CUDA(cudaStreamSynchronize(st));
        }
        if (pc.ok) {
            uint8_t token = 1;
            avifImage peerCanvas = *canvas;
            avifRGBImage peerRgb = *canvasRgb;
            if (!isRoot) {
                uint8_t * mapped[kPeerBufs] = { NULL, NULL, NULL, NULL, NULL };
                for (int b = 0; b < kPeerBufs; ++b) {
                    if (!pc.present[b]) continue;
                    uint8_t * base = (uint8_t *)peerOpen(pc.handle[b]);
                    if (!base) token = 0; /* reported to the root below; this rank converts nothing */
                    else mapped[b] = base + pc.offset[b];
                }
                peerRgb.pixels = mapped[0];
                peerCanvas.yuvPlanes[0] = mapped[1];
                peerCanvas.yuvPlanes[1] = mapped[2];
                peerCanvas.yuvPlanes[2] = mapped[3];
                peerCanvas.alphaPlane = mapped[4];
                /* strides and extents of the views are the root's: every rank was handed the same canvas description */
            }
            if (phaseMs) GR_CUDA(cudaEventRecord(ev[1], st));
            if (token) {
                for (uint32_t u = 0; u < n; ++u) {
                    const CellGeom * g = &cells[u];
                    if (g->owner != rank) continue;
                    avifRGBImage * r = &rgbs[mine];
                    *r = peerRgb;
                    r->width = g->r.width;
                    r->height = g->r.height;
                    r->pixels = peerRgb.pixels + (size_t)g->r.y * peerRgb.rowBytes + (size_t)g->r.x * pix;
                    viewOfCanvas(&peerCanvas, g, bps, sx, sy, &imgs[mine]);
                    pImgs[mine] = &imgs[mine];
                    pRgbs[mine] = r;
                    ++mine;
                }
                if (mine) {
                    res = avifB200ImageRGBToYUVBatchAsync(pImgs, pRgbs, mine, st);
                    if (res != AVIF_RESULT_OK) token = 0;
                }
            }
            if (phaseMs) GR_CUDA(cudaEventRecord(ev[2], st));
            /* tokens: stream-ordered behind each rank's kernel, so the root's stream continues only when every plane is complete */
            GR_CUDA((cudaError_t)abScratchAllocOn((void **)&dTok, (size_t)nranks + 64, st));
            if (!isRoot) GR_CUDA(cudaMemcpyAsync(dTok + rank, &token, 1, cudaMemcpyHostToDevice, st));
            GR_NCCL(nccl.GroupStart());
            for (int r = 0; r < nranks; ++r) {
                if (r == root) continue;
                if (isRoot) GR_NCCL(nccl.Recv(dTok + r, 1, kNcclUint8, r, comm, st));
                else if (r == rank) GR_NCCL(nccl.Send(dTok + rank, 1, kNcclUint8, root, comm, st));
            }
            GR_NCCL(nccl.GroupEnd());
            if (phaseMs) GR_CUDA(cudaEventRecord(ev[3], st));
            if (isRoot) {
                uint8_t toks[256];
                const int m = nranks < 256 ? nranks : 256;
                memset(toks, 1, sizeof(toks));
                GR_CUDA(cudaMemcpyAsync(toks, dTok, (size_t)m, cudaMemcpyDeviceToHost, st));
                GR_CUDA(cudaStreamSynchronize(st));
                for (int r = 0; r < m; ++r)
                    if (r != root && !toks[r]) res = AVIF_RESULT_UNKNOWN_ERROR; /* a rank could not map the canvas or failed to convert */
            } else if (!token && res == AVIF_RESULT_OK) {
                res = AVIF_RESULT_UNKNOWN_ERROR;
            }
            gGridLastMode = 1;
            if (phaseMs) {
                GR_CUDA(cudaEventRecord(ev[4], st));
                GR_CUDA(cudaEventSynchronize(ev[4]));
                for (int k = 0; k < 4; ++k) GR_CUDA(cudaEventElapsedTime(&phaseMs[k], ev[k], ev[k + 1])); /* descriptor, convert over NVLink, tokens, 0 */
            }
            goto cleanup;
        }
        /* the root could not (or was told not to) export its canvas: every rank takes the staged exchange */
    }

    /* ---- 1. scatter the pixels of remote cells (root packs each cell's rectangle into a contiguous staging buffer) ---- */
    for (uint32_t u = 0; u < n; ++u) {
        const CellGeom * g = &cells[u];
        if (g->owner == root) continue;
        if (isRoot || g->owner == rank) GR_CUDA((cudaError_t)abScratchAllocOn((void **)&stageRgb[u], g->rgbBytes, st));
        if (isRoot) {
            GR_CUDA(cudaMemcpy2DAsync(stageRgb[u], (size_t)g->r.width * pix, canvasRgb->pixels + (size_t)g->r.y * canvasRgb->rowBytes + (size_t)g->r.x * pix,
                                      canvasRgb->rowBytes, (size_t)g->r.width * pix, g->r.height, cudaMemcpyDeviceToDevice, st));
        }
    }
    if (nranks > 1) {
        GR_NCCL(nccl.GroupStart());
        for (uint32_t u = 0; u < n; ++u) {
            const CellGeom * g = &cells[u];
            if (g->owner == root) continue;
            if (isRoot) GR_NCCL(nccl.Send(stageRgb[u], g->rgbBytes, kNcclUint8, g->owner, comm, st));
            else if (g->owner == rank) GR_NCCL(nccl.Recv(stageRgb[u], g->rgbBytes, kNcclUint8, root, comm, st));
        }
        GR_NCCL(nccl.GroupEnd());
    }
    if (phaseMs) GR_CUDA(cudaEventRecord(ev[1], st));

    /* ---- 2. local conversion: ONE batched launch over this rank's cells (root: views of the canvas) ---- */
    for (uint32_t u = 0; u < n; ++u) {
        const CellGeom * g = &cells[u];
        if (g->owner != rank) continue;
        avifRGBImage * r = &rgbs[mine];
        *r = *canvasRgb;
        r->width = g->r.width;
        r->height = g->r.height;
        if (isRoot) {
            r->pixels = canvasRgb->pixels + (size_t)g->r.y * canvasRgb->rowBytes + (size_t)g->r.x * pix;
            viewOfCanvas(canvas, g, bps, sx, sy, &imgs[mine]);
        } else {
            r->pixels = stageRgb[u];
            r->rowBytes = g->r.width * pix;
            GR_CUDA((cudaError_t)abScratchAllocOn((void **)&stageYuv[u], g->yuvBytes, st));
            imageOverBuffer(canvas, g, bps, mono, alpha, stageYuv[u], &imgs[mine]);
        }
        pImgs[mine] = &imgs[mine];
        pRgbs[mine] = r;
        ++mine;
    }
    if (mine) {
        res = avifB200ImageRGBToYUVBatchAsync(pImgs, pRgbs, mine, st);
        if (res != AVIF_RESULT_OK) goto cleanup;
    }
    if (phaseMs) GR_CUDA(cudaEventRecord(ev[2], st));

    /* ---- 3. gather the planes of remote cells, 4. place them in the canvas planes (src/avif.c:187-225 per plane) ---- */
    if (nranks > 1) {
        if (isRoot) {
            for (uint32_t u = 0; u < n; ++u)
                if (cells[u].owner != root) GR_CUDA((cudaError_t)abScratchAllocOn((void **)&stageYuv[u], cells[u].yuvBytes, st));
        }
        GR_NCCL(nccl.GroupStart());
        for (uint32_t u = 0; u < n; ++u) {
            const CellGeom * g = &cells[u];
            if (g->owner == root) continue;
            if (isRoot) GR_NCCL(nccl.Recv(stageYuv[u], g->yuvBytes, kNcclUint8, g->owner, comm, st));
            else if (g->owner == rank) GR_NCCL(nccl.Send(stageYuv[u], g->yuvBytes, kNcclUint8, root, comm, st));
        }
        GR_NCCL(nccl.GroupEnd());
    }
    if (phaseMs) GR_CUDA(cudaEventRecord(ev[3], st));
    if (isRoot && nranks > 1) {
        for (uint32_t u = 0; u < n; ++u) {
            const CellGeom * g = &cells[u];
            if (g->owner == root) continue;
            avifImage packed, view;
            imageOverBuffer(canvas, g, bps, mono, alpha, stageYuv[u], &packed);
            viewOfCanvas(canvas, g, bps, sx, sy, &view);
            res = avifB200ImageCopySamplesAsync(&view, &packed, AVIF_PLANES_ALL, st);
            if (res != AVIF_RESULT_OK) goto cleanup;
        }
    }
    if (phaseMs) {
        GR_CUDA(cudaEventRecord(ev[4], st));
        GR_CUDA(cudaEventSynchronize(ev[4]));
        for (int k = 0; k < 4; ++k) GR_CUDA(cudaEventElapsedTime(&phaseMs[k], ev[k], ev[k + 1])); /* scatter, convert, gather, assemble */
    }

cleanup:
    if (dDesc) (void)cudaFreeAsync(dDesc, st);
    if (dTok) (void)cudaFreeAsync(dTok, st);
    if (stageRgb)
        for (uint32_t u = 0; u < n; ++u)
            if (stageRgb[u]) (void)cudaFreeAsync(stageRgb[u], st);
    if (stageYuv)
        for (uint32_t u = 0; u < n; ++u)
            if (stageYuv[u]) (void)cudaFreeAsync(stageYuv[u], st);
    for (int k = 0; k < 5; ++k)
        if (ev[k]) (void)cudaEventDestroy(ev[k]);
    free(cells);
    free(imgs);
    free(rgbs);
    free(pImgs);
    free((void *)pRgbs);
    free(stageRgb);
    free(stageYuv);
    return res;
}

/* ---- decode side: row bands with a chroma halo ---- */
avifResult avifB200GridYUVToRGB(const avifImage * canvas, avifRGBImage * canvasRgb, void * comm, int rank, int nranks, int root, void * streamPtr,
                                float phaseMs[3])
{
    if (!canvas || !canvasRgb) return AVIF_RESULT_INVALID_ARGUMENT;
    if (!comm) {
        nranks = 1;
        rank = root = 0;
    }
    if (nranks < 1 || rank < 0 || rank >= nranks || root < 0 || root >= nranks) return AVIF_RESULT_INVALID_ARGUMENT;
    if (nranks > 1 && !ncclLoad()) return AVIF_RESULT_NOT_IMPLEMENTED;
    if (canvasRgb->width != canvas->width || canvasRgb->height != canvas->height || !canvas->width || !canvas->height) return AVIF_RESULT_INVALID_ARGUMENT;
    cudaStream_t st = (cudaStream_t)streamPtr;
    avifPixelFormatInfo fi;
    avifGetPixelFormatInfo(canvas->yuvFormat, &fi);
    const int mono = fi.monochrome, sx = mono ? 0 : fi.chromaShiftX, sy = mono ? 0 : fi.chromaShiftY;
    const uint32_t bps = (canvas->depth > 8) ? 2 : 1, pix = avifRGBImagePixelSize(canvasRgb);
    const uint32_t W = canvas->width, H = canvas->height;
    const uint32_t cwid = mono ? 0 : (uint32_t)(((uint64_t)W + sx) >> sx), chei = mono ? 0 : (uint32_t)(((uint64_t)H + sy) >> sy);
    const int isRoot = (rank == root);
    /* which planes the conversion reads is the conversion's business (alpha may be ignored): ship what exists.  Every rank
     * passes the same descriptor; away from the root the row strides say which planes exist (non-zero), the pointers are unused. */
    const int haveA = (canvas->alphaRowBytes != 0);
    const int haveC = !mono && canvas->yuvRowBytes[1] != 0 && canvas->yuvRowBytes[2] != 0;
    if (isRoot && (!canvas->yuvPlanes[0] || !canvasRgb->pixels || (haveA && !canvas->alphaPlane) || (haveC && (!canvas->yuvPlanes[1] || !canvas->yuvPlanes[2])))) {
        return AVIF_RESULT_INVALID_ARGUMENT;
    }

    /* even bands: chroma rows and the 2-row bilinear units never straddle a seam */
    uint32_t per = (H + (uint32_t)nranks - 1) / (uint32_t)nranks;
    per = (per + 1) & ~1u;
    if (per < 2) per = 2;

    avifResult res = AVIF_RESULT_OK;
    uint8_t * inBuf[64] = { 0 }, * outBuf[64] = { 0 };
    uint8_t *dDesc = NULL, *dTok = NULL;
    cudaEvent_t ev[4] = { NULL, NULL, NULL, NULL };
    gGridLastMode = 0;
    if (nranks > 64) return AVIF_RESULT_INVALID_ARGUMENT;
    struct Band { uint32_t r0, r1, c0, c1; size_t yB, cB, aB, inB, outB; int owner; } band[64];
    int nb = 0;
    for (uint32_t r0 = 0; r0 < H && nb < nranks; r0 += per, ++nb) {
        struct Band * b = &band[nb];
        b->r0 = r0;
        b->r1 = (r0 + per > H) ? H : r0 + per;
        const uint32_t cLo = b->r0 >> sy, cHi = (uint32_t)(((uint64_t)b->r1 + sy) >> sy);
        b->c0 = (cLo > 0) ? cLo - 1 : 0;                 /* one halo row above ... */
        b->c1 = (cHi + 1 > chei) ? chei : cHi + 1;       /* ... and below, inside the plane */
        b->yB = (size_t)(b->r1 - b->r0) * W * bps;
        b->cB = haveC ? (size_t)(b->c1 - b->c0) * cwid * bps : 0;
        b->aB = haveA ? b->yB : 0;
        b->inB = b->yB + 2 * b->cB + b->aB;
        b->outB = (size_t)(b->r1 - b->r0) * W * pix;
        b->owner = (root + nb) % nranks;
    }
    if (phaseMs) {
        for (int k = 0; k < 4; ++k) GR_CUDA(cudaEventCreate(&ev[k]));
        GR_CUDA(cudaEventRecord(ev[0], st));
    }
    if (nranks > 1 && peerDecodeWanted()) {
        /* ---- peer-memory form (see avifB200GridRGBToYUV): every rank converts its band straight from the root's planes -- halo rows
         * included, they are just rows of the mapped plane -- into the root's pixel buffer.  OPT-IN (AVIF_B200_GRID_PEER_DECODE=1, to
         * be set on every rank alike): measured SLOWER than the staged exchange (2 GPUs, 8K RGBA16 output: 1.46 ms against 0.64 ms,
         * profiles/r02_variants.md) -- the pixel stores of a warp are 32 separate 16-byte pieces 64 bytes apart, which NVLink carries
         * at ~190 GB/s, where the encode side's plane stores are whole 256-byte runs ---- */
        PeerCanvas pc;
        memset(&pc, 0, sizeof(pc));
        if (isRoot) {
            const void * bufs[kPeerBufs] = { canvasRgb->pixels, canvas->yuvPlanes[0], haveC ? canvas->yuvPlanes[1] : NULL, haveC ? canvas->yuvPlanes[2] : NULL,
                                             haveA ? canvas->alphaPlane : NULL };
            pc.ok = (uint8_t)peerWanted();
            for (int b = 0; b < kPeerBufs && pc.ok; ++b) {
                pc.present[b] = (bufs[b] != NULL);
                if (bufs[b] && !peerExport(bufs[b], pc.handle[b], &pc.offset[b])) pc.ok = 0;
            }
        }
        GR_CUDA((cudaError_t)abScratchAllocOn((void **)&dDesc, sizeof(pc) + 64, st));
        if (isRoot) GR_CUDA(cudaMemcpyAsync(dDesc, &pc, sizeof(pc), cudaMemcpyHostToDevice, st));
        GR_NCCL(nccl.GroupStart());
        for (int r = 0; r < nranks; ++r) {
            if (r == root) continue;
            if (isRoot) GR_NCCL(nccl.Send(dDesc, sizeof(pc), kNcclUint8, r, comm, st));
            else if (r == rank) GR_NCCL(nccl.Recv(dDesc, sizeof(pc), kNcclUint8, root, comm, st));
        }
        GR_NCCL(nccl.GroupEnd());
        if (!isRoot) {
            GR_CUDA(cudaMemcpyAsync(&pc, dDesc, sizeof(pc), cudaMemcpyDeviceToHost, st));
            GR_CUDA(cudaStreamSynchronize(st));
        }
        if (pc.ok) {
            uint8_t token = 1;
            avifImage v = *canvas;
            avifRGBImage o = *canvasRgb;
            if (!isRoot) {
                uint8_t * mapped[kPeerBufs] = { NULL, NULL, NULL, NULL, NULL };
                for (int b = 0; b < kPeerBufs; ++b) {
                    if (!pc.present[b]) continue;
                    uint8_t * base = (uint8_t *)peerOpen(pc.handle[b]);
                    if (!base) token = 0;
                    else mapped[b] = base + pc.offset[b];
                }
                o.pixels = mapped[0];
                v.yuvPlanes[0] = mapped[1];
                v.yuvPlanes[1] = mapped[2];
                v.yuvPlanes[2] = mapped[3];
                v.alphaPlane = mapped[4];
            }
            if (phaseMs) GR_CUDA(cudaEventRecord(ev[1], st));
            for (int k = 0; k < nb && token; ++k) {
                if (band[k].owner != rank) continue;
                if (avifB200ImageYUVToRGBRowsAsync(&v, &o, band[k].r0, band[k].r1, st) != AVIF_RESULT_OK) token = 0;
            }
            if (phaseMs) GR_CUDA(cudaEventRecord(ev[2], st));
            GR_CUDA((cudaError_t)abScratchAllocOn((void **)&dTok, (size_t)nranks + 64, st));
            if (!isRoot) GR_CUDA(cudaMemcpyAsync(dTok + rank, &token, 1, cudaMemcpyHostToDevice, st));
            GR_NCCL(nccl.GroupStart());
            for (int r = 0; r < nranks; ++r) {
                if (r == root) continue;
                if (isRoot) GR_NCCL(nccl.Recv(dTok + r, 1, kNcclUint8, r, comm, st));
                else if (r == rank) GR_NCCL(nccl.Send(dTok + rank, 1, kNcclUint8, root, comm, st));
            }
            GR_NCCL(nccl.GroupEnd());
            if (isRoot) {
                uint8_t toks[64];
                memset(toks, 1, sizeof(toks));
                GR_CUDA(cudaMemcpyAsync(toks, dTok, (size_t)nranks, cudaMemcpyDeviceToHost, st));
                GR_CUDA(cudaStreamSynchronize(st));
                for (int r = 0; r < nranks; ++r)
                    if (r != root && !toks[r]) res = AVIF_RESULT_UNKNOWN_ERROR;
                if (!token) res = AVIF_RESULT_UNKNOWN_ERROR;
            } else if (!token) {
                res = AVIF_RESULT_UNKNOWN_ERROR;
            }
            gGridLastMode = 1;
            if (phaseMs) {
                GR_CUDA(cudaEventRecord(ev[3], st));
                GR_CUDA(cudaEventSynchronize(ev[3]));
                for (int k = 0; k < 3; ++k) GR_CUDA(cudaEventElapsedTime(&phaseMs[k], ev[k], ev[k + 1])); /* descriptor, convert over NVLink, tokens */
            }
            goto cleanup;
        }
    }
    /* ---- 1. scatter band inputs (packed [Y | U | V | A] rows) ---- */
    for (int k = 0; k < nb; ++k) {
        const struct Band * b = &band[k];
        if (b->owner == root) continue;
        if (isRoot || b->owner == rank) GR_CUDA((cudaError_t)abScratchAllocOn((void **)&inBuf[k], b->inB, st));
        if (isRoot) {
            uint8_t * p = inBuf[k];
            GR_CUDA(cudaMemcpy2DAsync(p, (size_t)W * bps, canvas->yuvPlanes[0] + (size_t)b->r0 * canvas->yuvRowBytes[0], canvas->yuvRowBytes[0], (size_t)W * bps,
                                      b->r1 - b->r0, cudaMemcpyDeviceToDevice, st));
            p += b->yB;
            if (haveC) {
                for (int c = 1; c < 3; ++c) {
                    GR_CUDA(cudaMemcpy2DAsync(p, (size_t)cwid * bps, canvas->yuvPlanes[c] + (size_t)b->c0 * canvas->yuvRowBytes[c], canvas->yuvRowBytes[c],
                                              (size_t)cwid * bps, b->c1 - b->c0, cudaMemcpyDeviceToDevice, st));
                    p += b->cB;
                }
            }
            if (haveA) {
                GR_CUDA(cudaMemcpy2DAsync(p, (size_t)W * bps, canvas->alphaPlane + (size_t)b->r0 * canvas->alphaRowBytes, canvas->alphaRowBytes, (size_t)W * bps,
                                          b->r1 - b->r0, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    if (nranks > 1) {
        GR_NCCL(nccl.GroupStart());
        for (int k = 0; k < nb; ++k) {
            const struct Band * b = &band[k];
            if (b->owner == root) continue;
            if (isRoot) GR_NCCL(nccl.Send(inBuf[k], b->inB, kNcclUint8, b->owner, comm, st));
            else if (b->owner == rank) GR_NCCL(nccl.Recv(inBuf[k], b->inB, kNcclUint8, root, comm, st));
        }
        GR_NCCL(nccl.GroupEnd());
    }
    if (phaseMs) GR_CUDA(cudaEventRecord(ev[1], st));
    /* ---- 2. convert my bands against whole-canvas geometry ---- */
    for (int k = 0; k < nb; ++k) {
        const struct Band * b = &band[k];
        if (b->owner != rank) continue;
        if (isRoot) {
            res = avifB200ImageYUVToRGBRowsAsync(canvas, canvasRgb, b->r0, b->r1, st);
        } else {
            GR_CUDA((cudaError_t)abScratchAllocOn((void **)&outBuf[k], b->outB, st));
            avifImage v = *canvas;
            avifRGBImage o = *canvasRgb;
            uint8_t * p = inBuf[k];
            /* virtual whole-frame base pointers: row j of the frame lives at base + j * rowBytes for every row the band touches */
            v.yuvRowBytes[0] = W * bps;
            v.yuvPlanes[0] = p - (size_t)b->r0 * v.yuvRowBytes[0];
            p += b->yB;
            v.yuvPlanes[1] = v.yuvPlanes[2] = NULL;
            v.yuvRowBytes[1] = v.yuvRowBytes[2] = 0;
            if (haveC) {
                for (int c = 1; c < 3; ++c) {
                    v.yuvRowBytes[c] = cwid * bps;
                    v.yuvPlanes[c] = p - (size_t)b->c0 * v.yuvRowBytes[c];
                    p += b->cB;
                }
            }
            v.alphaPlane = NULL;
            v.alphaRowBytes = 0;
            if (haveA) {
                v.alphaRowBytes = W * bps;
                v.alphaPlane = p - (size_t)b->r0 * v.alphaRowBytes;
            }
            o.rowBytes = W * pix;
            o.pixels = outBuf[k] - (size_t)b->r0 * o.rowBytes;
            res = avifB200ImageYUVToRGBRowsAsync(&v, &o, b->r0, b->r1, st);
        }
        if (res != AVIF_RESULT_OK) goto cleanup;
    }
    if (phaseMs) GR_CUDA(cudaEventRecord(ev[2], st));
    /* ---- 3. gather the RGB rows: a full-width band is contiguous in a tightly packed canvas, else via a staging copy ---- */
    if (nranks > 1) {
        const int tight = (canvasRgb->rowBytes == W * pix);
        if (isRoot && !tight) {
            for (int k = 0; k < nb; ++k)
                if (band[k].owner != root) GR_CUDA((cudaError_t)abScratchAllocOn((void **)&outBuf[k], band[k].outB, st));
        }
        GR_NCCL(nccl.GroupStart());
        for (int k = 0; k < nb; ++k) {
            const struct Band * b = &band[k];
            if (b->owner == root) continue;
            if (isRoot) {
                uint8_t * dst = tight ? canvasRgb->pixels + (size_t)b->r0 * canvasRgb->rowBytes : outBuf[k];
                GR_NCCL(nccl.Recv(dst, b->outB, kNcclUint8, b->owner, comm, st));
            } else if (b->owner == rank) {
                GR_NCCL(nccl.Send(outBuf[k], b->outB, kNcclUint8, root, comm, st));
            }
        }
        GR_NCCL(nccl.GroupEnd());
        if (isRoot && !tight) {
            for (int k = 0; k < nb; ++k) {
                const struct Band * b = &band[k];
                if (b->owner == root) continue;
                GR_CUDA(cudaMemcpy2DAsync(canvasRgb->pixels + (size_t)b->r0 * canvasRgb->rowBytes, canvasRgb->rowBytes, outBuf[k], (size_t)W * pix, (size_t)W * pix,
                                          b->r1 - b->r0, cudaMemcpyDeviceToDevice, st));
            }
        }
    }
    if (phaseMs) {
        GR_CUDA(cudaEventRecord(ev[3], st));
        GR_CUDA(cudaEventSynchronize(ev[3]));
        for (int k = 0; k < 3; ++k) GR_CUDA(cudaEventElapsedTime(&phaseMs[k], ev[k], ev[k + 1])); /* scatter, convert, gather */
    }

cleanup:
    if (dDesc) (void)cudaFreeAsync(dDesc, st);
    if (dTok) (void)cudaFreeAsync(dTok, st);
    for (int k = 0; k < 64; ++k) {
        if (inBuf[k]) (void)cudaFreeAsync(inBuf[k], st);
        if (outBuf[k]) (void)cudaFreeAsync(outBuf[k], st);
    }
    for (int k = 0; k < 4; ++k)
        if (ev[k]) (void)cudaEventDestroy(ev[k]);
    return res;
}
