/*
 * avif_oracle_scale.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of avifImageScale (ref: src/scale.c:24-207) and of the
 * libyuv loops it reaches with kFilterBox (ref: third_party/libyuv/source/scale.c, scale_common.c, scale_any.c, row_common.c
 * in the reference tree, v1880, C paths), written as plain sequential row loops over int arrays -- deliberately NOT the
 * table-plus-per-sample form of the product (libavif_b200/csrc/scale.c + scale_kernels.cu), so that the two check each other.
 * Pinned BIT-EXACT to the reference compiled from its own sources (tests/test_oracle_scale.py) -- parity status PINNED.
 *
 * Only tests/ may load this.  Nothing under libavif_b200/ links or calls it.
 */
#include "avif_b200.h"

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

enum { F_NONE, F_LINEAR, F_BILINEAR, F_BOX };

static int reduceFilter(int sw, int sh, int dw, int dh, int f) /* ref: scale_common.c:428-469 */
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
static int fdiv(int n, int d) { return (int)(((int64_t)n << 16) / d); }                         /* FixedDiv */
static int fdiv1(int n, int d) { return (int)((((int64_t)n << 16) - 0x00010001) / (d - 1)); } /* FixedDiv1 */
static int cstart(int d, int s) { return (d < 0) ? -((-d >> 1) + s) : ((d >> 1) + s); }
static int atLeast1(int v) { return v < 1 ? 1 : v; }

static void slope(int sw, int sh, int dw, int dh, int f, int * x, int * y, int * dx, int * dy) /* ref: scale_common.c:484-553 */
{
    if (dw == 1 && sw >= 32768) dw = sw;
    if (dh == 1 && sh >= 32768) dh = sh;
    if (f == F_BOX) {
        *dx = fdiv(sw, dw); *dy = fdiv(sh, dh); *x = 0; *y = 0;
    } else if (f == F_BILINEAR || f == F_LINEAR) {
        if (dw <= sw) { *dx = fdiv(sw, dw); *x = cstart(*dx, -32768); }
        else if (sw > 1 && dw > 1) { *dx = fdiv1(sw, dw); *x = 0; }
        if (f == F_BILINEAR) {
            if (dh <= sh) { *dy = fdiv(sh, dh); *y = cstart(*dy, -32768); }
            else if (sh > 1 && dh > 1) { *dy = fdiv1(sh, dh); *y = 0; }
        } else {
            *dy = fdiv(sh, dh); *y = *dy >> 1;
        }
    } else {
        *dx = fdiv(sw, dw); *dy = fdiv(sh, dh); *x = cstart(*dx, 0); *y = cstart(*dy, 0);
    }
}

typedef struct Plane { /* samples as ints; row r starts at v + r * w */
    int * v;
    int w, h;
} Plane;
static const int * rowOf(const Plane * p, int r)
{
    if (r < 0) r = 0;
    if (r >= p->h) r = p->h - 1; /* rows the reference touches only with zero weight */
    return p->v + (size_t)r * p->w;
}
static int at(const int * row, int w, int x) { return row[x < 0 ? 0 : (x >= w ? w - 1 : x)]; }

/* InterpolateRow_C / _16_C, ref: row_common.c:46-105 */
static void interpolateRow(int * dst, const int * r0, const int * r1, int w, int yf)
{
    for (int x = 0; x < w; ++x) {
        if (yf == 0) dst[x] = r0[x];
        else if (yf == 128) dst[x] = (r0[x] + r1[x] + 1) >> 1;
        else dst[x] = (r0[x] * (256 - yf) + r1[x] * yf + 128) >> 8;
    }
}
/* ScaleFilterCols_C / _16_C, ref: scale_common.c:188-314 */
static void filterCols(int * dst, const int * src, int sw, int dw, int x, int dx, int sixteen)
{
    for (int i = 0; i < dw; ++i) {
        const int xi = x >> 16, f = x & 0xffff;
        const int a = at(src, sw, xi), b = at(src, sw, xi + 1);
        if (sixteen) dst[i] = (uint16_t)(a + (int)((((int64_t)f * ((int64_t)b - a)) + 0x8000) >> 16));
        else dst[i] = (uint8_t)(a + ((((f >> 9) * (b - a)) + 0x40) >> 7));
        x += dx;
    }
}
/* ScaleRowUp2_Linear_Any_C, ref: scale_any.c:27-40 + scale_common.c:29-41 */
static void up2LinearRow(int * dst, const int * s, int dw)
{
    const int work = (dw - 1) & ~1;
    dst[0] = s[0];
    for (int x = 0; x < work / 2; ++x) {
        dst[1 + 2 * x] = (s[x] * 3 + s[x + 1] * 1 + 2) >> 2;
        dst[2 + 2 * x] = (s[x] * 1 + s[x + 1] * 3 + 2) >> 2;
    }
    dst[dw - 1] = s[(dw - 1) / 2];
}
/* ScaleRowUp2_Bilinear_Any_C writing the two destination rows da, db (which may be the same row), ref: scale_any.c:54-78 +
 * scale_common.c:51-74 */
static void up2BilinearRows(int * da, int * db, const int * sa, const int * sb, int dw)
{
    const int work = (dw - 1) & ~1;
    da[0] = (3 * sa[0] + sb[0] + 2) >> 2;
    db[0] = (sa[0] + 3 * sb[0] + 2) >> 2;
    for (int x = 0; x < work / 2; ++x) {
        int * d = da + 1, * e = db + 1;
        d[2 * x + 0] = (sa[x] * 9 + sa[x + 1] * 3 + sb[x] * 3 + sb[x + 1] * 1 + 8) >> 4;
        d[2 * x + 1] = (sa[x] * 3 + sa[x + 1] * 9 + sb[x] * 1 + sb[x + 1] * 3 + 8) >> 4;
        e[2 * x + 0] = (sa[x] * 3 + sa[x + 1] * 1 + sb[x] * 9 + sb[x + 1] * 3 + 8) >> 4;
        e[2 * x + 1] = (sa[x] * 1 + sa[x + 1] * 3 + sb[x] * 3 + sb[x + 1] * 9 + 8) >> 4;
    }
    const int k = (dw - 1) / 2;
    da[dw - 1] = (3 * sa[k] + sb[k] + 2) >> 2;
    db[dw - 1] = (sa[k] + 3 * sb[k] + 2) >> 2;
}

static int scalePlane(const Plane * S, Plane * D, int sixteen)
{
    const int sw = S->w, sh = S->h, dw = D->w, dh = D->h;
    const int f = reduceFilter(sw, sh, dw, dh, F_BOX);
    enum { COPY, VERT, BOX, UP2L, UP2B, BILUP, BILDOWN, SIMPLE } mode;
    int picked = 0;
    mode = COPY;
    if (sixteen) { /* ScalePlane_12, ref: scale.c:973-1007 */
        if ((dw + 1) / 2 == sw && f == F_LINEAR) { mode = UP2L; picked = 1; }
        else if ((dh + 1) / 2 == sh && (dw + 1) / 2 == sw && (f == F_BILINEAR || f == F_BOX)) { mode = UP2B; picked = 1; }
    }
    if (!picked) { /* ScalePlane / ScalePlane_16, ref: scale.c:829-971 */
        if (dw == sw && dh == sh) mode = COPY;
        else if (dw == sw && f != F_BOX) mode = VERT;
        else if (f == F_BOX && dh * 2 < sh) mode = BOX;
        else if ((dw + 1) / 2 == sw && f == F_LINEAR) mode = UP2L;
        else if ((dh + 1) / 2 == sh && (dw + 1) / 2 == sw && (f == F_BILINEAR || f == F_BOX)) mode = UP2B;
        else if (f && dh > sh) mode = BILUP;
        else if (f) mode = BILDOWN;
        else mode = SIMPLE;
    }
    int x = 0, y = 0, dx = 0, dy = 0;
    int * bufA = (int *)malloc(sizeof(int) * (size_t)((sw > dw ? sw : dw) + 2));
    int * bufB = (int *)malloc(sizeof(int) * (size_t)((sw > dw ? sw : dw) + 2));
    uint32_t * acc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(sw + 1));
    if (!bufA || !bufB || !acc) {
        free(bufA); free(bufB); free(acc);
        return 0;
    }
    switch (mode) {
        case COPY:
            for (int j = 0; j < dh; ++j) memcpy(D->v + (size_t)j * dw, rowOf(S, j), sizeof(int) * (size_t)dw);
            break;
        case VERT: { /* ref: scale.c:855-871, scale_common.c:348-386 */
            if (dh <= sh) { dy = fdiv(sh, dh); y = cstart(dy, -32768); }
            else if (sh > 1 && dh > 1) dy = fdiv1(sh, dh);
            const int maxY = (sh > 1) ? ((sh - 1) << 16) - 1 : 0;
            for (int j = 0; j < dh; ++j) {
                if (y > maxY) y = maxY;
                const int yi = y >> 16, yf = f ? ((y >> 8) & 255) : 0;
                interpolateRow(D->v + (size_t)j * dw, rowOf(S, yi), rowOf(S, yi + 1), dw, yf);
                y += dy;
            }
            break;
        }
        case BOX: { /* ref: scale.c:148-256 with ScaleAddRow / ScaleAddCols0,1,2 */
            slope(sw, sh, dw, dh, F_BOX, &x, &y, &dx, &dy);
            const int maxY = sh << 16;
            for (int j = 0; j < dh; ++j) {
                const int iy = y >> 16;
                y += dy;
                if (y > maxY) y = maxY;
                const int bh = atLeast1((y >> 16) - iy);
                memset(acc, 0, sizeof(uint32_t) * (size_t)sw);
                for (int k = 0; k < bh; ++k) {
                    const int * r = rowOf(S, iy + k);
                    for (int c = 0; c < sw; ++c) {
                        acc[c] += (uint32_t)r[c];
                        if (!sixteen) acc[c] &= 0xFFFFu; /* uint16_t row accumulators */
                    }
                }
                int * out = D->v + (size_t)j * dw;
                if (dx & 0xffff) {
                    const int minbw = dx >> 16;
                    int scaletbl[2];
                    scaletbl[0] = 65536 / (atLeast1(minbw) * bh);
                    scaletbl[1] = 65536 / (atLeast1(minbw + 1) * bh);
                    int xx = x;
                    for (int i = 0; i < dw; ++i) {
                        const int ix = xx >> 16;
                        xx += dx;
                        const int bw = atLeast1((xx >> 16) - ix);
                        uint32_t sum = 0;
                        for (int c = 0; c < bw; ++c) sum += acc[(ix + c) < sw ? (ix + c) : sw - 1];
                        const uint32_t v = (sum * (uint32_t)scaletbl[bw - minbw]) >> 16;
                        out[i] = sixteen ? (int)(uint16_t)v : (int)(uint8_t)v;
                    }
                } else {
                    const int bw = atLeast1(dx >> 16);
                    const int scaleval = 65536 / (bw * bh);
                    int xs = x >> 16;
                    for (int i = 0; i < dw; ++i) {
                        uint32_t sum = 0;
                        for (int c = 0; c < bw; ++c) sum += acc[(xs + c) < sw ? (xs + c) : sw - 1];
                        const uint32_t v = (sum * (uint32_t)scaleval) >> 16;
                        out[i] = sixteen ? (int)(uint16_t)v : (int)(uint8_t)v;
                        xs += bw;
                    }
                }
            }
            break;
        }
        case UP2L: { /* ref: scale.c:467-501 */
            if (dh == 1) {
                up2LinearRow(D->v, rowOf(S, (sh - 1) / 2), dw);
            } else {
                dy = fdiv(sh - 1, dh - 1);
                y = (1 << 15) - 1;
                for (int j = 0; j < dh; ++j) {
                    memcpy(bufA, rowOf(S, y >> 16), sizeof(int) * (size_t)sw);
                    bufA[sw] = bufA[sw - 1];
                    up2LinearRow(D->v + (size_t)j * dw, bufA, dw);
                    y += dy;
                }
            }
            break;
        }
        case UP2B: { /* ref: scale.c:503-538 */
            int * d = D->v;
            int srcRow = 0;
            memcpy(bufA, rowOf(S, 0), sizeof(int) * (size_t)sw);
            bufA[sw] = bufA[sw - 1];
            up2BilinearRows(d, d, bufA, bufA, dw); /* strides 0: both outputs are the first row */
            d += dw;
            for (int k = 0; k < sh - 1; ++k) {
                memcpy(bufA, rowOf(S, srcRow), sizeof(int) * (size_t)sw);
                bufA[sw] = bufA[sw - 1];
                memcpy(bufB, rowOf(S, srcRow + 1), sizeof(int) * (size_t)sw);
                bufB[sw] = bufB[sw - 1];
                up2BilinearRows(d, d + dw, bufA, bufB, dw);
                ++srcRow;
                d += 2 * (size_t)dw;
            }
            if (!(dh & 1)) {
                memcpy(bufA, rowOf(S, srcRow), sizeof(int) * (size_t)sw);
                bufA[sw] = bufA[sw - 1];
                up2BilinearRows(d, d, bufA, bufA, dw);
            }
            break;
        }
        case BILUP: { /* ref: scale.c:371-466 / :669-761 */
            slope(sw, sh, dw, dh, f, &x, &y, &dx, &dy);
            const int maxY = (sh - 1) << 16;
            if (y > maxY) y = maxY;
            int yi = y >> 16, src = yi, lasty = yi;
            int * rowptr = bufA, * other = bufB;
            filterCols(rowptr, rowOf(S, src), sw, dw, x, dx, sixteen);
            if (sh > 1) ++src;
            filterCols(other, rowOf(S, src), sw, dw, x, dx, sixteen);
            if (sh > 2) ++src;
            for (int j = 0; j < dh; ++j) {
                yi = y >> 16;
                if (yi != lasty) {
                    if (y > maxY) {
                        y = maxY;
                        yi = y >> 16;
                        src = yi;
                    }
                    if (yi != lasty) {
                        filterCols(rowptr, rowOf(S, src), sw, dw, x, dx, sixteen);
                        int * t = rowptr;
                        rowptr = other;
                        other = t;
                        lasty = yi;
                        if ((y + 65536) < maxY) ++src;
                    }
                }
                if (f == F_LINEAR) interpolateRow(D->v + (size_t)j * dw, rowptr, rowptr, dw, 0);
                else interpolateRow(D->v + (size_t)j * dw, rowptr, other, dw, (y >> 8) & 255);
                y += dy;
            }
            break;
        }
        case BILDOWN: { /* ref: scale.c:258-369 */
            slope(sw, sh, dw, dh, f, &x, &y, &dx, &dy);
            const int maxY = (sh - 1) << 16;
            if (y > maxY) y = maxY;
            for (int j = 0; j < dh; ++j) {
                const int yi = y >> 16;
                if (f == F_LINEAR) {
                    filterCols(D->v + (size_t)j * dw, rowOf(S, yi), sw, dw, x, dx, sixteen);
                } else {
                    interpolateRow(bufA, rowOf(S, yi), rowOf(S, yi + 1), sw, (y >> 8) & 255);
                    filterCols(D->v + (size_t)j * dw, bufA, sw, dw, x, dx, sixteen);
                }
                y += dy;
                if (y > maxY) y = maxY;
            }
            break;
        }
        default: { /* ScalePlaneSimple, ref: scale.c:763-827 */
            slope(sw, sh, dw, dh, F_NONE, &x, &y, &dx, &dy);
            const int up2 = (sw * 2 == dw && x < 0x8000);
            for (int j = 0; j < dh; ++j) {
                const int * r = rowOf(S, y >> 16);
                int xx = x;
                for (int i = 0; i < dw; ++i) {
                    D->v[(size_t)j * dw + i] = up2 ? r[i >> 1] : at(r, sw, xx >> 16);
                    xx += dx;
                }
                y += dy;
            }
            break;
        }
    }
    free(bufA);
    free(bufB);
    free(acc);
    return 1;
}

static void chromaShift(avifPixelFormat f, int * sx, int * sy, int * mono)
{
    *sx = (f == AVIF_PIXEL_FORMAT_YUV422 || f == AVIF_PIXEL_FORMAT_YUV420) ? 1 : 0;
    *sy = (f == AVIF_PIXEL_FORMAT_YUV420) ? 1 : 0;
    *mono = (f == AVIF_PIXEL_FORMAT_YUV400);
}

/* ref: src/scale.c:24-207.  New planes are malloc'ed and owned by the image; old ones are freed if the image owned them. */
ORACLE_API avifResult oracleImageScale(avifImage * image, uint32_t dstWidth, uint32_t dstHeight, avifDiagnostics * diag)
{
    if (diag) diag->error[0] = '\0';
    if (image->width == dstWidth && image->height == dstHeight) return AVIF_RESULT_OK;
    if (!dstWidth || !dstHeight) {
        if (diag) snprintf(diag->error, sizeof(diag->error), "invalid dst dimensions [%ux%u]", dstWidth, dstHeight);
        return AVIF_RESULT_INVALID_ARGUMENT;
    }
    if (dstWidth > (16384u * 16384u) / dstHeight || dstWidth > 32768u || dstHeight > 32768u) {
        if (diag) snprintf(diag->error, sizeof(diag->error), "dst dimensions too large [%ux%u]", dstWidth, dstHeight);
        return AVIF_RESULT_NOT_IMPLEMENTED;
    }
    avifImage old = *image;
    for (int c = 0; c < 3; ++c) { image->yuvPlanes[c] = NULL; image->yuvRowBytes[c] = 0; }
    image->imageOwnsYUVPlanes = AVIF_FALSE;
    image->alphaPlane = NULL;
    image->alphaRowBytes = 0;
    image->imageOwnsAlphaPlane = AVIF_FALSE;
    image->width = dstWidth;
    image->height = dstHeight;
    avifResult res = AVIF_RESULT_OK;
    if ((old.yuvPlanes[0] || old.alphaPlane) && (old.width > 16384 || old.height > 16384)) {
        if (diag) snprintf(diag->error, sizeof(diag->error), "invalid scale for libyuv");
        res = AVIF_RESULT_NOT_IMPLEMENTED;
    }
    int sx, sy, mono;
    chromaShift(image->yuvFormat, &sx, &sy, &mono);
    const int cs = (image->depth > 8) ? 2 : 1;
    for (int c = 0; c < 4 && res == AVIF_RESULT_OK; ++c) {
        const int chroma = (c == 1 || c == 2);
        const uint8_t * sp = (c < 3) ? old.yuvPlanes[c] : old.alphaPlane;
        const uint32_t srb = (c < 3) ? old.yuvRowBytes[c] : old.alphaRowBytes;
        const int wanted = (c < 3) ? (old.yuvPlanes[0] != NULL && image->yuvFormat != AVIF_PIXEL_FORMAT_NONE && !(chroma && mono)) : (old.alphaPlane != NULL);
        if (!wanted) continue;
        const int dw = chroma ? (int)((dstWidth + sx) >> sx) : (int)dstWidth, dh = chroma ? (int)((dstHeight + sy) >> sy) : (int)dstHeight;
        const int sw = chroma ? (int)((old.width + sx) >> sx) : (int)old.width, sh = chroma ? (int)((old.height + sy) >> sy) : (int)old.height;
        uint8_t * host = (uint8_t *)malloc((size_t)cs * dw * dh);
        if (!host) { res = AVIF_RESULT_OUT_OF_MEMORY; break; }
        if (c < 3) { image->yuvPlanes[c] = host; image->yuvRowBytes[c] = (uint32_t)(cs * dw); image->imageOwnsYUVPlanes = AVIF_TRUE; }
        else { image->alphaPlane = host; image->alphaRowBytes = (uint32_t)(cs * dw); image->imageOwnsAlphaPlane = AVIF_TRUE; }
        if (!sp) continue;
        Plane S = { (int *)malloc(sizeof(int) * (size_t)sw * sh), sw, sh }, D = { (int *)malloc(sizeof(int) * (size_t)dw * dh), dw, dh };
        if (!S.v || !D.v) { free(S.v); free(D.v); res = AVIF_RESULT_OUT_OF_MEMORY; break; }
        for (int j = 0; j < sh; ++j) {
            for (int i = 0; i < sw; ++i) {
                if (cs == 2) { uint16_t v; memcpy(&v, sp + (size_t)j * srb + 2 * (size_t)i, 2); S.v[(size_t)j * sw + i] = v; }
                else S.v[(size_t)j * sw + i] = sp[(size_t)j * srb + i];
            }
        }
        if (!scalePlane(&S, &D, cs == 2)) res = AVIF_RESULT_OUT_OF_MEMORY;
        for (int j = 0; j < dh && res == AVIF_RESULT_OK; ++j) {
            for (int i = 0; i < dw; ++i) {
                if (cs == 2) { const uint16_t v = (uint16_t)D.v[(size_t)j * dw + i]; memcpy(host + ((size_t)j * dw + i) * 2, &v, 2); }
                else host[(size_t)j * dw + i] = (uint8_t)D.v[(size_t)j * dw + i];
            }
        }
        free(S.v);
        free(D.v);
    }
    if (old.yuvPlanes[0] && old.imageOwnsYUVPlanes) {
        for (int c = 0; c < 3; ++c) free(old.yuvPlanes[c]);
    }
    if (old.alphaPlane && old.imageOwnsAlphaPlane) free(old.alphaPlane);
    return res;
}
