"""Tile-grid RGB->YUV across GPUs (BASELINE.json config C4; reference flow: apps/avifenc.c reads the whole canvas,
avifImageRGBToYUV converts it ONCE, avifImageSplitGrid cuts views, avifEncoderAddImageGrid encodes the cells --
SURVEY.md 3.4).  Here the canvas lives on the root GPU; every cell's pixels travel to the cell's owner over NCCL
point-to-point (NVLink), owners convert with the C-ABI (no collective inside the computation), planes travel back and are
assembled into the canvas planes with avifB200ImageCopySamplesAsync (the decode-side tile copy of src/read.c:1823).
Cell sizes are even when chroma is subsampled, so the result is bit-identical to converting the whole canvas.

torch / torch.distributed are plumbing (device memory, NCCL); the conversion itself is libavif_b200.so.
"""
import ctypes

from . import abi, sharding
from .images import RGBImage, YUVImage


def _wrap_rgb(t, width, height, depth, fmt, **kw):
    """avifRGBImage over an existing contiguous (height, width*pixelBytes) uint8 tensor."""
    rgb = RGBImage(width, height, depth, fmt, allocate=False, **kw)
    rgb.c.pixels = t.data_ptr()
    rgb.c.rowBytes = t.shape[1]
    rgb._keep = t
    return rgb


def grid_rgb_to_yuv(lib, dist, canvas_rgb, width, height, cols, rows, depth, yuv_format, yuv_range, mc, rgb_depth, rgb_fmt,
                    device, stream_ptr, root=0, timers=None):
    """Converts a cols x rows grid.  `canvas_rgb` is an RGBImage on `device` on the root rank (None elsewhere).
    Returns the assembled YUVImage (device-resident) on root, None elsewhere."""
    import torch
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rects = sharding.split_grid(width, height, cols, rows, yuv_format)
    pb = abi.rgb_pixel_size(rgb_fmt, rgb_depth)
    has_alpha = abi.rgb_format_has_alpha(rgb_fmt)
    bps = 2 if depth > 8 else 1

    def chroma_dims(r):
        sx = 0 if yuv_format == abi.AVIF_PIXEL_FORMAT_YUV444 else 1
        sy = 1 if yuv_format == abi.AVIF_PIXEL_FORMAT_YUV420 else 0
        return (r.width + sx) >> sx, (r.height + sy) >> sy

    def mark(name):
        if timers is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream())
            timers.append((name, e))

    mark("start")
    # ---- 1. scatter the source pixels of every cell to its owner (root keeps views of its own cells)
    mine = {}
    ops, keep = [], []
    canvas2d = None
    if rank == root:
        canvas2d = canvas_rgb.buf._t[:canvas_rgb.buf.nbytes].view(height, canvas_rgb.buf.row_bytes)
    for u, r in enumerate(rects):
        owner = sharding.owner_of(u, world)
        if rank == root:
            cell = canvas2d[r.y:r.y + r.height, r.x * pb:(r.x + r.width) * pb]
            if owner == root:
                mine[u] = cell  # strided view: converted in place, no copy
            else:
                t = cell.contiguous()
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, owner))
        elif owner == rank:
            t = torch.empty((r.height, r.width * pb), dtype=torch.uint8, device=device)
            mine[u] = t
            ops.append(dist.P2POp(dist.irecv, t, root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    mark("scatter")

    # ---- 2. local conversion of the owned cells: one batched launch
    from .images import pointer_array
    cells_rgb, cells_yuv, order = [], [], []
    for u, t in mine.items():
        r = rects[u]
        rgb = RGBImage(r.width, r.height, rgb_depth, rgb_fmt, allocate=False)
        rgb.c.pixels = t.data_ptr()
        rgb.c.rowBytes = t.stride(0)
        rgb._keep = t
        img = YUVImage(r.width, r.height, depth, yuv_format, yuv_range, mc, alpha=has_alpha, device=device)
        cells_rgb.append(rgb)
        cells_yuv.append(img)
        order.append(u)
    if order:
        res = lib.avifB200ImageRGBToYUVBatchAsync(pointer_array(cells_yuv), pointer_array(cells_rgb), len(order), stream_ptr)
        assert res == 0, res
    mark("convert")

    # ---- 3. gather the planes on root and assemble the canvas planes
    out = None
    ops = []
    pending = []
    if rank == root:
        out = YUVImage(width, height, depth, yuv_format, yuv_range, mc, alpha=has_alpha, device=device)
    for u, r in enumerate(rects):
        owner = sharding.owner_of(u, world)
        cw, ch = chroma_dims(r)
        shapes = [(r.height, r.width * bps)]
        if yuv_format != abi.AVIF_PIXEL_FORMAT_YUV400:
            shapes += [(ch, cw * bps), (ch, cw * bps)]
        if has_alpha:
            shapes.append((r.height, r.width * bps))
        if rank == root and owner != root:
            img = YUVImage(r.width, r.height, depth, yuv_format, yuv_range, mc, alpha=has_alpha, device=device)
            for p in img.present_planes():
                ops.append(dist.P2POp(dist.irecv, img.bufs[p]._t[:img.bufs[p].nbytes], owner))
            pending.append((u, img))
        elif rank != root and owner == rank:
            img = cells_yuv[order.index(u)]
            for p in img.present_planes():
                ops.append(dist.P2POp(dist.isend, img.bufs[p]._t[:img.bufs[p].nbytes], root))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    mark("gather")
    if rank == root:
        for u, img in [(u, cells_yuv[order.index(u)]) for u in order] + pending:
            view = sharding.view_yuv(out.c, rects[u])
            assert lib.avifB200ImageCopySamplesAsync(ctypes.byref(view), img.ptr(), abi.AVIF_PLANES_ALL, stream_ptr) == 0
        out._keep = (cells_yuv, pending)
    mark("assemble")
    return out
