"""Host-side partition rules of the multi-GPU paths and the tile scatter / gather plumbing.

The reformat path shards by independent units: whole frames (batch decode, config C5) or grid cells (config C4).
Cells are legal units because the reference's own split rule makes cell sizes even whenever chroma is subsampled
(apps/shared/avifutil.c:493-546, src/avif.c:1061-1063), so no 2x2 averaging block straddles a cell and converting
cell by cell is bit-identical to converting the canvas.  (The opposite direction with BILINEAR 4:2:0 upsampling is NOT
tile-local -- chroma taps cross cell seams -- so YUV->RGB of a grid shards by row bands with a one-chroma-row halo.)

Communication is torch.distributed point-to-point (NCCL over NVLink on the GPUs, gloo in the CPU tests): the rank that
holds the canvas sends each cell's pixels to its owner, owners convert locally (no collective inside the computation)
and send the planes back.  PyTorch is plumbing only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Sequence

from . import abi


@dataclass(frozen=True)
class Rect:
    x: int
    y: int
    width: int
    height: int


def best_cell_size(num_pixels: int, num_cells: int, subsampled: bool) -> int:
    """avifGetBestCellSize (apps/shared/avifutil.c:493-546): ceil split, >= 64, even when chroma is subsampled."""
    if num_pixels <= 0 or num_cells <= 0:
        raise ValueError("empty dimension")
    cell = (num_pixels + num_cells - 1) // num_cells
    if cell < 64:
        cell = 64
        if (num_cells - 1) * cell >= num_pixels:
            raise ValueError(f"too many cells ({num_cells}) to have at least 64 pixels per cell")
    if cell > 65536:
        raise ValueError(f"cell size {cell} exceeds the maximum AV1 frame size")
    if subsampled and (cell & 1):
        cell += 1
        if (num_cells - 1) * cell >= num_pixels:
            raise ValueError(f"odd cell size {cell - 1} is forbidden on a subsampled image")
    assert (num_cells - 1) * cell < num_pixels <= num_cells * cell
    return cell


def split_grid(width: int, height: int, cols: int, rows: int, yuv_format: int) -> List[Rect]:
    """avifImageSplitGrid's rectangles (apps/shared/avifutil.c:548-590), row-major; last column/row are clipped."""
    mono = yuv_format == abi.AVIF_PIXEL_FORMAT_YUV400
    sub_x = (not mono) and yuv_format in (abi.AVIF_PIXEL_FORMAT_YUV422, abi.AVIF_PIXEL_FORMAT_YUV420)
    sub_y = (not mono) and yuv_format == abi.AVIF_PIXEL_FORMAT_YUV420
    cw = best_cell_size(width, cols, sub_x)
    ch = best_cell_size(height, rows, sub_y)
    out = []
    for gy in range(rows):
        for gx in range(cols):
            x, y = gx * cw, gy * ch
            out.append(Rect(x, y, min(cw, width - x), min(ch, height - y)))
    return out


def split_rows(height: int, parts: int, halo_even: bool = True) -> List[Rect]:
    """Row bands for sharding ONE frame's YUV->RGB (bilinear 4:2:0 needs one extra chroma row per seam); even starts."""
    per = (height + parts - 1) // parts
    if halo_even:
        per = (per + 1) & ~1
    out, y = [], 0
    while y < height:
        out.append(Rect(0, y, 0, min(per, height - y)))
        y += per
    return out


def owner_of(unit: int, world_size: int) -> int:
    """unit f -> rank f mod G: round robin keeps every rank busy for any unit count."""
    return unit % world_size


def units_of(rank: int, n_units: int, world_size: int) -> List[int]:
    return [u for u in range(n_units) if owner_of(u, world_size) == rank]


def view_rgb(rgb: abi.avifRGBImage, rect: Rect) -> abi.avifRGBImage:
    """Shallow view of a pixel rectangle (what a caller does with rgb.pixels + offsets; cf. reformat.c:1717-1719)."""
    v = abi.avifRGBImage()
    for name, _ in abi.avifRGBImage._fields_:
        setattr(v, name, getattr(rgb, name))
    v.width, v.height = rect.width, rect.height
    v.pixels = rgb.pixels + rect.y * rgb.rowBytes + rect.x * abi.rgb_pixel_size(rgb.format, rgb.depth)
    return v


def view_yuv(img: abi.avifImage, rect: Rect) -> abi.avifImage:
    """avifImageSetViewRect (src/avif.c:325-358): planes are offset pointers into the parent, nothing is copied."""
    mono = img.yuvFormat == abi.AVIF_PIXEL_FORMAT_YUV400
    sx = 0 if img.yuvFormat == abi.AVIF_PIXEL_FORMAT_YUV444 else 1
    sy = 1 if img.yuvFormat in (abi.AVIF_PIXEL_FORMAT_YUV420, abi.AVIF_PIXEL_FORMAT_YUV400) else 0
    if rect.width > img.width or rect.height > img.height or rect.x > img.width - rect.width or rect.y > img.height - rect.height:
        raise ValueError("rectangle outside the image")
    if not mono and ((rect.x & sx) or (rect.y & sy)):
        raise ValueError("odd offset on a subsampled image")
    v = abi.avifImage()
    for name in ("depth", "yuvFormat", "yuvRange", "yuvChromaSamplePosition", "alphaPremultiplied", "colorPrimaries",
                 "transferCharacteristics", "matrixCoefficients"):
        setattr(v, name, getattr(img, name))
    v.width, v.height = rect.width, rect.height
    bps = 2 if img.depth > 8 else 1
    if img.yuvPlanes[0]:
        for p in range(3):
            if img.yuvRowBytes[p]:
                px = rect.x if p == 0 else rect.x >> sx
                py = rect.y if p == 0 else rect.y >> sy
                v.yuvPlanes[p] = img.yuvPlanes[p] + py * img.yuvRowBytes[p] + px * bps
                v.yuvRowBytes[p] = img.yuvRowBytes[p]
    if img.alphaPlane:
        v.alphaPlane = img.alphaPlane + rect.y * img.alphaRowBytes + rect.x * bps
        v.alphaRowBytes = img.alphaRowBytes
    return v


def scatter_convert_gather(dist, rank: int, world_size: int, rects: Sequence[Rect], send_cell: Callable, recv_cell: Callable,
                           convert_cell: Callable, send_result: Callable, recv_result: Callable, root: int = 0) -> Dict[int, object]:
    """Generic tile pipeline: `root` owns the canvas; cell u goes to rank u mod G, is converted there and its planes come
    back.  All five callables are supplied by the caller (tensors on the GPU with NCCL, numpy-backed tensors with gloo):
      send_cell(u, dst) / recv_cell(u, src) -> cell      move one cell's source pixels
      convert_cell(u, cell) -> result                     local conversion (the C-ABI call)
      send_result(u, result, dst) / recv_result(u, src)   move one cell's planes back
    Returns, on root, {unit: result} for every unit; elsewhere the units converted locally."""
    mine = {}
    # 1. scatter (root keeps its own cells without communication)
    for u in range(len(rects)):
        dst = owner_of(u, world_size)
        if rank == root:
            mine[u] = send_cell(u, dst) if dst != root else send_cell(u, root)
        elif dst == rank:
            mine[u] = recv_cell(u, root)
    # 2. local conversion -- no collective inside the computation
    done = {u: convert_cell(u, cell) for u, cell in mine.items() if owner_of(u, world_size) == rank}
    # 3. gather
    out = {}
    for u in range(len(rects)):
        src = owner_of(u, world_size)
        if rank == root:
            out[u] = done[u] if src == root else recv_result(u, src)
        elif src == rank:
            send_result(u, done[u], root)
    if dist is not None and world_size > 1:
        dist.barrier()
    return out if rank == root else done
