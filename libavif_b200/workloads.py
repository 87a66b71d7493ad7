"""BASELINE.json's configurations as data, shared by bench.py, the profiling tools, the golden-fixture
generator and the full-size parity tests.  No arithmetic happens here: these are image descriptions, a seeded input
generator (numpy RandomState: a frozen stream) and a byte digest used to compare outputs with the fixtures the compiled
reference produced (tests/golden/bench_vectors.json, written by tests/golden/make_golden.py).
"""
import hashlib

import numpy as np

from . import abi
from .images import RGBImage, YUVImage

A = abi

# name: image side, rgb side, algorithmic bytes per luma pixel (each sample read once, each output byte written once)
CONFIGS = {
    # C1: 1080p 4:2:0 8-bit BT.601 limited -> RGBA8 (default upsampling = bilinear without libyuv; nearest variant)
    "C1": dict(img=dict(width=1920, height=1080, depth=8, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT601),
               rgb=dict(depth=8, fmt=A.AVIF_RGB_FORMAT_RGBA, upsampling=A.AVIF_CHROMA_UPSAMPLING_BILINEAR), bytes_per_px=5.5, seed=1),
    "C1_nearest": dict(img=dict(width=1920, height=1080, depth=8, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT601),
                       rgb=dict(depth=8, fmt=A.AVIF_RGB_FORMAT_RGBA, upsampling=A.AVIF_CHROMA_UPSAMPLING_NEAREST), bytes_per_px=5.5, seed=1),
    # C2: 4K 4:4:4 10-bit BT.709 full + alpha (straight) <-> RGBA10 premultiplied: the headline round trip
    "C2": dict(img=dict(width=3840, height=2160, depth=10, yuv_format=A.AVIF_PIXEL_FORMAT_YUV444, yuv_range=A.AVIF_RANGE_FULL, mc=A.MC_BT709,
                        alpha=True, alpha_premultiplied=False),
               rgb=dict(depth=10, fmt=A.AVIF_RGB_FORMAT_RGBA, alpha_premultiplied=True), bytes_per_px=16, back=True, seed=2),
    # C3: 8K 4:2:0 12-bit BT.2020 limited + alpha (premultiplied) -> RGBA12 straight, bilinear (in-float unmultiply)
    "C3": dict(img=dict(width=7680, height=4320, depth=12, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT2020_NCL,
                        alpha=True, alpha_premultiplied=True),
               rgb=dict(depth=12, fmt=A.AVIF_RGB_FORMAT_RGBA, upsampling=A.AVIF_CHROMA_UPSAMPLING_BILINEAR, alpha_premultiplied=False),
               bytes_per_px=13, back=True, seed=3),
    # NS: the north-star target, 8K 4:4:4 10-bit -> RGBA10 (avifRGBImageSetDefaults)
    "NS": dict(img=dict(width=7680, height=4320, depth=10, yuv_format=A.AVIF_PIXEL_FORMAT_YUV444, yuv_range=A.AVIF_RANGE_FULL, mc=A.MC_BT709),
               rgb=dict(depth=10, fmt=A.AVIF_RGB_FORMAT_RGBA), bytes_per_px=14, seed=7),
    # C5: 4K 4:2:0 10-bit BT.709 limited -> RGB16, default (bilinear) and nearest
    "C5": dict(img=dict(width=3840, height=2160, depth=10, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT709),
               rgb=dict(depth=16, fmt=A.AVIF_RGB_FORMAT_RGB), bytes_per_px=9, seed=5),
    "C5_nearest": dict(img=dict(width=3840, height=2160, depth=10, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT709),
                       rgb=dict(depth=16, fmt=A.AVIF_RGB_FORMAT_RGB, upsampling=A.AVIF_CHROMA_UPSAMPLING_NEAREST), bytes_per_px=9, seed=5),
    # C4 cell: 1080p RGBA8 -> 4:2:0 8-bit BT.601 full + alpha plane (RGB -> YUV direction only)
    "C4_cell": dict(img=dict(width=1920, height=1080, depth=8, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_FULL, mc=A.MC_BT601, alpha=True),
                    rgb=dict(depth=8, fmt=A.AVIF_RGB_FORMAT_RGBA), bytes_per_px=6.5, back_only=True, seed=4),
    # Android's usual output (android_jni/avifandroidjni/src/main/jni/libavif_jni.cc:209-223): RGBA_F16 premultiplied from 4:2:0 8-bit + alpha
    "F16": dict(img=dict(width=3840, height=2160, depth=8, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT601, alpha=True),
                rgb=dict(depth=16, fmt=A.AVIF_RGB_FORMAT_RGBA, is_float=True, alpha_premultiplied=True), bytes_per_px=2.5 + 8, seed=8),
    "RGB565": dict(img=dict(width=3840, height=2160, depth=8, yuv_format=A.AVIF_PIXEL_FORMAT_YUV420, yuv_range=A.AVIF_RANGE_LIMITED, mc=A.MC_BT601),
                   rgb=dict(depth=8, fmt=A.AVIF_RGB_FORMAT_RGB_565), bytes_per_px=1.5 + 2, seed=9),
}


def dims(name):
    c = CONFIGS[name]["img"]
    return c["width"], c["height"]


def make_yuv(name, device=None, pinned=False, **over):
    kw = dict(CONFIGS[name]["img"])
    kw.update(over)
    return YUVImage(device=device, pinned=pinned, **kw)


def make_rgb(name, device=None, pinned=False, **over):
    c = CONFIGS[name]
    kw = dict(c["rgb"])
    kw.update(over)
    return RGBImage(c["img"]["width"], c["img"]["height"], device=device, pinned=pinned, **kw)


def seeded_plane(rows, row_bytes, bps, depth, seed):
    """(rows, row_bytes) uint8 array of uniform codes in [0, 2^depth): numpy RandomState(seed), one draw per sample."""
    rs = np.random.RandomState(seed)
    n = rows * row_bytes // bps
    if bps == 2:
        return rs.randint(0, 1 << depth, size=n, dtype=np.uint16).view(np.uint8).reshape(rows, row_bytes)
    return rs.randint(0, 1 << depth, size=n, dtype=np.uint8).reshape(rows, row_bytes)


def fill_yuv_seeded(img, seed):
    """Plane p of the image gets seeded_plane(seed * 16 + p)."""
    for p in img.present_planes():
        b = img.bufs[p]
        b.set(seeded_plane(b.rows, b.row_bytes, img.bps, img.c.depth, seed * 16 + p))


def fill_rgb_seeded(rgb, seed):
    bps = 2 if rgb.c.depth > 8 else 1
    rgb.buf.set(seeded_plane(rgb.buf.rows, rgb.buf.row_bytes, bps, rgb.c.depth, seed * 16 + 9))


def digest(*arrays):
    """blake2b-128 over the concatenated bytes (a checksum that runs at memory speed; FNV is byte-serial)."""
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is not None:
            h.update(np.ascontiguousarray(a).reshape(-1).view(np.uint8).data)
    return h.hexdigest()


def yuv_digest(img):
    return digest(*[img.bufs[p].host() for p in img.present_planes()])


def rgb_digest(rgb):
    return digest(rgb.buf.host())
