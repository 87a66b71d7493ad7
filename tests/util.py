"""Shared machinery of the parity tests: conversion specs, seeded inputs, runners for any of the three
ABI-compatible libraries (product / oracle / compiled reference)."""
import ctypes
import dataclasses
import itertools
import random

import numpy as np

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage

FNV_OFFSET = 0xcbf29ce484222325
FILL_BYTE = 0xA5  # destination buffers start as this pattern so that untouched bytes compare equal

YUV_FORMATS = [abi.AVIF_PIXEL_FORMAT_YUV444, abi.AVIF_PIXEL_FORMAT_YUV422, abi.AVIF_PIXEL_FORMAT_YUV420, abi.AVIF_PIXEL_FORMAT_YUV400]
FMT_NAME = {1: "444", 2: "422", 3: "420", 4: "400"}
DEPTHS = [8, 10, 12, 16]
COLOR_RGB_FORMATS = [abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_BGR,
                     abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ABGR]
GRAY_RGB_FORMATS = [abi.AVIF_RGB_FORMAT_GRAY, abi.AVIF_RGB_FORMAT_GRAYA, abi.AVIF_RGB_FORMAT_AGRAY]
ALL_RGB_FORMATS = COLOR_RGB_FORMATS + [abi.AVIF_RGB_FORMAT_RGB_565] + GRAY_RGB_FORMATS
COEFF_MCS = [abi.MC_BT709, abi.MC_UNSPECIFIED, abi.MC_FCC, abi.MC_BT470BG, abi.MC_BT601, abi.MC_SMPTE240, abi.MC_BT2020_NCL,
             abi.MC_CHROMA_DERIVED_NCL]


@dataclasses.dataclass(frozen=True)
class Spec:
    w: int
    h: int
    depth: int = 8
    fmt: int = abi.AVIF_PIXEL_FORMAT_YUV444
    rng: int = abi.AVIF_RANGE_FULL
    mc: int = abi.MC_BT601
    cp: int = 1
    alpha: bool = False            # image has an alpha plane
    img_premult: bool = False
    rgb_depth: int = 8
    rgb_fmt: int = abi.AVIF_RGB_FORMAT_RGBA
    ups: int = abi.AVIF_CHROMA_UPSAMPLING_AUTOMATIC
    downs: int = abi.AVIF_CHROMA_DOWNSAMPLING_AUTOMATIC
    ignore_alpha: bool = False
    rgb_premult: bool = False
    is_float: bool = False
    max_threads: int = 1
    yuv_pad: int = 0               # extra bytes per plane row
    rgb_pad: int = 0
    yuv_off: int = 0               # byte offset of every plane start inside its allocation (alignment tests)
    rgb_off: int = 0
    garbage: bool = False          # 16-bit containers carry bits above the nominal depth

    def ident(self):
        return (f"{self.w}x{self.h}_{FMT_NAME.get(self.fmt, self.fmt)}_d{self.depth}_r{self.rng}_mc{self.mc}"
                f"_a{int(self.alpha)}{int(self.img_premult)}__{abi.RGB_FORMAT_NAMES[self.rgb_fmt] if 0 <= self.rgb_fmt < 10 else self.rgb_fmt}"
                f"_d{self.rgb_depth}_u{self.ups}_i{int(self.ignore_alpha)}p{int(self.rgb_premult)}f{int(self.is_float)}"
                f"_pad{self.yuv_pad}.{self.rgb_pad}_off{self.yuv_off}.{self.rgb_off}_g{int(self.garbage)}")


def fnv1a64(oracle, arr, h=FNV_OFFSET):
    a = np.ascontiguousarray(arr)
    return oracle.raw.oracleFnv1a64(h, a.ctypes.data, a.size)


def xorshift_fill_image(oracle, img, seed):
    """SURVEY.md Appendix D generator: planes Y,U,V,A in order, sample = next() & mask."""
    state = seed
    mask = (1 << img.c.depth) - 1
    for p in img.present_planes():
        w, h = img.plane_dims(p)
        buf = img.bufs[p]
        assert buf.device is None and buf.row_bytes == w * img.bps
        state = oracle.raw.oracleXorshiftFill(state, buf.ptr, w * h, mask, img.bps)
    return state


def make_yuv(spec, seed, device=None, pinned=False, with_alpha=None):
    alpha = spec.alpha if with_alpha is None else with_alpha
    img = YUVImage(spec.w, spec.h, spec.depth, spec.fmt, spec.rng, spec.mc, alpha=alpha, alpha_premultiplied=spec.img_premult,
                   color_primaries=spec.cp, device=device, pinned=pinned, row_pad=spec.yuv_pad, offset=spec.yuv_off)
    rs = np.random.RandomState(seed)
    maxv = (1 << spec.depth) - 1
    for p in img.present_planes():
        w, h = img.plane_dims(p)
        kind = rs.randint(0, 4)
        if kind == 0:    # uniform noise
            s = rs.randint(0, maxv + 1, size=(h, w))
        elif kind == 1:  # extremes: exercises clamps and the alpha 0 / opaque branches
            s = rs.choice([0, 1, maxv // 2, maxv - 1, maxv], size=(h, w))
        elif kind == 2:  # smooth gradient
            s = (np.add.outer(np.arange(h), np.arange(w)) * max(1, maxv // max(1, (w + h)))) % (maxv + 1)
        else:
            s = rs.randint(0, maxv + 1, size=(h, w))
            s[rs.rand(h, w) < 0.3] = maxv if p == 3 else rs.randint(0, maxv + 1)
        s = s.astype(np.uint32)
        if spec.garbage and spec.depth in (10, 12):
            s = s | (rs.randint(0, 1 << (16 - spec.depth), size=(h, w)).astype(np.uint32) << spec.depth) * (rs.rand(h, w) < 0.2)
        full = np.full((h, img.bufs[p].row_bytes), FILL_BYTE, dtype=np.uint8)
        dt = np.uint16 if img.bps == 2 else np.uint8
        full[:, :w * img.bps] = s.astype(dt).view(np.uint8).reshape(h, w * img.bps)
        img.bufs[p].set(full)
    return img


def make_rgb_dst(spec, device=None, pinned=False, w=None, h=None):
    rgb = RGBImage(spec.w if w is None else w, spec.h if h is None else h, spec.rgb_depth, spec.rgb_fmt, upsampling=spec.ups,
                   downsampling=spec.downs, ignore_alpha=spec.ignore_alpha, alpha_premultiplied=spec.rgb_premult,
                   is_float=spec.is_float, max_threads=spec.max_threads, device=device, pinned=pinned, row_pad=spec.rgb_pad,
                   offset=spec.rgb_off)
    rgb.buf.set(np.full((rgb.buf.rows, rgb.buf.row_bytes), FILL_BYTE, dtype=np.uint8))
    return rgb


def make_rgb_src(spec, seed, device=None, pinned=False):
    rgb = RGBImage(spec.w, spec.h, spec.rgb_depth, spec.rgb_fmt, upsampling=spec.ups, downsampling=spec.downs,
                   ignore_alpha=spec.ignore_alpha, alpha_premultiplied=spec.rgb_premult, is_float=spec.is_float,
                   max_threads=spec.max_threads, device=device, pinned=pinned, row_pad=spec.rgb_pad, offset=spec.rgb_off)
    rs = np.random.RandomState(seed)
    nch = rgb.pixel_bytes // (2 if spec.rgb_depth > 8 else 1)
    maxv = (1 << spec.rgb_depth) - 1
    kind = rs.randint(0, 3)
    if kind == 0:
        s = rs.randint(0, maxv + 1, size=(spec.h, spec.w, nch))
    elif kind == 1:
        s = rs.choice([0, 1, maxv // 2, maxv - 1, maxv], size=(spec.h, spec.w, nch))
    else:
        s = rs.randint(0, maxv + 1, size=(spec.h, spec.w, nch))
        s[rs.rand(spec.h, spec.w) < 0.3] = maxv
    if spec.garbage and spec.rgb_depth in (10, 12):
        s = s | ((rs.randint(0, 1 << (16 - spec.rgb_depth), size=s.shape)) << spec.rgb_depth) * (rs.rand(*s.shape) < 0.2)
    dt = np.uint16 if spec.rgb_depth > 8 else np.uint8
    full = np.full((spec.h, rgb.buf.row_bytes), FILL_BYTE, dtype=np.uint8)
    full[:, :spec.w * rgb.pixel_bytes] = s.astype(dt).reshape(spec.h, -1).view(np.uint8)
    rgb.buf.set(full)
    return rgb


def make_yuv_dst(spec, device=None, pinned=False, preallocate=True, with_alpha=None):
    alpha = spec.alpha if with_alpha is None else with_alpha
    img = YUVImage(spec.w, spec.h, spec.depth, spec.fmt, spec.rng, spec.mc, alpha=alpha, alpha_premultiplied=spec.img_premult,
                   color_primaries=spec.cp, device=device, pinned=pinned, row_pad=spec.yuv_pad, offset=spec.yuv_off,
                   allocate=preallocate)
    if preallocate:
        for p in img.present_planes():
            b = img.bufs[p]
            b.set(np.full((b.rows, b.row_bytes), FILL_BYTE, dtype=np.uint8))
    return img


def yuv_to_rgb(lib, spec, seed, device=None, pinned=False):
    """Run avifImageYUVToRGB of `lib` on seeded inputs; returns (result code, every byte of the rgb allocation)."""
    img = make_yuv(spec, seed, device=device, pinned=pinned)
    rgb = make_rgb_dst(spec, device=device, pinned=pinned)
    res = lib.avifImageYUVToRGB(img.ptr(), rgb.ptr())
    return res, rgb.full_bytes().copy()


def rgb_to_yuv(lib, spec, seed, device=None, pinned=False, with_alpha_plane=None):
    """Run avifImageRGBToYUV on seeded inputs into pre-allocated planes; returns (code, [plane bytes...])."""
    rgb = make_rgb_src(spec, seed, device=device, pinned=pinned)
    want_alpha = abi.rgb_format_has_alpha(spec.rgb_fmt) and not spec.ignore_alpha
    img = make_yuv_dst(spec, device=device, pinned=pinned, with_alpha=(want_alpha if with_alpha_plane is None else with_alpha_plane))
    res = lib.avifImageRGBToYUV(img.ptr(), rgb.ptr())
    planes = [img.bufs[p].host().copy().reshape(-1) if img.bufs[p] is not None else None for p in range(4)]
    return res, planes


def first_mismatch(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    d = np.nonzero(a != b)[0]
    if d.size == 0:
        return None
    k = int(d[0])
    return f"{d.size} bytes differ; first at byte {k}: expected {int(a[k])} got {int(b[k])}"


def random_y2r_specs(n, seed, max_dim=70, sizes=None):
    """Seeded sample of the YUV->RGB parameter space of SURVEY.md Appendix B (only combinations the reference accepts
    are not filtered here: rejected ones must be rejected identically)."""
    r = random.Random(seed)
    out = []
    for _ in range(n):
        w, h = (r.choice(sizes) if sizes else (r.randint(1, max_dim), r.randint(1, max_dim)))
        depth = r.choice(DEPTHS)
        fmt = r.choice(YUV_FORMATS)
        mode = r.random()
        rgb_depth = r.choice(DEPTHS)
        rng = r.choice([0, 1])
        if mode < 0.70:
            mc = r.choice(COEFF_MCS)
        elif mode < 0.80:
            mc = abi.MC_IDENTITY
            fmt = r.choice([abi.AVIF_PIXEL_FORMAT_YUV444, abi.AVIF_PIXEL_FORMAT_YUV400])
            if r.random() < 0.5:
                depth = rgb_depth = 8
                rng = 1
        elif mode < 0.88:
            mc, rng = abi.MC_YCGCO, 1
        elif mode < 0.94:
            mc, rng = abi.MC_YCGCO_RE, 1
            depth = r.choice([10, 12])
            rgb_depth = depth - 2
        else:
            mc, rng = abi.MC_YCGCO_RO, 1
            depth = 16 if r.random() < 0.2 else r.choice([10, 12])
            rgb_depth = depth - 1 if depth == 16 else depth - 2  # depth-1 is only legal at 9/11/15 -> mostly rejected on purpose
        rgb_fmt = r.choice(ALL_RGB_FORMATS)
        if rgb_fmt == abi.AVIF_RGB_FORMAT_RGB_565 and r.random() < 0.9:
            rgb_depth = 8
            if mc == abi.MC_YCGCO_RE:
                depth = 10
        is_float = r.random() < 0.12
        if is_float and r.random() < 0.9:
            rgb_depth = 16
        out.append(Spec(
            w=w, h=h, depth=depth, fmt=fmt, rng=rng, mc=mc, cp=r.choice([1, 4, 5, 6, 9, 10, 11, 12, 22, 2]),
            alpha=r.random() < 0.6, img_premult=r.random() < 0.5, rgb_depth=rgb_depth, rgb_fmt=rgb_fmt,
            ups=r.choice([0, 1, 2, 3, 4]), ignore_alpha=r.random() < 0.25, rgb_premult=r.random() < 0.5, is_float=is_float,
            yuv_pad=r.choice([0, 0, 2, 6]), rgb_pad=r.choice([0, 0, 4, 10]), garbage=r.random() < 0.3))
    return out


def random_r2y_specs(n, seed, max_dim=70, sizes=None):
    r = random.Random(seed)
    out = []
    for _ in range(n):
        w, h = (r.choice(sizes) if sizes else (r.randint(1, max_dim), r.randint(1, max_dim)))
        depth = r.choice(DEPTHS)
        rgb_depth = r.choice(DEPTHS)
        fmt = r.choice(YUV_FORMATS)
        rng = r.choice([0, 1])
        mode = r.random()
        if mode < 0.70:
            mc = r.choice(COEFF_MCS)
        elif mode < 0.80:
            mc = abi.MC_IDENTITY
            fmt = r.choice([abi.AVIF_PIXEL_FORMAT_YUV444, abi.AVIF_PIXEL_FORMAT_YUV400])
        elif mode < 0.88:
            mc, rng = abi.MC_YCGCO, 1
        elif mode < 0.95:
            mc, rng = abi.MC_YCGCO_RE, 1
            depth = r.choice([10, 12])
            rgb_depth = depth - 2
        else:
            mc, rng = abi.MC_YCGCO_RO, 1
            depth, rgb_depth = 16, 15 if r.random() < 0.5 else 16
        rgb_fmt = r.choice(COLOR_RGB_FORMATS + GRAY_RGB_FORMATS)
        out.append(Spec(
            w=w, h=h, depth=depth, fmt=fmt, rng=rng, mc=mc, cp=r.choice([1, 4, 5, 6, 9, 10, 11, 12, 22, 2]),
            img_premult=r.random() < 0.5, rgb_depth=rgb_depth, rgb_fmt=rgb_fmt, downs=r.choice([0, 1, 2, 3, 3, 4]),
            ignore_alpha=r.random() < 0.25, rgb_premult=r.random() < 0.5, yuv_pad=r.choice([0, 0, 2, 6]),
            rgb_pad=r.choice([0, 0, 4, 10]), garbage=r.random() < 0.3))
    return out
