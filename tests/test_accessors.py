"""the reference's internal colour-space descriptions and per-pixel float accessors (src/reformat.c:32-159,
:1856-1939), exported by the product as host-side scalar helpers -- against the compiled reference, bit for bit.  No GPU involved."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage

A = abi


def _bind(lib):
    raw = getattr(lib, "raw", lib)
    raw.avifGetRGBColorSpaceInfo.restype = ctypes.c_int
    raw.avifGetRGBColorSpaceInfo.argtypes = [ctypes.POINTER(A.avifRGBImage), ctypes.POINTER(A.avifRGBColorSpaceInfo)]
    raw.avifGetYUVColorSpaceInfo.restype = ctypes.c_int
    raw.avifGetYUVColorSpaceInfo.argtypes = [ctypes.POINTER(A.avifImage), ctypes.POINTER(A.avifYUVColorSpaceInfo)]
    for n in ("avifGetRGBAPixel", "avifSetRGBAPixel"):
        f = getattr(raw, n)
        f.restype = None
        f.argtypes = [ctypes.POINTER(A.avifRGBImage), ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(A.avifRGBColorSpaceInfo), ctypes.POINTER(ctypes.c_float * 4)]
    return raw


def _fields(s, skip=()):
    return {n: (getattr(s, n) if not isinstance(getattr(s, n), ctypes.Structure) else tuple(getattr(getattr(s, n), m) for m, _ in getattr(s, n)._fields_))
            for n, _ in s._fields_ if n not in skip}


def test_rgb_color_space_info(product, reference):
    p, r = _bind(product), _bind(reference)
    for fmt in list(range(A.AVIF_RGB_FORMAT_COUNT)) + [A.AVIF_RGB_FORMAT_COUNT, 99, -1]:
        for depth in (8, 10, 12, 16, 9, 0):
            for is_float in (0, 1):
                rgb = A.avifRGBImage()
                rgb.format, rgb.depth, rgb.isFloat = fmt, depth, is_float
                ip, ir = A.avifRGBColorSpaceInfo(), A.avifRGBColorSpaceInfo()
                ok_p, ok_r = p.avifGetRGBColorSpaceInfo(ctypes.byref(rgb), ctypes.byref(ip)), r.avifGetRGBColorSpaceInfo(ctypes.byref(rgb), ctypes.byref(ir))
                assert bool(ok_p) == bool(ok_r), (fmt, depth, is_float)
                if ok_r:
                    assert _fields(ip) == _fields(ir), (fmt, depth, is_float)


def test_yuv_color_space_info(product, reference):
    p, r = _bind(product), _bind(reference)
    for mc in range(0, 20):
        for cp in (1, 2, 9, 12):
            for depth in (8, 10, 12, 16, 11):
                for fmt in range(0, 6):
                    for rng in (0, 1, 2):
                        img = A.avifImage()
                        img.depth, img.yuvFormat, img.yuvRange, img.matrixCoefficients, img.colorPrimaries = depth, fmt, rng, mc, cp
                        ip, ir = A.avifYUVColorSpaceInfo(), A.avifYUVColorSpaceInfo()
                        ip.mode = ir.mode = 77  # neither implementation assigns it
                        ok_p, ok_r = p.avifGetYUVColorSpaceInfo(ctypes.byref(img), ctypes.byref(ip)), r.avifGetYUVColorSpaceInfo(ctypes.byref(img), ctypes.byref(ir))
                        assert bool(ok_p) == bool(ok_r), (mc, depth, fmt, rng)
                        if ok_r:
                            a, b = _fields(ip), _fields(ir)
                            for k in ("kr", "kg", "kb"):  # bit patterns, not approximate equality
                                assert np.float32(a[k]).tobytes() == np.float32(b[k]).tobytes(), (mc, cp, k)
                            assert a == b, (mc, depth, fmt, rng)


@pytest.mark.parametrize("fmt", range(A.AVIF_RGB_FORMAT_COUNT))
def test_get_and_set_pixel(product, reference, fmt):
    p, r = _bind(product), _bind(reference)
    rs = np.random.RandomState(fmt)
    for depth, is_float in ((8, 0), (10, 0), (12, 0), (16, 0), (16, 1)):
        if fmt == A.AVIF_RGB_FORMAT_RGB_565 and depth != 8:
            continue
        w, h = 37, 5
        src = RGBImage(w, h, depth, fmt, is_float=bool(is_float), row_pad=6)
        raw = rs.randint(0, 256, size=(src.buf.rows, src.buf.row_bytes)).astype(np.uint8)
        if depth > 8 and not is_float:  # codes within the depth, plus a few above it (containers may hold garbage)
            v = raw.view(np.uint16) & ((1 << depth) - 1 if rs.rand() < 0.8 else 0xFFFF)
            raw = v.view(np.uint8)
        if is_float:  # finite halves only
            v = raw.view(np.uint16)
            v[(v & 0x7C00) == 0x7C00] = 0x3C00
            raw = v.view(np.uint8)
        src.buf.set(raw)
        info = A.avifRGBColorSpaceInfo()
        assert r.avifGetRGBColorSpaceInfo(src.ptr(), ctypes.byref(info))
        dst_p = RGBImage(w, h, depth, fmt, is_float=bool(is_float), row_pad=6)
        dst_r = RGBImage(w, h, depth, fmt, is_float=bool(is_float), row_pad=6)
        for y in range(h):
            for x in range(w):
                a, b = (ctypes.c_float * 4)(), (ctypes.c_float * 4)()
                p.avifGetRGBAPixel(src.ptr(), x, y, ctypes.byref(info), ctypes.byref(a))
                r.avifGetRGBAPixel(src.ptr(), x, y, ctypes.byref(info), ctypes.byref(b))
                assert bytes(a) == bytes(b), (fmt, depth, is_float, x, y, list(a), list(b))
                px = (ctypes.c_float * 4)(*rs.rand(4).astype(np.float32))
                if (x + y) % 7 == 0:
                    px = (ctypes.c_float * 4)(*rs.choice([0.0, 1.0, 0.5], size=4).astype(np.float32))
                p.avifSetRGBAPixel(dst_p.ptr(), x, y, ctypes.byref(info), ctypes.byref(px))
                r.avifSetRGBAPixel(dst_r.ptr(), x, y, ctypes.byref(info), ctypes.byref(px))
        assert np.array_equal(dst_p.full_bytes(), dst_r.full_bytes()), (fmt, depth, is_float)
