"""avifImageScale: the device scaler against the reference's src/scale.c + vendored libyuv compiled from their
sources (oracle/_ref/libavif_ref.so).  Integer arithmetic throughout: the planes must match bit for bit, for every one of
libyuv's seven loops (vertical-only, point, bilinear down / up, the two 2x paths, box), 8- and 16-bit samples, every chroma
layout, with and without alpha."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import YUVImage

A = abi
libc = ctypes.CDLL(None)
libc.free.argtypes = [ctypes.c_void_p]


def make_image(w, h, depth, fmt, alpha, seed):
    rs = np.random.RandomState(seed)
    img = YUVImage(w, h, depth, fmt, A.AVIF_RANGE_FULL, 1, alpha=alpha)
    for p in img.present_planes():
        b = img.bufs[p]
        if depth > 8:
            b.set(rs.randint(0, 1 << depth, size=(b.rows, b.row_bytes // 2)).astype(np.uint16).view(np.uint8))
        else:
            b.set(rs.randint(0, 256, size=(b.rows, b.row_bytes)).astype(np.uint8))
    return img


def scaled_planes(lib, w, h, depth, fmt, alpha, seed, dw, dh):
    img = make_image(w, h, depth, fmt, alpha, seed)
    diag = A.avifDiagnostics()
    res = lib.avifImageScale(img.ptr(), dw, dh, ctypes.byref(diag))
    out = []
    c = img.c
    if res == 0 and (c.width, c.height) != (w, h) or (res == 0 and (dw, dh) != (w, h)):
        bps = 2 if depth > 8 else 1
        sx, sy = {A.AVIF_PIXEL_FORMAT_YUV444: (0, 0), A.AVIF_PIXEL_FORMAT_YUV422: (1, 0), A.AVIF_PIXEL_FORMAT_YUV420: (1, 1), A.AVIF_PIXEL_FORMAT_YUV400: (0, 0)}[fmt]
        for p in range(4):
            ptr = c.yuvPlanes[p] if p < 3 else c.alphaPlane
            rb = c.yuvRowBytes[p] if p < 3 else c.alphaRowBytes
            if not ptr:
                out.append(None)
                continue
            pw = (dw + sx) >> sx if p in (1, 2) else dw
            ph = (dh + sy) >> sy if p in (1, 2) else dh
            raw = np.ctypeslib.as_array((ctypes.c_uint8 * (rb * ph)).from_address(ptr)).reshape(ph, rb)[:, :pw * bps].copy()
            out.append(raw)
            libc.free(ptr)  # both sides allocate the new planes with malloc (avifAlloc == malloc)
    return res, out, (c.width, c.height), diag.error.decode(errors="replace")


def check(gpu, ref, w, h, depth, fmt, alpha, seed, dw, dh):
    ra, a, da, _ = scaled_planes(ref, w, h, depth, fmt, alpha, seed, dw, dh)
    rb, b, db, err = scaled_planes(gpu, w, h, depth, fmt, alpha, seed, dw, dh)
    assert ra == rb, (w, h, dw, dh, ra, rb, err)
    assert da == db
    if ra != 0:
        return
    assert len(a) == len(b)
    for p, (x, y) in enumerate(zip(a, b)):
        assert (x is None) == (y is None), (w, h, dw, dh, p)
        if x is not None:
            if fmt != A.AVIF_PIXEL_FORMAT_YUV400 or p in (0, 3):
                assert np.array_equal(x, y), (w, h, depth, fmt, dw, dh, p, int((x != y).sum()))


CASES = [
    # (src w, h) -> (dst w, h): one per loop of libyuv's dispatcher
    ((64, 64), (64, 48)), ((64, 31), (64, 100)), ((64, 1), (64, 9)),             # vertical only (down, up, one source row)
    ((100, 80), (33, 27)), ((128, 96), (32, 24)), ((150, 149), (7, 5)),          # box: two widths, one width, big boxes
    ((33, 21), (66, 42)), ((33, 21), (65, 41)), ((2, 2), (3, 3)),                # 2x bilinear (even / odd destination)
    ((40, 30), (80, 30)), ((40, 30), (79, 30)),                                  # 2x linear (height unchanged)
    ((40, 30), (57, 45)), ((40, 30), (41, 31)), ((3, 2), (100, 90)), ((40, 1), (57, 4)),  # bilinear up
    ((100, 80), (70, 50)), ((40, 80), (60, 50)), ((100, 80), (99, 79)),          # bilinear down (also: wider but shorter)
    ((90, 60), (30, 60)), ((1, 50), (20, 30)), ((1, 1), (5, 7)), ((90, 61), (30, 61)),   # point sampling
    ((17, 13), (17, 13)),                                                          # nothing to do
]


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [8, 10, 12])
def test_every_loop_matches_the_reference(gpu_product, reference, depth):
    for k, ((w, h), (dw, dh)) in enumerate(CASES):
        for fmt, alpha in ((A.AVIF_PIXEL_FORMAT_YUV444, True), (A.AVIF_PIXEL_FORMAT_YUV420, False), (A.AVIF_PIXEL_FORMAT_YUV400, True)):
            check(gpu_product, reference, w, h, depth, fmt, alpha, 1000 + k, dw, dh)


@pytest.mark.gpu
def test_random_size_pairs(gpu_product, reference):
    rs = np.random.RandomState(99)
    fmts = [A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV422, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400]
    for k in range(160):
        w, h = int(rs.randint(1, 140)), int(rs.randint(1, 140))
        kind = k % 4
        if kind == 0:
            dw, dh = int(rs.randint(1, 140)), int(rs.randint(1, 140))
        elif kind == 1:
            dw, dh = max(1, w // int(rs.randint(2, 6))), max(1, h // int(rs.randint(2, 6)))
        elif kind == 2:
            dw, dh = w * 2 - int(rs.randint(0, 2)), h * 2 - int(rs.randint(0, 2))
        else:
            dw, dh = w + int(rs.randint(0, 40)), max(1, h - int(rs.randint(0, 40)))
        check(gpu_product, reference, w, h, (8, 10, 12, 8)[k % 4], fmts[(k // 4) % 4], bool(k & 1), 2000 + k, max(1, dw), max(1, dh))


@pytest.mark.gpu
def test_larger_planes(gpu_product, reference):
    check(gpu_product, reference, 1920, 1080, 8, A.AVIF_PIXEL_FORMAT_YUV420, True, 5, 640, 360)
    check(gpu_product, reference, 640, 360, 10, A.AVIF_PIXEL_FORMAT_YUV444, False, 6, 1920, 1080)
    check(gpu_product, reference, 1000, 700, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 7, 2000, 1400)


def test_argument_checks_need_no_device(product, reference):
    for dw, dh in ((0, 10), (10, 0), (40000, 10), (20000, 20000)):
        ra, _, _, ea = scaled_planes(reference, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, dw, dh)
        rb, _, _, eb = scaled_planes(product, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, dw, dh)
        assert ra == rb and ra != 0, (dw, dh, ra, rb)
        assert ea and eb
    # same size: nothing happens, nothing is allocated
    ra, _, _, _ = scaled_planes(reference, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, 16, 16)
    rb, _, _, _ = scaled_planes(product, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, 16, 16)
    assert ra == rb == 0


@pytest.mark.gpu
def test_committed_vectors(gpu_product, oracle):
    """tests/golden/scale_vectors.json: hashes of the planes the compiled reference produced (tests/golden/make_golden.py).
    Runs where neither /root/reference nor oracle/_ref exists."""
    import json
    import os
    from tests import util
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scale_vectors.json")))["scale"]
    for v in vec:
        (w, h), (dw, dh) = v["src"], v["dst"]
        res, planes, _, err = scaled_planes(gpu_product, w, h, v["depth"], v["fmt"], v["alpha"], v["seed"], dw, dh)
        assert res == v["result"], (v, err)
        if res != 0 or (w, h) == (dw, dh):
            continue
        for p, want in enumerate(v["planes"]):
            if want is None:
                continue
            assert planes[p] is not None
            assert f"{util.fnv1a64(oracle, np.ascontiguousarray(planes[p]).reshape(-1)):016x}" == want, (v["src"], v["dst"], v["depth"], v["fmt"], p)
