"""Pins oracle/avif_oracle_tonemap.c (the CPU restatement of gain-map tone mapping) to the reference's src/gainmap.c compiled
from its sources: both run the host libm with the same operation order, so equality is BIT-EXACT -- every transfer function,
RGB layout, depth, half float, primaries conversion, weight sign, and the content-light-level statistics.  No GPU."""
import ctypes

import numpy as np
import pytest

from tests import test_gpu_gainmap as tg

A = tg.A
libc = ctypes.CDLL(None)
libc.free.argtypes = [ctypes.c_void_p]


def apply_any(lib, base, base_cp, base_tc, gm, headroom, out_cp, out_tc, out_depth, out_fmt, out_float=False):
    out = A.avifRGBImage()
    out.depth, out.format, out.isFloat = out_depth, out_fmt, int(out_float)
    clli = A.avifContentLightLevelInformationBox()
    diag = A.avifDiagnostics()
    res = lib.avifRGBImageApplyGainMap(base.ptr(), base_cp, base_tc, ctypes.byref(gm), ctypes.c_float(headroom), out_cp, out_tc,
                                       ctypes.byref(out), ctypes.byref(clli), ctypes.byref(diag))
    data = None
    if res == 0:
        n = out.rowBytes * out.height
        data = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(out.pixels)).copy().reshape(out.height, out.rowBytes)
    if out.pixels:
        libc.free(out.pixels)  # malloc on both sides
    return res, data, (clli.maxCLL, clli.maxPALL)


@pytest.mark.parametrize("tc_in", tg.TCS)
def test_every_transfer_function_bit_exact(oracle, reference, tc_in):
    w, h = 61, 37
    gm, keep = tg.make_gain_map(w, h, seed=300 + tc_in)
    base = tg.make_base(w, h, 10, A.AVIF_RGB_FORMAT_RGBA, seed=400 + tc_in)
    for tc_out in (tc_in, 13, 16, 18):
        for depth in (8, 12, 16):
            ra, a, ca = apply_any(reference, base, 1, tc_in, gm, 1.5, 9, tc_out, depth, A.AVIF_RGB_FORMAT_RGBA)
            rb, b, cb = apply_any(oracle, base, 1, tc_in, gm, 1.5, 9, tc_out, depth, A.AVIF_RGB_FORMAT_RGBA)
            assert ra == rb == 0
            assert np.array_equal(a, b), (tc_in, tc_out, depth, int((a != b).sum()))
            assert ca == cb


def test_layouts_depths_primaries_weights_bit_exact(oracle, reference):
    w, h = 40, 24
    fmts = [A.AVIF_RGB_FORMAT_RGB, A.AVIF_RGB_FORMAT_RGBA, A.AVIF_RGB_FORMAT_ARGB, A.AVIF_RGB_FORMAT_BGR, A.AVIF_RGB_FORMAT_BGRA,
            A.AVIF_RGB_FORMAT_ABGR, A.AVIF_RGB_FORMAT_GRAY, A.AVIF_RGB_FORMAT_GRAYA, A.AVIF_RGB_FORMAT_AGRAY]
    k = 0
    for bfmt in fmts:
        for bdepth, bfloat in ((8, False), (10, False), (12, False), (16, False), (16, True)):
            k += 1
            ofmt = fmts[(k * 5) % len(fmts)]
            if ofmt == A.AVIF_RGB_FORMAT_AGRAY:
                ofmt = A.AVIF_RGB_FORMAT_GRAYA  # half of an AGRAY output is never written (see tests/test_gpu_gainmap.py)
            odepth, ofloat = ((8, False), (10, False), (12, False), (16, False), (16, True))[(k * 3) % 5]
            base_cp, out_cp = ((1, 1), (1, 9), (12, 1), (9, 12))[k % 4]
            use_base, alt_cp = ((1, 2), (0, 9), (0, 12))[k % 3]
            headroom = (0.0, 0.7, 3.0, 5.0)[k % 4]
            gm_fmt = (A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400)[k % 3]
            gm, keep = tg.make_gain_map(w, h, depth=(8, 10)[k % 2], fmt=gm_fmt, seed=500 + k, gamma=((1, 1), (22, 10))[k % 2], gm_min=(-1, 2),
                                        gm_max=(5, 2), use_base=use_base, alt_primaries=alt_cp, base_h=((0, 1), (4, 1))[k % 2],
                                        alt_h=((3, 1), (1, 1))[k % 2])
            base = tg.make_base(w, h, bdepth, bfmt, is_float=bfloat, seed=600 + k, row_pad=(0, 24)[k % 2])
            ra, a, ca = apply_any(reference, base, base_cp, 13, gm, headroom, out_cp, 13, odepth, ofmt, ofloat)
            rb, b, cb = apply_any(oracle, base, base_cp, 13, gm, headroom, out_cp, 13, odepth, ofmt, ofloat)
            assert ra == rb, (k, ra, rb)
            if ra == 0:
                assert np.array_equal(a, b), (k, bfmt, bdepth, ofmt, odepth, ofloat, int((a != b).sum()))
                assert ca == cb


def test_rgb565_zero_weight_and_errors_bit_exact(oracle, reference):
    w, h = 24, 16
    gm, keep = tg.make_gain_map(w, h, seed=9)
    base = tg.make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGB_565, seed=10)
    for ofmt in (A.AVIF_RGB_FORMAT_RGB_565, A.AVIF_RGB_FORMAT_RGBA):
        ra, a, _ = apply_any(reference, base, 1, 13, gm, 2.0, 1, 13, 8, ofmt)
        rb, b, _ = apply_any(oracle, base, 1, 13, gm, 2.0, 1, 13, 8, ofmt)
        assert ra == rb == 0 and np.array_equal(a, b)
    gm0, keep0 = tg.make_gain_map(w, h, seed=11, base_h=(1, 1), alt_h=(1, 1))  # weight 0
    base10 = tg.make_base(w, h, 10, A.AVIF_RGB_FORMAT_RGBA, seed=12)
    for odepth, ofmt, ocp, otc in ((10, A.AVIF_RGB_FORMAT_RGBA, 1, 13), (8, A.AVIF_RGB_FORMAT_BGR, 1, 13), (12, A.AVIF_RGB_FORMAT_ARGB, 9, 16)):
        ra, a, _ = apply_any(reference, base10, 1, 13, gm0, 2.0, ocp, otc, odepth, ofmt)
        rb, b, _ = apply_any(oracle, base10, 1, 13, gm0, 2.0, ocp, otc, odepth, ofmt)
        assert ra == rb == 0 and np.array_equal(a, b), (odepth, ofmt)
    bad, keepb = tg.make_gain_map(w, h, seed=13, gm_min=(-2000, 1), gm_max=(2000, 1), base_off=(0, 1), alt_off=(0, 1))
    b8 = tg.make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, seed=14)
    ra, _, _ = apply_any(reference, b8, 1, 13, bad, 3.0, 9, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
    rb, _, _ = apply_any(oracle, b8, 1, 13, bad, 3.0, 9, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
    assert ra == rb


def test_rescaled_gain_maps_bit_exact(oracle, reference):
    """src/gainmap.c:186-201: gain maps of another size are brought to the base image's size by avifImageScale first."""
    w, h = 64, 40
    base = tg.make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, seed=21)
    for gw, gh, fmt, depth in ((32, 20, A.AVIF_PIXEL_FORMAT_YUV444, 8), (16, 10, A.AVIF_PIXEL_FORMAT_YUV420, 10), (150, 99, A.AVIF_PIXEL_FORMAT_YUV400, 8),
                               (64, 13, A.AVIF_PIXEL_FORMAT_YUV444, 12), (21, 40, A.AVIF_PIXEL_FORMAT_YUV422, 8)):
        gm, keep = tg.make_gain_map(gw, gh, depth=depth, fmt=fmt, seed=gw * 5 + gh)
        ra, a, ca = apply_any(reference, base, 1, 13, gm, 2.0, 9, 16, 10, A.AVIF_RGB_FORMAT_RGBA)
        rb, b, cb = apply_any(oracle, base, 1, 13, gm, 2.0, 9, 16, 10, A.AVIF_RGB_FORMAT_RGBA)
        assert ra == rb == 0, (gw, gh, ra, rb)
        assert np.array_equal(a, b) and ca == cb, (gw, gh, fmt)
