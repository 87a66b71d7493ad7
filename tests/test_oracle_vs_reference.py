"""Pins the oracle (oracle/avif_oracle.c) against the UNMODIFIED reference compiled from its own sources
(oracle/_ref/libavif_ref.so): byte-for-byte equality of outputs and result codes over a seeded sweep of the whole
parameter space of SURVEY.md Appendix B.  Runs wherever _ref was built (skipped otherwise)."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from tests import util


def test_y2r_sweep(oracle, reference):
    n_ok = 0
    for k, spec in enumerate(util.random_y2r_specs(1500, seed=11)):
        ro, bo = util.yuv_to_rgb(oracle, spec, seed=1000 + k)
        rr, br = util.yuv_to_rgb(reference, spec, seed=1000 + k)
        assert ro == rr, f"result code {ro} vs reference {rr} for {spec.ident()}"
        if rr == abi.AVIF_RESULT_OK:
            n_ok += 1
            assert util.first_mismatch(br, bo) is None, f"{spec.ident()}: {util.first_mismatch(br, bo)}"
    assert n_ok > 1000


def test_r2y_sweep(oracle, reference):
    n_ok = 0
    for k, spec in enumerate(util.random_r2y_specs(1200, seed=12)):
        ro, po = util.rgb_to_yuv(oracle, spec, seed=2000 + k)
        rr, pr = util.rgb_to_yuv(reference, spec, seed=2000 + k)
        assert ro == rr, f"result code {ro} vs reference {rr} for {spec.ident()}"
        if rr == abi.AVIF_RESULT_OK:
            n_ok += 1
            for p in range(4):
                assert (po[p] is None) == (pr[p] is None)
                if pr[p] is not None:
                    assert util.first_mismatch(pr[p], po[p]) is None, f"{spec.ident()} plane {p}: {util.first_mismatch(pr[p], po[p])}"
    assert n_ok > 800


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
@pytest.mark.parametrize("fmt", [abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_ABGR,
                                 abi.AVIF_RGB_FORMAT_GRAYA, abi.AVIF_RGB_FORMAT_AGRAY, abi.AVIF_RGB_FORMAT_RGB])
def test_premultiply_roundtrip_codes(oracle, reference, depth, fmt):
    for un in (0, 1):
        spec = util.Spec(w=61, h=9, rgb_depth=depth, rgb_fmt=fmt, rgb_pad=6, garbage=True)
        outs = []
        for lib in (oracle, reference):
            rgb = util.make_rgb_src(spec, seed=77 + depth)
            fn = lib.avifRGBImageUnpremultiplyAlpha if un else lib.avifRGBImagePremultiplyAlpha
            outs.append((fn(rgb.ptr()), rgb.full_bytes().copy()))
        assert outs[0][0] == outs[1][0]
        assert util.first_mismatch(outs[1][1], outs[0][1]) is None


def test_coefficients_bits(oracle, reference):
    for mc in range(0, 20):
        for cp in [0, 1, 2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 22, 99]:
            img = abi.avifImage()
            img.matrixCoefficients, img.colorPrimaries = mc, cp
            a = [ctypes.c_float() for _ in range(3)]
            b = [ctypes.c_float() for _ in range(3)]
            oracle.raw.oracleCalcYUVCoefficients(ctypes.byref(img), *[ctypes.byref(x) for x in a])
            reference.raw.avifCalcYUVCoefficients(ctypes.byref(img), *[ctypes.byref(x) for x in b])
            fa = np.array([x.value for x in a], dtype=np.float32).view(np.uint32)
            fb = np.array([x.value for x in b], dtype=np.float32).view(np.uint32)
            assert (fa == fb).all(), (mc, cp, fa, fb)


def test_limited_full_exhaustive(oracle, reference):
    for depth in (8, 10, 12, 16, 9):
        for v in range(-5, (1 << min(depth, 12)) + 5):
            for name in ("avifLimitedToFullY", "avifLimitedToFullUV", "avifFullToLimitedY", "avifFullToLimitedUV"):
                assert getattr(oracle, name)(depth, v) == getattr(reference, name)(depth, v)
