"""Host-side logic of the shim that runs before any CUDA call: argument validation and result codes (the contract of the
reference's avif_fuzztest_yuvrgb.cc:20-54 and of src/reformat.c / src/alpha.c), helper functions, integer range helpers.
Compared with the oracle; needs no GPU."""
import ctypes

import pytest

from libavif_b200 import abi
from tests import util


def _code_y2r(lib, spec, max_threads=1, null_y=False, null_pixels=False):
    img = util.make_yuv(spec, 1)
    rgb = util.make_rgb_dst(spec)
    rgb.c.maxThreads = max_threads
    if null_y:
        img.c.yuvPlanes[0] = None
    if null_pixels:
        rgb.c.pixels = None
    return lib.avifImageYUVToRGB(img.ptr(), rgb.ptr())


def test_rejections_match_oracle(product, oracle):
    """Every spec the oracle rejects must be rejected with the same code by the product (accepted ones need a GPU)."""
    n = 0
    for k, spec in enumerate(util.random_y2r_specs(3000, seed=91)):
        ro = _code_y2r(oracle, spec)
        if ro != abi.AVIF_RESULT_OK:
            n += 1
            assert _code_y2r(product, spec) == ro, spec.ident()
    for k, spec in enumerate(util.random_r2y_specs(3000, seed=92)):
        ro, _ = util.rgb_to_yuv(oracle, spec, 1)
        if ro != abi.AVIF_RESULT_OK:
            n += 1
            rp, _ = util.rgb_to_yuv(product, spec, 1)
            assert rp == ro, spec.ident()
    assert n > 300


def test_fuzz_contract_codes(product):
    A = abi
    base = dict(w=8, h=4)
    # 565 needs depth 8; float needs depth 16; negative maxThreads; missing Y plane (avif_fuzztest_yuvrgb.cc:30-44)
    assert _code_y2r(product, util.Spec(rgb_fmt=A.AVIF_RGB_FORMAT_RGB_565, rgb_depth=10, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(is_float=True, rgb_depth=8, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(**base), max_threads=-1) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(**base), null_y=True) == A.AVIF_RESULT_REFORMAT_FAILED
    # unsupported matrices (reformat.c:129-144)
    for mc in (3, A.MC_BT2020_CL, A.MC_SMPTE2085, A.MC_CHROMA_DERIVED_CL, A.MC_ICTCP, 18, 200):
        assert _code_y2r(product, util.Spec(mc=mc, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(mc=A.MC_YCGCO, rng=0, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(mc=A.MC_IDENTITY, fmt=A.AVIF_PIXEL_FORMAT_YUV420, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    assert _code_y2r(product, util.Spec(mc=A.MC_YCGCO_RE, rng=1, depth=10, rgb_depth=10, **base)) == A.AVIF_RESULT_REFORMAT_FAILED
    # RGB -> YUV: 565 and missing pixels fail, float is NOT_IMPLEMENTED, sharp yuv on 4:2:0 is NOT_IMPLEMENTED (avif_fuzztest_yuvrgb.cc:46-54)
    assert util.rgb_to_yuv(product, util.Spec(rgb_fmt=A.AVIF_RGB_FORMAT_RGB_565, **base), 1)[0] == A.AVIF_RESULT_REFORMAT_FAILED
    assert util.rgb_to_yuv(product, util.Spec(is_float=True, rgb_depth=16, **base), 1)[0] == A.AVIF_RESULT_NOT_IMPLEMENTED
    assert util.rgb_to_yuv(product, util.Spec(fmt=A.AVIF_PIXEL_FORMAT_YUV420, downs=A.AVIF_CHROMA_DOWNSAMPLING_SHARP_YUV, **base), 1)[0] == A.AVIF_RESULT_NOT_IMPLEMENTED


def test_premultiply_argument_codes(product, oracle):
    for lib in (product, oracle):
        rgb = util.make_rgb_src(util.Spec(w=4, h=4, rgb_fmt=abi.AVIF_RGB_FORMAT_RGB), 1)
        assert lib.avifRGBImagePremultiplyAlpha(rgb.ptr()) == abi.AVIF_RESULT_INVALID_ARGUMENT   # alpha.c:159-161
        assert lib.avifRGBImageUnpremultiplyAlpha(rgb.ptr()) == abi.AVIF_RESULT_REFORMAT_FAILED  # alpha.c:346-348
        rgb = util.make_rgb_src(util.Spec(w=4, h=4, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA), 1)
        rgb.c.pixels = None
        assert lib.avifRGBImagePremultiplyAlpha(rgb.ptr()) == abi.AVIF_RESULT_REFORMAT_FAILED
        rgb = util.make_rgb_src(util.Spec(w=4, h=4, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA), 1)
        rgb.c.rowBytes = 0
        assert lib.avifRGBImageUnpremultiplyAlpha(rgb.ptr()) == abi.AVIF_RESULT_REFORMAT_FAILED


def test_rgb_helpers(product, reference):
    for fmt in range(abi.AVIF_RGB_FORMAT_COUNT):
        assert product.avifRGBFormatChannelCount(fmt) == reference.avifRGBFormatChannelCount(fmt) == abi.rgb_format_channel_count(fmt)
        assert product.avifRGBFormatHasAlpha(fmt) == reference.avifRGBFormatHasAlpha(fmt)
        assert product.avifRGBFormatIsGray(fmt) == reference.avifRGBFormatIsGray(fmt)
        for depth in (8, 10, 12, 16):
            r = abi.avifRGBImage()
            r.format, r.depth = fmt, depth
            assert product.avifRGBImagePixelSize(ctypes.byref(r)) == reference.avifRGBImagePixelSize(ctypes.byref(r))
    img = abi.avifImage()
    img.width, img.height, img.depth = 33, 17, 12
    a, b = abi.avifRGBImage(), abi.avifRGBImage()
    ctypes.memset(ctypes.byref(a), 0x5A, ctypes.sizeof(a))
    ctypes.memset(ctypes.byref(b), 0x5A, ctypes.sizeof(b))
    product.avifRGBImageSetDefaults(ctypes.byref(a), ctypes.byref(img))
    reference.avifRGBImageSetDefaults(ctypes.byref(b), ctypes.byref(img))
    assert bytes(a) == bytes(b)
    assert product.avifRGBImageAllocatePixels(ctypes.byref(a)) == 0 and a.rowBytes == 33 * 8 and a.pixels
    product.avifRGBImageFreePixels(ctypes.byref(a))
    assert not a.pixels and a.rowBytes == 0
    for f in range(0, 7):
        pa, pb = abi.avifPixelFormatInfo(), abi.avifPixelFormatInfo()
        product.avifGetPixelFormatInfo(f, ctypes.byref(pa))
        reference.avifGetPixelFormatInfo(f, ctypes.byref(pb))
        assert bytes(pa) == bytes(pb)
    assert product.avifLibYUVVersion() == 0


def test_limited_full_helpers_exhaustive(product, oracle):
    for depth in (8, 10, 12, 16, 9):
        for v in range(-3, (1 << min(depth, 12)) + 3):
            for name in ("avifLimitedToFullY", "avifLimitedToFullUV", "avifFullToLimitedY", "avifFullToLimitedUV"):
                assert getattr(product, name)(depth, v) == getattr(oracle, name)(depth, v)


def test_avifyuv_limited_round_trip(product):
    """tests/avifyuv.c:102-115 (-m limited): full -> limited -> full at depth 8 never drifts by more than 1."""
    for v in range(256):
        lim = product.avifFullToLimitedY(8, v)
        assert 16 <= lim <= 235
        assert abs(product.avifLimitedToFullY(8, lim) - v) <= 1


def test_verified_divisor_table_is_sorted_and_covers_supported_matrices(product):
    n = product.avifB200VerifiedDivisorCount()
    bits = [product.avifB200VerifiedDivisorBits(k) for k in range(n)]
    assert bits == sorted(bits) and len(set(bits)) == n
    import struct
    f = lambda x: struct.unpack("<I", struct.pack("<f", x))[0]
    import numpy as np
    for kr, kb in ((0.2126, 0.0722), (0.299, 0.114), (0.30, 0.11), (0.212, 0.087), (0.2627, 0.0593)):
        kr, kb = np.float32(kr), np.float32(kb)
        kg = np.float32(np.float32(1) - kr) - kb
        assert f(float(kg)) in bits
        assert f(float(np.float32(2) * (np.float32(1) - kb))) in bits
        assert f(float(np.float32(2) * (np.float32(1) - kr))) in bits
    for d in (255, 1023, 4095, 65535, 219, 876, 3504, 56064, 224, 896, 3584, 57344):
        assert f(float(d)) in bits


def test_premultiply_fma_constants_proven_on_host(product):
    """verified_premul.h: trunc(fma(c*a, r, k)) == floorf(c*a/max + 0.5f) for every code pair of depths 8..12 (host loop in
    the library, 22 M pairs in all); 16 bits has no constants and keeps the reference's operation sequence."""
    import ctypes
    for depth in (8, 9, 10, 11, 12):
        bad = ctypes.c_uint64(77)
        assert product.avifB200VerifyPremulFma(depth, ctypes.byref(bad)) == 0
        assert bad.value == 0, depth
    bad = ctypes.c_uint64(0)
    assert product.avifB200VerifyPremulFma(16, ctypes.byref(bad)) == 25  # AVIF_RESULT_NOT_IMPLEMENTED
