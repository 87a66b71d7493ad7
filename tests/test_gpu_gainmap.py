"""Gain-map tone mapping: avifRGBImageApplyGainMap / avifImageApplyGainMap on the GPU against the
reference's own src/gainmap.c compiled from its sources (oracle/_ref/libavif_ref.so).

The arithmetic is binary32 with powf / exp2f / logf: libm on the reference side, CUDA's on the device.  Both are accurate to
about an ulp but not identical, so this row is held to +-1 output code per channel (the tolerance SURVEY.md states for
it), with the overwhelming majority of samples equal, instead of the bit-exactness of the reformat rows.
"""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage

A = abi


def _frac(n, d):
    return (n, d)


def make_gain_map(w, h, depth=8, fmt=A.AVIF_PIXEL_FORMAT_YUV444, seed=1, gm_min=(0, 1), gm_max=(2, 1), gamma=(1, 1),
                  base_off=(1, 64), alt_off=(1, 64), base_h=(0, 1), alt_h=(3, 1), use_base=1, alt_primaries=2):
    rs = np.random.RandomState(seed)
    img = YUVImage(w, h, depth, fmt, A.AVIF_RANGE_FULL, 6, alpha=False)
    for p in img.present_planes():
        b = img.bufs[p]
        if depth > 8:
            v = rs.randint(0, 1 << depth, size=(b.rows, b.row_bytes // 2)).astype(np.uint16)
            b.set(v.view(np.uint8))
        else:
            b.set(rs.randint(0, 256, size=(b.rows, b.row_bytes)).astype(np.uint8))
    gm = A.avifGainMap()
    gm.image = ctypes.pointer(img.c)
    for c in range(3):
        gm.gainMapMin[c].n, gm.gainMapMin[c].d = gm_min
        gm.gainMapMax[c].n, gm.gainMapMax[c].d = gm_max
        gm.gainMapGamma[c].n, gm.gainMapGamma[c].d = gamma
        gm.baseOffset[c].n, gm.baseOffset[c].d = base_off
        gm.alternateOffset[c].n, gm.alternateOffset[c].d = alt_off
    gm.baseHdrHeadroom.n, gm.baseHdrHeadroom.d = base_h
    gm.alternateHdrHeadroom.n, gm.alternateHdrHeadroom.d = alt_h
    gm.useBaseColorSpace = use_base
    gm.altColorPrimaries = alt_primaries
    gm.altTransferCharacteristics = 2
    gm.altMatrixCoefficients = 2
    gm.altDepth = 8
    gm.altPlaneCount = 3
    return gm, img  # keep img alive


def make_base(w, h, depth, fmt, is_float=False, seed=2, row_pad=0):
    rs = np.random.RandomState(seed)
    rgb = RGBImage(w, h, depth, fmt, is_float=is_float, row_pad=row_pad)
    nch = rgb.pixel_bytes // (2 if depth > 8 else 1)
    if fmt == A.AVIF_RGB_FORMAT_RGB_565:
        v = rs.randint(0, 65536, size=(h, w)).astype(np.uint16)
    elif is_float:
        vals = rs.rand(h, w, nch).astype(np.float32)  # [0,1) as half floats
        v = vals.astype(np.float16).view(np.uint16)
    elif depth > 8:
        v = rs.randint(0, 1 << depth, size=(h, w, nch)).astype(np.uint16)
    else:
        v = rs.randint(0, 256, size=(h, w, nch)).astype(np.uint8)
    full = np.zeros((rgb.buf.rows, rgb.buf.row_bytes), dtype=np.uint8)
    full[:, :w * rgb.pixel_bytes] = v.reshape(h, -1).view(np.uint8)
    rgb.buf.set(full)
    return rgb


def apply(lib, base, base_cp, base_tc, gm, headroom, out_cp, out_tc, out_depth, out_fmt, out_float=False, want_clli=True):
    out = A.avifRGBImage()
    out.depth, out.format, out.isFloat = out_depth, out_fmt, int(out_float)
    clli = A.avifContentLightLevelInformationBox()
    diag = A.avifDiagnostics()
    res = lib.avifRGBImageApplyGainMap(base.ptr(), base_cp, base_tc, ctypes.byref(gm), ctypes.c_float(headroom), out_cp, out_tc,
                                       ctypes.byref(out), ctypes.byref(clli) if want_clli else None, ctypes.byref(diag))
    data = None
    if res == 0:
        n = out.rowBytes * out.height
        data = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(out.pixels)).copy().reshape(out.height, out.rowBytes)
    if out.pixels:
        lib.avifRGBImageFreePixels(ctypes.byref(out))
    return res, data, (clli.maxCLL, clli.maxPALL), diag.error.decode(errors="replace")


def compare(a, b, depth, is_float, tol=1):
    """max |code difference| and the fraction of differing samples"""
    if depth > 8:
        x, y = a.view(np.uint16).astype(np.int64), b.view(np.uint16).astype(np.int64)
    else:
        x, y = a.astype(np.int64), b.astype(np.int64)
    d = np.abs(x - y)
    return int(d.max()), float((d != 0).mean())


TCS = [1, 4, 5, 7, 8, 9, 10, 11, 12, 13, 16, 17, 18]


@pytest.mark.gpu
@pytest.mark.parametrize("tc_in", TCS)
def test_tone_map_matches_reference_for_every_transfer_function(gpu_product, reference, tc_in):
    w, h = 97, 61
    gm, keep = make_gain_map(w, h, seed=10 + tc_in)
    base = make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, seed=20 + tc_in)
    for tc_out in (tc_in, 13, 16):
        for depth in (8, 10, 16):
            ra, a, ca, _ = apply(reference, base, 1, tc_in, gm, 1.5, 1, tc_out, depth, A.AVIF_RGB_FORMAT_RGBA)
            rb, b, cb, err = apply(gpu_product, base, 1, tc_in, gm, 1.5, 1, tc_out, depth, A.AVIF_RGB_FORMAT_RGBA)
            assert ra == rb == 0, (ra, rb, err)
            worst, frac = compare(a, b, depth, False)
            # 16-bit outputs evaluate the transfer function in double and round once (csrc/tonemap.cu: xpow<true>): +-1 code like the
            # other depths; PQ raises the linear value to the 78.84th power, which multiplies what is left of the difference
            tol = 1 if (depth < 16 or 16 not in (tc_in, tc_out)) else 2
            assert worst <= tol, (tc_in, tc_out, depth, worst, frac)
            assert frac < 0.02, (tc_in, tc_out, depth, worst, frac)
            assert abs(ca[0] - cb[0]) <= 1 and abs(ca[1] - cb[1]) <= 1, (ca, cb)


@pytest.mark.gpu
def test_formats_depths_primaries_and_weights(gpu_product, reference):
    w, h = 64, 48
    fmts = [A.AVIF_RGB_FORMAT_RGB, A.AVIF_RGB_FORMAT_RGBA, A.AVIF_RGB_FORMAT_ARGB, A.AVIF_RGB_FORMAT_BGR, A.AVIF_RGB_FORMAT_BGRA,
            A.AVIF_RGB_FORMAT_ABGR, A.AVIF_RGB_FORMAT_GRAY, A.AVIF_RGB_FORMAT_GRAYA, A.AVIF_RGB_FORMAT_AGRAY]
    k = 0
    for bfmt in fmts:
        for bdepth, bfloat in ((8, False), (10, False), (12, False), (16, False), (16, True)):
            k += 1
            ofmt = fmts[(k * 5) % len(fmts)]
            if ofmt == A.AVIF_RGB_FORMAT_AGRAY:
                # avifSetRGBAPixel writes R, G, B and then A all at offset 0 for this format and never touches the gray
                # sample (src/reformat.c:100-103, :1899-1939): half of the output is uninitialised memory on both sides
                ofmt = A.AVIF_RGB_FORMAT_GRAYA
            odepth, ofloat = ((8, False), (10, False), (12, False), (16, False), (16, True))[(k * 3) % 5]
            base_cp, out_cp = ((1, 1), (1, 9), (12, 1), (9, 12))[k % 4]
            use_base, alt_cp = ((1, 2), (0, 9), (0, 12))[k % 3]
            headroom = (0.0, 0.7, 3.0, 5.0)[k % 4]
            gm_fmt = (A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400)[k % 3]
            gm, keep = make_gain_map(w, h, depth=(8, 10)[k % 2], fmt=gm_fmt, seed=100 + k, gamma=((1, 1), (22, 10))[k % 2],
                                     gm_min=(-1, 2), gm_max=(5, 2), use_base=use_base, alt_primaries=alt_cp,
                                     base_h=((0, 1), (4, 1))[k % 2], alt_h=((3, 1), (1, 1))[k % 2])
            base = make_base(w, h, bdepth, bfmt, is_float=bfloat, seed=200 + k, row_pad=(0, 24)[k % 2])
            ra, a, ca, _ = apply(reference, base, base_cp, 13, gm, headroom, out_cp, 13, odepth, ofmt, ofloat)
            rb, b, cb, err = apply(gpu_product, base, base_cp, 13, gm, headroom, out_cp, 13, odepth, ofmt, ofloat)
            assert ra == rb, (k, ra, rb, err)
            if ra != 0:
                continue
            assert a.shape == b.shape
            worst, frac = compare(a, b, odepth, ofloat)
            assert worst <= (2 if odepth == 16 else 1), (k, bfmt, bdepth, ofmt, odepth, ofloat, worst, frac)
            assert abs(ca[0] - cb[0]) <= 1 and abs(ca[1] - cb[1]) <= 1, (k, ca, cb)


@pytest.mark.gpu
def test_rgb565_base_and_output(gpu_product, reference):
    w, h = 40, 30
    gm, keep = make_gain_map(w, h, seed=5)
    base = make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGB_565, seed=6)
    for ofmt in (A.AVIF_RGB_FORMAT_RGB_565, A.AVIF_RGB_FORMAT_RGBA):
        ra, a, _, _ = apply(reference, base, 1, 13, gm, 2.0, 1, 13, 8, ofmt)
        rb, b, _, err = apply(gpu_product, base, 1, 13, gm, 2.0, 1, 13, 8, ofmt)
        assert ra == rb == 0, err
        if ofmt == A.AVIF_RGB_FORMAT_RGB_565:
            x, y = a.view(np.uint16), b.view(np.uint16)
            for shift, mask in ((11, 31), (5, 63), (0, 31)):
                assert np.abs(((x >> shift) & mask).astype(int) - ((y >> shift) & mask).astype(int)).max() <= 1
        else:
            assert compare(a, b, 8, False)[0] <= 1


@pytest.mark.gpu
def test_zero_weight_paths_are_exact(gpu_product, reference):
    """weight == 0: same layout -> byte copy (src/gainmap.c:120-128); other layouts -> a pure format conversion whose only
    float steps are code/max and (T)(0.5f + v*max): bit-exact."""
    w, h = 33, 17
    gm, keep = make_gain_map(w, h, seed=7, base_h=(1, 1), alt_h=(1, 1))  # equal headrooms -> weight 0
    base = make_base(w, h, 10, A.AVIF_RGB_FORMAT_RGBA, seed=8)
    for odepth, ofmt in ((10, A.AVIF_RGB_FORMAT_RGBA), (10, A.AVIF_RGB_FORMAT_BGR), (8, A.AVIF_RGB_FORMAT_ARGB), (16, A.AVIF_RGB_FORMAT_RGB)):
        ra, a, _, _ = apply(reference, base, 1, 13, gm, 2.0, 1, 13, odepth, ofmt)
        rb, b, _, err = apply(gpu_product, base, 1, 13, gm, 2.0, 1, 13, odepth, ofmt)
        assert ra == rb == 0, err
        assert np.array_equal(a, b), (odepth, ofmt)


@pytest.mark.gpu
def test_nan_and_unsupported_cases_report_like_the_reference(gpu_product, reference):
    w, h = 16, 16
    # degenerate metadata: huge log2 gain with a negative offset -> inf - inf = NaN somewhere
    gm, keep = make_gain_map(w, h, seed=9, gm_min=(-2000, 1), gm_max=(2000, 1), base_off=(0, 1), alt_off=(0, 1))
    base = make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, seed=11)
    ra, _, _, _ = apply(reference, base, 1, 13, gm, 3.0, 9, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
    rb, _, _, err = apply(gpu_product, base, 1, 13, gm, 3.0, 9, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
    assert ra == rb, (ra, rb, err)


@pytest.mark.gpu
def test_gain_map_of_another_size_is_rescaled_like_the_reference(gpu_product, reference):
    """src/gainmap.c:186-201: avifImageScale (libyuv box / bilinear) brings the gain map to the base image's size first."""
    w, h = 96, 64
    base = make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, seed=31)
    for gw, gh, fmt in ((48, 32, A.AVIF_PIXEL_FORMAT_YUV444), (24, 16, A.AVIF_PIXEL_FORMAT_YUV420), (200, 150, A.AVIF_PIXEL_FORMAT_YUV400),
                        (96, 20, A.AVIF_PIXEL_FORMAT_YUV444), (37, 64, A.AVIF_PIXEL_FORMAT_YUV422)):
        gm, keep = make_gain_map(gw, gh, fmt=fmt, seed=gw * 7 + gh)
        ra, a, ca, _ = apply(reference, base, 1, 13, gm, 2.0, 1, 13, 10, A.AVIF_RGB_FORMAT_RGBA)
        rb, b, cb, err = apply(gpu_product, base, 1, 13, gm, 2.0, 1, 13, 10, A.AVIF_RGB_FORMAT_RGBA)
        assert ra == rb == 0, (gw, gh, ra, rb, err)
        worst, frac = compare(a, b, 10, False)
        assert worst <= 1 and frac < 0.02, (gw, gh, fmt, worst, frac)
        assert abs(ca[0] - cb[0]) <= 1 and abs(ca[1] - cb[1]) <= 1


@pytest.mark.gpu
def test_image_entry_point(gpu_product, reference):
    """avifImageApplyGainMap: YUV base image -> RGB (device resident) -> tone map."""
    w, h = 80, 40
    gm, keep = make_gain_map(w, h, seed=13)
    rs = np.random.RandomState(14)
    img = YUVImage(w, h, 10, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_LIMITED, 1, alpha=True)
    for p in img.present_planes():
        b = img.bufs[p]
        b.set(rs.randint(0, 1024, size=(b.rows, b.row_bytes // 2)).astype(np.uint16).view(np.uint8))
    img.c.colorPrimaries, img.c.transferCharacteristics = 1, 13
    outs = []
    for lib in (reference, gpu_product):
        out = A.avifRGBImage()
        out.depth, out.format = 10, A.AVIF_RGB_FORMAT_RGBA
        clli = A.avifContentLightLevelInformationBox()
        diag = A.avifDiagnostics()
        res = lib.avifImageApplyGainMap(img.ptr(), ctypes.byref(gm), ctypes.c_float(2.0), 9, 16, ctypes.byref(out), ctypes.byref(clli),
                                        ctypes.byref(diag))
        assert res == 0, diag.error
        n = out.rowBytes * out.height
        outs.append((np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(out.pixels)).copy(), clli.maxCLL, clli.maxPALL))
        lib.avifRGBImageFreePixels(ctypes.byref(out))
    worst, frac = compare(outs[0][0], outs[1][0], 10, False)
    assert worst <= 1 and frac < 0.02, (worst, frac)
    assert abs(outs[0][1] - outs[1][1]) <= 1 and abs(outs[0][2] - outs[1][2]) <= 1


def test_argument_checks_need_no_device(product, reference):
    """Validation happens before any device work: same result codes as the reference, on a machine without a GPU."""
    w, h = 8, 8
    base = make_base(w, h, 8, A.AVIF_RGB_FORMAT_RGBA)
    cases = []
    gm, k0 = make_gain_map(w, h)
    cases.append((gm, -1.0))                                   # negative headroom
    gm1, k1 = make_gain_map(w, h, gamma=(0, 1))                # gamma numerator 0
    gm2, k2 = make_gain_map(w, h, gm_min=(3, 1), gm_max=(1, 1))  # max < min
    gm3, k3 = make_gain_map(w, h, base_off=(1, 0))             # zero denominator
    gm4, k4 = make_gain_map(w, h, base_h=(1, 0))               # zero headroom denominator
    gm5, k5 = make_gain_map(w, h, use_base=2)                  # not a boolean
    cases += [(gm1, 1.0), (gm2, 1.0), (gm3, 1.0), (gm4, 1.0), (gm5, 1.0)]
    for g, headroom in cases:
        ra, _, _, ea = apply(reference, base, 1, 13, g, headroom, 1, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
        rb, _, _, eb = apply(product, base, 1, 13, g, headroom, 1, 13, 8, A.AVIF_RGB_FORMAT_RGBA)
        assert ra == rb == 24, (ra, rb, ea, eb)
        assert ea and eb  # both explain themselves in diag


@pytest.mark.gpu
def test_committed_vectors(gpu_product):
    """tests/golden/tonemap_vectors.npz: the reference's tone-mapped pixels for three small cases (tests/golden/make_golden.py).
    Runs where neither /root/reference nor oracle/_ref exists."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tonemap_vectors.npz"))
    cases = {"srgb8_to_pq10": dict(tc_in=13, tc_out=16, depth=10, headroom=1.5, cp_out=9),
             "bt709_8_to_srgb8": dict(tc_in=1, tc_out=13, depth=8, headroom=3.0, cp_out=1),
             "hlg10_to_hlg12": dict(tc_in=18, tc_out=18, depth=12, headroom=0.7, cp_out=12)}
    for name, c in cases.items():
        gm, keep = make_gain_map(48, 32, seed=77)
        base = make_base(48, 32, 8, A.AVIF_RGB_FORMAT_RGBA, seed=78)
        res, data, clli, err = apply(gpu_product, base, 1, c["tc_in"], gm, c["headroom"], c["cp_out"], c["tc_out"], c["depth"], A.AVIF_RGB_FORMAT_RGBA)
        assert res == 0, err
        worst, frac = compare(g[name + "_pixels"], data, c["depth"], False)
        assert worst <= 1 and frac < 0.02, (name, worst, frac)
        want = g[name + "_clli"]
        assert abs(int(want[0]) - clli[0]) <= 1 and abs(int(want[1]) - clli[1]) <= 1


@pytest.mark.gpu
def test_per_pixel_form_of_the_output_transfer_function(gpu_product):
    """The default build answers <= 12-bit outputs from the threshold table (csrc/tonemap.cu); AVIF_B200_TONEMAP_TABLE=0 evaluates the
    output transfer function per pixel instead.  The switch is read once per process, so the other form runs in a child process:
    the whole transfer-function sweep against the reference, same tolerances."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, AVIF_B200_TONEMAP_TABLE="0")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_gainmap.py"), "-q", "-m", "gpu", "-x",
                          "-k", "every_transfer_function or rgb565 or formats_depths"],
                         env=env, cwd=root, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
