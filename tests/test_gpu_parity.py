"""Parity tests proper: the CUDA path, called through the C-ABI, against the oracle on the same seeded inputs.
Bit-exact (integer code values) everywhere."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from tests import util

pytestmark = pytest.mark.gpu


def _check_y2r(product, oracle, spec, seed, device=None, policy=0):
    ro, bo = util.yuv_to_rgb(oracle, spec, seed)
    product.avifB200SetKernelPolicy(policy)
    try:
        rp, bp = util.yuv_to_rgb(product, spec, seed, device=device)
    finally:
        product.avifB200SetKernelPolicy(0)
    assert rp == ro, f"result code {rp} vs oracle {ro} for {spec.ident()}"
    if ro == abi.AVIF_RESULT_OK:
        mm = util.first_mismatch(bo, bp)
        assert mm is None, f"{spec.ident()} [{product.avifB200LastKernelName().decode()}]: {mm}"
    return ro


def _check_r2y(product, oracle, spec, seed, device=None, policy=0):
    ro, po = util.rgb_to_yuv(oracle, spec, seed)
    product.avifB200SetKernelPolicy(policy)
    try:
        rp, pp = util.rgb_to_yuv(product, spec, seed, device=device)
    finally:
        product.avifB200SetKernelPolicy(0)
    assert rp == ro, f"result code {rp} vs oracle {ro} for {spec.ident()}"
    if ro == abi.AVIF_RESULT_OK:
        for p in range(4):
            assert (po[p] is None) == (pp[p] is None)
            if po[p] is not None:
                mm = util.first_mismatch(po[p], pp[p])
                assert mm is None, f"{spec.ident()} plane {p}: {mm}"
    return ro


@pytest.mark.parametrize("policy", [0, 1, 2, 3, 4], ids=["fast", "generic", "tiled-ieee-div", "generic-ieee-div", "tiled"])
def test_y2r_sweep_host_pointers(gpu_product, oracle, policy):
    n_ok = 0
    for k, spec in enumerate(util.random_y2r_specs(700, seed=21 + policy)):
        n_ok += _check_y2r(gpu_product, oracle, spec, 3000 + k, policy=policy) == abi.AVIF_RESULT_OK
    assert n_ok > 450


def test_y2r_sweep_device_pointers(gpu_product, oracle):
    n_ok = 0
    for k, spec in enumerate(util.random_y2r_specs(300, seed=23)):
        n_ok += _check_y2r(gpu_product, oracle, spec, 4000 + k, device="cuda:0") == abi.AVIF_RESULT_OK
    assert n_ok > 180


def test_y2r_wide_tiles(gpu_product, oracle):
    """Widths around the 256-pixel warp tile and the 16-byte copy-out granularity, odd heights, padded rows."""
    sizes = [(255, 3), (256, 2), (257, 5), (511, 4), (513, 3), (1023, 2), (1025, 3), (1920, 6), (33, 17)]
    for k, spec in enumerate(util.random_y2r_specs(160, seed=24, sizes=sizes)):
        _check_y2r(gpu_product, oracle, spec, 5000 + k)


def test_y2r_fast_kernels_are_used_and_exact(gpu_product, oracle):
    """Aligned, padded-free frames of tile-straddling sizes: the specialised kernels must serve them (name check)
    and match the oracle byte for byte, for every (depth, format, upsampling, alpha) combination they cover."""
    import dataclasses
    import itertools
    sizes = [(264, 6), (520, 5), (1000, 4), (257, 7), (8, 3), (2056, 3)]
    served = 0
    k = 0
    for depth, rgb_depth in itertools.product([8, 10, 12, 16], [8, 10, 16]):
        for fmt, ups in [(1, 0), (2, 3), (2, 4), (3, 3), (3, 4)]:
            for rgb_fmt in [abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_ABGR,
                            abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_BGR]:
                for alpha, ip, rp in [(False, False, False), (True, False, False), (True, False, True), (True, True, False)]:
                    k += 1
                    w, h = sizes[k % len(sizes)]
                    spec = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=k % 2, mc=[1, 6, 9, 12][k % 4], cp=[1, 9, 12][k % 3], alpha=alpha,
                                     img_premult=ip, rgb_depth=rgb_depth, rgb_fmt=rgb_fmt, ups=ups, rgb_premult=rp, garbage=(k % 3 == 0))
                    assert _check_y2r(gpu_product, oracle, spec, 9000 + k, device="cuda:0") == 0
                    served += gpu_product.avifB200LastKernelName().decode().startswith("y2r_fast")
    assert served > 0.3 * k, f"only {served} of {k} aligned conversions ran on the specialised kernels"


def test_y2r_fast_kernels_serve_android_outputs(gpu_product, oracle):
    """What android_jni/.../libavif_jni.cc:206-223 asks for -- RGBA_F16 premultiplied, RGB_565, RGBA8 premultiplied from 10-bit --
    plus every alpha-depth rescale between 8-bit and 16-bit containers: served by the specialised kernels, exact."""
    sizes = [(264, 6), (520, 5), (1000, 4), (257, 7), (8, 3), (2056, 3)]
    k = 0
    for depth in (8, 10, 12):
        for fmt, ups in [(1, 0), (2, 3), (2, 4), (3, 3), (3, 0)]:
            for alpha, ip in [(False, False), (True, False), (True, True)]:
                for (rgb_depth, rgb_fmt, is_float, rp) in [(16, abi.AVIF_RGB_FORMAT_RGBA, True, True), (16, abi.AVIF_RGB_FORMAT_BGRA, True, False),
                                                           (16, abi.AVIF_RGB_FORMAT_RGB, True, False), (8, abi.AVIF_RGB_FORMAT_RGB_565, False, True),
                                                           (8 if depth > 8 else 16, abi.AVIF_RGB_FORMAT_RGBA, False, True),
                                                           (8 if depth > 8 else 12, abi.AVIF_RGB_FORMAT_BGRA, False, False)]:
                    k += 1
                    w, h = sizes[k % len(sizes)]
                    spec = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=k % 2, mc=[1, 6, 9][k % 3], alpha=alpha, img_premult=ip, rgb_depth=rgb_depth,
                                     rgb_fmt=rgb_fmt, ups=ups, rgb_premult=rp, is_float=is_float, garbage=(k % 3 == 0))
                    assert _check_y2r(gpu_product, oracle, spec, 9900 + k, device="cuda:0") == 0
                    name = gpu_product.avifB200LastKernelName().decode()
                    unmul = alpha and ip and not rp and abi.rgb_format_has_alpha(rgb_fmt)  # F16 / rescaled unmultiply: general kernel
                    if not unmul and w != 257:  # 257-sample rows are not vector-aligned on the device: those stay on the general kernel
                        assert name.startswith("y2r_fast"), (spec.ident(), name)


def test_identity_matrix_runs_on_the_fast_kernels(gpu_product, oracle):
    """Lossless GBR (matrix coefficients 0, 4:4:4): both directions on the specialised kernels, every depth, limited and full range,
    colour-first and alpha-first layouts, straight alpha."""
    k = 0
    for depth, rgb_depth in [(8, 8), (10, 10), (12, 12), (16, 16), (8, 16), (12, 8)]:
        for rng in (0, 1):
            for rgb_fmt in [abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_BGR]:
                for alpha in (False, True):
                    k += 1
                    w, h = [(264, 6), (520, 5), (1000, 4), (2056, 3)][k % 4]
                    spec = util.Spec(w=w, h=h, depth=depth, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rng=rng, mc=abi.MC_IDENTITY, alpha=alpha, rgb_depth=rgb_depth,
                                     rgb_fmt=rgb_fmt, garbage=(k % 3 == 0))
                    assert _check_y2r(gpu_product, oracle, spec, 9700 + k, device="cuda:0") == 0
                    rescale_same_container = alpha and abi.rgb_format_has_alpha(rgb_fmt) and depth != rgb_depth and (depth > 8) == (rgb_depth > 8)
                    multiply_in_float = alpha and not abi.rgb_format_has_alpha(rgb_fmt)  # straight alpha folded into a colour-only output
                    if not rescale_same_container and not multiply_in_float:
                        assert gpu_product.avifB200LastKernelName().decode().startswith("y2r_fast"), spec.ident()
                    if not alpha:
                        assert _check_r2y(gpu_product, oracle, spec, 9800 + k, device="cuda:0") == 0
                        if depth == rgb_depth or not abi.rgb_format_has_alpha(rgb_fmt):
                            assert gpu_product.avifB200LastKernelName().decode().startswith("r2y_fast"), spec.ident()


def test_r2y_fast_kernels_are_used_and_exact(gpu_product, oracle):
    import itertools
    sizes = [(264, 6), (520, 5), (1000, 4), (257, 7), (8, 3), (2056, 3), (1, 1), (9, 2)]
    served = 0
    k = 0
    for depth, rgb_depth in itertools.product([8, 10, 12, 16], [8, 10, 16]):
        for fmt in [1, 2, 3, 4]:
            for rgb_fmt in [abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_ABGR,
                            abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_BGR]:
                for ip, rp, ign in [(False, False, False), (True, False, False), (False, True, False), (False, False, True)]:
                    k += 1
                    w, h = sizes[k % len(sizes)]
                    spec = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=k % 2, mc=[1, 6, 9, 12][k % 4], cp=[1, 9, 12][k % 3],
                                     img_premult=ip, rgb_depth=rgb_depth, rgb_fmt=rgb_fmt, rgb_premult=rp, ignore_alpha=ign, garbage=(k % 3 == 0))
                    assert _check_r2y(gpu_product, oracle, spec, 9500 + k, device="cuda:0") == 0
                    served += gpu_product.avifB200LastKernelName().decode().startswith("r2y_fast")
    assert served > 0.3 * k, f"only {served} of {k} aligned conversions ran on the specialised kernels"


def test_y2r_unaligned_device_views(gpu_product, oracle):
    """View-rectangle style inputs: plane / pixel pointers that are only sample-aligned."""
    import dataclasses
    for k, spec in enumerate(util.random_y2r_specs(120, seed=25, sizes=[(67, 9), (300, 4), (129, 6)])):
        bps_y = 2 if spec.depth > 8 else 1
        bps_r = 2 if spec.rgb_depth > 8 else 1
        spec = dataclasses.replace(spec, yuv_off=bps_y * (1 + k % 3), rgb_off=bps_r * (1 + k % 5), yuv_pad=2 * (k % 4), rgb_pad=2 * (k % 3))
        _check_y2r(gpu_product, oracle, spec, 6000 + k, device="cuda:0")


def test_sample_aligned_views_stay_on_the_specialised_kernels(gpu_product, oracle):
    """Rows that are only sample-aligned (a view into a larger image, a tightly packed 257-pixel row) keep the specialised kernels:
    every tile goes through their ragged, element-wise code (frame.vecWidth == 0) instead of falling to the general kernel."""
    import dataclasses
    k = served_y2r = served_r2y = 0
    for depth, rgb_depth in [(8, 8), (10, 10), (12, 16), (8, 16), (10, 8)]:
        for fmt, ups in [(abi.AVIF_PIXEL_FORMAT_YUV444, 0), (abi.AVIF_PIXEL_FORMAT_YUV420, abi.AVIF_CHROMA_UPSAMPLING_BILINEAR),
                         (abi.AVIF_PIXEL_FORMAT_YUV420, abi.AVIF_CHROMA_UPSAMPLING_NEAREST), (abi.AVIF_PIXEL_FORMAT_YUV422, abi.AVIF_CHROMA_UPSAMPLING_BILINEAR)]:
            for rgb_fmt in [abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_BGRA]:
                for alpha in (False, True):
                    k += 1
                    w, h = [(257, 5), (521, 4), (300, 7), (1031, 3)][k % 4]
                    bps_y = 2 if depth > 8 else 1
                    bps_r = 2 if rgb_depth > 8 else 1
                    spec = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=k % 2, mc=[1, 6, 9][k % 3], alpha=alpha, rgb_depth=rgb_depth, rgb_fmt=rgb_fmt,
                                     ups=ups, garbage=(k % 3 == 0), yuv_off=bps_y * (1 + k % 3), rgb_off=2 * (1 + k % 3), yuv_pad=2 * (k % 4),
                                     rgb_pad=2 * (k % 3))
                    assert _check_y2r(gpu_product, oracle, spec, 9300 + k, device="cuda:0") == 0
                    name = gpu_product.avifB200LastKernelName().decode()
                    rescale_same_container = alpha and abi.rgb_format_has_alpha(rgb_fmt) and depth != rgb_depth and (depth > 8) == (rgb_depth > 8)
                    if not rescale_same_container and not (alpha and not abi.rgb_format_has_alpha(rgb_fmt)):
                        assert name.startswith("y2r_fast"), (spec.ident(), name)
                        served_y2r += 1
                    if not alpha:
                        assert _check_r2y(gpu_product, oracle, spec, 9400 + k, device="cuda:0") == 0
                        if depth == rgb_depth or not abi.rgb_format_has_alpha(rgb_fmt):
                            assert gpu_product.avifB200LastKernelName().decode().startswith("r2y_fast"), spec.ident()
                            served_r2y += 1
    assert served_y2r > 60 and served_r2y > 30, (served_y2r, served_r2y, k)


@pytest.mark.parametrize("policy", [0, 1, 2, 4], ids=["fast", "generic", "vec-ieee-div", "vec"])
def test_r2y_sweep_host_pointers(gpu_product, oracle, policy):
    n_ok = 0
    for k, spec in enumerate(util.random_r2y_specs(700, seed=31 + policy)):
        n_ok += _check_r2y(gpu_product, oracle, spec, 7000 + k, policy=policy) == abi.AVIF_RESULT_OK
    assert n_ok > 450


def test_r2y_sweep_device_pointers(gpu_product, oracle):
    for k, spec in enumerate(util.random_r2y_specs(250, seed=32, sizes=[(64, 8), (65, 9), (127, 3), (258, 5), (1, 1), (2, 2), (513, 2)])):
        _check_r2y(gpu_product, oracle, spec, 8000 + k, device="cuda:0")


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
@pytest.mark.parametrize("un", [0, 1], ids=["premultiply", "unpremultiply"])
def test_premultiply_in_place(gpu_product, oracle, depth, un):
    for fmt in [abi.AVIF_RGB_FORMAT_RGBA, abi.AVIF_RGB_FORMAT_BGRA, abi.AVIF_RGB_FORMAT_ARGB, abi.AVIF_RGB_FORMAT_ABGR,
                abi.AVIF_RGB_FORMAT_GRAYA, abi.AVIF_RGB_FORMAT_AGRAY, abi.AVIF_RGB_FORMAT_RGB, abi.AVIF_RGB_FORMAT_GRAY]:
        for device in (None, "cuda:0"):
            spec = util.Spec(w=301, h=7, rgb_depth=depth, rgb_fmt=fmt, rgb_pad=6, garbage=True)
            outs = []
            for lib, dev in ((oracle, None), (gpu_product, device)):
                rgb = util.make_rgb_src(spec, seed=90 + depth, device=dev)
                fn = lib.avifRGBImageUnpremultiplyAlpha if un else lib.avifRGBImagePremultiplyAlpha
                outs.append((fn(rgb.ptr()), rgb.full_bytes().copy()))
            assert outs[0][0] == outs[1][0]
            assert util.first_mismatch(outs[0][1], outs[1][1]) is None


def _golden(product, oracle, w, h, depth, fmt, rng, mc, alpha, iprem, rgb_depth, rgb_fmt, ups, rprem, seed, device):
    from libavif_b200.images import RGBImage, YUVImage
    host = YUVImage(w, h, depth, fmt, rng, mc, alpha=alpha, alpha_premultiplied=iprem)
    util.xorshift_fill_image(oracle, host, seed)
    in_hash = util.fnv1a64(oracle, host.payload_bytes())
    img = host
    if device:
        img = YUVImage(w, h, depth, fmt, rng, mc, alpha=alpha, alpha_premultiplied=iprem, device=device)
        for p in host.present_planes():
            img.bufs[p].set(host.bufs[p].host())
    rgb = RGBImage(w, h, rgb_depth, rgb_fmt, upsampling=ups, alpha_premultiplied=rprem, device=device)
    assert product.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
    return in_hash, util.fnv1a64(oracle, rgb.payload_bytes()), rgb


GOLDEN = {  # SURVEY.md Appendix C (reference outputs recorded from the compiled reference)
    "G1a": ((1920, 1080, 8, 3, 0, 6, False, False, 8, 1, 4, False, 1), 0xd8c41c5edae2a9db, 0xbe2d9874701b4ce2),
    "G1b": ((1920, 1080, 8, 3, 0, 6, False, False, 8, 1, 3, False, 1), 0xd8c41c5edae2a9db, 0x0d2f2eb74b548364),
    "G2a": ((3840, 2160, 10, 1, 1, 1, True, False, 10, 1, 0, True, 2), 0x003e7888b2412f05, 0xf47a37093a2029be),
    "G3": ((7680, 4320, 12, 3, 0, 9, True, True, 12, 1, 4, False, 3), 0x1b1d7742e04261dc, 0xa520172ab021d8f7),
    "G5": ((3840, 2160, 10, 3, 0, 1, False, False, 16, 0, 0, False, 5), 0x90804e1e3df1c849, 0x8512bd6346d339c7),
    "GN": ((7680, 4320, 10, 1, 1, 1, False, False, 10, 1, 0, False, 7), 0x19734f45d63ebb3e, 0x3294f95469fdc527),
}


@pytest.mark.parametrize("name", list(GOLDEN))
@pytest.mark.parametrize("device", [None, "cuda:0"], ids=["host", "device"])
def test_full_size_golden_hashes(gpu_product, oracle, name, device):
    """BASELINE.json's full-size configurations: whole-frame FNV-1a-64 of the CUDA output == the reference's."""
    args, want_in, want_out = GOLDEN[name]
    got_in, got_out, _ = _golden(gpu_product, oracle, *args, device=device)
    assert got_in == want_in
    assert got_out == want_out, f"{name}: {got_out:016x} != {want_out:016x} ({gpu_product.avifB200LastKernelName().decode()})"


def test_c2_round_trip_golden(gpu_product, oracle):
    """C2: G2a output (RGBA10 premultiplied) -> 4:4:4 10-bit straight alpha; hash over Y,U,V,A == G2b."""
    from libavif_b200.images import YUVImage
    args, _, _ = GOLDEN["G2a"]
    _, _, rgb = _golden(gpu_product, oracle, *args, device=None)
    back = YUVImage(3840, 2160, 10, abi.AVIF_PIXEL_FORMAT_YUV444, abi.AVIF_RANGE_FULL, abi.MC_BT709, alpha=True, alpha_premultiplied=False)
    assert gpu_product.avifImageRGBToYUV(back.ptr(), rgb.ptr()) == 0
    assert util.fnv1a64(oracle, back.payload_bytes()) == 0x1f2ee19d048166ce


def test_exhaustive_8bit_yuv_cube(gpu_product, oracle):
    """Every (Y,U,V) triple at depth 8 (2^24 pixels) through the fast 4:4:4 path, limited BT.601 -> RGBA8."""
    from libavif_b200.images import RGBImage, YUVImage
    w, h = 4096, 4096
    idx = np.arange(w * h, dtype=np.uint32).reshape(h, w)
    outs = []
    for lib in (oracle, gpu_product):
        img = YUVImage(w, h, 8, abi.AVIF_PIXEL_FORMAT_YUV444, abi.AVIF_RANGE_LIMITED, abi.MC_BT601)
        img.bufs[0].set((idx & 255).astype(np.uint8))
        img.bufs[1].set(((idx >> 8) & 255).astype(np.uint8))
        img.bufs[2].set(((idx >> 16) & 255).astype(np.uint8))
        rgb = RGBImage(w, h, 8, abi.AVIF_RGB_FORMAT_RGBA)
        assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
        outs.append(rgb.payload_bytes().copy())
    assert util.first_mismatch(outs[0], outs[1]) is None


def _batch_case(gpu_product, oracle, specs, seed0):
    """Runs the batch entry points on device-resident frames and checks every frame against the oracle."""
    import ctypes
    import torch
    from libavif_b200.images import pointer_array
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    imgs = [util.make_yuv(sp, seed0 + k, device="cuda:0") for k, sp in enumerate(specs)]
    rgbs = [util.make_rgb_dst(sp, device="cuda:0") for sp in specs]
    want = [util.yuv_to_rgb(oracle, sp, seed0 + k) for k, sp in enumerate(specs)]
    codes = {w[0] for w in want}
    res = gpu_product.avifB200ImageYUVToRGBBatchAsync(pointer_array(imgs), pointer_array(rgbs), len(specs), stream)
    torch.cuda.synchronize()
    if codes == {0}:
        assert res == 0
        for k, sp in enumerate(specs):
            mm = util.first_mismatch(want[k][1], rgbs[k].full_bytes())
            assert mm is None, f"frame {k} {sp.ident()}: {mm}"
    else:
        assert res in codes - {0}
    return gpu_product.avifB200LastKernelName().decode()


def test_batch_yuv_to_rgb(gpu_product, oracle):
    import dataclasses
    gpu_product.avifB200ResetKernelLaunchCount()
    base = util.Spec(w=1920, h=64, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV420, rng=0, mc=abi.MC_BT709, rgb_depth=16,
                     rgb_fmt=abi.AVIF_RGB_FORMAT_RGB, ups=abi.AVIF_CHROMA_UPSAMPLING_BILINEAR)
    # 30 equal frames: two fused launches (24 + 6)
    name = _batch_case(gpu_product, oracle, [base] * 30, 100)
    assert name.startswith("y2r_fast") and gpu_product.avifB200KernelLaunchCount() == 2
    # same constants, different sizes (incl. ragged and tiny ones) still fuse; the 257-wide frame has 514-byte rows (not
    # 16-byte aligned): its tiles take the element-wise code of the same kernel (frame.vecWidth == 0), the batch stays ONE launch
    sizes = [(264, 6), (8, 3), (1000, 5), (257, 7), (2056, 2), (16, 16), (520, 9)]
    gpu_product.avifB200ResetKernelLaunchCount()
    _batch_case(gpu_product, oracle, [dataclasses.replace(base, w=w, h=h) for w, h in sizes], 200)
    assert gpu_product.avifB200KernelLaunchCount() == 1
    # alpha + premultiply, 4:4:4
    c2 = util.Spec(w=512, h=33, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rng=1, mc=abi.MC_BT709, alpha=True, rgb_depth=10,
                   rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA, rgb_premult=True)
    _batch_case(gpu_product, oracle, [c2] * 5, 300)
    # heterogeneous batch: falls back to compatible runs / per-image launches, still exact
    mixed = [base, c2, dataclasses.replace(base, mc=abi.MC_BT601), dataclasses.replace(c2, rgb_fmt=abi.AVIF_RGB_FORMAT_BGRA),
             dataclasses.replace(base, depth=8, rgb_depth=8), dataclasses.replace(base, mc=abi.MC_YCGCO, rng=1),
             dataclasses.replace(base, rgb_fmt=abi.AVIF_RGB_FORMAT_RGB_565, rgb_depth=8), dataclasses.replace(c2, is_float=True, rgb_depth=16)]
    _batch_case(gpu_product, oracle, mixed, 400)
    # one invalid frame: its error code comes back
    _batch_case(gpu_product, oracle, [base, dataclasses.replace(base, mc=3), base], 500)


def test_batch_rgb_to_yuv(gpu_product, oracle):
    import ctypes
    import dataclasses
    import torch
    from libavif_b200.images import pointer_array
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    base = util.Spec(w=1920, h=36, depth=8, fmt=abi.AVIF_PIXEL_FORMAT_YUV420, rng=1, mc=abi.MC_BT601, rgb_depth=8, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA)
    c2 = util.Spec(w=520, h=17, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rng=1, mc=abi.MC_BT709, rgb_depth=10, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA, rgb_premult=True)
    for specs in ([base] * 26, [dataclasses.replace(base, w=w, h=h) for w, h in [(264, 6), (9, 3), (1001, 5), (2056, 2)]], [c2] * 3,
                  [base, c2, dataclasses.replace(base, rgb_fmt=abi.AVIF_RGB_FORMAT_GRAY), dataclasses.replace(c2, mc=abi.MC_IDENTITY)]):
        srcs = [util.make_rgb_src(sp, 700 + k, device="cuda:0") for k, sp in enumerate(specs)]
        dsts = [util.make_yuv_dst(sp, device="cuda:0", with_alpha=abi.rgb_format_has_alpha(sp.rgb_fmt)) for sp in specs]
        assert gpu_product.avifB200ImageRGBToYUVBatchAsync(pointer_array(dsts), pointer_array(srcs), len(specs), stream) == 0
        torch.cuda.synchronize()
        for k, sp in enumerate(specs):
            ro, po = util.rgb_to_yuv(oracle, sp, 700 + k)
            assert ro == 0
            for p in range(4):
                if po[p] is not None:
                    mm = util.first_mismatch(po[p], dsts[k].bufs[p].host().reshape(-1))
                    assert mm is None, f"frame {k} plane {p} {sp.ident()}: {mm}"
