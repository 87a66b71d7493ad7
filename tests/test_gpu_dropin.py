"""Drop-in proof with the reference's own test program.

`oracle/_ref/avifyuv` is the reference's conversion stress tool (tests/avifyuv.c), compiled from the reference
tree where it lies and linked against the reference build `oracle/_ref/libavif_ref.so` (recipe: oracle/Makefile).
Run with `LD_PRELOAD=libavif_b200.so`, every entry point declared in include/avif_b200.h is interposed by the
B200 library while the rest of the tool (image allocation, printing) still comes from the reference build: that
is exactly the situation of a libavif whose reformat path was replaced.  The tool's output must stay
byte-identical to the output of the untouched reference, which is committed under tests/golden/avifyuv_*.txt
(written by the native run in the build container; tests/golden/make_golden.py documents the commands).

The binary is test infrastructure like the rest of oracle/_ref; the product is only the preloaded library.
"""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TOOL = os.path.join(ROOT, "oracle", "_ref", "avifyuv")
LIB = os.path.join(ROOT, "libavif_b200", "csrc", "libavif_b200.so")

pytestmark = pytest.mark.gpu


def _run(mode, preload, extra=()):
    env = dict(os.environ)
    if preload:
        env["LD_PRELOAD"] = LIB
        env["AVIF_B200_TRACE_CALLS"] = "1"
    out = subprocess.run([TOOL, "-m", mode, *extra], env=env, capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode(), out.stderr.decode()


def _golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return f.read()


def _strip_progress(text):
    return "".join(line + "\n" for line in text.replace("\r", "\n").split("\n") if line and not line.startswith("["))


@pytest.fixture(scope="module")
def tool():
    if not os.path.exists(TOOL):
        pytest.skip("oracle/_ref/avifyuv not built (reference tree absent at build time)")
    if not os.path.exists(LIB):
        pytest.fail("libavif_b200.so missing: run __graft_entry__.build()")
    return TOOL


def test_rgb_mode_identical_and_served_by_gpu(tool):
    native, _ = _run("rgb", preload=False)
    assert native == _golden("avifyuv_rgb.txt")
    preloaded, trace = _run("rgb", preload=True)
    assert preloaded == native
    # the interposed library reports how many kernels it launched when the process exits
    assert "avif_b200: kernel launches" in trace
    launches = int(trace.split("avif_b200: kernel launches")[-1].split()[0])
    assert launches >= 144 * 4


def test_limited_mode_identical(tool):
    preloaded, _ = _run("limited", preload=True)
    assert preloaded == _golden("avifyuv_limited.txt")


def test_premultiply_mode_identical(tool):
    # 2^16 + 2^12 + 2^10 + 2^8 images, one un-premultiply and one premultiply call each, every (r <= a) pair
    preloaded, trace = _run("premultiply", preload=True, extra=("-v",))
    assert _strip_progress(preloaded) == _golden("avifyuv_premultiply.txt")
    launches = int(trace.split("avif_b200: kernel launches")[-1].split()[0])
    assert launches >= 2 * (65536 + 4096 + 1024 + 256)


# ---- B-2: the reference's internal seam -----------------------------------------------------------------------------------
# oracle/_ref/libavif_ref_b2.so = the reference's own reformat.c / alpha.c / ... compiled WITHOUT reformat_libyuv.c and
# reformat_libsharpyuv.c and linked against libavif_b200_hooks.so, which provides those files' six hook functions on the GPU
# (libavif_b200/csrc/hooks.c).  avifyuv_b2 is the reference's stress tool linked against that build.
TOOL_B2 = os.path.join(ROOT, "oracle", "_ref", "avifyuv_b2")
LIB_B2 = os.path.join(ROOT, "oracle", "_ref", "libavif_ref_b2.so")


@pytest.fixture(scope="module")
def b2():
    if not os.path.exists(TOOL_B2) or not os.path.exists(LIB_B2):
        pytest.skip("oracle/_ref B-2 build absent (reference tree absent at build time)")
    return TOOL_B2


def _run_b2(mode, extra=()):
    env = dict(os.environ, AVIF_B200_TRACE_CALLS="1")
    out = subprocess.run([TOOL_B2, "-m", mode, *extra], env=env, capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    return out.stdout.decode(), out.stderr.decode()


def test_b2_reference_tool_on_hooks(b2):
    out, trace = _run_b2("rgb")
    assert out == _golden("avifyuv_rgb.txt")
    launches = int(trace.split("avif_b200: kernel launches")[-1].split()[0])
    assert launches >= 500  # 576 conversions; the few the hooks decline run the reference's own loop
    out, _ = _run_b2("premultiply", extra=("-v",))
    assert _strip_progress(out) == _golden("avifyuv_premultiply.txt")


def test_b2_reference_library_equals_plain_reference(b2, reference, monkeypatch):
    """Random specs through the hooked reference build and through the plain reference build: same bytes and result
    codes, and the hooked build actually launched kernels (it declines only alpha-multiply-in-the-general-loop cases)."""
    import libavif_b200 as L
    from oracle import loader
    from tests import util
    hooked = loader.load_reference_b2()
    gpu = L.lib()
    gpu.avifB200ResetKernelLaunchCount()
    n_y2r = n_r2y = 0
    monkeypatch.setattr(L.RGBImage, "AVOID_LIBYUV", 0)  # the reference consults its seam only when rgb->avoidLibYUV is 0
    for k, spec in enumerate(util.random_y2r_specs(120, seed=777, max_dim=64)):
        ca, a = util.yuv_to_rgb(reference, spec, seed=9000 + k)
        cb, b = util.yuv_to_rgb(hooked, spec, seed=9000 + k)
        assert ca == cb, (spec, ca, cb)
        if ca == 0:
            assert bytes(a) == bytes(b), spec
            n_y2r += 1
    for k, spec in enumerate(util.random_r2y_specs(120, seed=778, max_dim=64)):
        ca, pa = util.rgb_to_yuv(reference, spec, seed=9500 + k)
        cb, pb = util.rgb_to_yuv(hooked, spec, seed=9500 + k)
        assert ca == cb, (spec, ca, cb)
        if ca == 0:
            for x, y in zip(pa, pb):
                assert (x is None) == (y is None)
                if x is not None:
                    assert bytes(x) == bytes(y), spec
            n_r2y += 1
    assert n_y2r > 60 and n_r2y > 60
    assert gpu.avifB200KernelLaunchCount() >= (n_y2r + n_r2y) // 2


# ---- Seam 1, link flavour: the reference WITHOUT reformat.c / alpha.c -----------------------------------------------------------
# oracle/_ref/libavif_ref_s1.so = the reference's avif.c, colr.c, gainmap.c, scale.c ... compiled without src/reformat.c, src/alpha.c,
# src/reformat_libyuv.c and src/reformat_libsharpyuv.c and linked (-Wl,--no-undefined) against libavif_b200.so, which therefore has
# to provide avifImageYUVToRGB / avifImageRGBToYUV and the internal accessors of SURVEY row a14 (avifGetRGBColorSpaceInfo,
# avifGetRGBAPixel, avifSetRGBAPixel).  avifyuv_s1 = the reference's stress tool on that build.
TOOL_S1 = os.path.join(ROOT, "oracle", "_ref", "avifyuv_s1")
LIB_S1 = os.path.join(ROOT, "oracle", "_ref", "libavif_ref_s1.so")


@pytest.fixture(scope="module")
def s1():
    if not os.path.exists(TOOL_S1) or not os.path.exists(LIB_S1):
        pytest.skip("oracle/_ref seam-1 build absent (reference tree absent at build time)")
    return TOOL_S1


def test_s1_link_replacement_runs_the_reference_tool(s1):
    env = dict(os.environ, AVIF_B200_TRACE_CALLS="1")
    out = subprocess.run([TOOL_S1, "-m", "rgb"], env=env, capture_output=True, timeout=600)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    assert out.stdout.decode() == _golden("avifyuv_rgb.txt")
    launches = int(out.stderr.decode().split("avif_b200: kernel launches")[-1].split()[0])
    assert launches >= 144 * 4
    out = subprocess.run([TOOL_S1, "-m", "limited"], env=env, capture_output=True, timeout=600)
    assert out.returncode == 0 and out.stdout.decode() == _golden("avifyuv_limited.txt")
    # the build really has no conversion code of its own
    nm = subprocess.run(["nm", "-D", LIB_S1], capture_output=True, text=True, check=True).stdout
    for sym in ("avifImageYUVToRGB", "avifImageRGBToYUV", "avifGetRGBAPixel", "avifSetRGBAPixel", "avifGetRGBColorSpaceInfo"):
        assert f" U {sym}" in nm, sym


def test_s1_reference_gain_map_code_on_the_library(s1, reference):
    """The reference's OWN src/gainmap.c (CPU loops, host libm) inside the seam-1 build: its conversions and per-pixel accessors
    are the library's.  Same libm, same loops -> the tone-mapped bytes must equal the plain reference's bit for bit."""
    import ctypes
    import numpy as np
    from libavif_b200 import abi as A
    from oracle import loader
    from tests import test_gpu_gainmap as tg
    linked = loader._Facade(ctypes.CDLL(LIB_S1), lambda n: n)
    for k, (tc_in, tc_out, depth, fmt) in enumerate([(13, 16, 10, A.AVIF_RGB_FORMAT_RGBA), (1, 13, 8, A.AVIF_RGB_FORMAT_BGR), (18, 18, 12, A.AVIF_RGB_FORMAT_ARGB),
                                                     (13, 8, 16, A.AVIF_RGB_FORMAT_RGB)]):
        outs = []
        for lib in (reference, linked):
            gm, keep = tg.make_gain_map(64, 40, seed=300 + k)
            base = tg.make_base(64, 40, 8, A.AVIF_RGB_FORMAT_RGBA, seed=400 + k)
            res, data, clli, _ = tg.apply(lib, base, 1, tc_in, gm, 1.7, 9, tc_out, depth, fmt)
            assert res == 0
            outs.append((np.asarray(data).copy(), tuple(clli)))
        assert np.array_equal(outs[0][0], outs[1][0]), (tc_in, tc_out, depth, fmt)
        assert outs[0][1] == outs[1][1]
