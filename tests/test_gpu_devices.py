"""Device affinity of the host-pointer path (include/avif_b200.h: avifB200SetDevice / avifB200SetDevices), thread-exit cleanup,
and the ownership flags of resident planes."""
import ctypes
import threading

import numpy as np
import pytest

from libavif_b200 import abi
from tests import util

pytestmark = pytest.mark.gpu

SPEC = util.Spec(w=640, h=360, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV420, rng=0, mc=abi.MC_BT709, rgb_depth=10, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA,
                 ups=abi.AVIF_CHROMA_UPSAMPLING_BILINEAR)


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _mem_used(index):
    import pynvml
    pynvml.nvmlInit()
    return pynvml.nvmlDeviceGetMemoryInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).used


def test_threads_release_their_contexts(gpu_product, oracle):
    """200 short-lived threads, one conversion each: device memory must not grow by an arena per thread."""
    ro, want = util.yuv_to_rgb(oracle, SPEC, 7)
    big = util.Spec(w=3840, h=2160, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rng=1, mc=abi.MC_BT709, rgb_depth=10)
    errors = []

    def work(k):
        try:
            if k % 20 == 0:  # now and then a 4K frame: a 100 MB arena
                r, _ = util.yuv_to_rgb(gpu_product, big, 1)
                assert r == 0
            r, got = util.yuv_to_rgb(gpu_product, SPEC, 7)
            assert r == 0 and util.first_mismatch(want, got) is None
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    work(0)
    base = _mem_used(0)
    for k in range(200):
        t = threading.Thread(target=work, args=(k,))
        t.start()
        t.join()
    assert not errors, errors[:1]
    grown = _mem_used(0) - base
    assert grown < 400 << 20, f"device memory grew by {grown >> 20} MB over 200 thread lifetimes"


def test_set_device_is_per_thread_and_restores_the_current_device(gpu_product, oracle):
    import torch
    ro, want = util.yuv_to_rgb(oracle, SPEC, 11)
    n = _n_devices()
    seen = {}

    def work(d):
        assert gpu_product.avifB200SetDevice(d) == 0
        assert gpu_product.avifB200GetDevice() == d
        before = torch.cuda.current_device()
        r, got = util.yuv_to_rgb(gpu_product, SPEC, 11)
        seen[d] = (r, util.first_mismatch(want, got), before, torch.cuda.current_device())

    ts = [threading.Thread(target=work, args=(d,)) for d in range(n)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for d in range(n):
        r, mm, before, after = seen[d]
        assert r == 0 and mm is None, (d, mm)
        assert before == after
    assert gpu_product.avifB200SetDevice(n) == abi.AVIF_RESULT_INVALID_ARGUMENT
    assert gpu_product.avifB200SetDevice(-1) == 0


def test_policy_deals_devices_round_robin(gpu_product):
    n = _n_devices()
    order = (ctypes.c_int * n)(*reversed(range(n)))
    assert gpu_product.avifB200SetDevices(order, n) == 0
    got = []

    def work():
        got.append(gpu_product.avifB200GetDevice())
    for _ in range(2 * n):
        t = threading.Thread(target=work)
        t.start()
        t.join()
    assert got == list(reversed(range(n))) * 2
    assert gpu_product.avifB200SetDevices(None, 0) == 0


@pytest.mark.skipif("_n_devices() < 2")
def test_resident_planes_pin_the_call_to_their_device(gpu_product, oracle):
    ro, want = util.yuv_to_rgb(oracle, SPEC, 13)
    for d in range(_n_devices()):
        r, got = util.yuv_to_rgb(gpu_product, SPEC, 13, device=f"cuda:{d}")  # synchronous entry point, planes on device d
        assert r == 0 and util.first_mismatch(want, got) is None, d


@pytest.mark.skipif("_n_devices() < 2")
def test_host_calls_spread_over_two_gpus(gpu_product, oracle):
    """two threads on two devices, concurrently, many frames: both results exact"""
    ro, want = util.yuv_to_rgb(oracle, SPEC, 17)
    bad = []

    def work(d):
        gpu_product.avifB200SetDevice(d)
        for _ in range(20):
            r, got = util.yuv_to_rgb(gpu_product, SPEC, 17)
            if r != 0 or util.first_mismatch(want, got) is not None:
                bad.append(d)
    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not bad


def test_resident_planes_keep_their_ownership_flags(gpu_product):
    """The stream-ordered and batch entry points (and synchronous calls on resident planes) never claim the caller's device
    memory: imageOwnsYUVPlanes / imageOwnsAlphaPlane stay as they were (the reference would free() them, src/avif.c:352-372)."""
    import torch
    from libavif_b200.images import pointer_array
    spec = util.Spec(w=256, h=64, depth=8, fmt=abi.AVIF_PIXEL_FORMAT_YUV420, rgb_depth=8, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA)
    rgb = util.make_rgb_src(spec, 3, device="cuda:0")
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for call in ("async", "batch", "sync"):
        img = util.make_yuv_dst(spec, device="cuda:0", with_alpha=True)
        assert img.c.imageOwnsYUVPlanes == 0 and img.c.imageOwnsAlphaPlane == 0
        if call == "async":
            assert gpu_product.avifB200ImageRGBToYUVAsync(img.ptr(), rgb.ptr(), sptr) == 0
        elif call == "batch":
            assert gpu_product.avifB200ImageRGBToYUVBatchAsync(pointer_array([img]), pointer_array([rgb]), 1, sptr) == 0
        else:
            assert gpu_product.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
        torch.cuda.synchronize()
        assert img.c.imageOwnsYUVPlanes == 0 and img.c.imageOwnsAlphaPlane == 0, call


def test_stride_zero_planes_alias_row_zero(gpu_product, oracle):
    """yuvRowBytes == 0: every row reads row 0 (the reference's pointer arithmetic does exactly that)"""
    spec = util.Spec(w=300, h=40, depth=8, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rgb_depth=8, rgb_fmt=abi.AVIF_RGB_FORMAT_RGB)
    outs = []
    for lib in (oracle, gpu_product):
        img = util.make_yuv(spec, 5)
        rgb = util.make_rgb_dst(spec)
        img.c.yuvRowBytes[0] = 0
        assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
        outs.append(rgb.full_bytes().copy())
    assert util.first_mismatch(outs[0], outs[1]) is None
