"""The C grid entry points (include/avif_b200.h: avifB200Grid*): the split rule against the Python restatement of
avifImageSplitGrid, the one-GPU path (cell views, one batched launch) against the whole-canvas conversion, the row-range entry
point against whole-frame conversions, and -- on a box with two GPUs -- tools/grid_run.py under torchrun with 2 ranks."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from libavif_b200 import abi, sharding
from tests import util

pytestmark = pytest.mark.gpu
A = abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    import torch
    return torch.cuda.device_count()


def test_row_range_conversion_equals_whole_frame(gpu_product, oracle):
    import torch
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for k, spec in enumerate([
            util.Spec(w=700, h=98, depth=10, fmt=A.AVIF_PIXEL_FORMAT_YUV420, rng=0, mc=A.MC_BT709, rgb_depth=10, ups=A.AVIF_CHROMA_UPSAMPLING_BILINEAR),
            util.Spec(w=333, h=77, depth=8, fmt=A.AVIF_PIXEL_FORMAT_YUV420, alpha=True, img_premult=True, rgb_depth=8, ups=A.AVIF_CHROMA_UPSAMPLING_BILINEAR),
            util.Spec(w=512, h=64, depth=12, fmt=A.AVIF_PIXEL_FORMAT_YUV422, rgb_depth=16, rgb_fmt=A.AVIF_RGB_FORMAT_RGB),
            util.Spec(w=300, h=50, depth=8, fmt=A.AVIF_PIXEL_FORMAT_YUV444, mc=A.MC_YCGCO, rgb_depth=8, rgb_fmt=A.AVIF_RGB_FORMAT_BGRA)]):
        ro, want = util.yuv_to_rgb(oracle, spec, 60 + k)
        assert ro == 0
        img = util.make_yuv(spec, 60 + k, device="cuda:0")
        rgb = util.make_rgb_dst(spec, device="cuda:0")
        cuts = [0, 2, 30, 31 if spec.fmt != A.AVIF_PIXEL_FORMAT_YUV420 else 32, spec.h]
        for a, b in zip(cuts[:-1], cuts[1:]):
            assert gpu_product.avifB200ImageYUVToRGBRowsAsync(img.ptr(), rgb.ptr(), a, b, sptr) == 0
        torch.cuda.synchronize()
        assert util.first_mismatch(want, rgb.full_bytes()) is None, spec.ident()
    # argument checks
    assert gpu_product.avifB200ImageYUVToRGBRowsAsync(img.ptr(), rgb.ptr(), 10, 5, sptr) == A.AVIF_RESULT_INVALID_ARGUMENT


def test_cell_rects_follow_the_reference_split_rule(gpu_product):
    for (w, h, cols, rows, fmt) in [(7680, 4320, 4, 4, 3), (1000, 700, 3, 2, 3), (999, 701, 5, 3, 1), (130, 129, 2, 2, 3), (4096, 64, 9, 1, 2), (100, 100, 4, 4, 1)]:
        img = A.avifImage()
        img.width, img.height, img.yuvFormat, img.depth = w, h, fmt, 8
        try:
            want = sharding.split_grid(w, h, cols, rows, fmt)
        except ValueError:
            want = None
        for u in range(cols * rows):
            r = A.avifCropRect()
            res = gpu_product.avifB200GridCellRect(ctypes.byref(img), cols, rows, u, ctypes.byref(r))
            if want is None:
                assert res == A.AVIF_RESULT_INVALID_ARGUMENT
            else:
                assert res == 0 and (r.x, r.y, r.width, r.height) == (want[u].x, want[u].y, want[u].width, want[u].height)


@pytest.mark.parametrize("geom", [(640, 384, 4, 4, 3, 8), (1000, 700, 3, 2, 3, 10), (515, 260, 2, 2, 1, 12), (900, 300, 5, 1, 2, 8)])
def test_single_gpu_grid_equals_whole_canvas(gpu_product, oracle, geom):
    import torch
    w, h, cols, rows, fmt, depth = geom
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    spec = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=1, mc=A.MC_BT601, alpha=True, rgb_depth=8 if depth == 8 else depth, rgb_fmt=A.AVIF_RGB_FORMAT_RGBA)
    ro, want = util.rgb_to_yuv(oracle, spec, 5)
    rgb = util.make_rgb_src(spec, 5, device="cuda:0")
    canvas = util.make_yuv_dst(spec, device="cuda:0", with_alpha=True)
    phase = (ctypes.c_float * 4)()
    assert gpu_product.avifB200GridRGBToYUV(canvas.ptr(), rgb.ptr(), cols, rows, None, 0, 1, 0, sptr, ctypes.byref(phase)) == 0
    torch.cuda.synchronize()
    for p in range(4):
        assert util.first_mismatch(want[p], canvas.bufs[p].host().reshape(-1)) is None, (geom, p)
    # decode side, one rank: the row-band path degenerates to one band
    spec2 = util.Spec(w=w, h=h, depth=depth, fmt=fmt, rng=0, mc=A.MC_BT709, alpha=True, rgb_depth=depth, rgb_fmt=A.AVIF_RGB_FORMAT_RGBA, ups=A.AVIF_CHROMA_UPSAMPLING_BILINEAR)
    ro, want2 = util.yuv_to_rgb(oracle, spec2, 6)
    img = util.make_yuv(spec2, 6, device="cuda:0")
    dst = util.make_rgb_dst(spec2, device="cuda:0")
    assert gpu_product.avifB200GridYUVToRGB(img.ptr(), dst.ptr(), None, 0, 1, 0, sptr, None) == 0
    torch.cuda.synchronize()
    assert util.first_mismatch(want2, dst.full_bytes()) is None


@pytest.mark.skipif("_n_devices() < 2")
def test_two_gpu_grid_over_nccl():
    """tools/grid_run.py under torchrun, 2 ranks: cells / row bands travel over NCCL, result == one-GPU whole-canvas conversion"""
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tools", "grid_run.py"), "--check-only"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["encode_check"] == "ok" and line["encode_check_staged"] == "ok", line
    assert line["decode_check"] == "ok" and line["decode_check_peer"] == "ok", line
    assert line["encode_mode_staged"] == "staged NCCL exchange", line  # AVIF_B200_GRID_PEER=0 is honoured; the default may be either form
