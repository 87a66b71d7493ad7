"""Pins the oracle (and, on the GPU box, the CUDA path) against known answers that came out of the reference itself:
SURVEY.md Appendix C single-pixel KATs and whole-frame hashes, and the committed vectors of tests/golden/
(generated from the compiled reference by tests/golden/make_golden.py).  Runs anywhere -- needs neither /root/reference
nor oracle/_ref."""
import json
import os

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage
from tests import util

A = abi
VECTORS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))

KAT_Y2R = [  # yuv depth, MC, range, (Y,U,V), rgb depth -> (R,G,B)   
    (8, 6, 0, (16, 128, 128), 8, (0, 0, 0)), (8, 6, 0, (235, 128, 128), 8, (255, 255, 255)), (8, 6, 0, (81, 90, 240), 8, (254, 0, 0)),
    (8, 6, 0, (145, 54, 34), 8, (0, 255, 1)), (8, 6, 0, (41, 240, 110), 8, (0, 0, 255)), (8, 6, 0, (100, 100, 100), 8, (53, 132, 41)),
    (8, 6, 1, (100, 100, 100), 8, (61, 130, 50)), (10, 1, 1, (512, 300, 700), 10, (808, 464, 119)), (10, 1, 0, (512, 300, 700), 8, (215, 117, 18)),
    (12, 9, 0, (2000, 1500, 2500), 12, (2800, 1846, 860)), (12, 9, 1, (2000, 1500, 2500), 16, (42674, 29317, 15507)),
    (8, 0, 1, (10, 20, 30), 8, (30, 10, 20)), (8, 8, 1, (100, 140, 90), 8, (50, 112, 126)),
]
KAT_R2Y = [  # rgb depth, (R,G,B), yuv depth, MC, range -> (Y,U,V)
    (8, (255, 0, 0), 8, 6, 0, (81, 90, 240)), (8, (0, 255, 0), 8, 6, 0, (145, 54, 34)), (8, (0, 0, 255), 8, 6, 0, (41, 240, 110)),
    (8, (255, 128, 0), 8, 6, 1, (151, 43, 202)), (10, (1023, 512, 0), 10, 1, 1, (584, 197, 791)), (12, (4095, 2047, 1), 12, 9, 0, (2364, 902, 3016)),
    (8, (200, 100, 50), 10, 1, 0, (468, 384, 696)),
]
FRAME_HASHES = {  # name: (w,h,depth,fmt,range,mc,alpha,img_premult,rgb_depth,rgb_fmt,upsampling,rgb_premult,seed), in-hash, out-hash
    "G1a": ((1920, 1080, 8, 3, 0, 6, False, False, 8, 1, 4, False, 1), 0xd8c41c5edae2a9db, 0xbe2d9874701b4ce2),
    "G1b": ((1920, 1080, 8, 3, 0, 6, False, False, 8, 1, 3, False, 1), 0xd8c41c5edae2a9db, 0x0d2f2eb74b548364),
    "G2a": ((3840, 2160, 10, 1, 1, 1, True, False, 10, 1, 0, True, 2), 0x003e7888b2412f05, 0xf47a37093a2029be),
    "G5": ((3840, 2160, 10, 3, 0, 1, False, False, 16, 0, 0, False, 5), 0x90804e1e3df1c849, 0x8512bd6346d339c7),
    "GN": ((7680, 4320, 10, 1, 1, 1, False, False, 10, 1, 0, False, 7), 0x19734f45d63ebb3e, 0x3294f95469fdc527),
}


def _one_pixel_y2r(lib, depth, mc, rng, yuv, rgb_depth):
    img = YUVImage(1, 1, depth, A.AVIF_PIXEL_FORMAT_YUV444, rng, mc)
    for p in range(3):
        img.set_plane_samples(p, np.array([[yuv[p]]]))
    rgb = RGBImage(1, 1, rgb_depth, A.AVIF_RGB_FORMAT_RGB)
    assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
    raw = rgb.payload_bytes()
    return tuple(int(x) for x in (raw.view(np.uint16) if rgb_depth > 8 else raw))


def _one_pixel_r2y(lib, rgb_depth, rgbv, depth, mc, rng):
    rgb = RGBImage(1, 1, rgb_depth, A.AVIF_RGB_FORMAT_RGB)
    dt = np.uint16 if rgb_depth > 8 else np.uint8
    rgb.set_payload(np.array(rgbv, dtype=dt).view(np.uint8).reshape(1, -1))
    img = YUVImage(1, 1, depth, A.AVIF_PIXEL_FORMAT_YUV444, rng, mc)
    assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
    return tuple(int(img.plane_samples(p)[0, 0]) for p in range(3))


@pytest.mark.parametrize("kat", KAT_Y2R)
def test_kat_yuv_to_rgb(oracle, kat):
    depth, mc, rng, yuv, rgb_depth, want = kat
    assert _one_pixel_y2r(oracle, depth, mc, rng, yuv, rgb_depth) == want


@pytest.mark.parametrize("kat", KAT_R2Y)
def test_kat_rgb_to_yuv(oracle, kat):
    rgb_depth, rgbv, depth, mc, rng, want = kat
    assert _one_pixel_r2y(oracle, rgb_depth, rgbv, depth, mc, rng) == want


def _frame(lib, oracle, args):
    w, h, depth, fmt, rng, mc, alpha, ip, rd, rf, ups, rp, seed = args
    img = YUVImage(w, h, depth, fmt, rng, mc, alpha=alpha, alpha_premultiplied=ip)
    util.xorshift_fill_image(oracle, img, seed)
    rgb = RGBImage(w, h, rd, rf, upsampling=ups, alpha_premultiplied=rp)
    assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
    return util.fnv1a64(oracle, img.payload_bytes()), util.fnv1a64(oracle, rgb.payload_bytes()), rgb


@pytest.mark.parametrize("name", list(FRAME_HASHES))
def test_frame_hashes(oracle, name):
    args, want_in, want_out = FRAME_HASHES[name]
    got_in, got_out, _ = _frame(oracle, oracle, args)
    assert (got_in, got_out) == (want_in, want_out)


def test_c2_round_trip_hash(oracle):
    _, _, rgb = _frame(oracle, oracle, FRAME_HASHES["G2a"][0])
    back = YUVImage(3840, 2160, 10, A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_RANGE_FULL, A.MC_BT709, alpha=True, alpha_premultiplied=False)
    assert oracle.avifImageRGBToYUV(back.ptr(), rgb.ptr()) == 0
    assert util.fnv1a64(oracle, back.payload_bytes()) == 0x1f2ee19d048166ce


def _check_vectors(lib, oracle, device=None):
    n = 0
    for v in VECTORS["y2r"]:
        spec = util.Spec(**v["spec"])
        code, data = util.yuv_to_rgb(lib, spec, seed=v["seed"], device=device)
        assert code == v["result"], spec.ident()
        if code == 0:
            n += 1
            assert f"{util.fnv1a64(oracle, data):016x}" == v["fnv"], spec.ident()
    for v in VECTORS["r2y"]:
        spec = util.Spec(**v["spec"])
        code, planes = util.rgb_to_yuv(lib, spec, seed=v["seed"], device=device)
        assert code == v["result"], spec.ident()
        if code == 0:
            n += 1
            h = util.FNV_OFFSET
            for p in planes:
                if p is not None:
                    h = util.fnv1a64(oracle, p, h)
            assert f"{h:016x}" == v["fnv"], spec.ident()
    return n


def test_oracle_reproduces_reference_vectors(oracle):
    assert _check_vectors(oracle, oracle) > 450


@pytest.mark.gpu
@pytest.mark.parametrize("device", [None, "cuda:0"], ids=["host", "device"])
def test_cuda_reproduces_reference_vectors(gpu_product, oracle, device):
    assert _check_vectors(gpu_product, oracle, device=device) > 450
