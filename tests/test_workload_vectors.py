"""tests/golden/bench_vectors.json (digests of the compiled reference's outputs for BASELINE.json's full-size configurations on
seeded inputs, libavif_b200/workloads.py) against the oracle port -- the smaller configurations, so that the CPU suite stays short --
and, where /root/reference was compiled, against the reference again (the generator is deterministic)."""
import json
import os

import pytest

from libavif_b200 import workloads as wl

VECTORS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bench_vectors.json")))
SMALL = ["C1", "C1_nearest", "C4_cell"]


def run_config(lib, name, device=None, pinned=False):
    """Converts the seeded input of configuration `name` with `lib`; returns {"rgb": digest, "yuv_back": digest} as applicable."""
    cfg, out = wl.CONFIGS[name], {}
    if cfg.get("back_only"):
        rgb = wl.make_rgb(name, device=device, pinned=pinned)
        wl.fill_rgb_seeded(rgb, cfg["seed"])
    else:
        img = wl.make_yuv(name, device=device, pinned=pinned)
        wl.fill_yuv_seeded(img, cfg["seed"])
        rgb = wl.make_rgb(name, device=device, pinned=pinned)
        assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
        out["rgb"] = wl.rgb_digest(rgb)
    if cfg.get("back") or cfg.get("back_only"):
        dst = wl.make_yuv(name, device=device, pinned=pinned)
        assert lib.avifImageRGBToYUV(dst.ptr(), rgb.ptr()) == 0
        out["yuv_back"] = wl.yuv_digest(dst)
    return out


def test_every_configuration_has_a_vector():
    assert set(VECTORS) == set(wl.CONFIGS)


@pytest.mark.parametrize("name", SMALL)
def test_oracle_port_reproduces_the_vectors(oracle, name):
    got = run_config(oracle, name)
    for k, v in got.items():
        assert VECTORS[name][k] == v, (name, k)


@pytest.mark.parametrize("name", ["C1", "C4_cell", "RGB565"])
def test_compiled_reference_reproduces_the_vectors(reference, name):
    got = run_config(reference, name)
    for k, v in got.items():
        assert VECTORS[name][k] == v, (name, k)
