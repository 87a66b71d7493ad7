"""BASELINE.json's configurations at FULL size on the GPU: every byte of the output against the digest the compiled reference
produced for the same seeded input (tests/golden/bench_vectors.json) -- resident planes, pinned host buffers and pageable host
buffers (the pinned-ring staging path of shim.c)."""
import json
import os

import pytest

from libavif_b200 import workloads as wl
from tests.test_workload_vectors import VECTORS, run_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(wl.CONFIGS))
def test_full_size_resident(gpu_product, name):
    got = run_config(gpu_product, name, device="cuda:0")
    for k, v in got.items():
        assert VECTORS[name][k] == v, (name, k, gpu_product.avifB200LastKernelName().decode())


@pytest.mark.parametrize("name", ["C2", "C3", "C5", "C1", "C4_cell", "F16"])
@pytest.mark.parametrize("pinned", [True, False], ids=["pinned", "pageable"])
def test_full_size_host_buffers(gpu_product, name, pinned):
    got = run_config(gpu_product, name, pinned=pinned)
    for k, v in got.items():
        assert VECTORS[name][k] == v, (name, k)


def test_pageable_staging_ring_sizes(gpu_product, monkeypatch):
    """chunk sizes from smaller than a row to larger than the frame; rows that are not a multiple of the chunk"""
    for mb in ("1", "3", "64"):
        monkeypatch.setenv("AVIF_B200_RING_MB", mb)
        got = run_config(gpu_product, "C5")
        assert got["rgb"] == VECTORS["C5"]["rgb"], mb
    monkeypatch.setenv("AVIF_B200_STAGING", "0")
    # (read once per process: this only documents the switch; the driver's own pageable path is what round 1 shipped)


@pytest.mark.parametrize("threads", [1, 4])
def test_pageable_staging_keeps_row_padding_and_matches_oracle(gpu_product, oracle, threads):
    """regions large enough for the pinned ring (>= 256 KB), odd row paddings on both sides, with and without helper threads:
    every byte of the destination allocation, padding included, against the oracle"""
    from libavif_b200 import abi
    from tests import util
    specs = [util.Spec(w=1203, h=611, depth=10, fmt=abi.AVIF_PIXEL_FORMAT_YUV420, rng=0, mc=abi.MC_BT709, alpha=True, rgb_depth=10, rgb_fmt=abi.AVIF_RGB_FORMAT_RGBA,
                       ups=abi.AVIF_CHROMA_UPSAMPLING_BILINEAR, yuv_pad=6, rgb_pad=10, max_threads=threads),
             util.Spec(w=2048, h=300, depth=8, fmt=abi.AVIF_PIXEL_FORMAT_YUV444, rgb_depth=8, rgb_fmt=abi.AVIF_RGB_FORMAT_BGRA, yuv_pad=0, rgb_pad=4, max_threads=threads)]
    for k, spec in enumerate(specs):
        ro, want = util.yuv_to_rgb(oracle, spec, 70 + k)
        rp, got = util.yuv_to_rgb(gpu_product, spec, 70 + k)
        assert ro == rp == 0 and util.first_mismatch(want, got) is None, spec.ident()
        ro, wantp = util.rgb_to_yuv(oracle, spec, 80 + k)
        rp, gotp = util.rgb_to_yuv(gpu_product, spec, 80 + k)
        assert ro == rp == 0
        for a, b in zip(wantp, gotp):
            assert (a is None) == (b is None)
            if a is not None:
                assert util.first_mismatch(a, b) is None, spec.ident()
