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
