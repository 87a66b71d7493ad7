"""Pins oracle/avif_oracle_scale.c (sequential CPU restatement of avifImageScale + the libyuv loops behind it) to the reference
compiled from its sources, bit for bit: every loop, 8/10/12-bit, every chroma layout, alpha, random size pairs.  No GPU.
(The product's table-plus-kernel form is checked against the same reference in tests/test_gpu_scale.py and, host half only,
in tests/test_host_scale_plan.py.)"""
import numpy as np
import pytest

from tests import test_gpu_scale as ts

A = ts.A


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_every_loop_bit_exact(oracle, reference, depth):
    for k, ((w, h), (dw, dh)) in enumerate(ts.CASES):
        for fmt, alpha in ((A.AVIF_PIXEL_FORMAT_YUV444, True), (A.AVIF_PIXEL_FORMAT_YUV420, False), (A.AVIF_PIXEL_FORMAT_YUV400, True)):
            ts.check(oracle, reference, w, h, depth, fmt, alpha, 1000 + k, dw, dh)


def test_random_size_pairs_bit_exact(oracle, reference):
    rs = np.random.RandomState(4242)
    fmts = [A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV422, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400]
    for k in range(300):
        w, h = int(rs.randint(1, 120)), int(rs.randint(1, 120))
        kind = k % 4
        if kind == 0:
            dw, dh = int(rs.randint(1, 160)), int(rs.randint(1, 160))
        elif kind == 1:
            dw, dh = max(1, w // int(rs.randint(2, 7))), max(1, h // int(rs.randint(2, 7)))
        elif kind == 2:
            dw, dh = w * 2 - int(rs.randint(0, 2)), h * 2 - int(rs.randint(0, 2))
        else:
            dw, dh = w + int(rs.randint(0, 40)), max(1, h - int(rs.randint(0, 40)))
        ts.check(oracle, reference, w, h, (8, 10, 12, 8)[k % 4], fmts[(k // 4) % 4], bool(k & 1), 7000 + k, max(1, dw), max(1, dh))


def test_argument_codes(oracle, reference):
    for dw, dh in ((0, 10), (10, 0), (40000, 10), (20000, 20000), (16, 16)):
        ra, _, _, _ = ts.scaled_planes(reference, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, dw, dh)
        rb, _, _, _ = ts.scaled_planes(oracle, 16, 16, 8, A.AVIF_PIXEL_FORMAT_YUV444, False, 1, dw, dh)
        assert ra == rb, (dw, dh, ra, rb)
