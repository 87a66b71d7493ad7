"""Host half of the scaler, without a GPU: csrc/scale.c replays libyuv's stepping state machines into position tables; this
test applies the kernel's per-sample arithmetic to those tables in numpy (a restatement of scale_kernels.cu, TEST
INFRASTRUCTURE) and compares the planes with the reference's src/scale.c + vendored libyuv (oracle/_ref), bit for bit.
The device kernel itself is compared with the same reference in tests/test_gpu_scale.py."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import YUVImage

A = abi
VERTICAL, SIMPLE, BIL_DOWN, BIL_UP, UP2_LINEAR, UP2_BILINEAR, BOX = 1, 2, 3, 4, 5, 6, 7


def plan(product, sw, sh, dw, dh, sixteen):
    mode, lin, minbw = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rows = (ctypes.c_int32 * (4 * dh))()
    cols = (ctypes.c_int32 * (4 * dw))()
    assert product.avifB200DebugScalePlan(sw, sh, dw, dh, int(sixteen), ctypes.byref(mode), ctypes.byref(lin), ctypes.byref(minbw), rows, cols) == 0
    return mode.value, lin.value, minbw.value, np.array(rows, dtype=np.int64).reshape(dh, 4), np.array(cols, dtype=np.int64).reshape(dw, 4)


def interp(a, b, f):
    a, b, f = a.astype(np.int64), b.astype(np.int64), np.asarray(f, dtype=np.int64)
    out = (a * (256 - f) + b * f + 128) >> 8
    out = np.where(f == 128, (a + b + 1) >> 1, out)
    return np.where(f == 0, a, out)


def blend(a, b, f, sixteen):
    a, b, f = a.astype(np.int64), b.astype(np.int64), np.asarray(f, dtype=np.int64)
    if sixteen:
        return (a + ((f * (b - a) + 0x8000) >> 16)) & 0xFFFF
    return (a + ((((f >> 9) * (b - a)) + 0x40) >> 7)) & 0xFF


def emulate(src, dw, dh, sixteen, p):
    mode, lin, minbw, R, C = p
    sh, sw = src.shape
    s = src.astype(np.int64)

    def ld(y, x):
        return s[np.clip(y, 0, sh - 1)[:, None], np.clip(x, 0, sw - 1)[None, :]]
    ii = np.arange(dw)
    if mode == VERTICAL:
        yi, yf = R[:, 0], R[:, 1]
        return interp(ld(yi, ii), ld(yi + 1, ii), yf[:, None])
    if mode == SIMPLE:
        return ld(R[:, 0], C[:, 0])
    if mode == BIL_DOWN:
        yi, yf = R[:, 0], (np.zeros(dh, dtype=np.int64) if lin else R[:, 1])
        xi, xf = C[:, 0], C[:, 1]
        a = interp(ld(yi, xi), ld(yi + 1, xi), yf[:, None])
        b = interp(ld(yi, xi + 1), ld(yi + 1, xi + 1), yf[:, None])
        return blend(a, b, xf[None, :], sixteen)
    if mode == BIL_UP:
        ra, rb, yf = R[:, 0], R[:, 1], (np.zeros(dh, dtype=np.int64) if lin else R[:, 2])
        xi, xf = C[:, 0], C[:, 1]
        ha = blend(ld(ra, xi), ld(ra, xi + 1), xf[None, :], sixteen)
        hb = blend(ld(rb, xi), ld(rb, xi + 1), xf[None, :], sixteen)
        return interp(ha, hb, yf[:, None])
    if mode == UP2_LINEAR:
        yi = R[:, 0]
        k = (ii - 1) >> 1
        a, b = ld(yi, k), ld(yi, k + 1)
        out = np.where(((ii - 1) & 1)[None, :] == 1, (a + 3 * b + 2) >> 2, (3 * a + b + 2) >> 2)
        out[:, 0] = ld(yi, np.array([0]))[:, 0]
        out[:, dw - 1] = ld(yi, np.array([(dw - 1) // 2]))[:, 0]
        return out
    if mode == UP2_BILINEAR:
        jj = np.arange(dh)
        r = (jj - 1) >> 1
        odd = ((jj - 1) & 1) == 1
        rs = np.where(odd, r + 1, r)
        rt = np.where(odd, r, r + 1)
        single = (jj == 0) | (r >= sh - 1)
        rs = np.where(jj == 0, 0, np.where(r >= sh - 1, sh - 1, rs))
        rt = np.where(single, rs, rt)
        k = (ii - 1) >> 1
        oddc = ((ii - 1) & 1) == 1
        kn, kf = np.where(oddc, k + 1, k), np.where(oddc, k, k + 1)
        out = (ld(rs, kn) * 9 + ld(rs, kf) * 3 + ld(rt, kn) * 3 + ld(rt, kf) + 8) >> 4
        for i, kk in ((0, 0), (dw - 1, (dw - 1) // 2)):
            out[:, i] = (3 * ld(rs, np.array([kk]))[:, 0] + ld(rt, np.array([kk]))[:, 0] + 2) >> 2
        return out
    assert mode == BOX
    out = np.zeros((dh, dw), dtype=np.int64)
    for j in range(dh):
        iy, bh = int(R[j, 0]), int(R[j, 1])
        acc = s[np.clip(np.arange(iy, iy + bh), 0, sh - 1)].sum(axis=0)
        if not sixteen:
            acc &= 0xFFFF
        for i in range(dw):
            ix, bw, idx = int(C[i, 0]), int(C[i, 1]), int(C[i, 2])
            tot = int(acc[np.clip(np.arange(ix, ix + bw), 0, sw - 1)].sum()) & 0xFFFFFFFF
            scale = 65536 // (max(1, minbw + idx) * bh)
            out[j, i] = ((tot * scale) & 0xFFFFFFFF) >> 16
    return out


def reference_scale(reference, src, dw, dh, depth):
    sh, sw = src.shape
    img = YUVImage(sw, sh, depth, A.AVIF_PIXEL_FORMAT_YUV400, A.AVIF_RANGE_FULL, 1, alpha=False)
    b = img.bufs[0]
    full = np.zeros((b.rows, b.row_bytes), dtype=np.uint8)
    if depth > 8:
        full[:, :sw * 2] = src.astype(np.uint16).view(np.uint8).reshape(sh, sw * 2)
    else:
        full[:, :sw] = src.astype(np.uint8)
    b.set(full)
    diag = A.avifDiagnostics()
    assert reference.avifImageScale(img.ptr(), dw, dh, ctypes.byref(diag)) == 0, diag.error
    rb, ptr = img.c.yuvRowBytes[0], img.c.yuvPlanes[0]
    raw = np.ctypeslib.as_array((ctypes.c_uint8 * (rb * dh)).from_address(ptr)).reshape(dh, rb).copy()
    ctypes.CDLL(None).free(ctypes.c_void_p(ptr))
    return raw[:, :dw * 2].view(np.uint16).astype(np.int64) if depth > 8 else raw[:, :dw].astype(np.int64)


PAIRS = [((64, 64), (64, 48)), ((64, 31), (64, 100)), ((100, 80), (33, 27)), ((128, 96), (32, 24)), ((33, 21), (66, 42)), ((33, 21), (65, 41)),
         ((40, 30), (80, 30)), ((40, 30), (57, 45)), ((3, 2), (60, 50)), ((100, 80), (70, 50)), ((40, 80), (60, 50)), ((90, 60), (30, 60)),
         ((1, 50), (20, 30)), ((1, 1), (5, 7)), ((40, 1), (57, 4)), ((2, 2), (3, 3)), ((150, 149), (7, 5))]


@pytest.mark.parametrize("depth", [8, 12])
def test_position_tables_reproduce_the_reference(product, reference, depth):
    rs = np.random.RandomState(3)
    seen = set()
    pairs = list(PAIRS)
    for _ in range(60):
        pairs.append(((int(rs.randint(1, 90)), int(rs.randint(1, 90))), (int(rs.randint(1, 120)), int(rs.randint(1, 120)))))
    for (sw, sh), (dw, dh) in pairs:
        if (sw, sh) == (dw, dh):
            continue
        src = rs.randint(0, 1 << depth, size=(sh, sw))
        p = plan(product, sw, sh, dw, dh, depth > 8)
        seen.add(p[0])
        want = reference_scale(reference, src, dw, dh, depth)
        if p[0] == 0:
            got = src
        else:
            got = emulate(src, dw, dh, depth > 8, p) & (0xFFFF if depth > 8 else 0xFF)
        assert np.array_equal(got, want), ((sw, sh), (dw, dh), depth, p[0], int((got != want).sum()))
    assert seen >= {VERTICAL, SIMPLE, BIL_DOWN, BIL_UP, UP2_LINEAR, UP2_BILINEAR, BOX}


def test_primaries_matrices_equal_the_reference(product, reference):
    """The tone mapper's RGB -> RGB matrices (double precision, Bradford-adapted): same bits as src/colrconvert.c."""
    ref_fn = reference.raw.avifColorPrimariesComputeRGBToRGBMatrix
    ref_fn.restype = ctypes.c_int
    ref_fn.argtypes = [ctypes.c_uint16, ctypes.c_uint16, ctypes.POINTER(ctypes.c_double)]
    cps = [1, 4, 5, 6, 7, 8, 9, 10, 11, 12, 22, 2, 0, 99]
    for a in cps:
        for b in cps:
            ma, mb = (ctypes.c_double * 9)(), (ctypes.c_double * 9)()
            ra, rb = ref_fn(a, b, ma), product.avifB200DebugPrimariesMatrix(a, b, mb)
            assert bool(ra) == bool(rb), (a, b)
            if ra:
                assert bytes(ma) == bytes(mb), (a, b, list(ma), list(mb))


def test_position_tables_reproduce_the_committed_vectors(product, oracle):
    """The same check against tests/golden/scale_vectors.json (plane hashes recorded from the compiled reference by
    tests/golden/make_golden.py): needs neither /root/reference nor oracle/_ref nor a GPU."""
    import json
    import os
    from tests import test_gpu_scale as ts
    from tests import util
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scale_vectors.json")))["scale"]
    shifts = {A.AVIF_PIXEL_FORMAT_YUV444: (0, 0), A.AVIF_PIXEL_FORMAT_YUV422: (1, 0), A.AVIF_PIXEL_FORMAT_YUV420: (1, 1), A.AVIF_PIXEL_FORMAT_YUV400: (0, 0)}
    checked = 0
    for v in vec[::3]:  # every third vector keeps this under a few seconds; the GPU test walks all of them
        (w, h), (dw, dh), depth, fmt = v["src"], v["dst"], v["depth"], v["fmt"]
        if v["result"] != 0 or (w, h) == (dw, dh):
            continue
        img = ts.make_image(w, h, depth, fmt, v["alpha"], v["seed"])
        sx, sy = shifts[fmt]
        for p in range(4):
            if v["planes"][p] is None:
                continue
            chroma = p in (1, 2)
            sw, sh = ((w + sx) >> sx, (h + sy) >> sy) if chroma else (w, h)
            pw, ph = ((dw + sx) >> sx, (dh + sy) >> sy) if chroma else (dw, dh)
            raw = img.bufs[p].host()
            src = raw[:, :sw * 2].copy().view(np.uint16).astype(np.int64) if depth > 8 else raw[:, :sw].astype(np.int64)
            pl = plan(product, sw, sh, pw, ph, depth > 8)
            got = src if pl[0] == 0 else emulate(src, pw, ph, depth > 8, pl) & (0xFFFF if depth > 8 else 0xFF)
            data = got.astype(np.uint16).view(np.uint8) if depth > 8 else got.astype(np.uint8)
            assert f"{util.fnv1a64(oracle, np.ascontiguousarray(data).reshape(-1)):016x}" == v["planes"][p], (v["src"], v["dst"], depth, fmt, p)
            checked += 1
    assert checked > 100


def test_identities_behind_the_staged_kernel(product):
    """scale_kernels.cu's row-staged loops lean on four facts; each is checked here over its whole domain (no GPU):
    1. 2x bilinear factors through H = 3*near + far: (9a + 3b + 3c + d + 8) >> 4 == (3*(3a + b) + (3c + d) + 8) >> 4 (plain algebra), and
       its edge columns, which read ONE source column, fall out of near == far: (3*(4a) + 4c + 8) >> 4 == (3a + c + 2) >> 2;
    2. 2x linear likewise: (H + 2) >> 2 with H = 4a is a;
    3. 8-bit box column sums of at most 257 rows fit 16 bits, so pairs of them can be added as packed halves of one register;
    4. a box is minBoxWidth or minBoxWidth + 1 columns wide (cols[:, 2] in {0, 1}, cols[:, 1] == minBoxWidth + cols[:, 2]) for every
       geometry the box loop serves."""
    a = np.arange(0, 65536, dtype=np.int64)[:, None]
    c = np.concatenate([np.arange(0, 65536, 257), [1, 2, 3, 65534, 65535]]).astype(np.int64)[None, :]
    assert np.array_equal((3 * (4 * a) + 4 * c + 8) >> 4, (3 * a + c + 2) >> 2)
    assert np.array_equal((4 * a + 2) >> 2, a)
    assert 257 * 255 < 65536 <= 258 * 255
    rng = np.random.RandomState(5)
    seen = 0
    for _ in range(300):
        sw, sh = int(rng.randint(2, 4000)), int(rng.randint(2, 3000))
        dw, dh = int(rng.randint(1, sw)), int(rng.randint(1, sh))
        for sixteen in (False, True):
            mode, lin, minbw, R, C = plan(product, sw, sh, dw, dh, sixteen)
            if mode != BOX:
                continue
            seen += 1
            assert minbw >= 1
            assert set(np.unique(C[:, 2])) <= {0, 1}, (sw, sh, dw, dh)
            assert np.array_equal(C[:, 1], minbw + C[:, 2]), (sw, sh, dw, dh)
            assert np.all(np.diff(C[:, 0]) >= 0) and C[0, 0] >= 0 and C[-1, 0] + C[-1, 1] <= sw
            assert np.all(R[:, 1] >= 1) and np.all(R[:, 0] >= 0) and np.all(R[:, 0] + R[:, 1] <= sh)
    assert seen > 100


def test_column_positions_in_closed_form(product):
    """The staged kernels compute column positions as x0 + i * dx (16.16) when the host vouches for it (AbScaleJob::colsArith): the
    closed form has to reproduce the table for every geometry that claims it, and the everyday geometries have to claim it (else the
    kernels silently fall back to dependent table loads)."""
    def columns(sw, sh, dw, dh, sixteen):
        cf, x0, dx = ctypes.c_int(), ctypes.c_int32(), ctypes.c_int32()
        assert product.avifB200DebugScalePlanColumns(sw, sh, dw, dh, int(sixteen), ctypes.byref(cf), ctypes.byref(x0), ctypes.byref(dx)) == 0
        return cf.value, x0.value, dx.value
    rng = np.random.RandomState(11)
    geoms = [(3840, 2160, 1280, 720), (1920, 1080, 640, 360), (3840, 2160, 1920, 1080), (1280, 720, 3840, 2160), (640, 360, 1920, 1080), (1000, 700, 333, 100)]
    geoms += [(int(rng.randint(2, 3000)), int(rng.randint(2, 2000)), int(rng.randint(1, 3000)), int(rng.randint(1, 2000))) for _ in range(200)]
    claimed = 0
    for k, (sw, sh, dw, dh) in enumerate(geoms):
        for sixteen in (False, True):
            mode, lin, minbw, R, C = plan(product, sw, sh, dw, dh, sixteen)
            cf, x0, dx = columns(sw, sh, dw, dh, sixteen)
            if k < 6 and mode in (SIMPLE, BIL_DOWN, BIL_UP, BOX):
                assert cf == 1, (sw, sh, dw, dh, sixteen, mode)
            if not cf or mode == 0:
                continue
            claimed += 1
            t = x0 + np.arange(dw, dtype=np.int64) * dx
            assert t[-1] + dx < 2 ** 31
            assert np.array_equal(C[:, 0], t >> 16), (sw, sh, dw, dh, mode)
            if mode == BOX:
                assert np.array_equal(C[:, 1], ((t + dx) >> 16) - (t >> 16))
            if mode in (BIL_DOWN, BIL_UP):
                assert np.array_equal(C[:, 1], t & 0xFFFF)
    assert claimed > 150
