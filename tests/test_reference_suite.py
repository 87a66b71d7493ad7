"""The reference's own exact-equality / bounded-drift tests for this path, restated (gtest is not available here):
tests/gtest/avifrgbtoyuvtest.cc (identity :598-645, monochrome :665-712, YCgCo-Re :715-726, gray->4:2:0 half chroma
:415-466), avifalphapremtest.cc:15-88 (opaque alpha == no alpha), avifrgbtoyuvthreadingtest.cc:30-106 (thread count never
changes bytes), tests/avifyuv.c (-m drift :116-323, -m premultiply :506-616; MAX_DRIFT 5).
Run against the oracle on CPU and against the CUDA library (host and device pointers) on the GPU box."""
import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage

A = abi

TARGETS = [
    pytest.param(("oracle", None), id="oracle"),
    pytest.param(("product", None), id="cuda-host", marks=pytest.mark.gpu),
    pytest.param(("product", "cuda:0"), id="cuda-device", marks=pytest.mark.gpu),
]


@pytest.fixture(params=TARGETS)
def target(request, oracle):
    which, device = request.param
    if which == "oracle":
        return oracle, None
    import libavif_b200
    lib = libavif_b200.lib()
    assert libavif_b200.device_available(), "GPU test selected without a CUDA device"
    return lib, device


def _rgb_with(samples, depth, fmt, device, **kw):
    h, w, _ = samples.shape
    rgb = RGBImage(w, h, depth, fmt, device=device, **kw)
    dt = np.uint16 if depth > 8 else np.uint8
    rgb.set_payload(samples.astype(dt).reshape(h, -1).view(np.uint8))
    return rgb


def _rgb_samples(rgb):
    c = rgb.c
    nch = rgb.pixel_bytes // (2 if c.depth > 8 else 1)
    raw = rgb.payload_bytes()
    return (raw.view(np.uint16) if c.depth > 8 else raw).reshape(c.height, c.width, nch).astype(np.int64)


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
@pytest.mark.parametrize("fmt", [A.AVIF_RGB_FORMAT_RGB, A.AVIF_RGB_FORMAT_RGBA, A.AVIF_RGB_FORMAT_BGRA])
def test_identity_round_trip_is_lossless(target, depth, fmt):
    lib, device = target
    rs = np.random.RandomState(depth)
    nch = abi.rgb_format_channel_count(fmt)
    src = rs.randint(0, 1 << depth, size=(37, 53, nch))
    rgb = _rgb_with(src, depth, fmt, device)
    img = YUVImage(53, 37, depth, A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_RANGE_FULL, A.MC_IDENTITY, alpha=(nch == 4), device=device)
    assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
    out = RGBImage(53, 37, depth, fmt, device=device)
    assert lib.avifImageYUVToRGB(img.ptr(), out.ptr()) == 0
    assert (_rgb_samples(out) == src).all()


@pytest.mark.parametrize("depth", [8, 10])
def test_ycgco_re_round_trip_is_lossless(target, depth):
    lib, device = target
    rs = np.random.RandomState(depth + 100)
    src = rs.randint(0, 1 << depth, size=(19, 64, 3))
    rgb = _rgb_with(src, depth, A.AVIF_RGB_FORMAT_RGB, device)
    img = YUVImage(64, 19, depth + 2, A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_RANGE_FULL, A.MC_YCGCO_RE, device=device)
    assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
    out = RGBImage(64, 19, depth, A.AVIF_RGB_FORMAT_RGB, device=device)
    assert lib.avifImageYUVToRGB(img.ptr(), out.ptr()) == 0
    assert (_rgb_samples(out) == src).all()


@pytest.mark.parametrize("depth,rgb_depth", [(8, 8), (10, 10), (12, 16), (16, 8)])
@pytest.mark.parametrize("rng", [0, 1])
def test_monochrome_equals_neutral_chroma(target, depth, rgb_depth, rng):
    """4:0:0 -> RGB has R == G == B and equals the 4:4:4 conversion of the same luma with chroma at the half value."""
    lib, device = target
    rs = np.random.RandomState(depth * 3 + rng)
    luma = rs.randint(0, 1 << depth, size=(21, 70))
    mono = YUVImage(70, 21, depth, A.AVIF_PIXEL_FORMAT_YUV400, rng, A.MC_BT709, device=device)
    mono.set_plane_samples(0, luma)
    full = YUVImage(70, 21, depth, A.AVIF_PIXEL_FORMAT_YUV444, rng, A.MC_BT709, device=device)
    full.set_plane_samples(0, luma)
    full.set_plane_samples(1, np.full_like(luma, 1 << (depth - 1)))
    full.set_plane_samples(2, np.full_like(luma, 1 << (depth - 1)))
    a = RGBImage(70, 21, rgb_depth, A.AVIF_RGB_FORMAT_RGB, device=device)
    b = RGBImage(70, 21, rgb_depth, A.AVIF_RGB_FORMAT_RGB, device=device)
    assert lib.avifImageYUVToRGB(mono.ptr(), a.ptr()) == 0
    assert lib.avifImageYUVToRGB(full.ptr(), b.ptr()) == 0
    sa = _rgb_samples(a)
    assert (sa[..., 0] == sa[..., 1]).all() and (sa[..., 1] == sa[..., 2]).all()
    assert (sa == _rgb_samples(b)).all()


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
@pytest.mark.parametrize("fmt", [A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV422, A.AVIF_PIXEL_FORMAT_YUV444])
def test_gray_input_gives_half_chroma(target, depth, fmt):
    lib, device = target
    rs = np.random.RandomState(depth)
    src = rs.randint(0, 1 << depth, size=(23, 41, 1))
    rgb = _rgb_with(src, depth, A.AVIF_RGB_FORMAT_GRAY, device)
    img = YUVImage(41, 23, depth, fmt, A.AVIF_RANGE_FULL, A.MC_BT601, device=device)
    assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
    assert (img.plane_samples(1) == (1 << (depth - 1))).all()
    assert (img.plane_samples(2) == (1 << (depth - 1))).all()
    assert (img.plane_samples(0) == src[..., 0]).all()  # full range, same depth: luma is the gray code


def test_opaque_alpha_equals_no_alpha(target):
    """avifalphapremtest.cc: an image with an opaque alpha plane converts to the same RGB bytes as one without."""
    lib, device = target
    rs = np.random.RandomState(5)
    w, h = 128, 96
    planes = [rs.randint(0, 256, size=(h, w)) for _ in range(3)]
    outs = []
    for with_alpha in (False, True):
        img = YUVImage(w, h, 8, A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=with_alpha, device=device)
        for p in range(3):
            img.set_plane_samples(p, planes[p])
        if with_alpha:
            img.set_plane_samples(3, np.full((h, w), 255))
        for fmt, prem in ((A.AVIF_RGB_FORMAT_RGB, False), (A.AVIF_RGB_FORMAT_RGBA, True), (A.AVIF_RGB_FORMAT_RGBA, False)):
            rgb = RGBImage(w, h, 8, fmt, alpha_premultiplied=prem, device=device)
            assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
            outs.append(rgb.payload_bytes().copy())
    for k in range(3):
        assert (outs[k] == outs[k + 3]).all()


@pytest.mark.parametrize("fmt", [A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV420])
def test_thread_count_never_changes_bytes(target, fmt):
    lib, device = target
    rs = np.random.RandomState(0xAABB)
    w, h = 200, 127
    img = YUVImage(w, h, 8, fmt, A.AVIF_RANGE_LIMITED, A.MC_BT601, alpha=True, device=device)
    for p in img.present_planes():
        pw, ph = img.plane_dims(p)
        img.set_plane_samples(p, rs.randint(0, 256, size=(ph, pw)))
    outs = []
    for threads in (1, 2, 7, 0):
        rgb = RGBImage(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, upsampling=A.AVIF_CHROMA_UPSAMPLING_FASTEST, max_threads=threads, device=device)
        assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
        outs.append(rgb.payload_bytes().copy())
    assert all((o == outs[0]).all() for o in outs[1:])


@pytest.mark.parametrize("mc", [A.MC_BT709, A.MC_BT601, A.MC_BT2020_NCL, A.MC_YCGCO, A.MC_SMPTE240])
@pytest.mark.parametrize("rng", [0, 1])
def test_rgb_cube_drift_below_5(target, mc, rng):
    """avifyuv -m drift on the full 8-bit RGB lattice with step 3 (+ the corners): |RGB -> YUV444 -> RGB - RGB| < 5."""
    lib, device = target
    if mc == A.MC_YCGCO and rng == 0:
        pytest.skip("YCgCo is full range only")
    v = np.unique(np.concatenate([np.arange(0, 256, 3), [255]]))
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    src = np.stack([r.ravel(), g.ravel(), b.ravel()], axis=-1)
    n = src.shape[0]
    w = 1024
    hh = (n + w - 1) // w
    pad = np.zeros((hh * w, 3), dtype=np.int64)
    pad[:n] = src
    src = pad.reshape(hh, w, 3)
    rgb = _rgb_with(src, 8, A.AVIF_RGB_FORMAT_RGB, device)
    img = YUVImage(w, hh, 8, A.AVIF_PIXEL_FORMAT_YUV444, rng, mc, device=device)
    assert lib.avifImageRGBToYUV(img.ptr(), rgb.ptr()) == 0
    out = RGBImage(w, hh, 8, A.AVIF_RGB_FORMAT_RGB, device=device)
    assert lib.avifImageYUVToRGB(img.ptr(), out.ptr()) == 0
    assert np.abs(_rgb_samples(out) - src).max() < 5


@pytest.mark.parametrize("depth", [8, 10, 12])
def test_premultiply_round_trip_drift_below_5(target, depth):
    """avifyuv -m premultiply: for every alpha a and colour r <= a: unpremultiply then premultiply drifts by < 5."""
    lib, device = target
    maxv = (1 << depth) - 1
    step = 1 if depth == 8 else (5 if depth == 10 else 37)
    a_vals = np.unique(np.concatenate([np.arange(0, maxv + 1, step), [maxv]]))
    rows = []
    for a in a_vals:
        r = np.unique(np.concatenate([np.arange(0, a + 1, step), [a]]))
        rows.append(np.stack([r, r, r, np.full_like(r, a)], axis=-1))
    src = np.concatenate(rows)
    n = src.shape[0]
    w = 1000
    hh = (n + w - 1) // w
    pad = np.zeros((hh * w, 4), dtype=np.int64)
    pad[:n] = src
    pad[n:, 3] = maxv
    src = pad.reshape(hh, w, 4)
    rgb = _rgb_with(src, depth, A.AVIF_RGB_FORMAT_RGBA, device)
    assert lib.avifRGBImageUnpremultiplyAlpha(rgb.ptr()) == 0
    assert lib.avifRGBImagePremultiplyAlpha(rgb.ptr()) == 0
    assert np.abs(_rgb_samples(rgb) - src).max() < 5
