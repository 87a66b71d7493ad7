"""GPU tests of the entry points around the two big conversions: the alpha seam (avifReformatAlpha / avifFillAlpha), the
mismatched-extent composition, alpha range expansion (src/read.c:6724-6764) and the grid plane staging helpers
(src/avif.c:187-225, src/write.c:1151-1205)."""
import ctypes

import numpy as np
import pytest

from libavif_b200 import abi
from libavif_b200.images import RGBImage, YUVImage
from tests import util

pytestmark = pytest.mark.gpu
A = abi


def _alpha_params(w, h, src, sdepth, soff, spix, srb, dst, ddepth, doff, dpix, drb):
    p = abi.avifAlphaParams()
    p.width, p.height = w, h
    p.srcDepth, p.srcPlane, p.srcRowBytes, p.srcOffsetBytes, p.srcPixelBytes = sdepth, src, srb, soff, spix
    p.dstDepth, p.dstPlane, p.dstRowBytes, p.dstOffsetBytes, p.dstPixelBytes = ddepth, dst, drb, doff, dpix
    return p


@pytest.mark.parametrize("sdepth,ddepth", [(8, 8), (10, 10), (8, 12), (12, 8), (10, 16), (16, 10), (16, 16)])
@pytest.mark.parametrize("direction", ["plane_to_rgba", "rgba_to_plane", "fill"])
def test_alpha_seam(gpu_product, oracle, sdepth, ddepth, direction):
    w, h = 97, 13
    rs = np.random.RandomState(sdepth * 17 + ddepth)
    sb, db = (2 if sdepth > 8 else 1), (2 if ddepth > 8 else 1)
    outs = []
    for lib in (oracle, gpu_product):
        rs = np.random.RandomState(sdepth * 17 + ddepth)
        if direction == "rgba_to_plane":
            srb, spix, soff = w * 4 * sb + 6, 4 * sb, 3 * sb
            drb, dpix, doff = w * db + 4, db, 0
        else:
            srb, spix, soff = w * sb + 2, sb, 0
            drb, dpix, doff = w * 4 * db + 8, 4 * db, 3 * db
        src = rs.randint(0, 256, size=h * srb, dtype=np.uint8)
        if sb == 2:  # keep codes inside the nominal depth most of the time, with some garbage above it
            v = src[: (h * srb // 2) * 2].view(np.uint16)
            v &= (1 << sdepth) - 1 if sdepth < 16 else 0xFFFF
        dst = np.full(h * drb, util.FILL_BYTE, dtype=np.uint8)
        p = _alpha_params(w, h, src.ctypes.data, sdepth, soff, spix, srb, dst.ctypes.data, ddepth, doff, dpix, drb)
        if direction == "fill":
            lib.avifFillAlpha(ctypes.byref(p))
        else:
            lib.avifReformatAlpha(ctypes.byref(p))
        outs.append(dst.copy())
    assert util.first_mismatch(outs[0], outs[1]) is None


@pytest.mark.parametrize("case", ["rgb_larger_fill", "rgb_smaller", "rgb_larger_premult_fill"])
def test_mismatched_extents_follow_the_reference(gpu_product, oracle, case):
    """The reference iterates the alpha channel / post-passes over rgb->width x rgb->height and the colour loop over the
    image's (src/reformat.c:1467-1468 vs :686): when they differ the library composes its standalone passes."""
    iw, ih = 48, 20
    if case == "rgb_smaller":
        rw, rh, alpha, rp = 40, 16, True, True    # alpha + premultiply over the smaller rectangle, colour over the image
    elif case == "rgb_larger_fill":
        rw, rh, alpha, rp = 56, 24, False, False  # fill alpha over the larger rectangle
    else:
        rw, rh, alpha, rp = 56, 24, False, True
    spec = util.Spec(w=iw, h=ih, depth=10, fmt=A.AVIF_PIXEL_FORMAT_YUV444, rng=1, mc=A.MC_BT709, alpha=alpha, rgb_depth=10,
                     rgb_fmt=A.AVIF_RGB_FORMAT_RGBA, rgb_premult=rp)
    outs = []
    for lib, dev in ((oracle, None), (gpu_product, None), (gpu_product, "cuda:0")):
        img = util.make_yuv(spec, 5, device=dev)
        rgb = RGBImage(max(iw, rw), max(ih, rh), 10, A.AVIF_RGB_FORMAT_RGBA, alpha_premultiplied=rp, device=dev, row_pad=16)
        rgb.buf.set(np.full((rgb.buf.rows, rgb.buf.row_bytes), util.FILL_BYTE, dtype=np.uint8))
        rgb.c.width, rgb.c.height = rw, rh
        assert lib.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0
        outs.append(rgb.full_bytes().copy())
    assert util.first_mismatch(outs[0], outs[1]) is None
    assert util.first_mismatch(outs[0], outs[2]) is None


@pytest.mark.parametrize("depth", [8, 10, 12, 16])
def test_limited_to_full_alpha(gpu_product, oracle, depth):
    w, h = 301, 9
    outs = []
    for lib, dev in ((oracle, None), (gpu_product, None), (gpu_product, "cuda:0")):
        img = YUVImage(w, h, depth, A.AVIF_PIXEL_FORMAT_YUV400, alpha=True, device=dev, row_pad=6)
        rs = np.random.RandomState(depth)
        img.set_plane_samples(3, rs.randint(0, 1 << depth, size=(h, w)))
        if lib is oracle:
            assert lib.raw.oracleImageLimitedToFullAlpha(img.ptr()) == 0
        else:
            assert lib.avifB200ImageLimitedToFullAlpha(img.ptr()) == 0
        outs.append(img.plane_samples(3).copy())
    assert (outs[0] == outs[1]).all() and (outs[0] == outs[2]).all()
    if depth in (8, 10, 12):
        assert outs[0].max() == (1 << depth) - 1 and outs[0].min() == 0  # something was clamped at both ends


@pytest.mark.parametrize("fmt", [A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400])
@pytest.mark.parametrize("depth", [8, 12])
def test_grid_plane_staging(gpu_product, fmt, depth):
    """avifImageCopySamples into a view of a canvas (decode-side tile assembly) and avifImageCopyAndPad of a ragged last
    cell (encode side: replicate the right column / bottom row)."""
    import torch
    from libavif_b200 import sharding
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rs = np.random.RandomState(depth + fmt)
    cw, ch = 192, 130
    canvas = YUVImage(cw, ch, depth, fmt, alpha=True, device="cuda:0")
    want = {}
    for p in canvas.present_planes():
        pw, ph = canvas.plane_dims(p)
        want[p] = np.zeros((ph, pw), dtype=np.int64)
    # assemble the canvas from 2x2 tiles (64-aligned cells like avifDecoderDataCopyTileToImage, clipped at the canvas edge)
    for r in sharding.split_grid(cw, ch, 2, 2, fmt):
        tile = YUVImage(r.width, r.height, depth, fmt, alpha=True, device="cuda:0")
        for p in tile.present_planes():
            pw, ph = tile.plane_dims(p)
            s = rs.randint(0, 1 << depth, size=(ph, pw))
            tile.set_plane_samples(p, s)
            sx = 1 if (p in (1, 2) and fmt != A.AVIF_PIXEL_FORMAT_YUV444) else 0
            sy = 1 if (p in (1, 2) and fmt == A.AVIF_PIXEL_FORMAT_YUV420) else 0
            want[p][r.y >> sy:(r.y >> sy) + ph, r.x >> sx:(r.x >> sx) + pw] = s
        view = sharding.view_yuv(canvas.c, r)
        assert gpu_product.avifB200ImageCopySamplesAsync(ctypes.byref(view), tile.ptr(), A.AVIF_PLANES_ALL, stream) == 0
    torch.cuda.synchronize()
    for p in canvas.present_planes():
        assert (canvas.plane_samples(p) == want[p]).all(), f"plane {p}"
    # pad a ragged 100x70 cell to 128x96
    src = YUVImage(100, 70, depth, fmt, alpha=True, device="cuda:0")
    for p in src.present_planes():
        pw, ph = src.plane_dims(p)
        src.set_plane_samples(p, rs.randint(0, 1 << depth, size=(ph, pw)))
    dst = YUVImage(128, 96, depth, fmt, alpha=True, device="cuda:0")
    assert gpu_product.avifB200ImageCopyAndPadAsync(dst.ptr(), src.ptr(), stream) == 0
    torch.cuda.synchronize()
    for p in src.present_planes():
        s = src.plane_samples(p)
        dw, dh = dst.plane_dims(p)
        exp = np.pad(s, ((0, dh - s.shape[0]), (0, dw - s.shape[1])), mode="edge")
        assert (dst.plane_samples(p) == exp).all(), f"plane {p}"


@pytest.mark.parametrize("fmt", [A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_PIXEL_FORMAT_YUV422, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_PIXEL_FORMAT_YUV400])
@pytest.mark.parametrize("depth", [8, 10])
def test_copy_samples_against_the_compiled_reference(gpu_product, reference, fmt, depth):
    """avifB200ImageCopySamplesAsync against the reference's own avifImageCopySamples (src/avif.c:187-225, compiled into
    oracle/_ref): 3x2 tiles of a padded canvas, YUV-only / alpha-only / all planes, destination bytes outside the copied area untouched."""
    import torch
    from libavif_b200 import sharding
    raw = reference.raw
    raw.avifImageCopySamples.restype = None
    raw.avifImageCopySamples.argtypes = [ctypes.POINTER(A.avifImage), ctypes.POINTER(A.avifImage), ctypes.c_uint32]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rs = np.random.RandomState(100 + depth + fmt)
    cw, ch = 334, 150
    for planes in (A.AVIF_PLANES_ALL, A.AVIF_PLANES_YUV, A.AVIF_PLANES_A):
        host = YUVImage(cw, ch, depth, fmt, alpha=True, row_pad=6)
        dev = YUVImage(cw, ch, depth, fmt, alpha=True, device="cuda:0", row_pad=6)
        for p in host.present_planes():
            fill = rs.randint(0, 256, size=(host.bufs[p].rows, host.bufs[p].row_bytes)).astype(np.uint8)
            host.bufs[p].set(fill)
            dev.bufs[p].set(fill)
        for r in sharding.split_grid(cw, ch, 3, 2, fmt):
            t_host = YUVImage(r.width, r.height, depth, fmt, alpha=True, row_pad=2)
            t_dev = YUVImage(r.width, r.height, depth, fmt, alpha=True, device="cuda:0", row_pad=2)
            for p in t_host.present_planes():
                data = rs.randint(0, 256, size=(t_host.bufs[p].rows, t_host.bufs[p].row_bytes)).astype(np.uint8)
                t_host.bufs[p].set(data)
                t_dev.bufs[p].set(data)
            vh, vd = sharding.view_yuv(host.c, r), sharding.view_yuv(dev.c, r)
            raw.avifImageCopySamples(ctypes.byref(vh), t_host.ptr(), planes)
            assert gpu_product.avifB200ImageCopySamplesAsync(ctypes.byref(vd), t_dev.ptr(), planes, stream) == 0
        torch.cuda.synchronize()
        for p in host.present_planes():
            assert np.array_equal(host.bufs[p].host(), dev.bufs[p].host()), (planes, p)


def test_concurrent_host_threads(gpu_product, oracle):
    """The entry points are re-entrant (the reference is exercised under TSan with concurrent calls on different images):
    all library state is thread-local.  8 host threads convert different frames at once, host and device pointers mixed."""
    import concurrent.futures as cf
    specs = util.random_y2r_specs(64, seed=77, sizes=[(640, 48), (1921, 9), (264, 33), (96, 96)])
    want = [util.yuv_to_rgb(oracle, sp, 900 + k) for k, sp in enumerate(specs)]

    def work(k):
        dev = "cuda:0" if k % 3 == 0 else None
        return util.yuv_to_rgb(gpu_product, specs[k], 900 + k, device=dev)

    with cf.ThreadPoolExecutor(max_workers=8) as pool:
        got = list(pool.map(work, range(len(specs))))
    for k, (w, g) in enumerate(zip(want, got)):
        assert w[0] == g[0], specs[k].ident()
        if w[0] == 0:
            assert util.first_mismatch(w[1], g[1]) is None, specs[k].ident()
