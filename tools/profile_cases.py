#!/usr/bin/env python
"""Launches the kernel of every BASELINE.json configuration exactly as bench.py times it -- the BATCHED launch, same frames
per launch -- twice each, in a fixed order, so that `ncu -k regex:"y2r_|r2y_|tone_map|scale_"` captures them; writes the order
(case name, pixels and frames per launch, algorithmic bytes) to gpurun_out/<tag>/cases.json for tools/ncu_table.py.
Used by tools/gpu_round.sh."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import abi as A  # noqa: E402
from libavif_b200 import sharding, workloads as wl  # noqa: E402
from libavif_b200.images import RGBImage, YUVImage, pointer_array  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/cases.json"
lib = L.lib()
dev = "cuda:0"
torch.zeros(1, device=dev)
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
order = []


def fill(img, seed):
    for p in img.present_planes():
        b = img.bufs[p]
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 7 + p)
        if img.bps == 2:
            v = torch.randint(0, 1 << img.c.depth, (b.nbytes // 2,), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
            b._t[:b.nbytes].view(torch.int16).copy_(v)
        else:
            b._t[:b.nbytes].copy_(torch.randint(0, 256, (b.nbytes,), generator=g, device=dev, dtype=torch.int32).to(torch.uint8))


PLAN = [("C2", 8), ("NS", 4), ("C3", 4), ("C5", 8), ("C5_nearest", 8), ("C1", 24), ("C1_nearest", 24), ("F16", 8), ("RGB565", 8)]
ONLY = set(sys.argv[sys.argv.index("--only") + 1].split(",")) if "--only" in sys.argv else None
if ONLY:
    PLAN = [p for p in PLAN if p[0] in ONLY]
for name, nf in PLAN:
    cfg = wl.CONFIGS[name]
    w, h = wl.dims(name)
    imgs, rgbs = [], []
    for f in range(nf):
        img = wl.make_yuv(name, device=dev)
        fill(img, 100 + f)
        imgs.append(img)
        rgbs.append(wl.make_rgb(name, device=dev))
    pi, pr = pointer_array(imgs), pointer_array(rgbs)
    lib.avifB200ResetKernelLaunchCount()
    for _ in range(2):
        assert lib.avifB200ImageYUVToRGBBatchAsync(pi, pr, nf, sptr) == 0
    torch.cuda.synchronize()
    order.append({"case": name, "kernel": lib.avifB200LastKernelName().decode(), "frames": nf, "px_per_frame": w * h, "alg_bytes_per_px": cfg["bytes_per_px"],
                  "launches": int(lib.avifB200KernelLaunchCount())})
    if cfg.get("back"):
        dsts = [wl.make_yuv(name, device=dev) for _ in range(nf)]
        pd = pointer_array(dsts)
        lib.avifB200ResetKernelLaunchCount()
        for _ in range(2):
            assert lib.avifB200ImageRGBToYUVBatchAsync(pd, pr, nf, sptr) == 0
        torch.cuda.synchronize()
        order.append({"case": name + "_back", "kernel": lib.avifB200LastKernelName().decode(), "frames": nf, "px_per_frame": w * h, "alg_bytes_per_px": cfg["bytes_per_px"],
                      "launches": int(lib.avifB200KernelLaunchCount())})
        del dsts
    del imgs, rgbs
    torch.cuda.empty_cache()

if ONLY and not (ONLY & {"SC", "TM", "C4"}):
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump(order, open(out_path, "w"), indent=1)
    sys.exit(0)


def want(tag):
    return ONLY is None or tag in ONLY


# C4: the 16 cells of the 4x4 grid as views of one canvas, one batched RGB -> YUV launch
def run_c4():
    canvas_rgb = RGBImage(7680, 4320, 8, A.AVIF_RGB_FORMAT_RGBA, device=dev)
    canvas_rgb.buf._t[:canvas_rgb.buf.nbytes].copy_(torch.randint(0, 256, (canvas_rgb.buf.nbytes,), device=dev, dtype=torch.int32).to(torch.uint8))
    canvas = YUVImage(7680, 4320, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    rects = sharding.split_grid(7680, 4320, 4, 4, A.AVIF_PIXEL_FORMAT_YUV420)
    cy = [sharding.view_yuv(canvas.c, r) for r in rects]
    cr = [sharding.view_rgb(canvas_rgb.c, r) for r in rects]
    pi = (ctypes.POINTER(A.avifImage) * 16)(*[ctypes.pointer(c) for c in cy])
    pr = (ctypes.POINTER(A.avifRGBImage) * 16)(*[ctypes.pointer(c) for c in cr])
    lib.avifB200ResetKernelLaunchCount()
    for _ in range(2):
        assert lib.avifB200ImageRGBToYUVBatchAsync(pi, pr, 16, sptr) == 0
    torch.cuda.synchronize()
    order.append({"case": "C4_canvas", "kernel": lib.avifB200LastKernelName().decode(), "frames": 16, "px_per_frame": 1920 * 1080, "alg_bytes_per_px": 6.5,
                  "launches": int(lib.avifB200KernelLaunchCount())})



if want("C4"):
    run_c4()

if "--reformat-only" not in sys.argv and want("TM"):
    # gain-map tone mapping: 4K sRGB RGBA8 base + 8-bit 4:4:4 gain map -> PQ RGBA10, device resident
    import numpy as np  # noqa: E402
    g = np.random.RandomState(77)
    W, H = 3840, 2160
    base = RGBImage(W, H, 8, 1, device=dev)
    base.buf.set(g.randint(0, 256, size=(base.buf.rows, base.buf.row_bytes)).astype(np.uint8))
    gimg = YUVImage(W, H, 8, 1, 1, 6, alpha=False, device=dev)
    for p in gimg.present_planes():
        b = gimg.bufs[p]
        b.set(g.randint(0, 256, size=(b.rows, b.row_bytes)).astype(np.uint8))
    gm = A.avifGainMap()
    gm.image = ctypes.pointer(gimg.c)
    for c in range(3):
        gm.gainMapMin[c].n, gm.gainMapMin[c].d = 0, 1
        gm.gainMapMax[c].n, gm.gainMapMax[c].d = 3, 1
        gm.gainMapGamma[c].n, gm.gainMapGamma[c].d = 1, 1
        gm.baseOffset[c].n, gm.baseOffset[c].d = 1, 64
        gm.alternateOffset[c].n, gm.alternateOffset[c].d = 1, 64
    gm.baseHdrHeadroom.n, gm.baseHdrHeadroom.d = 0, 1
    gm.alternateHdrHeadroom.n, gm.alternateHdrHeadroom.d = 3, 1
    gm.useBaseColorSpace, gm.altColorPrimaries = 1, 2
    dst = RGBImage(W, H, 10, 1, device=dev)
    clli, diag = A.avifContentLightLevelInformationBox(), A.avifDiagnostics()
    for _ in range(2):
        assert lib.avifB200RGBImageApplyGainMapDevice(base.ptr(), 1, 13, ctypes.byref(gm), ctypes.c_float(2.0), 9, 16, dst.ptr(),
                                                      ctypes.byref(clli), ctypes.byref(diag)) == 0, diag.error
    torch.cuda.synchronize()
    order.append({"case": "TM_4K", "kernel": "tone_map (+ y2r of the gain map, + tone_lut)", "frames": 1, "px_per_frame": W * H, "alg_bytes_per_px": 15, "match": "tone_map_kernel"})

if "--reformat-only" not in sys.argv and want("SC"):
    # avifImageScale's device scaler: 4K 4:2:0 8-bit + alpha -> 720p (box) and 1080p -> 4K (2x bilinear)
    for (sw, sh, dw, dh) in ((3840, 2160, 1280, 720), (1920, 1080, 3840, 2160), (3840, 2160, 1920, 1080), (1280, 720, 3840, 2160)):
        s_img = YUVImage(sw, sh, 8, 3, 1, 1, alpha=True, device=dev)
        fill(s_img, 9)
        d_img = YUVImage(dw, dh, 8, 3, 1, 1, alpha=True, device=dev)
        for _ in range(2):
            assert lib.avifB200ImageScaleToDevice(s_img.ptr(), d_img.ptr()) == 0
        torch.cuda.synchronize()
        order.append({"case": f"SC_{sw}x{sh}_to_{dw}x{dh}", "kernel": "scale", "frames": 1, "px_per_frame": sw * sh + dw * dh, "alg_bytes_per_px": 2.5, "match": "scale_", "launches": 2})

os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
json.dump(order, open(out_path, "w"), indent=1)
for o in order:
    print(o["case"], o["kernel"])
