#!/usr/bin/env python
"""Launches each BASELINE.json configuration's kernel twice (device-resident planes) so that
`ncu -k regex:"y2r_|r2y_"` captures them; prints the kernel order.  Used by tools/gpu_round.sh."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import abi as A  # noqa: E402
from libavif_b200.images import RGBImage, YUVImage  # noqa: E402

lib = L.lib()
dev = "cuda:0"
torch.zeros(1, device=dev)
sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

CASES = [
    # name, w, h, depth, fmt, range, mc, alpha, img_premult, rgb_depth, rgb_fmt, upsampling, rgb_premult, back?
    ("NS", 7680, 4320, 10, 1, 1, 1, False, False, 10, 1, 0, False, False),
    ("C2", 3840, 2160, 10, 1, 1, 1, True, False, 10, 1, 0, True, True),
    ("C3", 7680, 4320, 12, 3, 0, 9, True, True, 12, 1, 4, False, True),
    ("C5bil", 3840, 2160, 10, 3, 0, 1, False, False, 16, 0, 0, False, False),
    ("C5near", 3840, 2160, 10, 3, 0, 1, False, False, 16, 0, 3, False, False),
    ("C1bil", 1920, 1080, 8, 3, 0, 6, False, False, 8, 1, 4, False, False),
    ("C4cell", 1920, 1080, 8, 3, 1, 6, True, False, 8, 1, 0, False, True),
]
for name, w, h, d, fmt, rng, mc, alpha, ip, rd, rf, ups, rp, back in CASES:
    img = YUVImage(w, h, d, fmt, rng, mc, alpha=alpha, alpha_premultiplied=ip, device=dev)
    for p in img.present_planes():
        b = img.bufs[p]
        if d > 8:
            v = torch.randint(0, 1 << d, (b.nbytes // 2,), device=dev, dtype=torch.int32).to(torch.int16)
            b._t[:b.nbytes].view(torch.int16).copy_(v)
        else:
            b._t[:b.nbytes].copy_(torch.randint(0, 256, (b.nbytes,), device=dev, dtype=torch.int32).to(torch.uint8))
    rgb = RGBImage(w, h, rd, rf, upsampling=ups, alpha_premultiplied=rp, device=dev)
    for _ in range(2):
        assert lib.avifB200ImageYUVToRGBAsync(img.ptr(), rgb.ptr(), sptr) == 0
    torch.cuda.synchronize()
    print(name, "fwd", lib.avifB200LastKernelName().decode())
    if back:
        dst = YUVImage(w, h, d, fmt, rng, mc, alpha=alpha, alpha_premultiplied=ip, device=dev)
        for _ in range(2):
            assert lib.avifB200ImageRGBToYUVAsync(dst.ptr(), rgb.ptr(), sptr) == 0
        torch.cuda.synchronize()
        print(name, "back", lib.avifB200LastKernelName().decode())

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
print("TM tone_map_kernel", clli.maxCLL, clli.maxPALL)

# avifImageScale's device scaler: 4K 4:2:0 8-bit + alpha -> 720p (box) and 1080p -> 4K (2x bilinear)
for (sw, sh, dw, dh) in ((3840, 2160, 1280, 720), (1920, 1080, 3840, 2160)):
    s_img = YUVImage(sw, sh, 8, 3, 1, 1, alpha=True, device=dev)
    for p in s_img.present_planes():
        b = s_img.bufs[p]
        b._t[:b.nbytes].copy_(torch.randint(0, 256, (b.nbytes,), device=dev, dtype=torch.int32).to(torch.uint8))
    d_img = YUVImage(dw, dh, 8, 3, 1, 1, alpha=True, device=dev)
    for _ in range(2):
        assert lib.avifB200ImageScaleToDevice(s_img.ptr(), d_img.ptr()) == 0
    torch.cuda.synchronize()
    print("SC", sw, sh, "->", dw, dh)
