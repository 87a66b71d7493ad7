#!/usr/bin/env python
"""GPU experiments batched into one gpurun call: (1) PCIe copy rates (uni/bi-directional, by chunk size) that bound the
host-pointer path; (2) kernel variants of the two headline 4:4:4 16-bit configurations (policy bits 3-4) timed with CUDA
events and checked byte-for-byte against the production kernel."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import abi as A  # noqa: E402
from libavif_b200.images import RGBImage, YUVImage  # noqa: E402

dev = "cuda:0"
out = {}

# ---------------- PCIe ----------------
pcie = {}
for mb in (1, 4, 16, 64, 256):
    n = mb << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(4, 512 // mb)

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d1.copy_(h1, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9
    run(True, True)
    pcie[f"{mb}MB"] = {"h2d_GBps": run(True, False), "d2h_GBps": run(False, True), "both_each_GBps": run(True, True)}
out["pcie"] = pcie

# ---------------- kernel variants ----------------
lib = L.lib()
torch.zeros(1, device=dev)
stream = torch.cuda.current_stream()
sptr = ctypes.c_void_p(stream.cuda_stream)
CASES = {
    "NS": (7680, 4320, 10, 1, 1, 1, False, False, 10, 1, 0, False, 14),
    "C2fwd": (3840, 2160, 10, 1, 1, 1, True, False, 10, 1, 0, True, 16),
    "C3": (7680, 4320, 12, 3, 0, 9, True, True, 12, 1, 4, False, 13),
    "C5bil": (3840, 2160, 10, 3, 0, 1, False, False, 16, 0, 0, False, 9),
}
variants = {}
for name, (w, h, d, fmt, rng, mc, alpha, ip, rd, rf, ups, rp, bpp) in CASES.items():
    frames = []
    for f in range(4):
        img = YUVImage(w, h, d, fmt, rng, mc, alpha=alpha, alpha_premultiplied=ip, device=dev)
        for p in img.present_planes():
            b = img.bufs[p]
            v = torch.randint(0, 1 << d, (b.nbytes // 2,), device=dev, dtype=torch.int32).to(torch.int16)
            b._t[:b.nbytes].view(torch.int16).copy_(v)
        frames.append((img, RGBImage(w, h, rd, rf, upsampling=ups, alpha_premultiplied=rp, device=dev)))
    ref_hash = None
    for pol, label in ((0, "production_persistent_prefetch"), (8, "one_tile_per_warp")):
        lib.avifB200SetKernelPolicy(pol)
        for img, rgb in frames:
            assert lib.avifB200ImageYUVToRGBAsync(img.ptr(), rgb.ptr(), sptr) == 0
        torch.cuda.synchronize()
        kname = lib.avifB200LastKernelName().decode()
        hsh = int(frames[0][1].buf._t.to(torch.int64).sum().item()) ^ int((frames[0][1].buf._t[::7].to(torch.int64) * 3).sum().item())
        if ref_hash is None:
            ref_hash = hsh
            full_ref = frames[0][1].buf._t.clone()
        same = bool(torch.equal(frames[0][1].buf._t, full_ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        reps = 10
        for _ in range(reps):
            for img, rgb in frames:
                lib.avifB200ImageYUVToRGBAsync(img.ptr(), rgb.ptr(), sptr)
        e1.record(stream)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(frames))
        variants[f"{name}/{label}"] = {"kernel": kname, "us": us, "alg_GBps": bpp * w * h / us / 1e3, "bytes_equal_production": same}
    lib.avifB200SetKernelPolicy(0)
    del frames
    torch.cuda.empty_cache()
out["variants"] = variants

# ---------------- host-pointer path: staged band pipeline vs zero-copy (kernels access pinned host memory) ----------------
e2e = {}
w, h, d = 3840, 2160, 10
hsrc = YUVImage(w, h, d, 1, 1, 1, alpha=True, pinned=True)
for p in hsrc.present_planes():
    import numpy as np
    b = hsrc.bufs[p]
    b.flat[b.offset:b.offset + b.nbytes].view(np.uint16)[:] = np.random.RandomState(p).randint(0, 1024, size=b.nbytes // 2, dtype=np.uint16)
hrgb = RGBImage(w, h, d, 1, alpha_premultiplied=True, pinned=True)
hback = YUVImage(w, h, d, 1, 1, 1, alpha=True, pinned=True)
ref_bytes = None
for pol, label in ((0, "band_pipeline_default"), (32, "zero_copy")):
    lib.avifB200SetKernelPolicy(pol)
    for _ in range(3):
        assert lib.avifImageYUVToRGB(hsrc.ptr(), hrgb.ptr()) == 0
        assert lib.avifImageRGBToYUV(hback.ptr(), hrgb.ptr()) == 0
    cur = hrgb.full_bytes().copy()
    if ref_bytes is None:
        ref_bytes = cur
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        lib.avifImageYUVToRGB(hsrc.ptr(), hrgb.ptr())
    t1 = time.perf_counter()
    for _ in range(n):
        lib.avifImageRGBToYUV(hback.ptr(), hrgb.ptr())
    t2 = time.perf_counter()
    e2e[label] = {"y2r_ms": (t1 - t0) / n * 1e3, "r2y_ms": (t2 - t1) / n * 1e3, "Mpx_s_round_trip": 2 * w * h / ((t2 - t0) / n) / 1e6,
                  "bytes_equal": bool((cur == ref_bytes).all())}
lib.avifB200SetKernelPolicy(0)
out["e2e_host_path"] = e2e
print(json.dumps(out, indent=1))
