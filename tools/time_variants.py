#!/usr/bin/env python
"""Times the batched launch of every BASELINE configuration for ONE build of the library (AVIF_B200_LIB selects it) and checks
the converted bytes of each against tests/golden/bench_vectors.json.  One JSON line on stdout.  Used by tools/variant_sweep.sh."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import workloads as wl  # noqa: E402
from libavif_b200.images import pointer_array  # noqa: E402

only = set(sys.argv[1].split(",")) if len(sys.argv) > 1 and sys.argv[1] != "all" else None
lib = L.lib()
dev = "cuda:0"
torch.zeros(1, device=dev)
stream = torch.cuda.current_stream()
sptr = ctypes.c_void_p(stream.cuda_stream)
gold = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_vectors.json")))


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


def timed(fn, reps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


out = {"lib": os.environ.get("AVIF_B200_LIB", "default")}
PLAN = [("C2", 8), ("NS", 4), ("C3", 4), ("C5", 8), ("C5_nearest", 8), ("C1", 24), ("C1_nearest", 24), ("F16", 8), ("RGB565", 8)]
for name, nf in PLAN:
    if only and name not in only:
        continue
    cfg = wl.CONFIGS[name]
    imgs, rgbs = [], []
    for f in range(nf):
        img = wl.make_yuv(name, device=dev)
        fill(img, 100 + f)
        imgs.append(img)
        rgbs.append(wl.make_rgb(name, device=dev))
    pi, pr = pointer_array(imgs), pointer_array(rgbs)
    ms = timed(lambda: lib.avifB200ImageYUVToRGBBatchAsync(pi, pr, nf, sptr))
    wl.fill_yuv_seeded(imgs[0], gold[name]["seed"])
    assert lib.avifB200ImageYUVToRGBBatchAsync(pi, pr, nf, sptr) == 0
    torch.cuda.synchronize()
    ok = wl.rgb_digest(rgbs[0]) == gold[name]["rgb"]
    out[name] = {"us": round(ms * 1e3 / nf, 2), "ok": ok, "k": lib.avifB200LastKernelName().decode()}
    if cfg.get("back"):
        dsts = [wl.make_yuv(name, device=dev) for _ in range(nf)]
        pd = pointer_array(dsts)
        assert lib.avifB200ImageRGBToYUVBatchAsync(pd, pr, nf, sptr) == 0
        torch.cuda.synchronize()
        ok = wl.yuv_digest(dsts[0]) == gold[name]["yuv_back"]
        ms = timed(lambda: lib.avifB200ImageRGBToYUVBatchAsync(pd, pr, nf, sptr))
        out[name + "_back"] = {"us": round(ms * 1e3 / nf, 2), "ok": ok, "k": lib.avifB200LastKernelName().decode()}
        del dsts
    del imgs, rgbs
    torch.cuda.empty_cache()
print(json.dumps(out), flush=True)
