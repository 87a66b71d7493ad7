#!/usr/bin/env python
"""BASELINE.json config C4: 4x4 grid of 1920x1080 cells (7680x4320 RGBA8 canvas on rank 0) -> 4:2:0 8-bit BT.601 full + alpha,
cells sharded across the ranks with NCCL point-to-point scatter / gather (libavif_b200/grid.py).
Run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_grid.py   (N = 1, 2, 4 ...)
Prints one JSON line on rank 0: whole-canvas Mpixels/s with and without the staging, and a hash check against the
single-GPU whole-canvas conversion (tile-by-tile must be bit-identical)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import abi as A, grid  # noqa: E402
from libavif_b200.images import RGBImage, YUVImage  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(dev))
lib = L.lib()
torch.zeros(1, device=dev)
stream = torch.cuda.current_stream()
sptr = ctypes.c_void_p(stream.cuda_stream)
W, H, COLS, ROWS = 7680, 4320, 4, 4
canvas = None
if rank == 0:
    canvas = RGBImage(W, H, 8, A.AVIF_RGB_FORMAT_RGBA, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    canvas.buf._t[:canvas.buf.nbytes].copy_(torch.randint(0, 256, (canvas.buf.nbytes,), generator=g, device=dev, dtype=torch.int32).to(torch.uint8))
args = (W, H, COLS, ROWS, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, 8, A.AVIF_RGB_FORMAT_RGBA, dev, sptr)
D = dist if world > 1 else None
out = None
for _ in range(3):
    out = grid.grid_rgb_to_yuv(lib, D, canvas, *args)
torch.cuda.synchronize()
steps, tot, parts = 10, 0.0, {}
for _ in range(steps):
    timers = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    out = grid.grid_rgb_to_yuv(lib, D, canvas, *args, timers=timers)
    torch.cuda.synchronize()
    ms = timers[0][1].elapsed_time(timers[-1][1])
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tot += ms
    for (n0, e0), (n1, e1) in zip(timers[:-1], timers[1:]):
        parts[n1] = parts.get(n1, 0.0) + e0.elapsed_time(e1)
if rank == 0:
    whole = YUVImage(W, H, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    assert lib.avifB200ImageRGBToYUVAsync(whole.ptr(), canvas.ptr(), sptr) == 0
    torch.cuda.synchronize()
    same = all(torch.equal(whole.bufs[p]._t, out.bufs[p]._t) for p in whole.present_planes())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10):
        lib.avifB200ImageRGBToYUVAsync(whole.ptr(), canvas.ptr(), sptr)
    e1.record(stream)
    torch.cuda.synchronize()
    print(json.dumps({
        "metric": "Mpixels/s RGB->YUV, 4x4 grid of 1920x1080 cells (C4), canvas on rank 0, NCCL p2p tile scatter + plane gather",
        "n_gpus": world, "value": W * H / (tot / steps * 1e-3) / 1e6, "unit": "Mpixels/s", "ms_per_canvas": tot / steps,
        "ms_parts_rank0": {k: v / steps for k, v in parts.items()},
        "single_gpu_whole_canvas_kernel_ms": e0.elapsed_time(e1) / 10,
        "tiles_equal_whole_canvas": bool(same), "alg_bytes_per_px": 6.5}))
if world > 1:
    dist.destroy_process_group()
