#!/usr/bin/env python
"""BASELINE config 4 through the C entry points avifB200GridRGBToYUV / avifB200GridYUVToRGB, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/grid_run.py [--check-only]
(or plainly `python tools/grid_run.py` for one GPU).  The root's result is compared byte for byte with the conversion of the whole
canvas on one GPU; then both directions are timed (CUDA events inside the library: scatter / convert / gather / assemble).
One JSON line on stdout (rank 0)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import libavif_b200 as L  # noqa: E402
from libavif_b200 import abi as A  # noqa: E402
from libavif_b200 import gridcomm, workloads as wl  # noqa: E402
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
comm = gridcomm.make_comm(lib, dist, rank, world, device=dev)
check_only = "--check-only" in sys.argv
out = {"n_gpus": world}


def fill(buf, seed, depth):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    if depth > 8:
        v = torch.randint(0, 1 << depth, (buf.nbytes // 2,), generator=g, device=dev, dtype=torch.int32).to(torch.int16)
        buf._t[:buf.nbytes].view(torch.int16).copy_(v)
    else:
        buf._t[:buf.nbytes].copy_(torch.randint(0, 256, (buf.nbytes,), generator=g, device=dev, dtype=torch.int32).to(torch.uint8))


def descriptor_only(img):
    """what a non-root rank passes: same fields, no planes (row strides kept: they say which planes exist)"""
    c = A.avifImage()
    ctypes.memmove(ctypes.byref(c), ctypes.byref(img.c), ctypes.sizeof(c))
    if rank != 0:
        for p in range(3):
            c.yuvPlanes[p] = None
        c.alphaPlane = None
    return c


# ---------------- encode side: 4x4 grid of 1920x1080 cells, RGBA8 -> 4:2:0 8-bit BT.601 full + alpha ----------------
W, H, cols, rows = 7680, 4320, 4, 4
rgb = RGBImage(W, H, 8, A.AVIF_RGB_FORMAT_RGBA, device=dev)
canvas = YUVImage(W, H, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
if rank == 0:
    fill(rgb.buf, 4, 8)
cdesc, rdesc = descriptor_only(canvas), A.avifRGBImage()
ctypes.memmove(ctypes.byref(rdesc), ctypes.byref(rgb.c), ctypes.sizeof(rdesc))
if rank != 0:
    rdesc.pixels = None
phase = (ctypes.c_float * 4)()


def grid_encode(timed=None):
    r = lib.avifB200GridRGBToYUV(ctypes.byref(cdesc), ctypes.byref(rdesc), cols, rows, comm, rank, world, 0, sptr, ctypes.byref(phase) if timed is not None else None)
    assert r == 0, r
    if timed is not None:
        timed.append(list(phase))


def zero_canvas():
    for p_ in canvas.present_planes():
        canvas.bufs[p_]._t.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


whole = None
if rank == 0:
    whole = YUVImage(W, H, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    assert lib.avifB200ImageRGBToYUVAsync(whole.ptr(), rgb.ptr(), sptr) == 0
    torch.cuda.synchronize()
# both forms of the multi-rank path: peer memory (default where the canvas can be exported) and the staged NCCL exchange
for form, env in (("", None), ("_staged", "0")):
    if env is not None:
        os.environ["AVIF_B200_GRID_PEER"] = env
    zero_canvas()
    grid_encode()
    torch.cuda.synchronize()
    mode = int(lib.avifB200GridLastMode())
    if rank == 0:
        out["encode_check" + form] = "ok" if wl.yuv_digest(whole) == wl.yuv_digest(canvas) else "MISMATCH"
        out["encode_mode" + form] = "peer memory" if mode == 1 else ("staged NCCL exchange" if world > 1 else "one rank")
    if not check_only:
        for _ in range(3):
            grid_encode()
        rec = []
        for _ in range(10):
            grid_encode(rec)
        t = torch.tensor([[sum(x[k] for x in rec) / len(rec) for k in range(4)]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sc, cv, ga, asm_ = [float(v) for v in t[0]]
        px = W * H
        out["encode" + form] = {"workload": "C4: 4x4 grid of 1920x1080 cells, RGBA8 -> 4:2:0 8-bit BT.601 full + alpha, cells dealt to (cell mod N)",
                                "mode": out.get("encode_mode" + form),
                                "ms_scatter_or_descriptor": sc, "ms_convert": cv, "ms_gather_or_tokens": ga, "ms_assemble": asm_, "ms_total": sc + cv + ga + asm_,
                                "Mpixels_per_s_kernel": px / (cv * 1e-3) / 1e6, "Mpixels_per_s_kernel_plus_nccl": px / ((sc + cv + ga + asm_) * 1e-3) / 1e6,
                                "timing": "CUDA events inside avifB200GridRGBToYUV, mean of 10 calls, max over ranks per phase"}
    if env is not None:
        del os.environ["AVIF_B200_GRID_PEER"]

# ---------------- decode side: 8K 4:2:0 12-bit + alpha (premultiplied) -> RGBA12, bilinear: row bands with chroma halo ----------------
img = wl.make_yuv("C3", device=dev)
dst = wl.make_rgb("C3", device=dev)
if rank == 0:
    for p in img.present_planes():
        fill(img.bufs[p], 30 + p, 12)
idesc = descriptor_only(img)
ddesc = A.avifRGBImage()
ctypes.memmove(ctypes.byref(ddesc), ctypes.byref(dst.c), ctypes.sizeof(ddesc))
if rank != 0:
    ddesc.pixels = None
phase3 = (ctypes.c_float * 3)()


def grid_decode(timed=None):
    r = lib.avifB200GridYUVToRGB(ctypes.byref(idesc), ctypes.byref(ddesc), comm, rank, world, 0, sptr, ctypes.byref(phase3) if timed is not None else None)
    assert r == 0, r
    if timed is not None:
        timed.append(list(phase3))


ref = None
if rank == 0:
    ref = wl.make_rgb("C3", device=dev)
    assert lib.avifB200ImageYUVToRGBAsync(img.ptr(), ref.ptr(), sptr) == 0
    torch.cuda.synchronize()
# default = staged exchange; the peer-memory form of the decode side is opt-in (measured slower: remote 16-byte strided pixel stores)
for form, env in (("", None), ("_peer", "1")):
    if env is not None:
        os.environ["AVIF_B200_GRID_PEER_DECODE"] = env
    dst.buf._t.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    grid_decode()
    torch.cuda.synchronize()
    mode = int(lib.avifB200GridLastMode())
    if rank == 0:
        out["decode_check" + form] = "ok" if wl.rgb_digest(ref) == wl.rgb_digest(dst) else "MISMATCH"
        out["decode_mode" + form] = "peer memory" if mode == 1 else ("staged NCCL exchange" if world > 1 else "one rank")
    if not check_only:
        for _ in range(3):
            grid_decode()
        rec = []
        for _ in range(10):
            grid_decode(rec)
        t = torch.tensor([[sum(x[k] for x in rec) / len(rec) for k in range(3)]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sc, cv, ga = [float(v) for v in t[0]]
        out["decode" + form] = {"workload": "8K 4:2:0 12-bit + alpha -> RGBA12 bilinear + unmultiply (C3 geometry), N even row bands with one chroma row of halo",
                                "mode": out.get("decode_mode" + form),
                                "ms_scatter": sc, "ms_convert": cv, "ms_gather": ga, "ms_total": sc + cv + ga,
                                "Mpixels_per_s_kernel": 7680 * 4320 / (cv * 1e-3) / 1e6, "Mpixels_per_s_kernel_plus_nccl": 7680 * 4320 / ((sc + cv + ga) * 1e-3) / 1e6}
    if env is not None:
        del os.environ["AVIF_B200_GRID_PEER_DECODE"]
gridcomm.destroy_comm(lib, comm)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
