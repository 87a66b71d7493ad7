#!/usr/bin/env python
"""bench.py -- measurement harness of the B200 reformat path (contract: see DESIGN.md "Measurement").

Workload (BASELINE.json configs[1], "C2"): avifImageYUVToRGB + avifImageRGBToYUV round trip of 3840x2160 4:4:4 10-bit
BT.709 full-range frames with an alpha plane (straight) <-> RGBA 10-bit premultiplied.  One step = one pass of the hot
path over a batch of B distinct synthetic frames (forward conversion of every frame, then the backward conversion of
every result); the batch working set (B x 265 MB) is far larger than the 126 MB L2, so no flush is needed.

  value      : Mpixels/s, device-resident planes, CUDA events on the launching stream (each direction counts W*H px)
  e2e        : same metric through the public C-ABI (avifImageYUVToRGB / avifImageRGBToYUV) with pinned HOST buffers,
               host<->device copies inside the timed region
  roofline   : dominant kernel's algorithmic bytes / its mean launch time vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference : the reference's own CPU code (oracle/_ref, else the oracle port) on the host cores

N > 1 (torchrun): frames shard across ranks with no data-path collective (weak scaling); time = max over ranks.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, DEPTH = 3840, 2160, 10
PX = W * H
ALG_BYTES_PER_PX = 16  # C2, either direction: 4 planes x 2 B read/written + 8 B RGBA16 written/read


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="frames per step (device-resident leg)")
    ap.add_argument("--e2e-batch", type=int, default=8, help="frames per step (host-buffer leg)")
    ap.add_argument("--e2e-threads", type=int, default=8,
                    help="host threads of the host-buffer leg, each converting its own frames through the synchronous C-ABI "
                         "(the reference arm uses every host core; independent calls from different threads are legal, avif.h:1012)")
    ap.add_argument("--stage-threads", type=int, default=3, help="rgb.maxThreads of the pageable host-buffer leg (calling thread + helpers of the staging copies)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-configuration legs (NS, C3, C5, C1, C4 canvas, C5 batch)")
    ap.add_argument("--extras", action="store_true", help="also time the rows beyond the reformat path (tone mapping, scaling) into 'extras'")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) -- runs during the timed regions


class ClockSampler:
    def __init__(self, index):
        self.index, self.samples, self.reasons, self._stop, self._thr = index, [], set(), threading.Event(), None
        self.max_sm = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)  # the timed region of the default run is only a few milliseconds long

    def start(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_sm, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------
# synthetic frames (SURVEY.md Appendix D distribution: uniform codes), generated with torch on the target memory

WORKLOAD = ("C2 round trip: avifImageYUVToRGB + avifImageRGBToYUV, 3840x2160 YUV444 10-bit BT.709 full + alpha, RGBA10 premultiplied")
METRIC = "Mpixels/s YUV<->RGB (C2 round trip 3840x2160 4:4:4 10-bit + alpha premultiply)"
CONFIG = {"workload": WORKLOAD, "pixel_count": "each direction counts W*H", "sharding": "frames across ranks, no collective",
          "l2": "inputs larger than L2: every step converts distinct frames, 265 MB each"}


def fill_uniform(img, seed):
    """uniform random codes in [0, 2^depth) written straight into the image's buffers (host numpy or device torch)."""
    import numpy as np
    import torch
    mask = (1 << img.c.depth) - 1
    for p in img.present_planes():
        b = img.bufs[p]
        if b.device is None:
            rs = np.random.RandomState(seed * 7 + p)
            if img.bps == 2:
                b.flat[b.offset:b.offset + b.nbytes].view(np.uint16)[:] = rs.randint(0, mask + 1, size=b.nbytes // 2, dtype=np.uint16)
            else:
                b.flat[b.offset:b.offset + b.nbytes] = rs.randint(0, mask + 1, size=b.nbytes, dtype=np.uint8)
        else:
            g = torch.Generator(device=b.device)
            g.manual_seed(seed * 7 + p)
            if img.bps == 2:
                v = torch.randint(0, mask + 1, (b.nbytes // 2,), generator=g, device=b.device, dtype=torch.int32).to(torch.int16)
                b._t[b.offset:b.offset + b.nbytes].view(torch.int16).copy_(v)
            else:
                v = torch.randint(0, mask + 1, (b.nbytes,), generator=g, device=b.device, dtype=torch.int32).to(torch.uint8)
                b._t[b.offset:b.offset + b.nbytes].copy_(v)


def fill_uniform_rgb(rgb, seed):
    import torch
    b = rgb.buf
    g = torch.Generator(device=b.device)
    g.manual_seed(seed)
    if rgb.c.depth > 8:
        v = torch.randint(0, 1 << rgb.c.depth, (b.nbytes // 2,), generator=g, device=b.device, dtype=torch.int32).to(torch.int16)
        b._t[b.offset:b.offset + b.nbytes].view(torch.int16).copy_(v)
    else:
        b._t[b.offset:b.offset + b.nbytes].copy_(torch.randint(0, 256, (b.nbytes,), generator=g, device=b.device, dtype=torch.int32).to(torch.uint8))


# ----------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU code on the host cores


def usable_cpus():
    """Host threads this process may really use: affinity mask, capped by the cgroup CPU quota if there is one."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def cpu_round_trips(seconds_budget, steps, warmup, max_threads=None):
    """Times C2 round trips of the reference CPU implementation with one frame per host thread.
    Returns dict(value Mpx/s, ms_per_step, cores, kind, sample)."""
    import concurrent.futures as cf
    from libavif_b200 import workloads as wl
    from oracle import loader
    ref = loader.load_reference()
    kind = "reference"
    if ref is None:
        ref, kind = loader.load_oracle(), "port"
    cores = usable_cpus()
    threads = max(1, min(cores, max_threads or cores))
    # one shared read-only input frame; per-thread outputs (the conversion never writes its inputs)
    src = wl.make_yuv("C2")
    fill_uniform(src, 1)
    work = []
    for _ in range(threads):
        work.append((wl.make_rgb("C2"), wl.make_yuv("C2")))

    def one(k):
        rgb, back = work[k]
        r1 = ref.avifImageYUVToRGB(src.ptr(), rgb.ptr())
        r2 = ref.avifImageRGBToYUV(back.ptr(), rgb.ptr())
        return r1 | r2

    pool = cf.ThreadPoolExecutor(max_workers=threads)

    def step(n=None):
        assert all(r == 0 for r in pool.map(one, range(n or threads)))

    t_w0 = time.perf_counter()
    for _ in range(max(1, warmup)):
        step()
    warm_s = time.perf_counter() - t_w0
    # the reference arm gets its best configuration: all usable threads, or one per physical core if SMT siblings only fight
    if threads >= 4:
        t0 = time.perf_counter(); step(); t_all_threads = time.perf_counter() - t0
        t0 = time.perf_counter(); step(threads // 2); t_half = time.perf_counter() - t0
        if (threads // 2) / t_half > threads / t_all_threads * 1.05:
            threads = threads // 2
    t_all, done = [], 0
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step(threads)
        t_all.append(time.perf_counter() - t0)
        done += 1
        if seconds_budget and time.perf_counter() - t_start > seconds_budget:
            break
    pool.shutdown()
    total = sum(t_all)
    px = done * threads * 2 * PX
    return {"value": px / total / 1e6, "unit": "Mpixels/s", "ms_per_step": 1e3 * total / done, "cores": threads, "kind": kind,
            "steps_done": done, "warmup_s": warm_s,
            "sample": f"{done} steps x {threads} concurrent C2 round trips (one 3840x2160 frame per host thread, maxThreads=1), "
                      f"{'reference sources compiled -O3 (oracle/_ref)' if kind == 'reference' else 'oracle port'}"}


def run_reference_arm(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    r = cpu_round_trips(seconds_budget=150, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC,
        "value": r["value"], "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": r["steps_done"], "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(CONFIG), "frames_per_step": r["cores"],
        "cpu_baseline": {"value": r["value"], "unit": "Mpixels/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------
# B200 arm


def gpu_local_cpus(index):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), intersected with this process's mask."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.lower().split(":", 1)
        txt = open(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def time_launches(stream, fn, reps):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def golden_vectors():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_vectors.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def traffic_table():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def run_configs(lib, dev, sptr, stream, peak_gbs, rank):
    """The other BASELINE.json configurations (north star NS, C3, C5, C1, C4 cells): kernel time per frame through the batch
    entry point (frames per launch as listed) and through one launch per frame, each against the HBM roofline, plus a check of
    the converted bytes against the digests the compiled reference produced for the same seeded input (tests/golden)."""
    import torch
    from libavif_b200 import workloads as wl
    from libavif_b200.images import pointer_array
    gold = golden_vectors()
    traffic = traffic_table()
    out = {}
    plan = [("NS", 4), ("C3", 4), ("C5", 8), ("C5_nearest", 8), ("C1", 24), ("C1_nearest", 24), ("F16", 8), ("RGB565", 8)]
    for name, nf in plan:
        cfg = wl.CONFIGS[name]
        w, h = wl.dims(name)
        fr = []
        for f in range(nf):
            img = wl.make_yuv(name, device=dev)
            fill_uniform(img, 1000 + 31 * f + rank)
            fr.append((img, wl.make_rgb(name, device=dev)))
        pi, pr = pointer_array([f[0] for f in fr]), pointer_array([f[1] for f in fr])

        def batch():
            assert lib.avifB200ImageYUVToRGBBatchAsync(pi, pr, nf, sptr) == 0

        def single():
            for img, rgb in fr:
                assert lib.avifB200ImageYUVToRGBAsync(img.ptr(), rgb.ptr(), sptr) == 0
        for _ in range(3):
            batch()
            single()
        torch.cuda.synchronize()
        reps = 5
        ms_b = time_launches(stream, batch, reps) / nf
        kname = lib.avifB200LastKernelName().decode()
        ms_s = time_launches(stream, single, reps) / nf
        alg = cfg["bytes_per_px"] * w * h
        entry = {"kernel": kname, "direction": "YUVToRGB", "frames_per_launch": nf, "us_per_frame": ms_b * 1e3, "us_per_frame_single_launch": ms_s * 1e3,
                 "Mpixels_per_s": w * h / (ms_b * 1e-3) / 1e6, "alg_bytes_per_px": cfg["bytes_per_px"],
                 "roofline": {"bound": "hbm", "achieved": alg / (ms_b * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s",
                              "frac": alg / (ms_b * 1e-3) / 1e9 / peak_gbs, "traffic": (traffic.get(name) or {}).get("dram_bytes_per_frame")},
                 "roofline_single_launch": {"achieved": alg / (ms_s * 1e-3) / 1e9, "frac": alg / (ms_s * 1e-3) / 1e9 / peak_gbs}}
        if name == "NS":
            entry["north_star"] = {"target": ">= 0.70 of 8 TB/s", "frac_of_8TBps": alg / (ms_b * 1e-3) / 1e9 / 8000.0,
                                   "frac_of_8TBps_single_launch": alg / (ms_s * 1e-3) / 1e9 / 8000.0}
        if cfg.get("back"):
            bi = [wl.make_yuv(name, device=dev) for _ in range(nf)]
            pb = pointer_array(bi)

            def back():
                assert lib.avifB200ImageRGBToYUVBatchAsync(pb, pr, nf, sptr) == 0
            for _ in range(3):
                back()
            torch.cuda.synchronize()
            ms = time_launches(stream, back, reps) / nf
            entry["back"] = {"kernel": lib.avifB200LastKernelName().decode(), "direction": "RGBToYUV", "us_per_frame": ms * 1e3,
                             "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak_gbs,
                                          "traffic": (traffic.get(name + "_back") or {}).get("dram_bytes_per_frame")}}
        # parity of what was just timed: frame 0 <- the seeded input of the fixture, same entry points, digest of the bytes
        g = gold.get(name)
        if g:
            wl.fill_yuv_seeded(fr[0][0], g["seed"])
            batch()
            torch.cuda.synchronize()
            entry["check"] = "ok" if wl.rgb_digest(fr[0][1]) == g["rgb"] else "MISMATCH"
            if cfg.get("back") and "yuv_back" in g:
                back()
                torch.cuda.synchronize()
                entry["back"]["check"] = "ok" if wl.yuv_digest(bi[0]) == g["yuv_back"] else "MISMATCH"
        out[name] = entry
        del fr
        torch.cuda.empty_cache()
    out["C4_canvas"] = run_c4_single_gpu(lib, dev, sptr, stream, peak_gbs, gold)
    return out


def run_c4_single_gpu(lib, dev, sptr, stream, peak_gbs, gold):
    """C4 on one GPU: the 16 cells of the 4x4 grid (1920x1080 RGBA8) are VIEWS of one 7680x4320 canvas -- source pixels and
    destination planes alike -- converted by ONE batched RGB->YUV launch: no per-cell allocation, no assembly pass."""
    import ctypes
    import torch
    from libavif_b200 import abi as A
    from libavif_b200 import sharding, workloads as wl
    from libavif_b200.images import RGBImage, YUVImage, pointer_array
    cw, ch, cols, rows = 1920, 1080, 4, 4
    W_, H_ = cw * cols, ch * rows
    canvas_rgb = RGBImage(W_, H_, 8, A.AVIF_RGB_FORMAT_RGBA, device=dev)
    fill_uniform_rgb(canvas_rgb, 4)
    canvas = YUVImage(W_, H_, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    rects = sharding.split_grid(W_, H_, cols, rows, A.AVIF_PIXEL_FORMAT_YUV420)
    cells_y = [sharding.view_yuv(canvas.c, r) for r in rects]
    cells_r = [sharding.view_rgb(canvas_rgb.c, r) for r in rects]
    pi = (ctypes.POINTER(A.avifImage) * 16)(*[ctypes.pointer(c) for c in cells_y])
    pr = (ctypes.POINTER(A.avifRGBImage) * 16)(*[ctypes.pointer(c) for c in cells_r])

    def go():
        assert lib.avifB200ImageRGBToYUVBatchAsync(pi, pr, 16, sptr) == 0
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    ms = time_launches(stream, go, 10)
    whole = YUVImage(W_, H_, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    assert lib.avifB200ImageRGBToYUVAsync(whole.ptr(), canvas_rgb.ptr(), sptr) == 0
    torch.cuda.synchronize()
    same = wl.yuv_digest(whole) == wl.yuv_digest(canvas)
    alg = 6.5 * W_ * H_
    return {"kernel": lib.avifB200LastKernelName().decode(), "direction": "RGBToYUV", "cells": 16, "us_per_canvas": ms * 1e3, "Mpixels_per_s": W_ * H_ / (ms * 1e-3) / 1e6,
            "alg_bytes_per_px": 6.5, "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak_gbs, "unit": "GB/s", "frac": alg / (ms * 1e-3) / 1e9 / peak_gbs},
            "check": "ok (cell views == whole-canvas conversion)" if same else "MISMATCH"}


def run_c5_batch(lib, dev, sptr, stream, rank, world, dist, nthreads):
    """BASELINE config 5: 256 synthetic 4K 4:2:0 10-bit frames -> RGB16 (default upsampling = bilinear), frames dealt to the
    ranks (f mod world), no collective; (a) resident: batch launches of <= 24 frames, CUDA events, max over ranks;
    (b) through the synchronous C-ABI with pinned host buffers on a sample of the rank's frames."""
    import torch
    from libavif_b200 import workloads as wl
    from libavif_b200.images import pointer_array
    total = 256
    mine = [f for f in range(total) if f % world == rank]
    w, h = wl.dims("C5")
    srcs, dsts = [], []
    for f in mine:
        img = wl.make_yuv("C5", device=dev)
        fill_uniform(img, 5 + f)
        srcs.append(img)
        dsts.append(wl.make_rgb("C5", device=dev))
    chunks = [(pointer_array(srcs[k:k + 24]), pointer_array(dsts[k:k + 24]), len(srcs[k:k + 24])) for k in range(0, len(srcs), 24)]

    def go():
        for pi, pr, n in chunks:
            assert lib.avifB200ImageYUVToRGBBatchAsync(pi, pr, n, sptr) == 0
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = time_launches(stream, go, 3)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = {"frames": total, "frames_per_gpu": len(mine), "ms_per_pass": ms, "Mpixels_per_s": total * w * h / (ms * 1e-3) / 1e6,
           "achieved_gbs_per_gpu": 9 * len(mine) * w * h / (ms * 1e-3) / 1e9, "kernel": lib.avifB200LastKernelName().decode()}
    del srcs, dsts, chunks
    torch.cuda.empty_cache()
    return res


def run_c4_grid(lib, dev, sptr, stream, rank, world, dist):
    """BASELINE config 4 across the ranks: avifB200GridRGBToYUV (C entry point, NCCL inside the library): the 7680x4320 RGBA8 canvas
    lives on rank 0, its 4x4 cells are dealt to the ranks, converted, and the planes gathered -- kernel time and kernel + NCCL time
    reported separately (CUDA events inside the library, max over ranks); output compared with the one-GPU whole-canvas conversion."""
    import ctypes
    import torch
    from libavif_b200 import abi as A
    from libavif_b200 import gridcomm, workloads as wl
    from libavif_b200.images import RGBImage, YUVImage
    comm = gridcomm.make_comm(lib, dist, rank, world, device=dev)
    W_, H_ = 7680, 4320
    rgb = RGBImage(W_, H_, 8, A.AVIF_RGB_FORMAT_RGBA, device=dev)
    canvas = YUVImage(W_, H_, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
    if rank == 0:
        fill_uniform_rgb(rgb, 4)
    phase = (ctypes.c_float * 4)()

    def go(timed):
        r = lib.avifB200GridRGBToYUV(canvas.ptr(), rgb.ptr(), 4, 4, comm, rank, world, 0, sptr, ctypes.byref(phase) if timed else None)
        return r, list(phase)

    def measure():
        """None if the form failed on any rank (every rank learns it: the verdict is all-reduced, nobody is left waiting)"""
        r, _ = go(False)
        torch.cuda.synchronize()
        bad = torch.tensor([1.0 if r != 0 else 0.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if float(bad.item()) != 0.0:
            return None, int(lib.avifB200GridLastMode())
        for _ in range(3):
            go(False)
        rec = [go(True)[1] for _ in range(10)]
        t = torch.tensor([sum(x[k] for x in rec) / len(rec) for k in range(4)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t], int(lib.avifB200GridLastMode())

    whole = None
    if rank == 0:
        whole = YUVImage(W_, H_, 8, A.AVIF_PIXEL_FORMAT_YUV420, A.AVIF_RANGE_FULL, A.MC_BT601, alpha=True, device=dev)
        assert lib.avifB200ImageRGBToYUVAsync(whole.ptr(), rgb.ptr(), sptr) == 0
        torch.cuda.synchronize()

    def entry(phases, mode):
        if phases is None:
            return {"mode": "peer memory" if mode == 1 else "staged", "error": "avifB200GridRGBToYUV failed on at least one rank in this form"}
        sc, cv, ga, asm_ = phases
        names = ("ms_descriptor", "ms_convert_over_nvlink", "ms_tokens", "ms_assemble") if mode == 1 else ("ms_scatter", "ms_convert", "ms_gather", "ms_assemble")
        e = dict(zip(names, phases))
        e.update({"mode": "peer memory (CUDA IPC views of the root's canvas; NCCL carries descriptor + tokens)" if mode == 1 else "staged (grouped ncclSend/ncclRecv)",
                  "ms_total": sc + cv + ga + asm_, "Mpixels_per_s_kernel": W_ * H_ / (cv * 1e-3) / 1e6,
                  "Mpixels_per_s_kernel_plus_nccl": W_ * H_ / ((sc + cv + ga + asm_) * 1e-3) / 1e6})
        if rank == 0:
            e["check"] = "ok (== whole-canvas conversion on one GPU)" if wl.yuv_digest(whole) == wl.yuv_digest(canvas) else "MISMATCH"
        return e

    # default form first (peer memory where the canvas can be exported), then the staged NCCL exchange for the "kernel + NCCL" column
    phases, mode = measure()
    first = entry(phases, mode)
    for p in canvas.present_planes():
        canvas.bufs[p]._t.zero_()  # the second form has to produce every byte again
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    os.environ["AVIF_B200_GRID_PEER"] = "0"
    phases2, mode2 = measure()
    second = entry(phases2, mode2)
    del os.environ["AVIF_B200_GRID_PEER"]
    res = None
    if rank == 0:
        res = {"workload": "4x4 grid of 1920x1080 RGBA8 cells -> 4:2:0 8-bit BT.601 full + alpha; canvas resident on rank 0, cells dealt round-robin",
               "api": "avifB200GridRGBToYUV", "n_gpus": world}
        res.update(first)
        res["staged"] = second
    gridcomm.destroy_comm(lib, comm)
    return res


def run_b200(args):
    rank, world, local = dist_env()
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist
    import libavif_b200 as L
    from libavif_b200 import workloads as wl
    from libavif_b200.images import pointer_array

    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    lib = L.lib()
    if not L.device_available():
        raise SystemExit("bench.py: libavif_b200 sees no CUDA device (there is no CPU fallback)")
    torch.zeros(1, device=dev)  # make torch's context current before the library's first call
    assert lib.avifB200SetDevice(local) == 0  # this rank's GPU serves the main thread's synchronous calls too
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)

    # ---------------- device-resident leg ----------------
    B = args.batch
    frames = []
    for f in range(B):
        src = wl.make_yuv("C2", device=dev)
        fill_uniform(src, 100 + rank * 1000 + f)
        frames.append((src, wl.make_rgb("C2", device=dev), wl.make_yuv("C2", device=dev)))
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    p_src, p_rgb, p_back = pointer_array([f[0] for f in frames]), pointer_array([f[1] for f in frames]), pointer_array([f[2] for f in frames])

    def step(record=None):
        # one step = the library's batch entry points over the B frames: forward conversion of all, then backward of all
        e0 = ev() if record is not None else None
        if e0:
            e0.record(stream)
        r = lib.avifB200ImageYUVToRGBBatchAsync(p_src, p_rgb, B, sptr)
        assert r == 0, r
        e1 = ev() if record is not None else None
        if e1:
            e1.record(stream)
        r = lib.avifB200ImageRGBToYUVBatchAsync(p_back, p_rgb, B, sptr)
        assert r == 0, r
        if record is not None:
            e2 = ev()
            e2.record(stream)
            record.append((e0, e1, e2))

    for _ in range(args.warmup):
        step()
    step()
    torch.cuda.synchronize()
    # names of the two kernels of the step
    lib.avifB200ImageYUVToRGBAsync(frames[0][0].ptr(), frames[0][1].ptr(), sptr)
    name_fwd = lib.avifB200LastKernelName().decode()
    lib.avifB200ImageRGBToYUVAsync(frames[0][2].ptr(), frames[0][1].ptr(), sptr)
    name_back = lib.avifB200LastKernelName().decode()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    lib.avifB200ResetKernelLaunchCount()
    rec = []
    sampler.start()
    t_begin, t_end = ev(), ev()
    t_begin.record(stream)
    for _ in range(args.steps):
        step(rec)
    t_end.record(stream)
    torch.cuda.synchronize()
    sampler.stop()
    if world > 1:
        dist.barrier()
    launches = int(lib.avifB200KernelLaunchCount())
    elapsed_ms = t_begin.elapsed_time(t_end)
    fwd_ms = sum(a.elapsed_time(b) for a, b, _ in rec) / len(rec)    # one launch converts the whole batch
    back_ms = sum(b.elapsed_time(c) for _, b, c in rec) / len(rec)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = world * args.steps * B * 2 * PX / (elapsed_ms * 1e-3) / 1e6

    # parity of the timed path: frame 0 <- the fixture's seeded input, same two launches, digests against the compiled reference's
    gold = golden_vectors().get("C2")
    check = None
    if gold:
        wl.fill_yuv_seeded(frames[0][0], gold["seed"])
        step()
        torch.cuda.synchronize()
        check = {"rgb": "ok" if wl.rgb_digest(frames[0][1]) == gold["rgb"] else "MISMATCH",
                 "yuv_back": "ok" if wl.yuv_digest(frames[0][2]) == gold["yuv_back"] else "MISMATCH",
                 "against": "tests/golden/bench_vectors.json (digests of the compiled reference's output for the same seeded frame)"}
    del frames, p_src, p_rgb, p_back
    torch.cuda.empty_cache()

    # ---------------- host-buffer (end-to-end) legs: pinned, then pageable ----------------
    Be = args.e2e_batch
    cpus_before = os.sched_getaffinity(0)
    total_usable = usable_cpus()  # affinity mask capped by the cgroup quota, before the NUMA pinning below
    local_cpus = gpu_local_cpus(local)
    if local_cpus:
        os.sched_setaffinity(0, local_cpus)  # pinned buffers are first-touched, and the workers run, next to this rank's GPU
    from concurrent.futures import ThreadPoolExecutor
    nthr = max(1, min(args.e2e_threads, Be, max(1, total_usable // world)))  # ranks share the host's cores

    # helper threads of the pageable leg must fit next to the callers: rgb.maxThreads = calling thread + helpers
    cpus_per_rank = max(1, total_usable // world)
    stage_threads = max(1, min(args.stage_threads, cpus_per_rank // max(1, nthr)))

    def worker_init():
        # every worker thread of this rank talks to this rank's GPU (shim.c: a thread without a choice is dealt a device round-robin)
        assert lib.avifB200SetDevice(local) == 0
    pool = ThreadPoolExecutor(max_workers=nthr, initializer=worker_init)

    def e2e_leg(pinned):
        hframes = []
        for f in range(Be):
            src = wl.make_yuv("C2", pinned=pinned)
            fill_uniform(src, 500 + rank * 1000 + f)
            # pageable leg: rgb.maxThreads = 3 lets the library share the staging copies with two helper threads per call
            # (the reference spends maxThreads threads on the conversion itself, src/reformat.c:1679-1747)
            hframes.append((src, wl.make_rgb("C2", pinned=pinned, max_threads=1 if pinned else stage_threads), wl.make_yuv("C2", pinned=pinned)))

        def e2e_frames(chunk):  # one host thread: its frames, one synchronous round trip after the other
            for src, rgb, back in chunk:
                assert lib.avifImageYUVToRGB(src.ptr(), rgb.ptr()) == 0
                assert lib.avifImageRGBToYUV(back.ptr(), rgb.ptr()) == 0

        def e2e_step():
            # ctypes drops the GIL inside the library; every thread has its own streams and staging arena (shim.c), so one
            # thread's upload overlaps another's kernel and download (PCIe is full duplex)
            list(pool.map(e2e_frames, [hframes[t::nthr] for t in range(nthr)]))

        for _ in range(max(1, args.warmup)):
            e2e_step()
        e2e_steps = max(3, min(args.steps, 20 if pinned else 6))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()  # synchronous API: returns when the results are in host memory
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        ok = None
        if gold:  # what came back over PCIe is checked as well
            src, rgb, back = hframes[0]
            wl.fill_yuv_seeded(src, gold["seed"])
            e2e_frames([hframes[0]])
            ok = "ok" if (wl.rgb_digest(rgb) == gold["rgb"] and wl.yuv_digest(back) == gold["yuv_back"]) else "MISMATCH"
        return world * e2e_steps * Be * 2 * PX / e2e_s / 1e6, e2e_steps, ok

    e2e_value, e2e_steps, e2e_ok = e2e_leg(True)
    pg_value, pg_steps, pg_ok = e2e_leg(False)
    pool.shutdown()
    os.sched_setaffinity(0, cpus_before)
    plane_bytes = PX * 2
    h2d = Be * (4 * plane_bytes + 4 * plane_bytes)  # fwd: Y,U,V,A ; back: RGBA16
    d2h = Be * (4 * plane_bytes + 4 * plane_bytes)  # fwd: RGBA16  ; back: Y,U,V,A

    c5 = run_c5_batch(lib, dev, sptr, stream, rank, world, dist, nthr) if not args.no_configs else None
    c4 = run_c4_grid(lib, dev, sptr, stream, rank, world, dist) if (world > 1 and not args.no_configs) else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    peaks, peak_src = None, "fallback 6650 GB/s (B200_PROFILING.md)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak_gbs, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        peak_gbs = 6650.0
    alg = ALG_BYTES_PER_PX * PX * B  # per launch: the fused batch of B frames
    kern = {
        name_fwd: {"direction": "YUVToRGB", "ms_per_launch": fwd_ms, "achieved_gbs": alg / (fwd_ms * 1e-3) / 1e9},
        name_back: {"direction": "RGBToYUV", "ms_per_launch": back_ms, "achieved_gbs": alg / (back_ms * 1e-3) / 1e9},
    }
    dom = name_fwd if fwd_ms >= back_ms else name_back
    tr = traffic_table().get("C2" if dom == name_fwd else "C2_back") or {}
    traffic = tr.get("dram_bytes_per_frame")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["achieved_gbs"], "peak": peak_gbs, "unit": "GB/s",
                "frac": kern[dom]["achieved_gbs"] / peak_gbs, "traffic": (traffic * B if traffic is not None else None),
                "traffic_source": tr.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "frames_per_launch": B, "kernels": kern}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_round_trips(seconds_budget=20, steps=3, warmup=1)
        cpu = {"value": r["value"], "unit": "Mpixels/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    configs = run_configs(lib, dev, sptr, stream, peak_gbs, rank) if not args.no_configs else None
    extras = run_extras(lib, dev, sptr, stream, peak_gbs) if args.extras else None

    line = {
        "metric": METRIC,
        "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": dict(CONFIG), "frames_per_step_per_gpu": B,
        "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "frames_per_step_per_gpu": Be, "steps": e2e_steps, "host_memory": "pinned", "host_threads": nthr,
                "device_affinity": "every worker thread: avifB200SetDevice(LOCAL_RANK); process pinned to the GPU's NUMA-local CPUs" if local_cpus else "every worker thread: avifB200SetDevice(LOCAL_RANK)",
                "api": "avifImageYUVToRGB / avifImageRGBToYUV (C-ABI, host pointers)", "check": e2e_ok},
        "e2e_pageable": {"value": pg_value, "unit": "Mpixels/s", "steps": pg_steps, "host_memory": "pageable (numpy / malloc), staged through the library's pinned ring", "rgb_maxThreads": stage_threads,
                         "host_threads": nthr, "frac_of_pinned": pg_value / e2e_value if e2e_value else None, "check": pg_ok},
        "gpu_launches": launches,
        "roofline": roofline,
        "check": check,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
    }
    if c5:
        line["c5_batch"] = c5
    if c4:
        line["c4_grid"] = c4
    if configs:
        line["configs"] = configs
    if extras:
        line["extras"] = extras
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extras(lib, dev, sptr, stream, peak_gbs):
    """Rows beyond the reformat path: gain-map tone mapping and plane scaling, device-resident."""
    out = {"TM_4K_gainmap_sRGB8_to_PQ_RGBA10": run_tone_map(lib, dev, stream, peak_gbs)}
    out.update(run_scale(lib, dev, stream, peak_gbs))
    return out


def run_scale(lib, dev, stream, peak_gbs):
    """avifImageScale's device scaler (avifB200ImageScaleToDevice, planes resident) for a box downscale and a 2x
    upscale of 4:2:0 8-bit + alpha, and the reference's src/scale.c + libyuv on one host core for the same sizes."""
    import ctypes
    import numpy as np
    import torch
    from libavif_b200 import abi as A
    from libavif_b200.images import YUVImage
    res = {}
    for name, (sw, sh, dw, dh) in {"SC_4K_to_720p_box_420_8b_alpha": (3840, 2160, 1280, 720),
                                   "SC_1080p_to_4K_up2_420_8b_alpha": (1920, 1080, 3840, 2160)}.items():
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        src = YUVImage(sw, sh, 8, A.AVIF_PIXEL_FORMAT_YUV420, 1, 1, alpha=True, device=dev)
        for p in src.present_planes():
            b = src.bufs[p]
            b._t[:b.nbytes].copy_(torch.randint(0, 256, (b.nbytes,), generator=g, device=dev, dtype=torch.int32).to(torch.uint8))
        dst = YUVImage(dw, dh, 8, A.AVIF_PIXEL_FORMAT_YUV420, 1, 1, alpha=True, device=dev)

        def go():
            assert lib.avifB200ImageScaleToDevice(src.ptr(), dst.ptr()) == 0
        for _ in range(3):
            go()
        torch.cuda.synchronize()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            go()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = 2.5 * (sw * sh + dw * dh)  # Y + U + V (4:2:0) + alpha, read once and written once
        entry = {"api": "avifB200ImageScaleToDevice (4 planes: position tables + one kernel each, synchronous)", "us_per_image": ms * 1e3,
                 "Mpixels_per_s_out": dw * dh / (ms * 1e-3) / 1e6, "alg_bytes": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
                 "frac_of_peak": nbytes / (ms * 1e-3) / 1e9 / peak_gbs}
        try:
            from oracle import loader
            ref = loader.load_reference()
            if ref is not None and hasattr(ref, "avifImageScale"):
                himg = YUVImage(sw, sh, 8, A.AVIF_PIXEL_FORMAT_YUV420, 1, 1, alpha=True)
                rs = np.random.RandomState(5)
                for p in himg.present_planes():
                    b = himg.bufs[p]
                    b.set(rs.randint(0, 256, size=(b.rows, b.row_bytes)).astype(np.uint8))
                diag = A.avifDiagnostics()
                t0 = time.perf_counter()
                r = ref.avifImageScale(himg.ptr(), dw, dh, ctypes.byref(diag))
                dt = time.perf_counter() - t0
                if r == 0:
                    entry["cpu_reference"] = {"Mpixels_per_s_out": dw * dh / dt / 1e6, "cores": 1, "sample": "one call of src/scale.c + vendored libyuv (oracle/_ref)"}
        except Exception as e:
            entry["cpu_reference"] = {"unavailable": str(e)[:100]}
        res[name] = entry
    return res


def run_tone_map(lib, dev, stream, peak_gbs):
    """gain-map tone mapping of a 4K sRGB RGBA8 base with an 8-bit 4:4:4 gain map into PQ RGBA10, everything
    device-resident (avifB200RGBImageApplyGainMapDevice), and the reference's src/gainmap.c on one host core on a 1080p
    crop of the same data (its loop is single-threaded)."""
    import ctypes
    import numpy as np
    import torch
    from libavif_b200 import abi as A
    from libavif_b200.images import RGBImage, YUVImage

    def build(w, h, device):
        g = np.random.RandomState(77)
        base = RGBImage(w, h, 8, A.AVIF_RGB_FORMAT_RGBA, device=device)
        base.buf.set(g.randint(0, 256, size=(base.buf.rows, base.buf.row_bytes)).astype(np.uint8))
        gimg = YUVImage(w, h, 8, A.AVIF_PIXEL_FORMAT_YUV444, A.AVIF_RANGE_FULL, 6, alpha=False, device=device)
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
        return base, gimg, gm

    W, H = 3840, 2160
    base, gimg, gm = build(W, H, dev)
    dst = RGBImage(W, H, 10, A.AVIF_RGB_FORMAT_RGBA, device=dev)
    clli, diag = A.avifContentLightLevelInformationBox(), A.avifDiagnostics()

    def go():
        r = lib.avifB200RGBImageApplyGainMapDevice(base.ptr(), 1, 13, ctypes.byref(gm), ctypes.c_float(2.0), 9, 16, dst.ptr(),
                                                   ctypes.byref(clli), ctypes.byref(diag))
        assert r == 0, diag.error
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        go()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res = {"api": "avifB200RGBImageApplyGainMapDevice (gain map YUV->RGB + tone-map kernel + statistics read-back per call)",
           "us_per_frame": ms * 1e3, "Mpixels_per_s": W * H / (ms * 1e-3) / 1e6, "alg_bytes_per_px": 15,
           "achieved_gbs": 15 * W * H / (ms * 1e-3) / 1e9, "frac_of_peak": 15 * W * H / (ms * 1e-3) / 1e9 / peak_gbs,
           "bound": "SFU/FMA issue (the output transfer function, 3-6 powf per pixel, after every per-code function is tabulated), not HBM", "maxCLL": clli.maxCLL, "maxPALL": clli.maxPALL}
    try:
        from oracle import loader
        ref = loader.load_reference()
        if ref is not None and hasattr(ref, "avifRGBImageApplyGainMap"):
            cw, ch = 1920, 1080
            hb, hg, hgm = build(cw, ch, None)
            out = A.avifRGBImage()
            out.depth, out.format = 10, A.AVIF_RGB_FORMAT_RGBA
            t0 = time.perf_counter()
            r = ref.avifRGBImageApplyGainMap(hb.ptr(), 1, 13, ctypes.byref(hgm), ctypes.c_float(2.0), 9, 16, ctypes.byref(out), None, ctypes.byref(diag))
            dt = time.perf_counter() - t0
            ref.avifRGBImageFreePixels(ctypes.byref(out))
            if r == 0:
                res["cpu_reference"] = {"Mpixels_per_s": cw * ch / dt / 1e6, "cores": 1, "sample": "one 1920x1080 call of src/gainmap.c (oracle/_ref), single-threaded as in the reference"}
    except Exception as e:  # the checker is optional here
        res["cpu_reference"] = {"unavailable": str(e)[:100]}
    return res


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
