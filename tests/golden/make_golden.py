#!/usr/bin/env python
"""Generates tests/golden/reference_vectors.json from the UNMODIFIED reference (oracle/_ref/libavif_ref.so, built from
/root/reference by oracle/Makefile).  Each vector = a conversion spec + seed (inputs are regenerated from the seed by
tests/util.py, numpy RandomState) and the FNV-1a-64 of every output byte the reference produced.
Run in the container that has /root/reference:   python tests/golden/make_golden.py
"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from tests import util  # noqa: E402


def main():
    ref = loader.load_reference()
    assert ref is not None, "build oracle/_ref first (make -C oracle)"
    orc = loader.load_oracle()
    out = {"y2r": [], "r2y": []}
    for k, spec in enumerate(util.random_y2r_specs(400, seed=4242, max_dim=96)):
        code, data = util.yuv_to_rgb(ref, spec, seed=50000 + k)
        out["y2r"].append({"spec": dataclasses.asdict(spec), "seed": 50000 + k, "result": code,
                           "fnv": f"{util.fnv1a64(orc, data):016x}" if code == 0 else None})
    for k, spec in enumerate(util.random_r2y_specs(300, seed=4343, max_dim=96)):
        code, planes = util.rgb_to_yuv(ref, spec, seed=60000 + k)
        h = util.FNV_OFFSET
        if code == 0:
            for p in planes:
                if p is not None:
                    h = util.fnv1a64(orc, p, h)
        out["r2y"].append({"spec": dataclasses.asdict(spec), "seed": 60000 + k, "result": code, "fnv": f"{h:016x}" if code == 0 else None})
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, len(out["y2r"]), len(out["r2y"]))
    tool_outputs()
    next_rows()
    bench_vectors()


def bench_vectors(ref=None):
    """Digests of what the compiled reference produces for BASELINE.json's full-size configurations on the seeded inputs of
    libavif_b200/workloads.py: what bench.py checks its own outputs against after the timed region, and what the full-size GPU
    parity tests compare with (tests/test_gpu_workloads.py)."""
    from libavif_b200 import workloads as wl
    ref = ref or loader.load_reference()
    out = {}
    for name, cfg in wl.CONFIGS.items():
        entry = {"seed": cfg["seed"]}
        if cfg.get("back_only"):
            rgb = wl.make_rgb(name)
            wl.fill_rgb_seeded(rgb, cfg["seed"])
        else:
            img = wl.make_yuv(name)
            wl.fill_yuv_seeded(img, cfg["seed"])
            rgb = wl.make_rgb(name)
            assert ref.avifImageYUVToRGB(img.ptr(), rgb.ptr()) == 0, name
            entry["rgb"] = wl.rgb_digest(rgb)
        if cfg.get("back") or cfg.get("back_only"):
            dst = wl.make_yuv(name)
            assert ref.avifImageRGBToYUV(dst.ptr(), rgb.ptr()) == 0, name
            entry["yuv_back"] = wl.yuv_digest(dst)
        out[name] = entry
        print(name, entry, flush=True)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "bench_vectors.json"), "w"), indent=1)


def next_rows():
    """Fixtures of the rows beyond the reformat path, from the compiled reference: FNV-1a-64 of every plane avifImageScale
    produces for the size pairs of tests/test_gpu_scale.py (bit-exact row), and the tone-mapped pixels of three small gain-map
    cases (the +-1-code row keeps the pixels themselves)."""
    import numpy as np
    from tests import test_gpu_scale as ts
    from tests import test_gpu_gainmap as tg
    ref = loader.load_reference()
    orc = loader.load_oracle()
    out = {"scale": []}
    A = ts.A
    for depth in (8, 10, 12):
        for k, ((w, h), (dw, dh)) in enumerate(ts.CASES):
            for fmt, alpha in ((A.AVIF_PIXEL_FORMAT_YUV444, True), (A.AVIF_PIXEL_FORMAT_YUV420, False), (A.AVIF_PIXEL_FORMAT_YUV400, True)):
                res, planes, dims, _ = ts.scaled_planes(ref, w, h, depth, fmt, alpha, 1000 + k, dw, dh)
                hs = []
                for p_i, pl in enumerate(planes):
                    if pl is None or (fmt == A.AVIF_PIXEL_FORMAT_YUV400 and p_i in (1, 2)):
                        hs.append(None)
                    else:
                        hs.append(f"{util.fnv1a64(orc, np.ascontiguousarray(pl).reshape(-1)):016x}")
                out["scale"].append({"src": [w, h], "dst": [dw, dh], "depth": depth, "fmt": int(fmt), "alpha": alpha, "seed": 1000 + k,
                                     "result": res, "planes": hs})
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "scale_vectors.json"), "w"), indent=0)
    tm = {}
    cases = {"srgb8_to_pq10": dict(tc_in=13, tc_out=16, depth=10, headroom=1.5, cp_out=9),
             "bt709_8_to_srgb8": dict(tc_in=1, tc_out=13, depth=8, headroom=3.0, cp_out=1),
             "hlg10_to_hlg12": dict(tc_in=18, tc_out=18, depth=12, headroom=0.7, cp_out=12)}
    for name, c in cases.items():
        gm, keep = tg.make_gain_map(48, 32, seed=77)
        base = tg.make_base(48, 32, 8, A.AVIF_RGB_FORMAT_RGBA, seed=78)
        res, data, clli, _ = tg.apply(ref, base, 1, c["tc_in"], gm, c["headroom"], c["cp_out"], c["tc_out"], c["depth"], A.AVIF_RGB_FORMAT_RGBA)
        assert res == 0
        tm[name + "_pixels"] = data
        tm[name + "_clli"] = np.array(clli, dtype=np.int64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tonemap_vectors.npz"), **tm)
    print("wrote scale_vectors.json", len(out["scale"]), "and tonemap_vectors.npz", len(tm))


def tool_outputs():
    """stdout of the reference's own stress tool (oracle/_ref/avifyuv = tests/avifyuv.c of the reference, linked
    against the reference build), run natively on the CPU; tests/test_gpu_dropin.py replays the tool with the
    B200 library preloaded and requires the same bytes."""
    import subprocess
    tool = os.path.join(ROOT, "oracle", "_ref", "avifyuv")
    for mode, extra in (("rgb", []), ("limited", []), ("premultiply", ["-v"])):
        text = subprocess.run([tool, "-m", mode, *extra], capture_output=True, check=True).stdout.decode()
        if extra:  # drop the "[ n/ N]" progress lines, keep the per-depth drift histograms
            text = "".join(l + "\n" for l in text.replace("\r", "\n").split("\n") if l and not l.startswith("["))
        open(os.path.join(ROOT, "tests", "golden", f"avifyuv_{mode}.txt"), "w").write(text)
        print("wrote avifyuv_%s.txt" % mode)


if __name__ == "__main__":
    main()
