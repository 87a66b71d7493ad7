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
