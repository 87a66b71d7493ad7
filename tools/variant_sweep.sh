#!/bin/bash
# Times every tuning variant built by `make -C libavif_b200/csrc variant NAME=.. VFLAGS=..` against the default build:
# one `bench.py --extras` (device-resident kernel timings of every BASELINE configuration) per library.
# usage: gpurun --timeout 1200 -- 'bash tools/variant_sweep.sh <tag>'
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
run() { # name, lib path ("" = default)
  AVIF_B200_LIB=$2 timeout 300 python bench.py --extras --no-cpu-baseline --steps 10 > $OUT/variant_$1.json 2> $OUT/variant_$1.err
  python - "$OUT/variant_$1.json" "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
k = d["roofline"]["kernels"]
row = {"C2fwd": k["y2r_fast_444"]["ms_per_launch"] * 1e3 / d["roofline"]["frames_per_launch"], "C2back": k["r2y_fast"]["ms_per_launch"] * 1e3 / d["roofline"]["frames_per_launch"]}
for n, e in d["extras"].items():
    if not isinstance(e, dict) or "us_per_frame" not in e:
        continue
    row[n.split("_")[0] + ("_" + n.split("_")[-1] if n.startswith(("C1", "C5")) else "") + ("_back" if "back" in n else "")] = e["us_per_frame"]
print(sys.argv[2].ljust(10), " ".join(f"{a}={b:.1f}" for a, b in row.items()))
PY
}
run default ""
for lib in libavif_b200/csrc/variants/libavif_b200_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#libavif_b200_}
  run $n $PWD/$lib
done
