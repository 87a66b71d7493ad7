#!/bin/bash
# Times every tuning variant built by `make -C libavif_b200/csrc variant NAME=.. VFLAGS=..` against the default build with
# tools/time_variants.py (batched launches of every BASELINE configuration, CUDA events, output digests checked).
# usage: gpurun --timeout 1200 -- 'bash tools/variant_sweep.sh <tag> [cases]'
TAG=${1:-sweep}
CASES=${2:-all}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
run() { # name, lib path ("" = default)
  AVIF_B200_LIB=$2 timeout 300 python tools/time_variants.py $CASES > $OUT/variant_$1.json 2> $OUT/variant_$1.err || tail -3 $OUT/variant_$1.err
  python - "$OUT/variant_$1.json" "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[2].ljust(12), "FAILED", e); sys.exit(0)
print(sys.argv[2].ljust(12), " ".join(f"{k}={v['us']:.1f}{'' if v['ok'] else '!BAD'}" for k, v in d.items() if isinstance(v, dict)))
PY
}
run default ""
for lib in libavif_b200/csrc/variants/libavif_b200_*.so; do
  [ -e "$lib" ] || continue
  n=$(basename $lib .so); n=${n#libavif_b200_}
  run $n $PWD/$lib
done
