#!/bin/bash
# host-buffer leg of bench.py by host thread count and band size. usage: gpurun -- 'bash tools/e2e_threads.sh <tag>'
TAG=${1:-e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
for cfg in "1 4 24" "2 4 24" "4 4 24" "4 8 24" "8 8 24" "4 4 64" "4 4 128" "8 8 128" "16 16 128"; do
  set -- $cfg
  AVIF_B200_BAND_MB=$3 timeout 300 python bench.py --no-cpu-baseline --steps 10 --e2e-threads $1 --e2e-batch $2 > $OUT/t$1_b$2_band$3.json 2>> $OUT/err.log
  python - "$OUT/t$1_b$2_band$3.json" "$cfg" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("threads/batch/bandMB", sys.argv[2], "e2e Mpx/s %.0f" % d["e2e"]["value"], "device %.0f" % d["value"])
PY
done
