# usage: gpurun -- 'bash tools/run_j.sh <tag>': the whole GPU suite (no -x), then the rows beyond the reformat path timed
TAG=${1:-r02j}
mkdir -p gpurun_out/$TAG
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/$TAG/pytest_gpu.log 2>&1
tail -40 gpurun_out/$TAG/pytest_gpu.log
timeout 600 python bench.py --extras --no-configs --no-cpu-baseline > gpurun_out/$TAG/bench_extras.json 2> gpurun_out/$TAG/bench_extras.err
tail -3 gpurun_out/$TAG/bench_extras.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$TAG/bench_extras.json').read().strip().splitlines()[-1])
for k,v in d["extras"].items(): print(k, v.get("us_per_frame") or v.get("us_per_image"), v.get("frac_of_peak"))
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], d["e2e_pageable"]["frac_of_pinned"])
PY
