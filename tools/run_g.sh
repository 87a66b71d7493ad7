timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
bash tools/variant_sweep.sh r02g
python bench.py --extras --no-configs --no-cpu-baseline > gpurun_out/r02g/bench_extras.json 2> gpurun_out/r02g/bench_extras.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g/bench_extras.json').read().strip().splitlines()[-1])
print({k:(v.get("us_per_frame") or v.get("us_per_image")) for k,v in d["extras"].items()})
print("e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], d["e2e_pageable"]["frac_of_pinned"])
PY
