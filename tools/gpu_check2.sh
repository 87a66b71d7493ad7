#!/bin/bash
# quick 2-GPU validation: new tests + grid runner + pageable e2e (bench without per-config legs)
TAG=${1:-chk}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_grid.py tests/test_gpu_dropin.py tests/test_gpu_devices.py tests/test_gpu_workloads.py tests/test_gpu_division.py -q -m gpu -x 2>&1 | tail -15
NG=$(nvidia-smi -L | wc -l)
timeout 300 python tools/grid_run.py > $OUT/grid_n1.json 2> $OUT/grid_n1.err; tail -c 1500 $OUT/grid_n1.json; tail -3 $OUT/grid_n1.err
for N in 2 4 8; do
  if [ "$NG" -ge "$N" ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/grid_run.py > $OUT/grid_n$N.json 2> $OUT/grid_n$N.err
    tail -c 1500 $OUT/grid_n$N.json; tail -3 $OUT/grid_n$N.err
  fi
done
timeout 600 python bench.py --no-configs --no-cpu-baseline > $OUT/bench_nc.json 2> $OUT/bench_nc.err
python -c "import json; d=json.loads(open('$OUT/bench_nc.json').read().strip().splitlines()[-1]); print('e2e', d['e2e']['value'], 'pageable', d['e2e_pageable'])" || tail -5 $OUT/bench_nc.err
AVIF_B200_NT_COPY=0 timeout 600 python bench.py --no-configs --no-cpu-baseline > $OUT/bench_nc_memcpy.json 2> $OUT/bench_nc_memcpy.err
python -c "import json; d=json.loads(open('$OUT/bench_nc_memcpy.json').read().strip().splitlines()[-1]); print('memcpy staging: e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'])"
AVIF_B200_STAGING=0 timeout 600 python bench.py --no-configs --no-cpu-baseline > $OUT/bench_nc_driver.json 2> $OUT/bench_nc_driver.err
python -c "import json; d=json.loads(open('$OUT/bench_nc_driver.json').read().strip().splitlines()[-1]); print('driver pageable path: e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'])"
