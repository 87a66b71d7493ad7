# multi-GPU evidence: bench + grid driver at every N the box offers
TAG=${1:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "$2" == "tests" ]; then timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6; fi
for N in ${3:-2 4 8}; do
  if [ "$NG" -ge "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
    python - $OUT/bench_n$N.json $N <<'PY' || tail -5 $OUT/bench_n$N.err
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N=%s value %.0f e2e %.0f pageable %.0f c5 %.0f' % (sys.argv[2], d['value'], d['e2e']['value'], d['e2e_pageable']['value'], d['c5_batch']['Mpixels_per_s']))
print('   c4_grid', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['c4_grid'].items() if k.startswith('ms_') or k in ('check','mode') or k.startswith('Mpix')})
print('   c4_grid staged', {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['c4_grid'].get('staged',{}).items() if k.startswith('ms_') or k in ('check','mode')})
PY
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/grid_run.py > $OUT/grid_n$N.json 2> $OUT/grid_n$N.err
    python -c "import json; d=json.loads(open('$OUT/grid_n$N.json').read().strip().splitlines()[-1]); print('   grid N=$N', d['encode_check'], d.get('encode_mode'), d.get('encode_check_staged'), d['decode_check'], {k:round(v,3) for k,v in d['encode'].items() if k.startswith('ms_')}, {k:round(v,3) for k,v in d.get('encode_staged',{}).items() if k.startswith('ms_')}, {k:round(v,3) for k,v in d['decode'].items() if k.startswith('ms_')}, d.get('decode_mode'), d.get('decode_check_peer'), {k:round(v,3) for k,v in d.get('decode_peer',{}).items() if k.startswith('ms_')})" || tail -3 $OUT/grid_n$N.err
  fi
done
