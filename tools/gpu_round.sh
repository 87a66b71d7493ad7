#!/bin/bash
# One gpurun call = tests + bench + ncu evidence (each call costs minutes of box time: batch everything).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [skiptests] [noprof]'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt
nvidia-smi topo -m >> $OUT/gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA" >> $OUT/gpu.txt
NG=$(nvidia-smi -L | wc -l)
if [ "$2" != "skiptests" ]; then
  timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1
  tail -25 $OUT/pytest_gpu.log
fi
timeout 900 python bench.py --extras > $OUT/bench.json 2> $OUT/bench.err
tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$NG" -ge 2 ]; then
  for N in 2 4 8; do
    if [ "$NG" -ge "$N" ]; then
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N \
          > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
      python -c "import json,sys; d=json.loads(open('$OUT/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'], 'c5', (d.get('c5_batch') or {}).get('Mpixels_per_s'))" || tail -5 $OUT/bench_n$N.err
    fi
  done
fi
if [ "$3" == "noprof" ]; then ls -la $OUT; exit 0; fi
# launch list (cold-cache, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --batch 8 --e2e-batch 1 --e2e-threads 1 > $OUT/ncu_launches.log 2>&1
# full capture of every configuration's BATCHED kernel (2 launches each; the 2nd of each pair is the warm one).  The report itself
# can exceed what gpurun copies back: keep it in /tmp and bring home the raw-page CSV (+ the SASS pages of selected kernels)
REP=/tmp/prof_cases_$TAG
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"y2r_|r2y_|tone_map|scale_" -c 200 -o $REP \
    python tools/profile_cases.py $OUT/cases.json > $OUT/ncu_cases.log 2>&1
tail -30 $OUT/ncu_cases.log
ncu -i $REP.ncu-rep --page raw --csv --print-units base > $OUT/prof_cases_raw.csv 2>/dev/null
python tools/ncu_table.py $OUT/prof_cases_raw.csv $OUT/cases.json --traffic $OUT/traffic.json > $OUT/ncu_cases.md 2> $OUT/ncu_table.err
cat $OUT/ncu_cases.md $OUT/ncu_table.err
# SASS/source pages: C2 fwd (kernel id 1), C2 back (3), C3 fwd (7), C5 (11)
for ID in 1 3 7 11; do ncu -i $REP.ncu-rep --page source --csv --kernel-id :::$ID > $OUT/src_k$ID.csv 2>/dev/null; done
if [ $(stat -c %s $REP.ncu-rep) -lt 40000000 ]; then cp $REP.ncu-rep $OUT/prof_cases.ncu-rep; fi
# the TMA-staged variant of the bilinear 4:2:0 kernel (libavif_b200/csrc/variants/libavif_b200_tma.so, if built): same capture
if [ -e libavif_b200/csrc/variants/libavif_b200_tma.so ]; then
  AVIF_B200_LIB=$PWD/libavif_b200/csrc/variants/libavif_b200_tma.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"y2r_" -c 8 -o ${REP}_tma \
      python tools/profile_cases.py $OUT/cases_tma.json --only C3,C5 > $OUT/ncu_cases_tma.log 2>&1
  ncu -i ${REP}_tma.ncu-rep --page raw --csv --print-units base > $OUT/prof_cases_tma_raw.csv 2>/dev/null
  python tools/ncu_table.py $OUT/prof_cases_tma_raw.csv $OUT/cases_tma.json > $OUT/ncu_cases_tma.md 2>> $OUT/ncu_table.err
  cat $OUT/ncu_cases_tma.md
  ncu -i ${REP}_tma.ncu-rep --page source --csv --kernel-id :::1 > $OUT/src_tma_c3.csv 2>/dev/null
fi
ls -la $OUT
# the rows beyond the reformat path: their tests, kernel times of the scaling cases, tone mapping with and without the threshold table
bash tools/gpu_rows.sh $TAG > $OUT/rows.log 2>&1
tail -30 $OUT/rows.log
