#!/bin/bash
# One gpurun call = tests + bench + ncu evidence (each call costs minutes of box time: batch everything).
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [skiptests]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/gpu.txt
if [ "$2" != "skiptests" ]; then
  timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1
  tail -25 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --extras > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
# launch list (cold-cache, serialised: compare shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --batch 4 --e2e-batch 1 > $OUT/ncu_launches.log 2>&1
# full capture of every configuration's kernel (2 launches each; the 2nd of each pair is the warm one).  The report itself
# can exceed what gpurun copies back: keep it in /tmp and bring home the raw-page CSV (+ the SASS pages of the C2 kernels)
REP=/tmp/prof_cases_$TAG
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"y2r_|r2y_|tone_map|scale_" -c 80 -o $REP \
    python tools/profile_cases.py > $OUT/ncu_cases.log 2>&1
tail -12 $OUT/ncu_cases.log
ncu -i $REP.ncu-rep --page raw --csv > $OUT/prof_cases_raw.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --kernel-id :::4 > $OUT/src_c2_fwd.csv 2>/dev/null
ncu -i $REP.ncu-rep --page source --csv --kernel-id :::6 > $OUT/src_c2_back.csv 2>/dev/null
if [ $(stat -c %s $REP.ncu-rep) -lt 40000000 ]; then cp $REP.ncu-rep $OUT/prof_cases.ncu-rep; fi
ls -la $OUT
