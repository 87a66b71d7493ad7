# usage: gpurun -- 'bash tools/run_n.sh <tag>': whole GPU suite, kernel times of the scaling cases (ncu launch list), extras
TAG=${1:-r02n}
mkdir -p gpurun_out/$TAG
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/$TAG/pytest_gpu.log 2>&1
tail -12 gpurun_out/$TAG/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scale_" --csv --log-file gpurun_out/$TAG/scale_launches.csv \
    python tools/profile_cases.py gpurun_out/$TAG/cases_sc.json --only SC > gpurun_out/$TAG/ncu_sc.log 2>&1
grep -E "scale_batch" gpurun_out/$TAG/scale_launches.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160
timeout 600 python bench.py --extras --no-configs --no-cpu-baseline > gpurun_out/$TAG/bench_extras.json 2> gpurun_out/$TAG/bench_extras.err
tail -3 gpurun_out/$TAG/bench_extras.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$TAG/bench_extras.json').read().strip().splitlines()[-1])
for k,v in d["extras"].items(): print(k, v.get("us_per_frame") or v.get("us_per_image"), v.get("frac_of_peak"))
PY
REP=/tmp/prof_sc_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scale_" -c 8 -o $REP python tools/profile_cases.py gpurun_out/$TAG/cases_sc2.json --only SC > gpurun_out/$TAG/ncu_sc_full.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv --print-units base > gpurun_out/$TAG/prof_sc_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/$TAG/prof_sc_raw.csv gpurun_out/$TAG/cases_sc2.json > gpurun_out/$TAG/ncu_sc.md 2> gpurun_out/$TAG/ncu_table.err
cat gpurun_out/$TAG/ncu_sc.md gpurun_out/$TAG/ncu_table.err
for ID in 1 3 7; do ncu -i $REP.ncu-rep --page source --csv --kernel-id :::$ID > gpurun_out/$TAG/src_sc_k$ID.csv 2>/dev/null; done
