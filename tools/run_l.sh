# usage: gpurun -- 'bash tools/run_l.sh <tag>': y2r parity tests on the default build, the variant sweep, ncu --set full of the scaling cases
TAG=${1:-r02l}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -q -m gpu -x 2>&1 | tail -6
bash tools/variant_sweep.sh $TAG C3,C5,C1,F16,RGB565,NS
REP=/tmp/prof_sc_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scale_" -c 8 -o $REP python tools/profile_cases.py gpurun_out/$TAG/cases_sc.json --only SC > gpurun_out/$TAG/ncu_sc.log 2>&1
ncu -i $REP.ncu-rep --page raw --csv --print-units base > gpurun_out/$TAG/prof_sc_raw.csv 2>/dev/null
python tools/ncu_table.py gpurun_out/$TAG/prof_sc_raw.csv gpurun_out/$TAG/cases_sc.json > gpurun_out/$TAG/ncu_sc.md 2> gpurun_out/$TAG/ncu_table.err
cat gpurun_out/$TAG/ncu_sc.md gpurun_out/$TAG/ncu_table.err
for ID in 1 3 7; do ncu -i $REP.ncu-rep --page source --csv --kernel-id :::$ID > gpurun_out/$TAG/src_sc_k$ID.csv 2>/dev/null; done
ls -la gpurun_out/$TAG
