# usage: gpurun -- 'bash tools/gpu_rows.sh <tag>': scaler / tone-map tests, kernel times of the scaling cases (ncu launch list), extras
TAG=${1:-rows}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_gainmap.py tests/test_gpu_aux.py -q -m gpu 2>&1 | tail -15
AVIF_B200_TONEMAP_TABLE=0 timeout 900 python -m pytest tests/test_gpu_gainmap.py -q -m gpu 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scale_" --csv --log-file gpurun_out/$TAG/scale_launches.csv \
    python tools/profile_cases.py gpurun_out/$TAG/cases_sc.json --only SC > gpurun_out/$TAG/ncu_sc.log 2>&1
grep -E "scale_batch" gpurun_out/$TAG/scale_launches.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-160
for TABLE in 1 0; do
  AVIF_B200_TONEMAP_TABLE=$TABLE timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tone_" --csv --log-file gpurun_out/$TAG/tone_launches_$TABLE.csv \
      python tools/profile_cases.py gpurun_out/$TAG/cases_tm.json --only TM > gpurun_out/$TAG/ncu_tm_$TABLE.log 2>&1
  echo "tone-map kernels, AVIF_B200_TONEMAP_TABLE=$TABLE:"
  grep -E "tone_" gpurun_out/$TAG/tone_launches_$TABLE.csv | awk -F'","' '{print $5, $(NF)}' | cut -c1-200
done
timeout 600 python bench.py --extras --no-configs --no-cpu-baseline > gpurun_out/$TAG/bench_extras.json 2> gpurun_out/$TAG/bench_extras.err
tail -3 gpurun_out/$TAG/bench_extras.err
python - <<PY
import json
d=json.loads(open('gpurun_out/$TAG/bench_extras.json').read().strip().splitlines()[-1])
for k,v in d["extras"].items(): print(k, v.get("us_per_frame") or v.get("us_per_image"), v.get("frac_of_peak"))
PY
