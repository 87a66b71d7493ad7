mkdir -p gpurun_out/r02h
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 600 python bench.py --extras > gpurun_out/r02h/bench.json 2> gpurun_out/r02h/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h/bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'pageable',round(d['e2e_pageable']['value']),d['e2e_pageable']['frac_of_pinned'],'roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3),'cpu',d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)
for k,v in d['configs'].items():
    print('  ',k,v.get('kernel'),round(v.get('us_per_frame',v.get('us_per_canvas',0)),1),round(v['roofline']['frac'],3),v.get('check'),('back %.1f %.3f %s'%(v['back']['us_per_frame'],v['back']['roofline']['frac'],v['back'].get('check'))) if 'back' in v else '')
print({k:(v.get("us_per_frame") or v.get("us_per_image")) for k,v in d["extras"].items()})
PY
for T in 5 8; do
timeout 300 python bench.py --no-configs --no-cpu-baseline --stage-threads $T > gpurun_out/r02h/bench_st$T.json 2>/dev/null
python -c "import json; d=json.loads(open('gpurun_out/r02h/bench_st$T.json').read().strip().splitlines()[-1]); print('stage threads $T: pageable', d['e2e_pageable']['value'], d['e2e_pageable']['frac_of_pinned'])"
done
