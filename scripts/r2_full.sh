#!/bin/bash
# full GPU validation: all -m gpu tests, smoke, bench (with the config sub-records), reference arm
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; ( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 600 gpurun_out/bench_r2.err
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['roofline']['frac'], l['e2e']['value'], l.get('cpu_baseline',{}).get('value'))
    for k,v in l.get('configs',{}).items():
        print(k, v.get('error') or (round(v['value']), round(v['ms_per_step'],3), round(v['roofline']['frac'],3), round(v['e2e']['value'])))
except Exception as e: print('bench parse failed', e)
PY
} > gpurun_out/r2_full.txt 2>&1
tail -40 gpurun_out/r2_full.txt
