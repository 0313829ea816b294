#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "pair or prefill" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l23.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/l23.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-12:]: print(r[4][:60].ljust(60), r[-1])
PY
timeout 600 python scripts/bench_configs.py cfg3 --trials 5 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), r['clocks'])
"
} > gpurun_out/r2_exp23.txt 2>&1
cat gpurun_out/r2_exp23.txt
