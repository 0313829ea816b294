#!/bin/bash
# warp-uniform MMA / TMA issue: parity of every tensor-core kernel, then cfg3 / cfg4 / cfg5 and the prefill attention time
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_skinny.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -6
for c in cfg3 cfg4 cfg5; do
timeout 900 python scripts/bench_configs.py $c --trials 5 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), r['roofline'].get('frac_of_burst'), r['clocks'])
    else: print(ln.rstrip()[:300])
"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prefill_v3.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/launches_prefill_v3.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-12:]: print(r[4][:60].ljust(60), r[-1])
PY
} > gpurun_out/r2_exp15.txt 2>&1
tail -40 gpurun_out/r2_exp15.txt
