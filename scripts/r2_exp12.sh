#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "prefill" 2>&1 | tail -6
for v1 in "" 1; do
echo "ZG_ATTN_V1=$v1"
env ${v1:+ZG_ATTN_V1=1} timeout 900 python scripts/bench_configs.py cfg3 --trials 5 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), round(r['roofline']['frac_of_burst'],3), r['clocks'])
    else: print(ln.rstrip()[:300])
"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prefill_v2.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
grep attn_prefill gpurun_out/launches_prefill_v2.csv | tail -2 | cut -c1-60,200-
python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/launches_prefill_v2.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-10:]: print(r[4][:60].ljust(60), r[-1])
PY
} > gpurun_out/r2_exp12.txt 2>&1
tail -30 gpurun_out/r2_exp12.txt
