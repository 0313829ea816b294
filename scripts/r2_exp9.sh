#!/bin/bash
mkdir -p gpurun_out
{
for lvl in 47; do
echo "ZG_PDL=$lvl"
env ZG_PDL=$lvl timeout 900 python scripts/bench_configs.py cfg4 --trials 3 --steps 4 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), round(r['e2e']['value']))
" | grep t1024
done
echo "cfg5 small (128 sequences on one GPU)"
timeout 900 python scripts/bench_configs.py cfg5 --small --trials 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3))
    else: print(ln.rstrip()[:200])
"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_decode_128.csv python scripts/profile_batch.py decode 128 > /dev/null 2>&1
python - <<PY
import csv
rows=[]
for ln in csv.reader(open('gpurun_out/launches_batch_decode_128.csv')):
    if len(ln)>5 and ln[0].isdigit(): rows.append(ln)
for r in rows[-14:]:
    print(r[4][:70].ljust(70), r[-1])
PY
} > gpurun_out/r2_exp9.txt 2>&1
tail -32 gpurun_out/r2_exp9.txt
