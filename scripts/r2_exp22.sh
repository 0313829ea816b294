#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_batch.py -m gpu -x -q -k "prefill" 2>&1 | tail -5
for v in "" ZG_ATTN_V2=1; do
echo "variant '$v'"
env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_prefill -c 6 --csv --log-file gpurun_out/l22.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
grep attn_prefill gpurun_out/l22.csv | tail -3 | awk -F'","' '{print $NF}'
done
timeout 600 python scripts/bench_configs.py cfg3 --trials 5 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), r['clocks'])
"
} > gpurun_out/r2_exp22.txt 2>&1
cat gpurun_out/r2_exp22.txt
