#!/bin/bash
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests/test_gpu_batch.py tests/test_gpu_skinny.py -m gpu -x -q -k "16bit or f16_operands" 2>&1 | tail -8
timeout 900 python scripts/bench_configs.py cfg4 --trials 3 --steps 4 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3), round(r['e2e']['value']))
    else: print(ln.rstrip()[:300])
"
} > gpurun_out/r2_exp11.txt 2>&1
tail -20 gpurun_out/r2_exp11.txt
