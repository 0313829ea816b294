#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python scripts/split_err.py
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_skinny.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -6
for c in cfg5 cfg4; do
timeout 600 python scripts/bench_configs.py $c --trials 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3))
"
done
} > gpurun_out/r2_exp19.txt 2>&1
cat gpurun_out/r2_exp19.txt
