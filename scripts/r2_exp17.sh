#!/bin/bash
mkdir -p gpurun_out
{
for v in "" storehi; do
  echo "variant '$v'"
  if [ -z "$v" ]; then L=""; else L="ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_$v.so"; fi
  env $L timeout 300 python scripts/split_err.py
  env $L timeout 600 python scripts/bench_configs.py cfg5 --trials 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3))
"
  env $L timeout 600 python scripts/bench_configs.py cfg4 --trials 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3))
"
done
timeout 900 python -m pytest tests/test_gpu_skinny.py tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -4
} > gpurun_out/r2_exp17.txt 2>&1
cat gpurun_out/r2_exp17.txt
