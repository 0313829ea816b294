#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_loaders.py -m gpu -x -q 2>&1 | tail -3
AB_POS=24,60,200 bash scripts/ab.sh
for s in 355M 1.5B; do timeout 300 python scripts/ab_time.py $s 24,200; done
} > gpurun_out/r2_exp10.txt 2>&1
grep "us/token\|passed\|failed\|error" gpurun_out/r2_exp10.txt | head -20
