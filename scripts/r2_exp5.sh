#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
AB_POS=24,60,200 bash scripts/ab.sh
for s in 355M 1.5B; do timeout 300 python scripts/ab_time.py $s 24,200; ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_r1.so timeout 300 python scripts/ab_time.py $s 24,200; done
timeout 200 python scripts/clock_profile.py 124M 16 100 24 | grep -A8 "P5 reduce"
} > gpurun_out/r2_exp5.txt 2>&1
grep "us/token\|passed\|failed\|error" gpurun_out/r2_exp5.txt | head -20
