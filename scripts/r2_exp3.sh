#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
AB_POS=24,60,200 bash scripts/ab.sh
timeout 200 python scripts/clock_profile.py 124M 16 0 24
} > gpurun_out/r2_exp3.txt 2>&1
grep "us/token" gpurun_out/r2_exp3.txt | head -12
