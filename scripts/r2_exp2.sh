#!/bin/bash
mkdir -p gpurun_out
{
for cta in 0 5 100; do
timeout 200 python scripts/clock_profile.py 124M 16 $cta 24
done
timeout 200 python scripts/clock_profile.py 124M 16 0 60
} > gpurun_out/r2_exp2.txt 2>&1
grep -A8 "P2 attn" gpurun_out/r2_exp2.txt
