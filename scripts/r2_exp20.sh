#!/bin/bash
mkdir -p gpurun_out
{
echo "== pair kernel"; ZG_DEBUG_PAIR=1 timeout 600 python scripts/gpu_check_tc.py 2>&1 | grep -v "'prec': 2" | grep 'zg:\|timing\|ALL\|FAIL\|4096' | head -24
} > gpurun_out/r2_exp20.txt 2>&1
cat gpurun_out/r2_exp20.txt
