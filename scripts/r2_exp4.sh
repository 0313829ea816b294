#!/bin/bash
mkdir -p gpurun_out
{ AB_POS=24,60,200 bash scripts/ab.sh; } > gpurun_out/r2_exp4.txt 2>&1
grep "us/token" gpurun_out/r2_exp4.txt | head -12
