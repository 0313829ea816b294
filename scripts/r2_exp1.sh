#!/bin/bash
# round 2, GPU call 1: exchange replicas / warp-count A/B + phase timelines (attention CTA and a plain CTA)
mkdir -p gpurun_out
{
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
echo "== tests model"; timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
echo "== A/B"
AB_POS=24,60,200 bash scripts/ab.sh
echo "== timelines"
for v in "" x4 w15x4; do
  for cta in 100 0; do
    echo "-- variant '${v}' cta ${cta}"
    if [ -z "$v" ]; then timeout 200 python scripts/clock_profile.py 124M 16 $cta 24
    else ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_$v.so timeout 200 python scripts/clock_profile.py 124M 16 $cta 24; fi
  done
done
} > gpurun_out/r2_exp1.txt 2>&1
tail -60 gpurun_out/r2_exp1.txt
