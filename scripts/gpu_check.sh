#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU tests, short bench, phase profile. Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== phase profile" ; timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | tail -12
echo "== bench" ; timeout 600 python bench.py --steps 256 --warmup 8 2>&1 | tail -3
