#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU tests, phase profile, short bench. Logs under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== pytest gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== phase profile" ; timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | tail -30
echo "== bench" ; timeout 600 python bench.py --steps 256 --warmup 8 2>&1 | tail -3
