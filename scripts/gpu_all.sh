#!/bin/bash
# Runs on the GPU box under gpurun: smoke, GPU tests, headline bench, BASELINE configs 3-5. Logs under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -c 1500 gpurun_out/bench_now.json; timeout 600 python bench.py --steps 256 --warmup 8 --no-cpu-baseline > gpurun_out/bench_k256.json 2>>gpurun_out/bench_now.err
for c in cfg3 cfg4 cfg5; do echo "== $c"; timeout 900 python scripts/bench_configs.py $c 2>&1 | tail -2 | cut -c1-1500; done
