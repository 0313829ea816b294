#!/bin/bash
# Round-1 profiling pass (run under gpurun): launch list of the bench command + one full ncu capture of the decode kernel.
mkdir -p gpurun_out
echo "== bench (no profiler)"; timeout 600 python bench.py --steps 256 --warmup 8 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 600 gpurun_out/bench_r01.json
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 64 --warmup 4 --trials 2 --no-cpu-baseline > gpurun_out/launches_r01.log 2>&1; tail -3 gpurun_out/launches_r01.log | cut -c1-300
echo "== full capture"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 3 -c 1 -o gpurun_out/decode_r01 python bench.py --steps 64 --warmup 4 --trials 2 --no-cpu-baseline > gpurun_out/ncu_full_r01.log 2>&1; tail -3 gpurun_out/ncu_full_r01.log | cut -c1-300
ls -la gpurun_out | tail -8
