#!/bin/bash
# Runs on the GPU box under gpurun AFTER gpu_all.sh: phase timeline, ncu launch list of the bench command and one
# `ncu --set full` capture of the persistent decode kernel.  Everything lands under gpurun_out/.
mkdir -p gpurun_out
echo "== clock profile" ; timeout 300 python scripts/clock_profile.py 124M 16 100 > gpurun_out/clock_profile_124M.txt 2>&1; tail -75 gpurun_out/clock_profile_124M.txt
echo "== pos sweep" ; timeout 300 python scripts/pos_sweep.py 124M 2>&1 | tee gpurun_out/pos_sweep_124M.txt | tail -12
echo "== ncu launch list (bench)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 64 --warmup 4 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c decode_persistent gpurun_out/launches_bench.csv
echo "== ncu full (decode kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 2 -c 1 -f -o gpurun_out/decode_full \
    python scripts/ncu_decode.py 124M 24 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/ | tail -12
echo "== ncu launch lists (batched paths: cfg 3 / 5 / 4 shapes, 2 layers)"
for w in prefill decode decode_xl; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_$w.csv \
      python scripts/profile_batch.py $w > gpurun_out/profile_batch_$w.log 2>&1
done
ls gpurun_out | tr '\n' ' '
