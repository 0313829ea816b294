#!/bin/bash
# round-2 evidence: phase timeline, position sweep, ncu launch lists and `ncu --set full` captures.  Lands under gpurun_out/;
# the CSV / text summaries are copied to profiles/ by hand (the .ncu-rep files stay in gpurun_out/).
mkdir -p gpurun_out
echo "== clock profile"; timeout 300 python scripts/clock_profile.py 124M 16 100 24 > gpurun_out/r02_phase_timeline_124M.txt 2>&1
timeout 300 python scripts/clock_profile.py 124M 16 0 24 >> gpurun_out/r02_phase_timeline_124M.txt 2>&1
echo "== pos sweep"; timeout 300 python scripts/pos_sweep.py 124M > gpurun_out/r02_pos_sweep_124M.txt 2>&1; tail -8 gpurun_out/r02_pos_sweep_124M.txt
echo "== ncu launch list (bench, headline only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/bench_under_ncu.log 2>&1
grep -c decode_persistent gpurun_out/r02_launches_bench.csv
echo "== ncu full (decode kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 2 -c 1 -f -o gpurun_out/r02_decode_full \
    python scripts/ncu_decode.py 124M 24 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full (stream-K GEMMs + decode attention, cfg4 shapes, 2 layers)"
for mode in 0 1; do
ZG_TF32=$mode timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_skinny|attn_decode_batch" -s 24 -c 8 -f -o gpurun_out/r02_skinny_full_tf32_$mode \
    python scripts/profile_batch.py decode_xl > gpurun_out/ncu_full_skinny.log 2>&1
done
echo "== ncu full (prefill: CTA-pair GEMMs + attention, cfg3 shapes, 2 layers)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair|gemm_tc|attn_prefill" -s 11 -c 6 -f -o gpurun_out/r02_prefill_full \
    python scripts/profile_batch.py prefill > gpurun_out/ncu_full_prefill.log 2>&1
echo "== ncu full (3xTF32 general GEMM at 1024 rows + decode attention, cfg5 shapes, 2 layers)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|attn_decode_batch" -s 24 -c 8 -f -o gpurun_out/r02_decode1024_full \
    python scripts/profile_batch.py decode > gpurun_out/ncu_full_decode1024.log 2>&1
echo "== ncu launch lists (batched paths)"
for w in prefill decode; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch_$w.csv \
      python scripts/profile_batch.py $w > gpurun_out/profile_batch_$w.log 2>&1
done
for mode in 0 1; do
  ZG_TF32=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch_decode_xl_tf32_$mode.csv \
      python scripts/profile_batch.py decode_xl > /dev/null 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_batch_decode_128.csv python scripts/profile_batch.py decode 128 > /dev/null 2>&1
ls -la gpurun_out/ | tail -24
