#!/bin/bash
# where the batched GEMMs lose time: launch list of a cfg5 step + full captures of the 3xTF32 GEMM (M = 1024) and the f16 prefill GEMM / attention
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l16_decode.csv python scripts/profile_batch.py decode > /dev/null 2>&1
python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/l16_decode.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-20:]: print(r[4][:70].ljust(70), r[-1])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 20 -c 5 -f -o gpurun_out/r02_gemm3x_full python scripts/profile_batch.py decode > gpurun_out/ncu16a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_prefill" -s 12 -c 6 -f -o gpurun_out/r02_prefill_full python scripts/profile_batch.py prefill > gpurun_out/ncu16b.log 2>&1
ls -la gpurun_out | tail -5
