#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
  ZG_TF32=$mode timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_batch_decode_xl_tf32_$mode.csv \
      python scripts/profile_batch.py decode_xl > gpurun_out/profile_batch_decode_xl.log 2>&1
  python - <<PY
import csv
rows=[]
for ln in csv.reader(open('gpurun_out/launches_batch_decode_xl_tf32_$mode.csv')):
    if len(ln)>5 and ln[0].isdigit(): rows.append(ln)
print("mode tf32=$mode")
for r in rows[-14:]:
    print(r[4][:70].ljust(70), r[-1])
PY
done
