#!/bin/bash
mkdir -p gpurun_out
{
for v in "" "ZG_HEAD_GENERAL=1"; do
  echo "variant '$v'"
  env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l26.csv python scripts/profile_batch.py decode 128 > /dev/null 2>&1
  python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/l26.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-4:-1]: print('  B=128', r[4][:60].ljust(60), r[-1])
PY
  for m in 1 0; do
  env $v ZG_TF32=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l26.csv python scripts/profile_batch.py decode_xl > /dev/null 2>&1
  python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/l26.csv')) if len(ln)>5 and ln[0].isdigit()]
for r in rows[-4:-1]: print('  xl tf32=$m', r[4][:60].ljust(60), r[-1])
PY
  done
done
} > gpurun_out/r2_exp26.txt 2>&1
cat gpurun_out/r2_exp26.txt
