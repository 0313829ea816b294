#!/bin/bash
mkdir -p gpurun_out
{
for v in "" hint20 hint200; do
  echo "variant '$v'"
  if [ -z "$v" ]; then L=""; else L="ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_$v.so"; fi
  env $L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/l24_$v.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
  python - <<PY
import csv
rows=[ln for ln in csv.reader(open('gpurun_out/l24_$v.csv')) if len(ln)>5 and ln[0].isdigit()]
print(' '.join(r[4].split('(')[0].split('::')[-1][:18]+'='+r[-1] for r in rows[-12:-1]))
PY
  for c in cfg3 cfg4 cfg5; do
  env $L timeout 600 python scripts/bench_configs.py $c --trials 3 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        r = json.loads(ln); print(r['record'], round(r['value']), round(r['ms_per_step'],3), round(r['roofline']['frac'],3))
"
  done
done
} > gpurun_out/r2_exp24.txt 2>&1
cat gpurun_out/r2_exp24.txt
