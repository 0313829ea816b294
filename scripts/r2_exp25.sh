#!/bin/bash
mkdir -p gpurun_out
{
for v in "" poly2 poly4 poly8; do
  echo "variant '$v'"
  if [ -z "$v" ]; then L=""; else L="ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_$v.so"; fi
  env $L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:attn_prefill -c 6 --csv --log-file gpurun_out/l25.csv python scripts/profile_batch.py prefill > /dev/null 2>&1
  grep attn_prefill gpurun_out/l25.csv | tail -4 | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
done
} > gpurun_out/r2_exp25.txt 2>&1
cat gpurun_out/r2_exp25.txt
