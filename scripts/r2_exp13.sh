#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_prefill -s 3 -c 1 -f -o gpurun_out/r02_attn_prefill_full python scripts/profile_batch.py prefill > gpurun_out/ncu_attn.log 2>&1
tail -3 gpurun_out/ncu_attn.log
