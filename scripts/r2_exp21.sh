#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair" -s 8 -c 3 -f -o gpurun_out/r02_pair_full python scripts/profile_batch.py prefill > gpurun_out/ncu21.log 2>&1
tail -3 gpurun_out/ncu21.log
