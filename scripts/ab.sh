#!/bin/bash
# A/B run on one GPU box: the in-tree build, then every zig_gpt2_b200/variants/libzg_*.so (scripts/build_variant.py).
mkdir -p gpurun_out
{
for rep in 1 2; do
  timeout 120 python scripts/ab_time.py 124M ${AB_POS:-24,200,460}
  for so in zig_gpt2_b200/variants/libzg_*.so; do
    [ -e "$so" ] && ZG_B200_LIB=$PWD/$so timeout 120 python scripts/ab_time.py 124M ${AB_POS:-24,200,460}
  done
done
} 2>&1 | tee gpurun_out/ab.txt
