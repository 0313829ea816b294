#!/bin/bash
# us/token of the in-tree decode engine and of every variant for the wider GPT-2 sizes (short context)
for size in 355M 1.5B; do
  timeout 300 python scripts/ab_time.py $size 24
  for so in zig_gpt2_b200/variants/libzg_off.so; do [ -e "$so" ] && ZG_B200_LIB=$PWD/$so timeout 300 python scripts/ab_time.py $size 24; done
done
