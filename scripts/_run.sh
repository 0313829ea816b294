ZG_DEBUG=8 timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | tail -60
for s in 355M 1.5B; do timeout 600 python scripts/phase_profile.py $s 16 2>&1 | grep -E "unprofiled|sum|lm_head"; done
