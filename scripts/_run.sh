timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | grep -E "unprofiled"
ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_prof.so timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | grep -E "unprofiled|P[1-5]|lm_head|sum"
echo "== nowait"; ZG_B200_LIB=$PWD/zig_gpt2_b200/variants/libzg_prof_nowait.so timeout 300 python scripts/phase_profile.py 124M 32 2>&1 | grep -E "unprofiled|P[1-5]|lm_head|sum"
for s in 355M 1.5B; do timeout 600 python scripts/phase_profile.py $s 16 2>&1 | grep -E "unprofiled"; done
