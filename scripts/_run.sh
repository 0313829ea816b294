timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/pos_sweep.py 124M 2>&1
timeout 300 python scripts/clock_profile.py 124M 16 100 2>&1 | tail -70
