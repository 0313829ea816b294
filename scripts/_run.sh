timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/pos_sweep.py 124M
