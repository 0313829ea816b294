for d in 0 1 2 4 7; do echo "== ZG_DEBUG=$d"; ZG_DEBUG=$d timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "forward_state or generate_greedy_64" 2>&1 | tail -4; done
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke 2>&1 | tail -15
