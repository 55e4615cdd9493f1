set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_full.py -x -q -m gpu -k "delta or state_place or c1" 2>&1 | tail -2
timeout 600 python profiles/dev_paths.py 1.0 2>&1 | tail -3
