set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_full.py tests/test_golden.py -x -q -m gpu -k "delta or c1 or fixtures or node" 2>&1 | tail -2
timeout 600 python profiles/dev_paths.py 1.0 2>&1 | tail -3
