set -x
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -x -q -m gpu -k "group" 2>&1 | tail -2
