set -x
timeout 600 python -m pytest tests/test_cli_gpu.py -x -q -m gpu -k "nomask" 2>&1 | tail -2
