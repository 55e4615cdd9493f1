set -x
timeout 900 python -m pytest tests/test_cli_gpu.py -x -q -m gpu -k "sharded" 2>&1 | grep -v "^$" | tail -40 | cut -c1-1500
