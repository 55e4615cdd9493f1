set -x
timeout 900 python -m pytest tests/test_golden.py tests/test_cli_gpu.py -x -q -m gpu -k "peak_loop or selective or sharded" 2>&1 | tail -3
WEPP_TIMING=1 timeout 900 python profiles/peaks_run.py 1.0 2>&1 | grep "initial filter\|peak loop\|neighbour\|filter_peaks_s" | cut -c1-330
