set -x
timeout 900 python -m pytest tests/test_golden.py -x -q -m gpu -k "peak_loop" 2>&1 | tail -3
WEPP_TIMING=1 timeout 900 python profiles/peaks_run.py 1.0 2>&1 | grep "initial filter\|peak loop\|neighbour\|filter_peaks_s" | cut -c1-330
