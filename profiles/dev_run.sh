set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiny0 and delta" 2>&1 | tail -1
