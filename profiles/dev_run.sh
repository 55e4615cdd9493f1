set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_full.py -x -q -m gpu -k "delta or state_place or c1" 2>&1 | tail -3
timeout 600 python profiles/dev_paths.py 1.0 2>&1 | tail -3
cp wepp_b200/libwepp_b200.so /tmp/lib_orig.so
cp profiles/tmp_libs/lib_plain.so wepp_b200/libwepp_b200.so; echo PLAIN; timeout 600 python profiles/dev_paths.py 1.0 2>&1 | grep "^delta"
cp /tmp/lib_orig.so wepp_b200/libwepp_b200.so
