set -x
cp wepp_b200/libwepp_b200.so /tmp/lib_orig.so
for V in 3x6 4x4; do cp profiles/tmp_libs/lib_$V.so wepp_b200/libwepp_b200.so; echo VARIANT $V; timeout 600 python profiles/dev_paths.py 1.0 2>&1 | grep "^delta" | cut -c1-110; done
cp /tmp/lib_orig.so wepp_b200/libwepp_b200.so
