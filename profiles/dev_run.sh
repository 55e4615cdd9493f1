set -x
cp wepp_b200/libwepp_b200.so /tmp/lib_orig.so
for U in 512 128; do cp profiles/tmp_libs/lib_unit_$U.so wepp_b200/libwepp_b200.so; echo UNIT $U; timeout 600 python profiles/dev_paths.py 1.0 2>&1 | grep "^delta {" ; done
cp /tmp/lib_orig.so wepp_b200/libwepp_b200.so
