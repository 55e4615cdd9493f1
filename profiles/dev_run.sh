set -x
cd profiles/microbench && g++ -O2 -std=c++17 -pthread uncondense_bench.cpp ../../wepp_b200/build/host_io.o -lz -o uncondense_bench && WEPP_TIMING=1 ./uncondense_bench 8000000 2>&1 | tail -6
