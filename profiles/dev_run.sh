set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'delta_place_kernel|node_tile_kernel' -s 4 -c 2 -o gpurun_out/r02_top -f python profiles/dev_one.py 1.0 4 > gpurun_out/r02_top.log 2>&1
tail -1 gpurun_out/r02_top.log
