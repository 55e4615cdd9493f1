set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'delta_|node_tile|sacc_pack' -s 12 -c 8 --csv --log-file gpurun_out/dev_launches.csv python profiles/dev_one.py 1.0 4 > gpurun_out/dev_one.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:delta_place_kernel -s 2 -c 1 -o gpurun_out/delta_v7 -f python profiles/dev_one.py 1.0 4 > gpurun_out/dev_ncu.log 2>&1
tail -3 gpurun_out/dev_ncu.log
