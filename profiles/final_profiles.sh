# Round-2 profile artefacts (run on the GPU box: gpurun -- 'bash profiles/final_profiles.sh'); summaries are written
# here with profiles/ncu_summary.py and copied into profiles/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-parity --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
tail -c 600 gpurun_out/r02_launches_bench.log
ncu --set full --clock-control none --import-source on -k regex:'delta_place_kernel|node_tile_kernel' -s 4 -c 2 -o gpurun_out/r02_top -f python profiles/dev_one.py 1.0 4 > gpurun_out/r02_top.log 2>&1
tail -2 gpurun_out/r02_top.log
