set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_scale_v2_8gpu.json 2> gpurun_out/scale_8.err
tail -c 400 gpurun_out/scale_8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_scale_v2_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','exchange_ms','n_gpus')}, d['e2e']['ms_per_step'], d.get('parity',{}).get('all_green'))
PY
WEPP_TIMING=1 timeout 600 python profiles/peaks_run.py 1.0 --gpus 8 2>&1 | grep "peak loop\|neighbour\|filter_peaks" | tail -4
