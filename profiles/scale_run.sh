# torchrun bench at N GPUs (the driver's launch line): bash profiles/scale_run.sh N
set -x
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_scale_v2_${N}gpu.json 2> gpurun_out/scale_$N.err
tail -c 300 gpurun_out/scale_$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r02_scale_v2_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','exchange_ms','n_gpus')}, d['e2e']['ms_per_step'], d.get('parity',{}).get('all_green'))
PY
