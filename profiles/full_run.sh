set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r02_bench_v3_1gpu.json 2> gpurun_out/r02_bench_v3_1gpu.err; tail -c 1500 gpurun_out/r02_bench_v3_1gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_v3_1gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('e2e_cold'), d['cpu_baseline'])
PY
