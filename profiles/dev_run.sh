set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
for C in C1 C2; do timeout 900 python bench.py --config $C > gpurun_out/r02_config_$C.json 2> gpurun_out/config_$C.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_config_$C.json').read().strip().splitlines()[-1])
print('$C', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['parity'].get('all_green'), d['stats']['place_path'])
PY
done
