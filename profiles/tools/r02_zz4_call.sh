# round 2, last call (1 GPU): the default bench line of the final tree (timed loop without a device-wide sync per step)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r02_zz4_bench.json 2> gpurun_out/r02_zz4_bench.err; tail -3 gpurun_out/r02_zz4_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_zz4_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'frac', round(r['frac'],4), 'call', round(r['call']['frac'],4), 'compress', round(r['compress']['frac'],4), 'e2e', round(e['value']), 'cpu', round(d['cpu_baseline']['value']), 'parity', d['parity'].get('checked_chunks'), 'clocks', d['clocks'])
PY
