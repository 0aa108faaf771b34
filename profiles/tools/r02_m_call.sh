# round 2, call m: sliced host<->device copies - duplex probe with persistent threads, bench with the e2e trace
mkdir -p gpurun_out
timeout 300 python profiles/tools/e2e_duplex_probe.py 2>&1 | tail -9 | tee gpurun_out/r02_m_e2e_probe.txt
timeout 500 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err; tail -c 900 gpurun_out/r02_m_bench.json; tail -5 gpurun_out/r02_m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_m_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print('e2e', e['value'], e['ms_per_step'], 'single', e['single_call']['ms_per_step']); print(e['trace_ms'])
PY
