# round 2, call n: transfer-function tANS fallback + throttled sliced copies
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_n_pytest.txt
timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -16 | tee gpurun_out/r02_n_wide_spans.txt
timeout 300 python profiles/tools/e2e_duplex_probe.py 2>&1 | tail -9 | tee gpurun_out/r02_n_e2e_probe.txt
timeout 500 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r02_n_bench.json 2> gpurun_out/r02_n_bench.err; tail -c 600 gpurun_out/r02_n_bench.json; tail -5 gpurun_out/r02_n_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_n_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print('e2e', e['value'], e['ms_per_step'], 'single', e['single_call']['ms_per_step']); print(e['trace_ms'])
PY
