# round 2, call p: deep-guess tANS fallback, 128-thread fused-kernel variants, phase trace of the streamed e2e calls
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_p_pytest.txt
timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -3 | tee gpurun_out/r02_p_wide_spans.txt
DTYPE=int32 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_p_wide_spans.txt
DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_p_wide_spans.txt
bash profiles/tools/r01_variants.sh default t128w1b2r64x7 t128w1b2r32x7 t128w1b2r128x6 2>&1 | tee gpurun_out/r02_p_variants.txt
PCOB200_TRACE=1 timeout 500 python bench.py --no-cpu-baseline --no-index-free --steps 3 > gpurun_out/r02_p_bench.json 2> gpurun_out/r02_p_bench.err
tail -40 gpurun_out/r02_p_bench.err > gpurun_out/r02_p_trace_tail.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_p_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print('value', d['value'], 'frac', d['roofline']['frac'], 'call', d['roofline']['call']['frac'])
print('e2e', e['value'], e['ms_per_step'], 'single', e['single_call']['ms_per_step'], e.get('pass_wall_ms')); print(e['trace_ms'])
PY
