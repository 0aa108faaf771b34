# round 2, call u: NaN sign fix of the FloatMult joins (property hunt over 14 seeds), tANS fallback policy, default bench
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_u_pytest.txt
for seed in 5 7 11 21 22 23 24 25 26 27 28 29 30 31; do
  timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -x -q --hypothesis-seed=$seed -p no:cacheprovider 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r02_u_prop_hunt.txt
ls gpurun_out/prop_fail_* 2>/dev/null | head -4
for d in int64 int32 float64 float32; do DTYPE=$d timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1; done | tee gpurun_out/r02_u_wide_spans.txt
timeout 500 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r02_u_bench.json 2> gpurun_out/r02_u_bench.err; tail -3 gpurun_out/r02_u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_u_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print('value', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'call', round(d['roofline']['call']['frac'],4), d['kernel_ms'])
print('e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
print('index free', d['index_free_decompress']['ms'], d['index_free_decompress']['kernel_ms'])
PY
