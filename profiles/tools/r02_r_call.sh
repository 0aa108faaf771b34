# round 2, call r: one slice in flight + 2 + 2 streaming threads in the e2e leg, tANS fallbacks (iteration cap 32, two-phase in-order pass), 16-byte gather copy
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_r_pytest.txt
timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee gpurun_out/r02_r_wide_spans.txt
DTYPE=int32 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_r_wide_spans.txt
DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_r_wide_spans.txt
for t in 2 3 1; do
timeout 400 python bench.py --no-cpu-baseline --no-index-free --steps 5 --e2e-threads $t > gpurun_out/r02_r_bench_t$t.json 2> gpurun_out/r02_r_bench_t$t.err
done
timeout 400 python bench.py --no-cpu-baseline --no-index-free --steps 5 --e2e-threads 2 --e2e-groups 16 > gpurun_out/r02_r_bench_t2g16.json 2> gpurun_out/r02_r_bench_t2g16.err
python - <<'PY'
import json
for t in ('t2','t3','t1','t2g16'):
    try:
        d=json.loads(open(f'gpurun_out/r02_r_bench_{t}.json').read().strip().splitlines()[-1]); e=d['e2e']
        print(t, 'value', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'call', round(d['roofline']['call']['frac'],4), 'solve', d['kernel_ms'].get('plan_solve_kernel'))
        print('   e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
    except Exception as ex: print(t, 'ERR', ex)
PY
