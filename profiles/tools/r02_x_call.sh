# round 2, call x (1 GPU): big host<->device copies take turns per direction (copy_sliced FIFO) - suite, streamed e2e probe for three
# slice sizes, default bench
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r02_x_pytest.txt; tail -4 gpurun_out/r02_x_pytest.txt
for mb in 32 16 8; do
  PCOB200_COPY_SLICE_MB=$mb timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 2>&1 | grep -v "^$" | tee -a gpurun_out/r02_x_e2e_fifo.txt
done
PCOB200_COPY_FIFO=0 timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 2>&1 | tee -a gpurun_out/r02_x_e2e_fifo.txt
timeout 600 python bench.py > gpurun_out/r02_x_bench.json 2> gpurun_out/r02_x_bench.err; tail -3 gpurun_out/r02_x_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_x_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
    print('value', round(d['value']), 'frac', round(r['frac'],4), 'call', round(r['call']['frac'],4), 'compress', round(r['compress']['frac'],4))
    print('e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
    for t in e['trace_ms']: print(t)
    print('abi3', e.get('reference_abi'))
except Exception as ex: print('bench line unreadable', ex)
PY
