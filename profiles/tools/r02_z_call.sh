# round 2, call z (1 GPU): small readbacks by kernel into a mapped bounce buffer (no copy-queue entry) - suite, streamed e2e A/B, bench
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r02_z_pytest.txt; tail -4 gpurun_out/r02_z_pytest.txt
for rk in 1 0; do
  PCOB200_READBACK_KERNEL=$rk timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 16x2,16x3,32x3,8x2 2>/dev/null | sed "s/^/readback_kernel=$rk /" | tee -a gpurun_out/r02_z_e2e_readback.txt
done
PCOB200_TRACE=1 timeout 300 python profiles/tools/e2e_in_place_probe.py 1024 0 16x2 > /dev/null 2> gpurun_out/r02_z_trace_16x2.err
awk '/=== timed passes/{f=1} f' gpurun_out/r02_z_trace_16x2.err | head -70 > gpurun_out/r02_z_trace_16x2.txt; rm -f gpurun_out/r02_z_trace_16x2.err
timeout 600 python bench.py > gpurun_out/r02_z_bench.json 2> gpurun_out/r02_z_bench.err; tail -3 gpurun_out/r02_z_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_z_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
    print('value', round(d['value']), 'frac', round(r['frac'],4), 'call', round(r['call']['frac'],4), 'compress', round(r['compress']['frac'],4))
    print('e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
    print('abi3', e.get('reference_abi'))
except Exception as ex: print('bench line unreadable', ex)
PY
