# round 2, final evidence call (1 GPU) on the library the round ends on: ncu launch list of the bench command, default bench, reference arm,
# config sweep at 512 chunks per row
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_zz_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_zz_launches.log 2>&1; tail -1 gpurun_out/r02_zz_launches.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/r02_zz_bench.json 2> gpurun_out/r02_zz_bench.err; tail -3 gpurun_out/r02_zz_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_zz_bench_reference.json 2> gpurun_out/r02_zz_bench_reference.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_zz_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
    print('value', round(d['value']), 'frac', round(r['frac'],4), 'call', round(r['call']['frac'],4), 'compress', round(r['compress']['frac'],4), 'traffic', r['traffic'], 'launches', d['gpu_launches'])
    print('e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2))
    print('abi3', e.get('reference_abi'))
    ref=json.loads(open('gpurun_out/r02_zz_bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', round(ref['value']), ref['cpu_baseline']['cores'])
except Exception as ex: print('bench line unreadable', ex)
PY
N_CHUNKS=512 timeout 500 python profiles/tools/config_sweep.py gpurun_out/r02_zz_config_sweep.md > gpurun_out/r02_zz_sweep.log 2>&1; tail -1 gpurun_out/r02_zz_sweep.log | cut -c1-100
