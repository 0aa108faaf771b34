# round 2, call w (1 GPU): the state the round ends on - full GPU suite (speculative index-free walk and in-place host buffers included),
# in-place e2e probe, default bench + reference arm, ncu launch list of the bench command, ncu --set full of fused_narrow_kernel and of
# the five compress kernels.  Everything lands in gpurun_out/r02_w_*; profiles/tools/r02_w_post.sh turns the reports into profiles/ files.
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r02_w_pytest.txt; tail -6 gpurun_out/r02_w_pytest.txt
timeout 300 python profiles/tools/e2e_in_place_probe.py > gpurun_out/r02_w_e2e_in_place.txt 2> gpurun_out/r02_w_e2e_in_place.err; cat gpurun_out/r02_w_e2e_in_place.txt; tail -3 gpurun_out/r02_w_e2e_in_place.err
timeout 600 python bench.py > gpurun_out/r02_w_bench.json 2> gpurun_out/r02_w_bench.err; tail -3 gpurun_out/r02_w_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_w_bench.json').read().strip().splitlines()[-1]); e=d['e2e']; r=d['roofline']
    print('value', round(d['value']), 'frac', round(r['frac'],4), 'call', round(r['call']['frac'],4), 'compress', round(r['compress']['frac'],4), d['kernel_ms'])
    print('e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
    print('abi3', e.get('reference_abi'))
    print('index free', d['index_free_decompress']['ms'], d['index_free_decompress']['kernel_ms'])
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], 'clocks', d['clocks'])
except Exception as ex: print('bench line unreadable', ex)
PY
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_w_bench_reference.json 2> gpurun_out/r02_w_bench_reference.err; tail -c 400 gpurun_out/r02_w_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_w_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_w_launches.log 2>&1; tail -2 gpurun_out/r02_w_launches.log | cut -c1-300
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'fused_narrow_kernel' -s 2 -c 1 -f -o gpurun_out/r02_w_fused python profiles/tools/decompress_time.py > gpurun_out/r02_w_ncu.log 2>&1; tail -2 gpurun_out/r02_w_ncu.log
REPS=2 timeout 600 ncu --set full --clock-control none -k regex:'split_count_kernel|plan_solve_kernel|bin_lut_kernel|ans_encode_kernel|pack_kernel' -s 5 -c 5 -f -o gpurun_out/r02_w_comp python profiles/tools/compress_once.py >> gpurun_out/r02_w_ncu.log 2>&1; tail -2 gpurun_out/r02_w_ncu.log
ls -la gpurun_out | tail -20
