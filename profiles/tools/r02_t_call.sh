# round 2, call t: the falsifying examples of the property test (seeds 5, 7, 11 of call s), full text + dumps; tANS fallback policy; e2e slice sizes
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
for seed in 5 7 11; do
  timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -x -q --hypothesis-seed=$seed -p no:cacheprovider 2>&1 | grep -v "^E    *[0-9.e+-]*,$" | tail -60 > gpurun_out/r02_t_prop_seed$seed.txt
  tail -5 gpurun_out/r02_t_prop_seed$seed.txt
done
ls gpurun_out/prop_fail_* 2>/dev/null | head
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_properties.py::test_gpu_bytes_equal_the_oracles 2>&1 | tail -4 | tee gpurun_out/r02_t_pytest.txt
timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee gpurun_out/r02_t_wide_spans.txt
DTYPE=int32 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_t_wide_spans.txt
DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_t_wide_spans.txt
DTYPE=float32 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_t_wide_spans.txt
for mb in 16 32; do
PCOB200_COPY_SLICE_MB=$mb timeout 400 python bench.py --no-cpu-baseline --no-index-free --steps 4 --e2e-threads 2 --e2e-groups 16 > gpurun_out/r02_t_bench_s$mb.json 2> gpurun_out/r02_t_bench_s$mb.err
done
python - <<'PY'
import json
for t in ('s16','s32'):
    try:
        d=json.loads(open(f'gpurun_out/r02_t_bench_{t}.json').read().strip().splitlines()[-1]); e=d['e2e']
        print(t, 'value', round(d['value']), 'e2e', round(e['value']), round(e['ms_per_step'],2), 'single', round(e['single_call']['ms_per_step'],2), e.get('pass_wall_ms'))
    except Exception as ex: print(t, 'ERR', ex)
PY
