# round 2, call b: the fused decode kernel - tests first, then the bench with both paths (A/B), then the launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02_b_pytest.txt; cat gpurun_out/r02_b_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_b_bench_fused.json 2> gpurun_out/r02_b_bench_fused.err; tail -c 2500 gpurun_out/r02_b_bench_fused.json; tail -5 gpurun_out/r02_b_bench_fused.err
PCOB200_FUSED=0 timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-index-free > gpurun_out/r02_b_bench_unfused.json 2> gpurun_out/r02_b_bench_unfused.err; tail -c 1500 gpurun_out/r02_b_bench_unfused.json
