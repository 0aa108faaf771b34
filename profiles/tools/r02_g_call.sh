# round 2, call g: whole GPU suite after the attribute fix (every step under a timeout), bench, config sweep
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 ) > gpurun_out/r02_g_pytest.txt; tail -30 gpurun_out/r02_g_pytest.txt
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r02_g_pytest_all.txt; tail -12 gpurun_out/r02_g_pytest_all.txt
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.err; tail -c 1800 gpurun_out/r02_g_bench.json; tail -5 gpurun_out/r02_g_bench.err
N_CHUNKS=256 timeout 900 python profiles/tools/config_sweep.py gpurun_out/r02_g_config_sweep.md > gpurun_out/r02_g_sweep.log 2>&1; tail -3 gpurun_out/r02_g_sweep.log; cat gpurun_out/r02_g_config_sweep.md | cut -c1-200 | head -60
