# round 2, call i: suite (lookback candidate, walker, sort), bench with the e2e call trace, index-free timing, wide-range sweep rows
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r02_i_pytest.txt; tail -14 gpurun_out/r02_i_pytest.txt
timeout 200 python profiles/tools/walk_once.py 2>&1 | tail -1
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_i_bench.json 2> gpurun_out/r02_i_bench.err; tail -c 600 gpurun_out/r02_i_bench.json; tail -5 gpurun_out/r02_i_bench.err
N_CHUNKS=128 timeout 600 python profiles/tools/config_sweep.py gpurun_out/r02_i_config_sweep.md > gpurun_out/r02_i_sweep.log 2>&1; grep -E "order 0|order 2|C1|C3" gpurun_out/r02_i_config_sweep.md | cut -c1-150
