# round 2, call f: find the failures of call e (every step under a timeout; full tracebacks kept)
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 300 python -m pytest tests/test_c_dropin.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02_f_1_dropin.txt; tail -25 gpurun_out/r02_f_1_dropin.txt
( timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_narrow.py tests/test_gpu_encode.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02_f_2_core.txt; tail -25 gpurun_out/r02_f_2_core.txt
( timeout 600 python -m pytest tests/test_gpu_baseline_fullsize.py tests/test_gpu_auto_mode_search.py tests/test_gpu_auto.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02_f_3_full.txt; tail -25 gpurun_out/r02_f_3_full.txt
( timeout 400 python -m pytest tests/test_gpu_device_gather.py -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02_f_4_gather.txt; tail -25 gpurun_out/r02_f_4_gather.txt
( timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_device_gather.py 2>&1 | tail -40 ) > gpurun_out/r02_f_5_all.txt; tail -15 gpurun_out/r02_f_5_all.txt
timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_f_bench.json 2> gpurun_out/r02_f_bench.err; tail -c 1500 gpurun_out/r02_f_bench.json; tail -5 gpurun_out/r02_f_bench.err
