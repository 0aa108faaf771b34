# First GPU call of round 2 (gpurun --timeout 1500 -- 'bash profiles/tools/r02_first_call.sh'): everything written after round 1's GPU
# budget ran out, in the order that decides what to do next.  Each step writes under gpurun_out/ and the script never stops on a failure.
mkdir -p gpurun_out
# 1. the suite the driver runs (includes tests/test_gpu_zz_reference_pages.py: GPU bytes against pages the Rust crate wrote)
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_a_pytest.txt; cat gpurun_out/r02_a_pytest.txt
# 2. the opt-in Auto mode search inside compress and the GPU-vs-oracle property test
PCOB200_RUN_UNVALIDATED=1 python -m pytest tests/test_gpu_auto_mode_search.py tests/test_gpu_properties.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_a_unvalidated.txt; cat gpurun_out/r02_a_unvalidated.txt
# 3. both bench arms: the CPU arm is now one worker process per host thread - never yet run on the GPU box's host
python bench.py --results-csv gpurun_out/r02_a_results.csv > gpurun_out/r02_a_bench.json 2> gpurun_out/r02_a_bench.err; tail -c 1500 gpurun_out/r02_a_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_a_bench_reference.json 2> gpurun_out/r02_a_bench_reference.err; tail -c 900 gpurun_out/r02_a_bench_reference.json
# 4. launch list of the bench command (share of each kernel in the step)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_a_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-index-free > /dev/null 2>&1
tail -3 gpurun_out/r02_a_launches.csv
