# round-1 "k" evidence: full GPU test suite, bench (both arms), ncu launch list of the bench command, ncu --set full of the decompress kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r01_k_pytest.txt; cat gpurun_out/r01_k_pytest.txt
python bench.py > gpurun_out/r01_k_bench.json 2> gpurun_out/r01_k_bench.err; tail -c 1500 gpurun_out/r01_k_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_k_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_k_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'decode_narrow_kernel|symwalk_kernel' -s 2 -c 2 -f -o gpurun_out/r01_k_decomp python profiles/tools/decompress_time.py > gpurun_out/r01_k_ncu.log 2>&1
tail -2 gpurun_out/r01_k_ncu.log
