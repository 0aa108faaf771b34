# round-1 "m" evidence (the state the round ends on): full GPU test suite, bench (both arms), ncu launch list of the bench command, ncu --set full of the
# decompress kernels and of the two heaviest compress kernels, config sweep (BASELINE configs 1, 3, 5)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r01_m_pytest.txt; cat gpurun_out/r01_m_pytest.txt
python bench.py > gpurun_out/r01_m_bench.json 2> gpurun_out/r01_m_bench.err; tail -c 1200 gpurun_out/r01_m_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01_m_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_m_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'decode_narrow_kernel|symwalk_kernel' -s 2 -c 2 -f -o gpurun_out/r01_m_decomp python profiles/tools/decompress_time.py > gpurun_out/r01_m_ncu.log 2>&1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:'split_count_kernel|pack_kernel' -s 2 -c 2 -f -o gpurun_out/r01_m_comp python profiles/tools/compress_once.py >> gpurun_out/r01_m_ncu.log 2>&1
tail -2 gpurun_out/r01_m_ncu.log
timeout 900 python profiles/tools/config_sweep.py gpurun_out/r01_m_config_sweep.md > gpurun_out/sweep.log 2>&1; tail -2 gpurun_out/sweep.log
