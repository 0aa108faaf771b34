# round 2, call d: full GPU suite on the thread-local-context library, ncu --set full of fused_narrow_kernel, launch list of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_d_pytest.txt; cat gpurun_out/r02_d_pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fused_narrow_kernel' -s 2 -c 1 -f -o gpurun_out/r02_d_fused python profiles/tools/decompress_time.py > gpurun_out/r02_d_ncu.log 2>&1; tail -2 gpurun_out/r02_d_ncu.log
timeout 300 python profiles/tools/decompress_time.py
