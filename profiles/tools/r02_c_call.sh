# round 2, call c: fused-kernel variants (walker warps x ring buffers x stage row length), A/B against the round-1 pair
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_narrow.py tests/test_gpu_decode.py tests/test_gpu_baseline_fullsize.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r02_c_pytest.txt; cat gpurun_out/r02_c_pytest.txt
( bash profiles/tools/r01_variants.sh default w1b2r64 w2b3r128 w1b2r128 w3b3r64 w2b2r64 w2b3r256; PCOB200_FUSED=0 bash profiles/tools/r01_variants.sh default ) > gpurun_out/r02_c_variants.txt 2>&1
cat gpurun_out/r02_c_variants.txt
