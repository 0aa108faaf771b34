# round 2, call s: hunt the rare property-test failure (falsifying examples are dumped to gpurun_out/prop_fail_*), wait-hint variants of the fused kernel
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
for seed in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 300 python -m pytest tests/test_gpu_properties.py -m gpu -x -q --hypothesis-seed=$seed -p no:cacheprovider 2>&1 | tail -3 | head -2
done 2>&1 | tee gpurun_out/r02_s_prop_hunt.txt
ls gpurun_out/prop_fail_* 2>/dev/null | head
bash profiles/tools/r01_variants.sh default hint500 hint2000 hint20000 link40 hint2000link40 2>&1 | tee gpurun_out/r02_s_variants.txt
