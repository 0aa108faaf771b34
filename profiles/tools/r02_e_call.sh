# round 2, call e: whole GPU suite (failures in full), fused-kernel variants (4 / 3 walkers, wait back-off), the new bench.py
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_e_pytest.txt; tail -40 gpurun_out/r02_e_pytest.txt
bash profiles/tools/r01_variants.sh default w4b5r64 w4b5r64s w3b4r64 w2b3r128s w4b4r64 w4b5r64s5 w3b4r64s > gpurun_out/r02_e_variants.txt 2>&1; cat gpurun_out/r02_e_variants.txt
timeout 900 python bench.py > gpurun_out/r02_e_bench.json 2> gpurun_out/r02_e_bench.err; tail -c 3500 gpurun_out/r02_e_bench.json; tail -15 gpurun_out/r02_e_bench.err
