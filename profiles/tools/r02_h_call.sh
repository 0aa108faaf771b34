# round 2, call h: suite after the conv1 / test fixes, PCIe duplex check, walk_kernel source-level profile, bench
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r02_h_pytest.txt; tail -14 gpurun_out/r02_h_pytest.txt
timeout 120 python profiles/tools/pcie_duplex.py 2>&1 | tail -2 | tee gpurun_out/r02_h_pcie.txt
timeout 200 python profiles/tools/walk_once.py 2>&1 | tail -1
N_CHUNKS=148 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'walk_kernel' -s 1 -c 1 -f -o gpurun_out/r02_h_walk python profiles/tools/walk_once.py > gpurun_out/r02_h_ncu.log 2>&1; tail -2 gpurun_out/r02_h_ncu.log
timeout 400 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_h_bench.json 2> gpurun_out/r02_h_bench.err; tail -c 1200 gpurun_out/r02_h_bench.json; tail -5 gpurun_out/r02_h_bench.err
N_CHUNKS=128 timeout 600 python profiles/tools/config_sweep.py gpurun_out/r02_h_config_sweep.md > gpurun_out/r02_h_sweep.log 2>&1; grep -E "order 0|C1|C3" gpurun_out/r02_h_config_sweep.md | cut -c1-150
