# round 2, last GPU seconds: the full GPU suite on the tree with per-run unoptimized_bins_log
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 100"
( timeout 80 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r02_zz7_pytest.txt; tail -4 gpurun_out/r02_zz7_pytest.txt
