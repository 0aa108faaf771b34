# round 2, last verification (1 GPU): the full GPU suite and smoke() on the final tree (chain logic factored into follow_chunk_chain)
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r02_zz5_pytest.txt; tail -3 gpurun_out/r02_zz5_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02_zz5_smoke.txt
