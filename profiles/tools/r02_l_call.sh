# round 2, call l: suite; e2e duplex probe (library calls against plain copies); wide-range spans after the tANS fallback; walker
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 ) > gpurun_out/r02_l_pytest.txt; tail -12 gpurun_out/r02_l_pytest.txt | cut -c1-1500
timeout 300 python profiles/tools/e2e_duplex_probe.py 2>&1 | tail -9 | tee gpurun_out/r02_l_e2e_probe.txt
( DTYPE=int64 timeout 200 python profiles/tools/wide_spans.py; DTYPE=int32 timeout 200 python profiles/tools/wide_spans.py; DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py ) 2>&1 | grep chunks | tee gpurun_out/r02_l_wide_spans.txt
timeout 200 python profiles/tools/walk_once.py 2>&1 | tail -1 | tee gpurun_out/r02_l_walk.txt
