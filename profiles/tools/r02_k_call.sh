# round 2, call k: diagnostics - PCIe duplex from two host threads, per-kernel spans of the wide-range compress, walker launch shapes; tests on the 8/16-bit fused path
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 150"
( timeout 900 python -m pytest tests/test_gpu_decode.py tests/test_gpu_narrow.py tests/test_gpu_properties.py tests/test_gpu_baseline_fullsize.py -m gpu -q 2>&1 | tail -20 ) > gpurun_out/r02_k_pytest.txt; tail -8 gpurun_out/r02_k_pytest.txt
timeout 200 python profiles/tools/pcie_duplex.py 2>&1 | tail -2 | tee gpurun_out/r02_k_pcie.txt
( DTYPE=int64 timeout 200 python profiles/tools/wide_spans.py; DTYPE=int32 timeout 200 python profiles/tools/wide_spans.py; DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py ) 2>&1 | grep chunks | tee gpurun_out/r02_k_wide_spans.txt
for v in default wk128x8 wk64x16 wk64x12 wk32x8; do if [ "$v" = default ]; then f=libcpcodec.so; else f=libcpcodec_$v.so; fi; echo -n "$v: "; PCOB200_LIB=$PWD/pcodec_b200/$f timeout 200 python profiles/tools/walk_once.py 2>&1 | tail -1; done | tee gpurun_out/r02_k_walk_variants.txt
N_CHUNKS=128 timeout 600 python profiles/tools/config_sweep.py gpurun_out/r02_k_config_sweep.md > gpurun_out/r02_k_sweep.log 2>&1; grep -E "uint8|uint16" gpurun_out/r02_k_config_sweep.md | cut -c1-150
