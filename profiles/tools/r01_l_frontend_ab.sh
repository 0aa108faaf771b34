# A/B of the one-pass front end (split_count_kernel) against the two-kernel path, after the encode parity tests
mkdir -p gpurun_out
export PCOB200_UNVALIDATED=1 PCOB200_ONE_PASS_FRONT_END=1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['compress_mb_s'], d['decompress_mb_s'], d['kernel_ms'])"
PCOB200_ONE_PASS_FRONT_END=0 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['compress_mb_s'], d['decompress_mb_s'], d['kernel_ms'])"
for v in binl256 binl512 binl1024; do PCOB200_LIB=$PWD/pcodec_b200/libcpcodec_$v.so python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['compress_mb_s'], d['kernel_ms'].get('bin_lut_kernel'))"; done
