# A/B of the one-pass front end (split_count_kernel): 8 vs 4 latents per thread and step, and the two-kernel path
mkdir -p gpurun_out
python -m pytest tests/test_gpu_encode.py tests/test_gpu_wrapped.py -x -q 2>&1 | tail -3
for v in default per4 default per4; do if [ "$v" = default ]; then f=libcpcodec.so; else f=libcpcodec_$v.so; fi; PCOB200_LIB=$PWD/pcodec_b200/$f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['compress_mb_s'], d['kernel_ms'].get('split_count_kernel'))"; done
PCOB200_ONE_PASS_FRONT_END=0 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('two-kernel', d['value'], d['compress_mb_s'], d['kernel_ms'])"
