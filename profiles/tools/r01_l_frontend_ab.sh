# wrapped tests + a short bench (kernel spans, index-free decompress timing)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_wrapped.py -x -q 2>&1 | tail -5
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['index_free_decompress'])"
