# round 2, final 2-GPU call: the device page gather test and the default bench under torchrun (pages gathered; then sizes only)
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
timeout 600 python -m pytest tests/test_gpu_device_gather.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02_zz_pytest_gather.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_zz_bench_2gpu.json 2> gpurun_out/r02_zz_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-index-free --no-gather-pages > gpurun_out/r02_zz_bench_2gpu_sizes.json 2> gpurun_out/r02_zz_bench_2gpu_sizes.err
python - <<'PY'
import json
for t in ('', '_sizes'):
    f=f'gpurun_out/r02_zz_bench_2gpu{t}.json'
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get('gather') or {}; e=d.get('e2e') or {}
        print(t or 'pages', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'gather', g, 'e2e', e.get('value'))
    except Exception as ex: print(t, 'ERR', ex, open(f.replace('.json','.err')).read()[-800:])
PY
