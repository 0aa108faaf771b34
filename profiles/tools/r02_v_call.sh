# round 2, call v (2 GPUs): page gather with the 16-byte / 4-in-flight copy - CTA count sweep, then the default line (e2e included)
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
timeout 600 python -m pytest tests/test_gpu_device_gather.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/r02_v_pytest.txt
port=29510
for ctas in 16 32 64 128; do
port=$((port+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free --gather-ctas $ctas > gpurun_out/r02_v_bench_2gpu_c$ctas.json 2> gpurun_out/r02_v_bench_2gpu_c$ctas.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free --no-gather-pages > gpurun_out/r02_v_bench_2gpu_sizes.json 2> gpurun_out/r02_v_bench_2gpu_sizes.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_v_bench_2gpu.json 2> gpurun_out/r02_v_bench_2gpu.err
python - <<'PY'
import json
for t in ('c16','c32','c64','c128','sizes',''):
    f=f'gpurun_out/r02_v_bench_2gpu_{t}.json' if t else 'gpurun_out/r02_v_bench_2gpu.json'
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get('gather') or {}; e=d.get('e2e') or {}
        print(t or 'default', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'gather_ms', g.get('device_ms'), 'fused', d['kernel_ms'].get('fused_narrow_kernel'), 'e2e', e.get('value'))
    except Exception as ex: print(t, 'ERR', ex, open(f.replace('.json','.err')).read()[-600:])
PY
