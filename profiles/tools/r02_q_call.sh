# round 2, call q (2 GPUs): the scaling bench with the device-side page gather in the step (NCCL sizes + cudaIpc P2P stores), sizes only,
# the reference arm under torchrun, the two-process gather test on two GPUs; plus the encoder / copy changes since call p
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 200"
nvidia-smi topo -m 2>/dev/null | head -8 > gpurun_out/r02_q_topo.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02_q_pytest.txt
timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee gpurun_out/r02_q_wide_spans.txt
DTYPE=float64 timeout 200 python profiles/tools/wide_spans.py 2>&1 | tail -1 | tee -a gpurun_out/r02_q_wide_spans.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_q_bench_2gpu_gather.json 2> gpurun_out/r02_q_bench_2gpu_gather.err; tail -c 1800 gpurun_out/r02_q_bench_2gpu_gather.json; tail -8 gpurun_out/r02_q_bench_2gpu_gather.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free --no-gather-pages > gpurun_out/r02_q_bench_2gpu_sizes.json 2> gpurun_out/r02_q_bench_2gpu_sizes.err; tail -c 600 gpurun_out/r02_q_bench_2gpu_sizes.json; tail -5 gpurun_out/r02_q_bench_2gpu_sizes.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r02_q_ref_2gpu.json 2> gpurun_out/r02_q_ref_2gpu.err; tail -c 700 gpurun_out/r02_q_ref_2gpu.json; tail -3 gpurun_out/r02_q_ref_2gpu.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-index-free > gpurun_out/r02_q_bench_1gpu.json 2>gpurun_out/r02_q_bench_1gpu.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_q_bench_2gpu_gather.json','gpurun_out/r02_q_bench_2gpu_sizes.json','gpurun_out/r02_q_bench_1gpu.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); e=d.get('e2e') or {}
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'gather', d.get('gather'), 'kernel_ms', d.get('kernel_ms'))
        if e.get('value'): print('   e2e', round(e['value']), e['ms_per_step'], 'single', e['single_call']['ms_per_step'], e.get('pass_wall_ms'))
    except Exception as ex: print(f, 'ERR', ex)
PY
