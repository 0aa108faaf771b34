# round 2, call j (2 GPUs): the scaling bench with the device-side page gather in the step, and with sizes only
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -8 > gpurun_out/r02_j_topo.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_j_bench_2gpu_gather.json 2> gpurun_out/r02_j_bench_2gpu_gather.err; tail -c 2500 gpurun_out/r02_j_bench_2gpu_gather.json; tail -8 gpurun_out/r02_j_bench_2gpu_gather.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free --no-gather-pages > gpurun_out/r02_j_bench_2gpu_sizes.json 2> gpurun_out/r02_j_bench_2gpu_sizes.err; tail -c 900 gpurun_out/r02_j_bench_2gpu_sizes.json; tail -5 gpurun_out/r02_j_bench_2gpu_sizes.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free > gpurun_out/r02_j_bench_1gpu.json 2>/dev/null; tail -c 400 gpurun_out/r02_j_bench_1gpu.json
