# round 2, 2-GPU call: page gather without a device-wide sync per step, on a high-priority stream - CTA count sweep (8 units in flight per thread), the
# 4-unit library at two CTA counts, then the default line
mkdir -p gpurun_out
export PYTEST_ADDOPTS="--timeout 300"
timeout 300 python -m pytest tests/test_gpu_device_gather.py -m gpu -x -q 2>&1 | tail -1
port=29700
run() { port=$((port+1)); timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-index-free "$@"; }
for c in 8 16 32 64; do run --gather-ctas $c > gpurun_out/r02_zz3_u8_c$c.json 2> gpurun_out/r02_zz3_u8_c$c.err; done
for c in 16 64; do PCOB200_LIB=$PWD/pcodec_b200/libcpcodec_gu4.so run --gather-ctas $c > gpurun_out/r02_zz3_u4_c$c.json 2> gpurun_out/r02_zz3_u4_c$c.err; done
run --no-gather-pages > gpurun_out/r02_zz3_sizes.json 2> gpurun_out/r02_zz3_sizes.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_zz3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); g=d.get('gather') or {}
        print(f.split('/')[-1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'gather_ms', g.get('device_ms'), 'ctas', g.get('ctas'), 'fused', round(d['kernel_ms'].get('fused_narrow_kernel',0),3), 'split', round(d['kernel_ms'].get('split_count_kernel',0),3), 'pack', round(d['kernel_ms'].get('pack_kernel',0),3))
    except Exception as ex: print(f, 'ERR', ex, open(f.replace('.json','.err')).read()[-500:])
PY
