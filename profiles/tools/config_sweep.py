"""BASELINE.json configs other than the bench workload, measured on one B200 next to the CPU port:
  C1  u32 lomax, classic, no delta, 1 chunk          (round trip bit-exact vs the oracle's bytes)
  C3  f64 decimal sinusoid, FloatMult(0.01) + consecutive order 2
  C5  {u8,u16,i32,i64,f32,f64} x consecutive orders 0..7, classic: compression ratio + MB/s, GPU vs CPU port
Every row: N_CHUNKS chunks of 2^18 numbers, buffers resident in HBM, compress_ex + decompress_ex through the C-ABI,
kernel time = sum of the library's CUDA-event spans (pco_b200_profile_last), GPU bytes compared with the oracle's on
chunk 0, decode compared with the input.  Writes a markdown table to argv[1].
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from oracle import pyoracle
from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, _lib, datagen

L = _lib.lib()
CH = 1 << 18
N_CHUNKS = int(os.environ.get("N_CHUNKS", "512"))
dev = torch.device("cuda")
TORCH_VIEW = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}


def spans():
    buf = C.create_string_buffer(4096)
    L.pco_b200_profile_last(buf, 4096)
    out = {}
    for item in buf.value.decode().split(";"):
        if "=" in item:
            k, v = item.split("=")
            out[k] = out.get(k, 0.0) + float(v)
    return out


_cache = {}


def run(name, dtype, gen, cfg, ocfg, n_chunks=N_CHUNKS, key=None):
    dt = np.dtype(dtype)
    if key is None or key not in _cache:
        _cache.clear()
        chunks = [gen(s) for s in range(n_chunks)]
        host = np.concatenate(chunks)
        dev_nums = torch.from_numpy(host.view(np.dtype(f"u{dt.itemsize}")).view(np.dtype(f"i{dt.itemsize}")) if dt.itemsize > 1 else host.view(np.uint8)).to(dev)
        _cache[key] = (chunks[: os.cpu_count() or 1], host.size, dev_nums)
        del host
    chunks, n, nums = _cache[key]
    dbyte = _lib.dtype_byte(dt)
    cap = L.pco_standalone_guarantee_file_size(n, dbyte)
    icap = L.pco_b200_index_size_bound(n, n_chunks)
    d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_idx = torch.empty(icap, dtype=torch.uint8, device=dev)
    d_out = torch.empty_like(nums)
    nw, il = C.c_size_t(), C.c_size_t()
    prog = _lib._CProgress()
    ccfg = cfg._to_c()
    L.pco_b200_profile_enable(1)
    tc, td = [], []
    for it in range(4):
        _lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(dbyte), C.byref(ccfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()),
                                          C.c_size_t(cap), C.byref(nw), C.c_void_p(d_idx.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(7), None))
        sc = spans()
        _lib.check(L.pco_b200_decompress_ex(C.c_void_p(d_comp.data_ptr()), nw, C.c_ubyte(dbyte), C.c_void_p(d_out.data_ptr()), C.c_size_t(n), C.byref(prog),
                                            C.c_void_p(d_idx.data_ptr()), il, C.c_uint32(7), None))
        sd = spans()
        if it:
            tc.append(sum(sc.values()))
            td.append(sum(sd.values()))
    exact = bool(torch.equal(d_out, nums))
    cls = (C.c_uint * 8)()
    L.pco_b200_profile_chunk_classes(cls)
    # oracle: bytes of chunk 0 (the file of one chunk)
    g0 = torch.empty(L.pco_standalone_guarantee_file_size(CH, dbyte), dtype=torch.uint8, device=dev)
    nw0 = C.c_size_t()
    _lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(chunks[0].size), C.c_ubyte(dbyte), C.byref(ccfg), C.c_int(0), C.c_void_p(g0.data_ptr()),
                                      C.c_size_t(g0.numel()), C.byref(nw0), None, C.c_size_t(0), None, C.c_uint32(3), None))
    gpu_bytes = g0[: nw0.value].cpu().numpy().tobytes()
    ref_bytes = pyoracle.simple_compress(chunks[0], ocfg)
    same = gpu_bytes == ref_bytes
    # CPU side: ONE core of the port on a few chunks (native call, no Python in the timed region) - per-core figures compare
    # directly with the reference's published per-core numbers; a thread pool around the port does not scale
    threads = 1
    sample = chunks[: min(4, len(chunks))]
    cpu_c, cpu_d, _ = pyoracle.bench_roundtrip(np.concatenate(sample), len(sample), CH, ocfg, 1)
    t0, t1, t2 = 0.0, cpu_c, cpu_c + cpu_d
    mb = n * dt.itemsize / 1e6
    smb = len(sample) * CH * dt.itemsize / 1e6
    row = dict(name=name, ratio=n * dt.itemsize / nw.value, c_gpu=mb / (np.median(tc) / 1e3), d_gpu=mb / (np.median(td) / 1e3), c_cpu=smb / (t1 - t0), d_cpu=smb / (t2 - t1),
               exact=exact, same=same, cls=list(cls)[1:5], threads=threads)
    print(row, flush=True)
    return row


rows = []
o = pyoracle
rows.append(run("C1 u32 lomax, classic, no delta (1 chunk)", np.uint32, lambda s: datagen.c1_u32_lomax(seed=s),
                ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.no_op()), o.make_config(mode=o.MODE_CLASSIC, delta=o.DELTA_NOOP), n_chunks=1))
rows.append(run("C3 f64 sinusoid, FloatMult(0.01), order 2", np.float64, lambda s: datagen.c3_f64_decimal_sinusoid(seed=s),
                ChunkConfig(mode_spec=ModeSpec.try_float_mult(0.01), delta_spec=DeltaSpec.try_consecutive(2)),
                o.make_config(mode=o.MODE_FLOAT_MULT, float_mult_base=0.01, delta=o.DELTA_CONSECUTIVE, delta_order=2)))
for dtype in (np.uint8, np.uint16, np.int32, np.int64, np.float32, np.float64):
    for order in range(8):
        cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(order) if order else DeltaSpec.no_op(), enable_8_bit=True)
        ocfg = o.make_config(mode=o.MODE_CLASSIC, delta=o.DELTA_CONSECUTIVE if order else o.DELTA_NOOP, delta_order=order, enable_8_bit=True)
        rows.append(run(f"C5 {np.dtype(dtype).name} order {order}", dtype, lambda s, d=dtype: datagen.c5_sweep(d, seed=s), cfg, ocfg, key=np.dtype(dtype).name))

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/config_sweep.md"
with open(out, "w") as f:
    f.write(f"| config ({N_CHUNKS} chunks x 2^18 unless noted) | ratio | GPU compress MB/s | GPU decompress MB/s | CPU port compress MB/s (1 core) | CPU port decompress MB/s (1 core) | CPU cores | decode == input | chunk 0 bytes == oracle | decode classes [gen1, gen2, narrow0, narrow1] |\n")
    f.write("|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r['name']} | {r['ratio']:.2f} | {r['c_gpu']:.0f} | {r['d_gpu']:.0f} | {r['c_cpu']:.0f} | {r['d_cpu']:.0f} | {r['threads']} | {r['exact']} | {r['same']} | {r['cls']} |\n")
print("wrote", out)
