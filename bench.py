#!/usr/bin/env python
"""bench.py — the hot-path benchmark (BASELINE.json metric: compress+decompress MB/s, % HBM roofline).

A "step" is one pass of the hot path over one batch of synthetic input: compress, then decompress, 1024 independent
chunks of 2^18 u64 (BASELINE config 2: classic mode, consecutive delta order 1; C2(i) data).  Per rank, N ranks = weak
scaling (8 ranks = config 4's 8192 chunks).  MB = 10^6 uncompressed bytes (pco_cli/src/bench/mod.rs:233-241).

  value     resident: inputs already in HBM, device buffers in and out, through the C-ABI (*_ex, flags DEVICE).
  e2e       same calls with pinned HOST buffers (H2D of the inputs and D2H of the results inside the timed region).
  roofline  decompress kernels (symwalk_kernel + the decode kernels, decode_narrow_kernel on this data): (U + C + side index) bytes / their CUDA-event durations vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: oracle/ (C++ restatement of pco 1.0.3; the Rust reference cannot be built here)
            on the host cores - one worker process per host thread - on a bounded sample of the same chunks.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHUNK_N = 1 << 18
N_CHUNKS = 1024
METRIC = "compress+decompress MB/s per chunk (u64, 2^18 elems)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chunks", type=int, default=N_CHUNKS, help="chunks per rank (default: BASELINE config 2)")
    ap.add_argument("--cpu-sample-chunks", type=int, default=1024, help="chunks of the CPU arm per step, spread over one worker process per host thread (about 12 s of core time)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-index-free", action="store_true", help="skip the pco_b200_decompress_chunks timing (not part of `value`)")
    ap.add_argument("--results-csv", default=None, help="also merge this run into a CSV with the reference bench tool's schema and codec naming (pcodec_b200/benchfmt.py)")
    ap.add_argument("--gather-pages", action="store_true", help="N > 1: also all-gather the compressed page bytes (every rank ends up with the whole file)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (the one place bench.py may execute oracle/)
# ------------------------------------------------------------------------------------------------
# One worker PROCESS per host thread, each compressing and decompressing its own chunks with the oracle
# (oracle/c_api.cpp pco_oracle_bench_roundtrip, one native thread).  Processes, not threads: the port allocates MiB-sized
# scratch per chunk and threads of one process serialise in the kernel's address-space lock (8 threads: 1.6x one thread for
# compress), and a Python thread pool around per-chunk ctypes calls adds GIL-held buffer copies on top.  Measured on the
# authoring box: 8 processes = 8x one.
_W = {}
CPU_RUN_TIMEOUT_S = 120  # one timed pass of the pool; far above the ~1 s it takes


def host_threads():
    """Host threads this process may really use: the affinity mask, capped by a cgroup CPU quota when one is set."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, n)


def _cpu_worker_init(counter, barrier, chunks_per_worker):
    from oracle import pyoracle
    from pcodec_b200 import datagen

    with counter.get_lock():
        wid = counter.value
        counter.value += 1
    cfg = pyoracle.make_config(level=8, mode=pyoracle.MODE_CLASSIC, delta=pyoracle.DELTA_CONSECUTIVE, delta_order=1)
    nums = np.concatenate([datagen.c2_u64_cumsum_geometric(CHUNK_N, seed=wid * chunks_per_worker + i) for i in range(chunks_per_worker)])
    pyoracle.bench_roundtrip(nums[:CHUNK_N], 1, CHUNK_N, cfg, 1)  # load the library, touch the allocator
    _W.update(nums=nums, cfg=cfg, k=chunks_per_worker, barrier=barrier)


def _cpu_worker_run(_):
    from oracle import pyoracle

    _W["barrier"].wait(timeout=180)  # every worker takes exactly one task and they start together (a lost worker breaks the barrier, it does not hang)
    tc, td, cbytes = pyoracle.bench_roundtrip(_W["nums"], _W["k"], CHUNK_N, _W["cfg"], 1)  # raises if a chunk does not round-trip
    return tc, td, cbytes


class CpuArm:
    """`workers` processes with `chunks_per_worker` C2(i) chunks each; run() = one timed compress + decompress of all of them."""

    def __init__(self, workers, chunks_per_worker):
        import multiprocessing as mp

        from oracle import pyoracle

        pyoracle.build()  # once, here: the workers must find the library up to date instead of racing to rebuild it
        ctx = mp.get_context("spawn")  # the GPU arm's process holds a CUDA context and helper threads: no fork
        self.workers, self.k = workers, chunks_per_worker
        self.pool = ctx.Pool(workers, initializer=_cpu_worker_init, initargs=(ctx.Value("i", 0), ctx.Barrier(workers), chunks_per_worker))

    def run(self):
        """Returns (MB/s round trip, compress MB/s, decompress MB/s, compressed bytes): bytes of all workers over the slowest
        worker's time (they run concurrently from a common barrier)."""
        # map_async + timeout: a pool whose workers die at start-up is respawned forever and a plain map() would never return
        try:
            res = self.pool.map_async(_cpu_worker_run, range(self.workers), chunksize=1).get(timeout=CPU_RUN_TIMEOUT_S + 2 * self.k)
        except Exception:
            self.broken = True
            raise
        mb = self.workers * self.k * CHUNK_N * 8 / 1e6
        tc, td = max(r[0] for r in res), max(r[1] for r in res)
        return mb / max(r[0] + r[1] for r in res), mb / tc, mb / td, sum(r[2] for r in res)

    def close(self):
        if getattr(self, "broken", False):
            self.pool.terminate()
        else:
            self.pool.close()
        self.pool.join()


def cpu_roundtrip(n_chunks, threads, repeats=1):
    """Compress + decompress about `n_chunks` C2(i) chunks on `threads` host threads (one worker process each, at least one
    chunk per worker).  Returns (MB/s round trip, compress MB/s, decompress MB/s, compressed bytes, chunks actually run)."""
    workers = max(1, threads)
    k = max(1, n_chunks // workers)
    arm = CpuArm(workers, k)
    try:
        best = None
        for _ in range(repeats):
            cur = arm.run()
            if best is None or cur[0] > best[0]:
                best = cur
    finally:
        arm.close()
    return best + (workers * k,)


def run_reference_arm(args, rank):
    """--impl reference: the reference's own CPU path of the same workload.  pco is Rust and there is no cargo/rustc in the
    image, so this runs the oracle port and says so (cpu_baseline.kind = "port")."""
    if rank != 0:
        return
    vals, err = None, None
    for threads in sorted({host_threads(), min(host_threads(), 8), 1}, reverse=True):
        # all host threads; if a pool of that size cannot be brought up on this host, a smaller one still gives a line (and says so)
        k = max(1, args.cpu_sample_chunks // threads)
        sample = threads * k
        try:
            arm = CpuArm(threads, k)
            try:
                for _ in range(args.warmup):
                    arm.run()
                vals, t0 = [], time.perf_counter()
                for _ in range(args.steps):
                    vals.append(arm.run())
                ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
            finally:
                arm.close()
            break
        except Exception as ex:  # noqa: BLE001
            vals, err = None, f"{threads} workers: {type(ex).__name__}: {ex}"
    if vals is None:
        print(json.dumps({"impl": "reference", "unavailable": f"CPU port could not be run: {err}"}))
        return
    v = float(np.median([x[0] for x in vals]))
    try:
        one = cpu_roundtrip(4, 1)  # one thread and all host cores in the same run
    except Exception:  # noqa: BLE001
        one = (None, None, None, 0, 0)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "C2: 2^18-element u64 chunks, classic, consecutive delta order 1, level 8 (cumsum of geometric(0.001))",
                   "chunks_per_step": sample, "chunk_n": CHUNK_N},
        "cpu_baseline": {"value": v, "unit": "MB/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} chunks of 2^18 u64 per step, {k} per worker process, one process per host thread; C++ restatement of pco 1.0.3 (oracle/), not the Rust crate",
                         "compress_mb_s": float(np.median([x[1] for x in vals])), "decompress_mb_s": float(np.median([x[2] for x in vals])),
                         "single_core": {"value": one[0], "compress_mb_s": one[1], "decompress_mb_s": one[2], "sample": f"{one[4]} chunks, one process"},
                         **({"note": f"fewer workers than host threads ({host_threads()}) after: {err}"} if err else {})},
        "e2e": {"value": v, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm, smax, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = max(smax, float(s[1]))
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": len(sm)}


def parse_profile(L):
    buf = C.create_string_buffer(4096)
    L.pco_b200_profile_last(buf, C.c_size_t(4096))
    out = {}
    for item in buf.value.decode().split(";"):
        if "=" in item:
            k, v = item.split("=")
            out[k] = out.get(k, 0.0) + float(v)
    return out


def run_gpu_arm(args, rank, world):
    import torch

    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, _lib, datagen

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION prints it to stdout) off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if not L.pco_b200_device_available():
        raise RuntimeError("libcpcodec.so found no usable CUDA device: " + L.pco_b200_last_error_message().decode())

    n_chunks = args.chunks
    n = n_chunks * CHUNK_N
    U = n * 8
    cfg = ChunkConfig(compression_level=8, mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
    nums = datagen.c2_u64_torch(n_chunks, CHUNK_N, seed=1000 + rank, device=dev)
    cap = L.pco_standalone_guarantee_file_size(n, 2)
    icap = L.pco_b200_index_size_bound(n, n_chunks)
    d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_index = torch.empty(icap, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    SRC, DST, IDX = 1, 2, 4
    n_written, ilen = C.c_size_t(), C.c_size_t()
    prog = _lib._CProgress()

    def compress_resident():
        rc = L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()),
                                    C.c_size_t(cap), C.byref(n_written), C.c_void_p(d_index.data_ptr()), C.c_size_t(icap), C.byref(ilen),
                                    C.c_uint32(SRC | DST | IDX), sp)
        _lib.check(rc)

    def decompress_resident():
        rc = L.pco_b200_decompress_ex(C.c_void_p(d_comp.data_ptr()), n_written, C.c_ubyte(2), C.c_void_p(d_out.data_ptr()), C.c_size_t(n), C.byref(prog),
                                      C.c_void_p(d_index.data_ptr()), ilen, C.c_uint32(SRC | DST | IDX), sp)
        _lib.check(rc)
        assert prog.n_processed == n and prog.finished

    gathered = None
    file_offset = [0]

    def gather_pages():
        # The exchange step of the sharded path: chunks are independent, so a rank only needs to know WHERE its
        # pages go in the logical standalone file - one NCCL all-gather of the per-rank compressed byte counts; the exclusive
        # prefix is this rank's byte offset (a sharded writer pwrite()s there; decompress stays sharded and needs nothing).
        # --gather-pages additionally all-gathers the page bytes themselves so that every rank holds the whole file.
        nonlocal gathered
        if world == 1:
            return
        import torch.distributed as dist

        sizes = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([n_written.value], dtype=torch.int64, device=dev))
        sizes_h = sizes.cpu()
        file_offset[0] = int(sizes_h[:rank].sum().item())
        if args.gather_pages:
            mx = (int(sizes_h.max().item()) + 255) // 256 * 256
            if gathered is None or gathered.numel() < world * mx:
                gathered = torch.empty(world * mx, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(gathered[: world * mx], d_comp[:mx])

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also verifies the bit-exact round trip)
    L.pco_b200_profile_enable(1)
    for w in range(max(args.warmup, 3)):
        compress_resident()
        gather_pages()
        decompress_resident()
    torch.cuda.synchronize()
    assert torch.equal(d_out, nums), "GPU round trip is not bit-exact"
    Cbytes, Ibytes = n_written.value, ilen.value

    # ---- timed region: resident
    sampler = ClockSampler(local_rank)
    if rank == 0:  # only rank 0's samples go into the line; N nvidia-smi pollers would only add host noise
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_c, t_d, t_g, prof_c, prof_d = [], [], [], [], []
    barrier()
    ev_all0, ev_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_all0.record(stream)
    for _ in range(args.steps):
        ev[0].record(stream)
        compress_resident()
        ev[1].record(stream)
        prof_c.append(parse_profile(L))
        gather_pages()
        ev[2].record(stream)
        decompress_resident()
        ev[3].record(stream)
        prof_d.append(parse_profile(L))
        torch.cuda.synchronize()
        t_c.append(ev[0].elapsed_time(ev[1]))
        t_g.append(ev[1].elapsed_time(ev[2]))
        t_d.append(ev[2].elapsed_time(ev[3]))
    ev_all1.record(stream)
    barrier()
    total_ms = ev_all0.elapsed_time(ev_all1)
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / max(args.steps, 1)
    value = world * U / 1e6 / (ms_per_step / 1e3)

    # ---- the index-free path (not part of `value`): the same chunks decoded from their byte offsets alone
    # (pco_b200_decompress_chunks: one tANS walk per chunk builds the per-batch index on the device, all chunks in parallel)
    chunks_free = None
    if hasattr(L, "pco_b200_decompress_chunks") and not args.no_index_free:
        import struct

        ih = bytes(d_index[:64].cpu().numpy())
        n_idx_chunks, chunks_off = struct.unpack_from("<Q", ih, 8)[0], struct.unpack_from("<Q", ih, 32)[0]
        recs = bytes(d_index[chunks_off:chunks_off + 32 * n_idx_chunks].cpu().numpy())
        offs = np.array([struct.unpack_from("<Q", recs, 32 * i)[0] for i in range(n_idx_chunks)], dtype=np.uint64)
        cns = np.array([struct.unpack_from("<I", recs, 32 * i + 8)[0] for i in range(n_idx_chunks)], dtype=np.uint32)
        nw_free = C.c_size_t()

        def decompress_chunks():
            rc = L.pco_b200_decompress_chunks(C.c_void_p(d_comp.data_ptr()), n_written, C.c_ubyte(2), offs.ctypes.data_as(C.c_void_p), cns.ctypes.data_as(C.c_void_p),
                                              C.c_size_t(n_idx_chunks), C.c_void_p(d_out.data_ptr()), C.c_size_t(n), C.byref(nw_free), C.c_uint32(SRC | DST), sp)
            _lib.check(rc)

        d_out.zero_()
        decompress_chunks()
        torch.cuda.synchronize()
        assert nw_free.value == n and torch.equal(d_out, nums), "index-free decompress is not bit-exact"
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(3):
            decompress_chunks()
        f1.record(stream)
        torch.cuda.synchronize()
        free_ms = f0.elapsed_time(f1) / 3
        chunks_free = {"decompress_mb_s": U / 1e6 / (free_ms / 1e3), "ms": free_ms, "kernel_ms": parse_profile(L),
                       "api": "pco_b200_decompress_chunks (chunk byte offsets only, no side index), buffers resident in HBM"}

    # ---- e2e: the same calls with pinned host buffers (H2D / D2H inside the timed region)
    e2e = None
    e2e_ok = 0 if args.no_e2e else 1
    if e2e_ok:
        # the pinned staging buffers (2 x U + compressed + index per rank) can fail on a crowded host: every rank then
        # skips the e2e leg together instead of losing the whole line (the collectives below need all ranks)
        try:
            h_nums = torch.empty(n, dtype=torch.int64, pin_memory=True)
            h_nums.copy_(nums)
            h_comp = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            h_index = torch.empty(icap, dtype=torch.uint8, pin_memory=True)
            h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
        except Exception as ex:  # noqa: BLE001
            e2e_ok = 0
            e2e = {"value": None, "unit": "MB/s", "error": f"pinned host buffers: {ex}"}
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([e2e_ok], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if int(t.item()) == 0 and e2e_ok:
                e2e = {"value": None, "unit": "MB/s", "error": "pinned host buffers failed on another rank"}
            e2e_ok = int(t.item())
    if e2e_ok:
        nw2, il2 = C.c_size_t(), C.c_size_t()

        def e2e_step():
            rc = L.pco_b200_compress_ex(C.c_void_p(h_nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(h_comp.data_ptr()),
                                        C.c_size_t(cap), C.byref(nw2), C.c_void_p(h_index.data_ptr()), C.c_size_t(icap), C.byref(il2), C.c_uint32(0), sp)
            _lib.check(rc)
            rc = L.pco_b200_decompress_ex(C.c_void_p(h_comp.data_ptr()), nw2, C.c_ubyte(2), C.c_void_p(h_out.data_ptr()), C.c_size_t(n), C.byref(prog),
                                          C.c_void_p(h_index.data_ptr()), il2, C.c_uint32(0), sp)
            _lib.check(rc)

        e2e_step()
        assert torch.equal(h_out, h_nums), "e2e round trip is not bit-exact"
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2e_steps = max(1, min(args.steps, 3))
        e0.record(stream)
        for _ in range(e2e_steps):
            e2e_step()
        e1.record(stream)
        barrier()
        e2e_ms = e0.elapsed_time(e1) / e2e_steps
        if world > 1:
            import torch.distributed as dist

            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": world * U / 1e6 / (e2e_ms / 1e3), "unit": "MB/s", "h2d_bytes_per_step": int(U + nw2.value + il2.value),
               "d2h_bytes_per_step": int(nw2.value + il2.value + U), "ms_per_step": e2e_ms,
               "api": "pco_b200_compress_ex + pco_b200_decompress_ex (C-ABI), pinned host buffers"}
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)

    if rank != 0:
        return
    # ---- roofline of the decompress kernels: algorithmic bytes / CUDA-event duration
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    # the decompress path is two kernels: symwalk_kernel (tANS symbol walk) + decode_kernel (offsets, un-delta, join)
    dk_ms = float(np.mean([p.get("fused_narrow_kernel", 0.0) + p.get("decode_kernel", 0.0) + p.get("symwalk_kernel", 0.0) for p in prof_d]))
    dec_spans = {}
    for p in prof_d:
        for k, v in p.items():
            dec_spans.setdefault(k, []).append(v)
    dec_spans = {k: float(np.mean(v)) for k, v in dec_spans.items()}
    alg_bytes = U + Cbytes + Ibytes
    achieved = alg_bytes / 1e9 / (dk_ms / 1e3) if dk_ms > 0 else None
    # DRAM traffic of the same two kernels from the committed `ncu --set full` capture (dram__bytes_read.sum +
    # dram__bytes_write.sum per launch, same workload); never measured under this run
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath) and n_chunks == N_CHUNKS:
        tj = json.load(open(tpath))
        traffic, traffic_src = int(sum(tj["dram_bytes_per_launch"].values())), tj["source"]
    comp_spans = {}
    for p in prof_c:
        for k, v in p.items():
            comp_spans.setdefault(k, []).append(v)
    comp_spans = {k: float(np.mean(v)) for k, v in comp_spans.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # the CPU arm beside the GPU number: rank 0 at N = 1 only
        threads = host_threads()
        try:  # a CPU-side failure must not cost the GPU line
            v = cpu_roundtrip(max(threads, args.cpu_sample_chunks), threads)
            one = cpu_roundtrip(4, 1)  # one thread and all host cores in the same run
            cpu = {"value": v[0], "unit": "MB/s", "cores": threads, "kind": "port",
                   "sample": f"{v[4]} chunks of 2^18 u64 (same generator), one worker process per host thread; C++ restatement of pco 1.0.3 (oracle/), not the Rust crate",
                   "compress_mb_s": v[1], "decompress_mb_s": v[2],
                   "single_core": {"value": one[0], "compress_mb_s": one[1], "decompress_mb_s": one[2], "sample": f"{one[4]} chunks, one process"}}
        except Exception as ex:  # noqa: BLE001
            cpu = {"value": None, "unit": "MB/s", "cores": threads, "kind": "port", "sample": "failed", "error": f"{type(ex).__name__}: {ex}"}

    line = {
        "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"C2: {n_chunks} chunks x 2^18 u64 per GPU, classic mode, consecutive delta order 1, level 8 (cumsum of geometric(0.001)); "
                               "step = compress + decompress of every chunk, buffers resident in HBM",
                   "chunks_per_gpu": n_chunks, "chunk_n": CHUNK_N, "l2": "inputs (2 GiB per GPU) exceed the 126 MB L2",
                   "multi_gpu": ("independent chunk shards per rank; one NCCL all-gather per step of the per-rank compressed sizes (file offsets of the shards)" + ("; plus an all-gather of the page bytes" if args.gather_pages else "; pages stay sharded")) if world > 1 else "single GPU",
                   "side_index": "decompress uses the per-batch side index emitted by the compressor (bytes counted in the roofline)"},
        "compress_mb_s": world * U / 1e6 / (float(np.mean(t_c)) / 1e3), "decompress_mb_s": world * U / 1e6 / (float(np.mean(t_d)) / 1e3),
        "gather_ms": float(np.mean(t_g)), "compressed_bytes_per_gpu": Cbytes, "index_bytes_per_gpu": Ibytes, "ratio": U / Cbytes,
        "kernel_ms": {**dec_spans, **comp_spans},
        "roofline": {"bound": "hbm", "kernel": "decompress path: symwalk_kernel + decode_kernel (sum of both durations)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                     "peak_source": peak_src},
        "index_free_decompress": chunks_free,
        "cpu_baseline": cpu, "e2e": e2e, "clocks": sampler.summary(),
        # per step (profiles/r01_l_launches.csv): compress = init_chunks, split_count, plan_solve, fallback, bin_lut, ans_encode,
        # layout, chunk_offsets, pack, header_footer, emit_index; decompress = symwalk_kernel + decode_narrow_kernel +
        # decode_kernel<L,1> + decode_kernel<L,2> (a chunk is decoded by exactly one of the three; the others' CTAs exit at once)
        "gpu_launches": 15,
    }
    print(json.dumps(line))
    if args.results_csv:
        # the same run in the reference bench tool's format (docs/benchmark_results/*.csv): seconds per pass over the whole job
        from pcodec_b200 import benchfmt

        benchfmt.merge_results_csv(args.results_csv, [dict(
            input=f"c2_u64_cumsum_geometric_{world * n_chunks}x2^18", codec=benchfmt.PcoCodec(level=8, delta=DeltaSpec.try_consecutive(1), mode=ModeSpec.classic()),
            compress_dt=float(np.mean(t_c)) / 1e3, decompress_dt=float(np.mean(t_d)) / 1e3, compressed_size=world * Cbytes, uncompressed_size=world * U)])


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    run_gpu_arm(args, rank, world)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
