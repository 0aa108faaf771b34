"""e2e (pinned host buffers in, pinned host buffers out) of BASELINE config 2 for every in-place mask of pco_b200_zero_copy, as one
whole-array call pair and streamed in chunk groups by P + P host threads - the same two C-ABI calls bench.py's e2e leg makes.
Usage: python profiles/tools/e2e_in_place_probe.py [chunks] [masks, e.g. 0,1,2,3,7] [variants GxP, e.g. 16x2,32x3]      (prints one line per variant; host wall clock around
synchronous calls; PCOB200_COPY_FIFO / PCOB200_COPY_SLICE_MB select the library's copy policy for the whole process)"""
import ctypes as C
import os
import queue
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, _lib, datagen  # noqa: E402

CHUNK_N = 1 << 18
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = _lib.lib()
n = n_chunks * CHUNK_N
U = n * 8
cfg = ChunkConfig(compression_level=8, mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
h_nums = torch.empty(n, dtype=torch.int64, pin_memory=True)
h_np = h_nums.numpy().view(np.uint64)
from concurrent.futures import ThreadPoolExecutor  # noqa: E402


def _gen(c):
    h_np[c * CHUNK_N:(c + 1) * CHUNK_N] = datagen.c2_u64_cumsum_geometric(CHUNK_N, seed=c)


with ThreadPoolExecutor(8) as ex:
    list(ex.map(_gen, range(n_chunks)))
torch.cuda.set_device(0)
h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)


def layout(G):
    bounds = [n_chunks * g // G for g in range(G + 1)]
    g_n = [(bounds[g + 1] - bounds[g]) * CHUNK_N for g in range(G)]
    g_cap = [L.pco_standalone_guarantee_file_size(g_n[g], 2) for g in range(G)]
    g_icap = [L.pco_b200_index_size_bound(g_n[g], bounds[g + 1] - bounds[g]) for g in range(G)]
    g_coff = np.concatenate([[0], np.cumsum(g_cap)]).astype(np.int64)
    g_ioff = np.concatenate([[0], np.cumsum([(x + 63) // 64 * 64 for x in g_icap])]).astype(np.int64)
    return bounds, g_n, g_cap, g_icap, g_coff, g_ioff


cap_all = int(layout(64)[4][-1]) + 4096
icap_all = int(layout(64)[5][-1]) + 4096
h_comp = torch.empty(cap_all, dtype=torch.uint8, pin_memory=True)
h_index = torch.empty(icap_all, dtype=torch.uint8, pin_memory=True)


class Worker(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.jobs, self.done = queue.Queue(), queue.Queue()
        self.start()

    def run(self):
        torch.cuda.set_device(0)
        while True:
            fn = self.jobs.get()
            if fn is None:
                L.pco_b200_thread_release()
                return
            try:
                fn()
                self.done.put(None)
            except Exception as ex:  # noqa: BLE001
                self.done.put(ex)


MAXP = 3
workers = [Worker() for _ in range(2 * MAXP)]
streams = [torch.cuda.Stream() for _ in range(2 * MAXP)]


def streamed(G, P):
    bounds, g_n, g_cap, g_icap, g_coff, g_ioff = layout(G)
    g_nw = [C.c_size_t() for _ in range(G)]
    g_il = [C.c_size_t() for _ in range(G)]
    hand = [queue.Queue() for _ in range(P)]

    def produce(p):
        def f():
            try:
                sa = C.c_void_p(streams[p].cuda_stream)
                for g in range(p, G, P):
                    _lib.check(L.pco_b200_compress_ex(C.c_void_p(h_nums.data_ptr() + 8 * bounds[g] * CHUNK_N), C.c_size_t(g_n[g]), C.c_ubyte(2), C.byref(cfg), C.c_int(0),
                                                      C.c_void_p(h_comp.data_ptr() + int(g_coff[g])), C.c_size_t(g_cap[g]), C.byref(g_nw[g]),
                                                      C.c_void_p(h_index.data_ptr() + int(g_ioff[g])), C.c_size_t(g_icap[g]), C.byref(g_il[g]), C.c_uint32(0), sa))
                    hand[p].put(g)
            finally:
                hand[p].put(None)
        return f

    def consume(p):
        def f():
            sb = C.c_void_p(streams[MAXP + p].cuda_stream)
            pr = _lib._CProgress()
            while True:
                g = hand[p].get()
                if g is None:
                    return
                _lib.check(L.pco_b200_decompress_ex(C.c_void_p(h_comp.data_ptr() + int(g_coff[g])), g_nw[g], C.c_ubyte(2),
                                                    C.c_void_p(h_out.data_ptr() + 8 * bounds[g] * CHUNK_N), C.c_size_t(g_n[g]), C.byref(pr),
                                                    C.c_void_p(h_index.data_ptr() + int(g_ioff[g])), g_il[g], C.c_uint32(0), sb))
                assert pr.n_processed == g_n[g] and pr.finished
        return f

    def one_pass():
        ws = workers[:P] + workers[MAXP:MAXP + P]
        for p in range(P):
            workers[p].jobs.put(produce(p))
            workers[MAXP + p].jobs.put(consume(p))
        for w in ws:
            e = w.done.get()
            if e is not None:
                raise e
    return one_pass


def single():
    nw, il = C.c_size_t(), C.c_size_t()
    pr = _lib._CProgress()
    cap = L.pco_standalone_guarantee_file_size(n, 2)
    icap = L.pco_b200_index_size_bound(n, n_chunks)
    t0 = time.perf_counter()
    _lib.check(L.pco_b200_compress_ex(C.c_void_p(h_nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(h_comp.data_ptr()),
                                      C.c_size_t(cap), C.byref(nw), C.c_void_p(h_index.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(0), None))
    t1 = time.perf_counter()
    _lib.check(L.pco_b200_decompress_ex(C.c_void_p(h_comp.data_ptr()), nw, C.c_ubyte(2), C.c_void_p(h_out.data_ptr()), C.c_size_t(n), C.byref(pr),
                                        C.c_void_p(h_index.data_ptr()), il, C.c_uint32(0), None))
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return ts


variants = [(16, 2), (32, 2), (16, 3), (32, 3), (8, 1), (16, 1), (64, 2)]
if len(sys.argv) > 3:  # e.g. 16x2,32x3
    variants = [tuple(int(v) for v in t.split('x')) for t in sys.argv[3].split(',')]
MASKS = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2, 3, 7]
print('copy policy: FIFO', os.environ.get('PCOB200_COPY_FIFO', 'default'), 'slice MB', os.environ.get('PCOB200_COPY_SLICE_MB', 'default'), flush=True)
for mask in MASKS:
    L.pco_b200_zero_copy(C.c_int(mask))
    h_out.zero_()
    single()
    ok = torch.equal(h_out, h_nums)
    ss = [single() for _ in range(2)]
    print(f"mask {mask} single-call: compress {min(s[0] for s in ss):.1f} ms, decompress {min(s[1] for s in ss):.1f} ms, pair {min(s[0] + s[1] for s in ss):.1f} ms"
          f" = {U / 1e6 / min(s[0] + s[1] for s in ss):.1f} GB/s, exact {ok}", flush=True)
    for G, P in variants:
        fn = streamed(G, P)
        h_out.zero_()
        fn()
        ok = torch.equal(h_out, h_nums)
        fn()
        print(f'=== timed passes G={G} P={P} mask {mask}', file=sys.stderr, flush=True)
        ts = timed(fn)
        print(f"mask {mask} streamed G={G} P={P}: {min(ts):.1f} ms (all {[round(t, 1) for t in ts]}) = {U / 1e6 / min(ts):.1f} GB/s, exact {ok}", flush=True)
L.pco_b200_zero_copy(C.c_int(0))
for w in workers:
    w.jobs.put(None)
