"""Why does compress(g+1) beside decompress(g) not use both PCIe directions?  Two host threads, 8 groups of 128 chunks (C2 data, pinned
host buffers); thread A / thread B run either the library call or a plain torch copy of the same bytes.  Prints ms per pass."""
import ctypes as C, os, sys, threading, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L = _lib.lib()
G, per = 8, 128; CH = 1 << 18; gn = per * CH; n = G * gn
dev = torch.device("cuda")
nums = datagen.c2_u64_torch(G * per, CH, seed=7, device=dev)
h_nums = torch.empty(n, dtype=torch.int64, pin_memory=True); h_nums.copy_(nums)
h_out = torch.empty(n, dtype=torch.int64, pin_memory=True)
cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
cap = L.pco_standalone_guarantee_file_size(gn, 2); icap = L.pco_b200_index_size_bound(gn, per)
h_comp = torch.empty(G * cap, dtype=torch.uint8, pin_memory=True); h_idx = torch.empty(G * icap, dtype=torch.uint8, pin_memory=True)
nw = [C.c_size_t() for _ in range(G)]; il = [C.c_size_t() for _ in range(G)]
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
d_stage = torch.empty(gn, dtype=torch.int64, device=dev); d_src = torch.empty(gn, dtype=torch.int64, device=dev)
def lib_compress(g):
    _lib.check(L.pco_b200_compress_ex(C.c_void_p(h_nums.data_ptr() + 8 * g * gn), C.c_size_t(gn), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(h_comp.data_ptr() + g * cap), C.c_size_t(cap),
                                      C.byref(nw[g]), C.c_void_p(h_idx.data_ptr() + g * icap), C.c_size_t(icap), C.byref(il[g]), C.c_uint32(0), C.c_void_p(sA.cuda_stream)))
def lib_decompress(g):
    pr = _lib._CProgress()
    _lib.check(L.pco_b200_decompress_ex(C.c_void_p(h_comp.data_ptr() + g * cap), nw[g], C.c_ubyte(2), C.c_void_p(h_out.data_ptr() + 8 * g * gn), C.c_size_t(gn), C.byref(pr),
                                        C.c_void_p(h_idx.data_ptr() + g * icap), il[g], C.c_uint32(0), C.c_void_p(sB.cuda_stream)))
def torch_h2d(g):
    with torch.cuda.stream(sA):
        d_stage.copy_(h_nums[g * gn:(g + 1) * gn], non_blocking=True); sA.synchronize()
def torch_d2h(g):
    with torch.cuda.stream(sB):
        h_out[g * gn:(g + 1) * gn].copy_(d_src, non_blocking=True); sB.synchronize()
import queue
class Worker(threading.Thread):
    """A persistent host thread (the library keeps its scratch per calling thread: a thread per pass would re-allocate it every time)."""
    def __init__(self):
        super().__init__(daemon=True); self.jobs, self.done = queue.Queue(), queue.Queue(); self.start()
    def run(self):
        while True:
            f = self.jobs.get()
            if f is None: return
            for g in range(G): f(g)
            self.done.put(None)
wa, wb = Worker(), Worker()
def run(fa, fb):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if fa: wa.jobs.put(fa)
    if fb: wb.jobs.put(fb)
    if fa: wa.done.get()
    if fb: wb.done.get()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
run(lib_compress, None); run(None, lib_decompress)
for name, fa, fb in [("lib compress alone", lib_compress, None), ("lib decompress alone", None, lib_decompress), ("torch H2D alone", torch_h2d, None), ("torch D2H alone", None, torch_d2h),
                     ("torch H2D || torch D2H", torch_h2d, torch_d2h), ("lib compress || torch D2H", lib_compress, torch_d2h), ("torch H2D || lib decompress", torch_h2d, lib_decompress),
                     ("lib compress || lib decompress", lib_compress, lib_decompress)]:
    run(fa, fb); print("%-34s %7.1f ms" % (name, run(fa, fb)), flush=True)
