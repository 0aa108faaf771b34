"""Experiment tool: per-phase clock64 sums inside plan_probe_kernel / plan_solve_kernel (build variant -DPCOB_ENC_TIMING).
Run on the GPU box with PCOB200_LIB=pcodec_b200/libcpcodec_enctiming.so."""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L = _lib.lib()
n_chunks = 1024; CH = 1 << 18; n = n_chunks * CH
dev = torch.device('cuda')
nums = datagen.c2_u64_torch(n_chunks, CH, seed=1000, device=dev)
cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
cap = L.pco_standalone_guarantee_file_size(n, 2); icap = L.pco_b200_index_size_bound(n, n_chunks)
d_comp = torch.empty(cap, dtype=torch.uint8, device=dev); d_idx = torch.empty(icap, dtype=torch.uint8, device=dev)
nw, il = C.c_size_t(), C.c_size_t()
def comp():
    _lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()), C.c_size_t(cap), C.byref(nw), C.c_void_p(d_idx.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(7), None))
for _ in range(2): comp()
buf = (C.c_ulonglong * 32)()
L.pco_b200_debug_enc_timing(buf)
comp()
L.pco_b200_debug_enc_timing(buf)
names = ['probe: zero+count', 'probe: scan', 'probe: probes', 'solve: stage + histogram state machine', 'solve: DP', 'solve: rewind+quantize', 'solve: tables']
tot = sum(buf[:7])
for i, nm in enumerate(names):
    print(f"{nm:32s} {buf[i] / n_chunks / 1e3:10.1f} kcycles/CTA  {100 * buf[i] / tot:5.1f}%")
print("total per CTA", tot / n_chunks / 1e3, "kcycles")
