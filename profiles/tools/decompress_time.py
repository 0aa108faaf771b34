"""Experiment tool: decompress kernel times (profile spans) for a library variant; no output check (for builds that skip work)."""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L = _lib.lib()
n_chunks = int(os.environ.get("N_CHUNKS", "1024")); CH = 1 << 18; n = n_chunks * CH
dev = torch.device('cuda')
nums = datagen.c2_u64_torch(n_chunks, CH, seed=1000, device=dev)
cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(int(os.environ.get("ORDER", "1"))))._to_c()
cap = L.pco_standalone_guarantee_file_size(n, 2); icap = L.pco_b200_index_size_bound(n, n_chunks)
d_comp = torch.empty(cap, dtype=torch.uint8, device=dev); d_idx = torch.empty(icap, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.int64, device=dev)
nw, il = C.c_size_t(), C.c_size_t(); prog = _lib._CProgress()
_lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()), C.c_size_t(cap), C.byref(nw), C.c_void_p(d_idx.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(7), None))
L.pco_b200_profile_enable(1)
buf = C.create_string_buffer(4096)
call_ms = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.pco_b200_decompress_ex(C.c_void_p(d_comp.data_ptr()), nw, C.c_ubyte(2), C.c_void_p(d_out.data_ptr()), C.c_size_t(n), C.byref(prog), C.c_void_p(d_idx.data_ptr()), il, C.c_uint32(7), None))
    e1.record(); torch.cuda.synchronize()
    call_ms.append(e0.elapsed_time(e1))
    L.pco_b200_profile_last(buf, 4096)
cls = (C.c_uint * 8)()
if hasattr(L, "pco_b200_profile_chunk_classes"): L.pco_b200_profile_chunk_classes(cls)
print(os.path.basename(os.environ.get("PCOB200_LIB", "default")), "compressed", nw.value, buf.value.decode(), "classes", list(cls)[1:5], "call_ms %.4f" % (sum(call_ms[2:]) / len(call_ms[2:])),
      "exact" if torch.equal(d_out, nums) else "OUTPUT DIFFERS")
