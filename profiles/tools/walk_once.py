"""Experiment tool: the index-free path (pco_b200_decompress_chunks: walk_kernel + decode) on N_CHUNKS chunks of C2 data; prints kernel spans."""
import ctypes as C, os, struct, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L = _lib.lib()
n_chunks = int(os.environ.get("N_CHUNKS", "1024")); CH = 1 << 18; n = n_chunks * CH
dev = torch.device("cuda")
nums = datagen.c2_u64_torch(n_chunks, CH, seed=1000, device=dev)
cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
cap = L.pco_standalone_guarantee_file_size(n, 2); icap = L.pco_b200_index_size_bound(n, n_chunks)
d_comp = torch.empty(cap, dtype=torch.uint8, device=dev); d_idx = torch.empty(icap, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.int64, device=dev)
nw, il = C.c_size_t(), C.c_size_t()
_lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()), C.c_size_t(cap), C.byref(nw), C.c_void_p(d_idx.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(7), None))
ih = bytes(d_idx[:64].cpu().numpy()); nch, coff = struct.unpack_from("<Q", ih, 8)[0], struct.unpack_from("<Q", ih, 32)[0]
recs = bytes(d_idx[coff:coff + 32 * nch].cpu().numpy())
offs = np.array([struct.unpack_from("<Q", recs, 32 * i)[0] for i in range(nch)], dtype=np.uint64)
cns = np.array([struct.unpack_from("<I", recs, 32 * i + 8)[0] for i in range(nch)], dtype=np.uint32)
L.pco_b200_profile_enable(1)
buf = C.create_string_buffer(4096); nwf = C.c_size_t()
for it in range(3):
    d_out.zero_()
    _lib.check(L.pco_b200_decompress_chunks(C.c_void_p(d_comp.data_ptr()), nw, C.c_ubyte(2), offs.ctypes.data_as(C.c_void_p), cns.ctypes.data_as(C.c_void_p), C.c_size_t(nch),
                                            C.c_void_p(d_out.data_ptr()), C.c_size_t(n), C.byref(nwf), C.c_uint32(3), None))
    L.pco_b200_profile_last(buf, 4096)
print(n_chunks, "chunks", buf.value.decode(), "exact" if torch.equal(d_out, nums) else "OUTPUT DIFFERS")
