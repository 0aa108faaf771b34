"""Experiment tool: per-kernel spans of one compress call on wide-range data (C5 int64 order 0: the sort path of the planner)."""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L = _lib.lib()
n_chunks = int(os.environ.get("N_CHUNKS", "128")); CH = 1 << 18; n = n_chunks * CH
dt = np.dtype(os.environ.get("DTYPE", "int64"))
host = np.concatenate([datagen.c5_sweep(dt, seed=s) for s in range(n_chunks)])
dev = torch.device("cuda")
nums = torch.from_numpy(host.view(np.dtype(f"i{dt.itemsize}"))).to(dev)
dbyte = _lib.dtype_byte(dt)
cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.no_op())._to_c()
cap = L.pco_standalone_guarantee_file_size(n, dbyte); d_comp = torch.empty(cap, dtype=torch.uint8, device=dev)
nw = C.c_size_t(); buf = C.create_string_buffer(4096)
L.pco_b200_profile_enable(1)
for it in range(3):
    _lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()), C.c_size_t(n), C.c_ubyte(dbyte), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()), C.c_size_t(cap), C.byref(nw), None, C.c_size_t(0), None, C.c_uint32(3), None))
    L.pco_b200_profile_last(buf, 4096)
print(dt.name, n_chunks, "chunks:", buf.value.decode())
