import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch
from pcodec_b200 import _lib, datagen, ChunkConfig, ModeSpec, DeltaSpec
L=_lib.lib()
n_chunks=1024; CH=1<<18; n=n_chunks*CH
dev=torch.device('cuda')
nums=datagen.c2_u64_torch(n_chunks, CH, seed=1000, device=dev)
cfg=ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1))._to_c()
cap=L.pco_standalone_guarantee_file_size(n,2); icap=L.pco_b200_index_size_bound(n,n_chunks)
d_comp=torch.empty(cap,dtype=torch.uint8,device=dev); d_idx=torch.empty(icap,dtype=torch.uint8,device=dev); d_out=torch.empty(n,dtype=torch.int64,device=dev)
nw,il=C.c_size_t(),C.c_size_t(); prog=_lib._CProgress()
_lib.check(L.pco_b200_compress_ex(C.c_void_p(nums.data_ptr()),C.c_size_t(n),C.c_ubyte(2),C.byref(cfg),C.c_int(0),C.c_void_p(d_comp.data_ptr()),C.c_size_t(cap),C.byref(nw),C.c_void_p(d_idx.data_ptr()),C.c_size_t(icap),C.byref(il),C.c_uint32(7),None))
def dec():
    _lib.check(L.pco_b200_decompress_ex(C.c_void_p(d_comp.data_ptr()),nw,C.c_ubyte(2),C.c_void_p(d_out.data_ptr()),C.c_size_t(n),C.byref(prog),C.c_void_p(d_idx.data_ptr()),il,C.c_uint32(7),None))
for _ in range(3): dec()
buf=(C.c_ulonglong*16)()
L.pco_b200_debug_dec_timing(buf)
dec()
L.pco_b200_debug_dec_timing(buf)
names=['prologue','phaseA work','barrier after A','B: window staging','B: syms/bins/scan/extract','(unused)','B: scans+chain wait+link+fold','B: join+store','loop tail','barrier after B']
tot=sum(buf[:10])
for i,nm in enumerate(names): print(f"{nm:32s} {buf[i]/1e6:10.1f} Mcycles  {100*buf[i]/tot:5.1f}%")
print("total warp-cycles", tot/1e6, "M; per warp", tot/ (1024*8)/1e3, "kcycles")
assert torch.equal(d_out, nums)
