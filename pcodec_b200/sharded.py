"""Multi-GPU sharding of independent chunks (BASELINE config 4).

Chunks of a standalone file are fully independent (own ChunkMeta, own page: pco/src/standalone/simple.rs:36-44,73-87),
so chunk c goes to rank c mod G, each rank compresses its shard with the single-GPU path, and ONE all-gather of the
compressed pages (sizes first, then the padded bytes) lets every rank assemble the same standalone file
`header | chunk_0 | chunk_1 | ... | 0x00`, byte-identical to a single-GPU / reference compress of the whole array.
There is no collective on the per-chunk critical path.  Works with any torch.distributed backend (NCCL on GPUs;
gloo on CPU tensors for the host-logic tests).
"""
import numpy as np


def chunk_sizes(n, max_page_n=1 << 18):
    """PagingSpec::EqualPagesUpTo(max_page_n).n_per_page(n) (pco/src/chunk_config.rs:134-183)."""
    if n == 0:
        return []
    n_pages = -(-n // max_page_n)
    low, r = divmod(n, n_pages)
    return [low + 1] * r + [low] * (n_pages - r)


def shard_plan(n, world, max_page_n=1 << 18):
    """Round-robin chunk -> rank map.  Returns per rank the list of (chunk_id, start, end) element ranges."""
    sizes = chunk_sizes(n, max_page_n)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    plan = [[] for _ in range(world)]
    for c, sz in enumerate(sizes):
        plan[c % world].append((c, int(starts[c]), int(starts[c] + sz)))
    return plan


def standalone_header(n_hint, uniform_type=0):
    """Standalone header bytes (pco/src/standalone/compressor.rs:12-16,85-105): magic, version 3, uniform type, varint(n), format 4.1."""
    power = 1 if n_hint == 0 else int(n_hint).bit_length()
    v = (power - 1) | ((n_hint & ((1 << power) - 1)) << 6)
    nbytes = (6 + power + 7) // 8
    return b"pco!" + bytes([3, uniform_type]) + v.to_bytes(nbytes, "little") + bytes([4, 1])


def compress_local_shard(local_nums, local_chunk_ns, config):
    """Compress this rank's chunks on its GPU: returns (bytes of the chunks back to back, per-chunk byte sizes)."""
    import ctypes as C
    import struct

    from . import _lib
    from ._lib import ChunkConfig, PagingSpec

    L = _lib.lib()
    arr = np.ascontiguousarray(local_nums)
    dt = _lib.dtype_byte(arr.dtype)
    cfg = ChunkConfig(config.compression_level, config.mode_spec, config.delta_spec, PagingSpec.exact_page_sizes(local_chunk_ns), config.enable_8_bit)._to_c()
    cap = L.pco_standalone_guarantee_file_size(arr.size, dt) + 160 * len(local_chunk_ns)
    dst = np.empty(cap, dtype=np.uint8)
    icap = L.pco_b200_index_size_bound(arr.size, len(local_chunk_ns)) + 64 * len(local_chunk_ns)
    idx = np.empty(icap, dtype=np.uint8)
    nw, il = C.c_size_t(), C.c_size_t()
    rc = L.pco_b200_compress_ex(arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.size), C.c_ubyte(dt), C.byref(cfg), C.c_int(0),
                                dst.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(nw), idx.ctypes.data_as(C.c_void_p), C.c_size_t(icap),
                                C.byref(il), C.c_uint32(8), None)  # PCO_B200_CHUNKS_ONLY
    _lib.check(rc)
    # chunk byte offsets from the side index (IndexHeader 64 B, IndexChunk 32 B; csrc/codec_common.cuh)
    buf = idx[: il.value].tobytes()
    n_chunks, chunks_offset = struct.unpack_from("<Q", buf, 8)[0], struct.unpack_from("<Q", buf, 32)[0]
    offs = [struct.unpack_from("<Q", buf, chunks_offset + 32 * c)[0] for c in range(n_chunks)] + [nw.value]
    sizes = [offs[i + 1] - offs[i] for i in range(n_chunks)]
    return dst[: nw.value].tobytes(), sizes


def gather_standalone_file(local_bytes, local_sizes, n_total, world, rank, group=None, device="cpu", uniform_type=0):
    """One all-gather of compressed pages; every rank returns the full standalone file (bytes).

    local_bytes: this rank's chunks back to back (chunk ids rank, rank + world, ...); local_sizes: their byte sizes.
    """
    import torch
    import torch.distributed as dist

    n_local = len(local_sizes)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    mine = torch.tensor([n_local], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(counts, mine, group=group)
    else:
        counts[0] = n_local
    max_chunks = int(counts.max().item())
    # per-chunk sizes, padded to the largest shard
    sz = torch.zeros(max(max_chunks, 1), dtype=torch.int64, device=device)
    if n_local:
        sz[:n_local] = torch.tensor(local_sizes, dtype=torch.int64, device=device)
    all_sz = torch.zeros(world * max(max_chunks, 1), dtype=torch.int64, device=device)
    if world > 1:
        dist.all_gather_into_tensor(all_sz, sz, group=group)
    else:
        all_sz.copy_(sz)
    all_sz = all_sz.view(world, -1)
    shard_bytes = all_sz.sum(dim=1)
    pad = int(shard_bytes.max().item())
    pad = max((pad + 255) // 256 * 256, 256)
    payload = torch.zeros(pad, dtype=torch.uint8, device=device)
    if len(local_bytes):
        payload[: len(local_bytes)] = torch.frombuffer(bytearray(local_bytes), dtype=torch.uint8).to(device)
    gathered = torch.zeros(world * pad, dtype=torch.uint8, device=device)
    if world > 1:
        dist.all_gather_into_tensor(gathered, payload, group=group)  # THE collective: compressed pages of every shard
    else:
        gathered.copy_(payload)
    gathered = gathered.view(world, pad)
    # assemble in chunk order: chunk c lives in shard c % world at local position c // world
    all_sz_h = all_sz.cpu().numpy()
    offs = np.concatenate([np.zeros((world, 1), dtype=np.int64), np.cumsum(all_sz_h, axis=1)], axis=1)
    counts_h = counts.cpu().numpy()
    n_chunks = int(counts_h.sum())
    g = gathered.cpu().numpy()
    parts = [standalone_header(n_total, uniform_type)]
    for c in range(n_chunks):
        r, j = c % world, c // world
        parts.append(g[r, offs[r, j]: offs[r, j + 1]].tobytes())
    parts.append(b"\x00")
    return b"".join(parts)


class DevicePageGather:
    """The page gather done by the GPUs themselves (pcodec_b200/csrc/gather_kernels.cuh; include/pco_b200.h "sharded writers").

    Every rank owns a file buffer in HBM that the other ranks map through cudaIpc handles (exchanged once, here, with
    all_gather_object); per call the per-chunk byte sizes are all-gathered as a DEVICE tensor (no size visits the host),
    one scan kernel turns them into file offsets and one copy kernel stores this rank's chunks into every rank's buffer
    over NVLink at their final positions.  `gather()` is asynchronous on the given stream; `wait()` reports the outcome.
    One process per GPU (several ranks may also share one GPU: the peer buffers are then ordinary local memory)."""

    def __init__(self, file_cap, world, rank, group=None):
        import ctypes as C

        import torch.distributed as dist

        from . import _lib

        self.L, self.C = _lib.lib(), C
        self._check = _lib.check
        self.world, self.rank, self.group, self.file_cap = world, rank, group, int(file_cap)
        L = self.L
        L.pco_b200_gather_pages.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t, C.c_ubyte, C.c_void_p, C.c_size_t, C.c_void_p,
                                            C.c_uint32, C.c_void_p]
        L.pco_b200_chunk_sizes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        L.pco_b200_gather_status.argtypes = [C.c_void_p]
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self._check(L.pco_b200_ipc_alloc(C.c_size_t(self.file_cap), C.byref(ptr), handle))
        self.own = ptr.value
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle), group=group)
        else:
            handles[0] = bytes(handle)
        self.peers = (C.c_void_p * world)()
        self._opened = []
        for r in range(world):
            if r == rank:
                self.peers[r] = self.own
            else:
                p = C.c_void_p()
                self._check(L.pco_b200_ipc_open((C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p)))
                self.peers[r] = p.value
                self._opened.append(p.value)

    def gather(self, chunks_ptr, all_sizes_ptr, n_local, n_total_numbers, file_len_ptr, stream_ptr, uniform_type=0, max_ctas=0):
        """chunks_ptr: this rank's chunk bytes (device); all_sizes_ptr: device u64 [world][n_local]; file_len_ptr: device u64 out."""
        C = self.C
        self._check(self.L.pco_b200_gather_pages(C.c_void_p(chunks_ptr), C.c_void_p(all_sizes_ptr), C.c_uint32(self.world), C.c_uint32(self.rank), C.c_size_t(n_local),
                                                 C.c_size_t(n_total_numbers), C.c_ubyte(uniform_type), C.cast(self.peers, C.c_void_p), C.c_size_t(self.file_cap),
                                                 C.c_void_p(file_len_ptr), C.c_uint32(max_ctas), C.c_void_p(stream_ptr)))

    def chunk_sizes(self, index_ptr, index_len, sizes_ptr, n_chunks, stream_ptr):
        C = self.C
        self._check(self.L.pco_b200_chunk_sizes(C.c_void_p(index_ptr), C.c_size_t(index_len), C.c_void_p(sizes_ptr), C.c_size_t(n_chunks), C.c_void_p(stream_ptr)))

    def wait(self, stream_ptr):
        self._check(self.L.pco_b200_gather_status(self.C.c_void_p(stream_ptr)))

    def file_tensor(self):
        """This rank's file buffer as a uint8 torch tensor view (no copy)."""
        import torch

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (self.file_cap,), "typestr": "|u1", "data": (self.own, False), "version": 2}
        return torch.as_tensor(h, device="cuda")

    def close(self):
        """Collective in spirit: call on every rank after a barrier (a peer must not be writing into a freed buffer)."""
        for p in self._opened:
            self.L.pco_b200_ipc_close(self.C.c_void_p(p))
        self._opened = []
        if self.own:
            self.L.pco_b200_ipc_free(self.C.c_void_p(self.own))
            self.own = None
