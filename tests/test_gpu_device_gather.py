"""The device-side page gather (pcodec_b200/csrc/gather_kernels.cuh, sharded.DevicePageGather): world_size-2 processes compress their
round-robin shards on the GPU, the GPUs store the pages into every rank's file buffer at their final offsets, and every rank's buffer
must be byte-equal to the single-GPU file and to the oracle's (pco/src/standalone/simple.rs:62-91).

The two ranks use two GPUs when the box has them and share GPU 0 otherwise (the peers' buffers are mapped through cudaIpc either way;
NCCL refuses two ranks on one GPU, so the rendezvous and the size exchange of this TEST go through gloo - the product's bench uses NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(n):
    rng = np.random.default_rng(7)
    return np.cumsum(rng.geometric(0.002, size=n)).astype(np.uint64)


def _worker(rank, world, port, n, max_page_n, q):
    sys.path.insert(0, ROOT)
    import ctypes as C

    import torch
    import torch.distributed as dist

    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec, _lib, sharded

    try:
        dev_idx = rank if torch.cuda.device_count() >= world else 0
        torch.cuda.set_device(dev_idx)
        dev = torch.device("cuda", dev_idx)
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        L = _lib.lib()
        nums = _data(n)
        plan = sharded.shard_plan(n, world, max_page_n)
        n_local_max = max(len(p) for p in plan)
        mine = plan[rank]
        local = np.concatenate([nums[s:e] for (_, s, e) in mine]) if mine else np.zeros(0, dtype=np.uint64)
        local_ns = [e - s for (_, s, e) in mine]
        cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1), paging_spec=PagingSpec.exact_page_sizes(local_ns))._to_c()
        d_nums = torch.from_numpy(local.view(np.int64)).to(dev)
        cap = L.pco_standalone_guarantee_file_size(max(local.size, 1), 2) + 160 * len(local_ns) + 64
        icap = L.pco_b200_index_size_bound(local.size, len(local_ns)) + 64 * len(local_ns) + 64
        d_comp = torch.zeros(cap, dtype=torch.uint8, device=dev)
        d_index = torch.zeros(icap, dtype=torch.uint8, device=dev)
        nw, il = C.c_size_t(), C.c_size_t()
        stream = torch.cuda.current_stream()
        sp = C.c_void_p(stream.cuda_stream)
        if local.size:
            _lib.check(L.pco_b200_compress_ex(C.c_void_p(d_nums.data_ptr()), C.c_size_t(local.size), C.c_ubyte(2), C.byref(cfg), C.c_int(0), C.c_void_p(d_comp.data_ptr()),
                                              C.c_size_t(cap), C.byref(nw), C.c_void_p(d_index.data_ptr()), C.c_size_t(icap), C.byref(il),
                                              C.c_uint32(1 | 2 | 4 | 8), sp))  # SRC | DST | INDEX on device, CHUNKS_ONLY
        file_cap = L.pco_standalone_guarantee_file_size(n, 2) + 64
        gather = sharded.DevicePageGather(file_cap, world, rank)
        sizes = torch.zeros(n_local_max, dtype=torch.int64, device=dev)
        if local_ns:
            gather.chunk_sizes(d_index.data_ptr(), il.value, sizes.data_ptr(), len(local_ns), stream.cuda_stream)
        # the test's size exchange: gloo on host copies (two ranks may share one GPU, which NCCL refuses)
        sizes_h = sizes.cpu()
        all_h = torch.zeros(world * n_local_max, dtype=torch.int64)
        dist.all_gather_into_tensor(all_h, sizes_h)
        all_sizes = all_h.to(dev)
        file_len = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.barrier()
        gather.gather(d_comp.data_ptr(), all_sizes.data_ptr(), n_local_max, n, file_len.data_ptr(), stream.cuda_stream, max_ctas=8)
        gather.wait(stream.cuda_stream)
        dist.barrier()  # every rank's stores into this rank's buffer have completed
        torch.cuda.synchronize()
        flen = int(file_len.item())
        out = bytes(gather.file_tensor()[:flen].cpu().numpy())
        q.put((rank, out, None))
        dist.barrier()
        gather.close()
        dist.destroy_process_group()
    except Exception as ex:  # noqa: BLE001
        import traceback

        q.put((rank, None, traceback.format_exc() + str(ex)))


@pytest.mark.parametrize("n,max_page_n", [(40000, 4096), (11 * 3000 + 5, 3000), (3 * (1 << 18) + 17, 1 << 18)])
def test_device_gather_equals_whole_file(oracle, n, max_page_n):
    import torch.multiprocessing as mp

    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec, standalone

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, max_page_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(world):
        rank, out, err = q.get(timeout=300)
        assert err is None, f"rank {rank}: {err}"
        results[rank] = out
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    nums = _data(n)
    cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1, max_page_n=max_page_n)
    expected = oracle.simple_compress(nums, cfg)
    single = standalone.simple_compress(nums, ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1),
                                                          paging_spec=PagingSpec.equal_pages_up_to(max_page_n)))
    assert single == expected
    for r in range(world):
        assert len(results[r]) == len(expected), (r, len(results[r]), len(expected))
        assert results[r] == expected, f"rank {r}: gathered file differs from the single-GPU / oracle file"
