"""Host logic of the multi-GPU path on CPU: world_size-2 gloo processes shard chunks round-robin,
all-gather their compressed pages once, and must assemble the same standalone file the whole-array compress gives.
The per-shard compress is the oracle here (the GPU library needs a device); sharding, gather and assembly are the product's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(n):
    rng = np.random.default_rng(42)
    return np.cumsum(rng.geometric(0.01, size=n)).astype(np.uint64)


def _worker(rank, world, port, n, max_page_n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import pyoracle
    from pcodec_b200 import sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    nums = _data(n)
    plan = sharded.shard_plan(n, world, max_page_n)
    cfg = pyoracle.make_config(mode=pyoracle.MODE_CLASSIC, delta=pyoracle.DELTA_CONSECUTIVE, delta_order=1)
    local_bytes, sizes = b"", []
    for (c, s, e) in plan[rank]:
        whole = pyoracle.simple_compress(nums[s:e], cfg)
        info = pyoracle.inspect(whole, np.uint64)["chunks"][0]
        chunk = whole[info["chunk_start"]: info["chunk_end"]]
        local_bytes += chunk
        sizes.append(len(chunk))
    out = sharded.gather_standalone_file(local_bytes, sizes, n, world, rank)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,max_page_n", [(10000, 1024), (5 * 700 + 3, 700), (100, 1 << 18)])
def test_sharded_gather_equals_whole_file(oracle, n, max_page_n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, max_page_n, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1, max_page_n=max_page_n)
    expected = oracle.simple_compress(_data(n), cfg)
    assert results[0] == expected and results[1] == expected


def test_shard_plan_round_robin():
    from pcodec_b200 import sharded

    plan = sharded.shard_plan(10 * 100 + 7, 4, 100)
    sizes = sharded.chunk_sizes(1007, 100)
    assert sum(sizes) == 1007 and len(sizes) == 11 and max(sizes) - min(sizes) <= 1
    seen = sorted(c for r in plan for (c, _, _) in r)
    assert seen == list(range(11))
    for r, chunks in enumerate(plan):
        assert all(c % 4 == r for (c, _, _) in chunks)
    assert sharded.standalone_header(1 << 18) == bytes([0x70, 0x63, 0x6F, 0x21, 0x03, 0x00, 0x12, 0x00, 0x00, 0x01, 0x04, 0x01])
