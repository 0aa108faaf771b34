"""Page-locked host buffers in place (pco_b200_zero_copy, include/pco_b200.h): compress reads a pinned `nums` over PCIe inside its
one pass, the bit-pack kernel stores the file into a pinned `dst`, decompress stores the numbers into a pinned `dst` - same bytes and
same numbers as the staged path and as the oracle, for every mask, including the cases that must leave the in-place route again
(wide-range chunks redo the front end from a staged copy; pageable buffers never take it)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import pyoracle as oracle  # noqa: E402
from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec, PcoError, _lib  # noqa: E402

DT = {np.dtype(np.uint64): torch.int64, np.dtype(np.int64): torch.int64, np.dtype(np.uint32): torch.int32, np.dtype(np.int32): torch.int32,
      np.dtype(np.uint16): torch.int16, np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}


def _pinned(host, lead=0):
    """a page-locked copy of `host`, `lead` elements into its allocation (alignment cases)"""
    t = torch.empty(host.size + lead, dtype=DT[host.dtype], pin_memory=True)
    v = t[lead:]
    v.numpy().view(host.dtype)[:] = host
    return t, v


def _roundtrip(host, cfg, ocfg, mask, page_n, lead=0, dst_short=0, comp_off=0):
    L = _lib.lib()
    n = host.size
    dt = _lib.dtype_byte(host.dtype)
    keep, src = _pinned(host, lead)
    n_chunks = (n + page_n - 1) // page_n
    cap = L.pco_standalone_guarantee_file_size(n, dt) + 160 * n_chunks
    icap = L.pco_b200_index_size_bound(n, n_chunks)
    comp = torch.zeros(cap + 64, dtype=torch.uint8, pin_memory=True)
    idx = torch.zeros(icap, dtype=torch.uint8, pin_memory=True)
    out_keep = torch.zeros(n + lead + 8, dtype=DT[host.dtype], pin_memory=True)
    out = out_keep[lead:]
    nw, il = C.c_size_t(), C.c_size_t()
    prog = _lib._CProgress()
    prev = L.pco_b200_zero_copy(C.c_int(mask))
    try:
        assert L.pco_b200_zero_copy(C.c_int(-1)) == mask
        c = cfg._to_c()
        _lib.check(L.pco_b200_compress_ex(C.c_void_p(src.data_ptr()), C.c_size_t(n), C.c_ubyte(dt), C.byref(c), C.c_int(0), C.c_void_p(comp.data_ptr() + comp_off),
                                          C.c_size_t(cap), C.byref(nw), C.c_void_p(idx.data_ptr()), C.c_size_t(icap), C.byref(il), C.c_uint32(0), None))
        data = bytes(comp.numpy()[comp_off : comp_off + nw.value])
        assert data == oracle.simple_compress(host, ocfg), f"mask {mask}: compressed bytes differ from the oracle's"
        assert not comp.numpy()[comp_off + nw.value :].any() and not comp.numpy()[:comp_off].any(), "bytes outside the file were written"
        want = n - dst_short
        _lib.check(L.pco_b200_decompress_ex(C.c_void_p(comp.data_ptr() + comp_off), nw, C.c_ubyte(dt), C.c_void_p(out.data_ptr()), C.c_size_t(want), C.byref(prog),
                                            C.c_void_p(idx.data_ptr()), il, C.c_uint32(0), None))
        got = out.numpy().view(host.dtype)
        assert prog.n_processed == want
        assert np.array_equal(got[:want].view(np.uint8), host[:want].view(np.uint8)), f"mask {mask}: decoded numbers differ"
        assert not out_keep.numpy().view(np.uint8)[(lead + want) * host.itemsize :].any(), "numbers behind the destination's end were written"
    finally:
        L.pco_b200_zero_copy(C.c_int(prev))


def _cfgs(order, page_n):
    cfg = ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(order) if order else DeltaSpec.no_op(), paging_spec=PagingSpec.equal_pages_up_to(page_n))
    ocfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE if order else oracle.DELTA_NOOP, delta_order=order, max_page_n=page_n)
    return cfg, ocfg


@pytest.mark.parametrize("mask", [0, 1, 2, 4, 7])
def test_in_place_matches_staged_and_oracle_u64(mask):
    rng = np.random.default_rng(5)
    host = np.cumsum(rng.geometric(0.001, size=9 * 8192 + 77)).astype(np.uint64)
    cfg, ocfg = _cfgs(1, 8192)
    _roundtrip(host, cfg, ocfg, mask, 8192)


@pytest.mark.parametrize("dtype,order", [(np.uint32, 0), (np.int32, 2), (np.uint16, 1), (np.int64, 7), (np.float32, 0), (np.float64, 1)])
def test_in_place_other_types_and_orders(dtype, order):
    rng = np.random.default_rng(6)
    n = 5 * 4096 + 300
    if np.dtype(dtype).kind == "f":
        host = np.cumsum(rng.normal(0, 1, size=n)).astype(dtype)
    else:
        host = (np.cumsum(rng.integers(0, 40, size=n)) % (1 << min(62, 8 * np.dtype(dtype).itemsize - 1))).astype(dtype)
    cfg, ocfg = _cfgs(order, 4096)
    _roundtrip(host, cfg, ocfg, 7, 4096)


def test_in_place_wide_range_chunks_restage():
    """uniform 64-bit numbers: split_count_kernel's speculation fails, the front end is redone from a staged copy of the pinned input"""
    rng = np.random.default_rng(7)
    host = rng.integers(0, 1 << 63, size=3 * 4096 + 5, dtype=np.uint64)
    cfg, ocfg = _cfgs(0, 4096)
    _roundtrip(host, cfg, ocfg, 7, 4096)


@pytest.mark.parametrize("lead", [1, 3])
def test_in_place_unaligned_buffers_and_short_destination(lead):
    rng = np.random.default_rng(8)
    host = np.cumsum(rng.geometric(0.01, size=4 * 4096 + 9)).astype(np.uint64)
    cfg, ocfg = _cfgs(1, 4096)
    _roundtrip(host, cfg, ocfg, 7, 4096, lead=lead, comp_off=lead)
    _roundtrip(host, cfg, ocfg, 7, 4096, lead=lead, dst_short=4096 + 100, comp_off=2 * lead + 1)


def test_in_place_destination_too_small_is_io_and_writes_nothing_behind_cap():
    L = _lib.lib()
    rng = np.random.default_rng(9)
    host = np.cumsum(rng.geometric(0.01, size=6 * 4096)).astype(np.uint64)
    keep, src = _pinned(host)
    cfg, _ = _cfgs(1, 4096)
    c = cfg._to_c()
    full = C.c_size_t()
    big = torch.zeros(L.pco_standalone_guarantee_file_size(host.size, 2) + 1000, dtype=torch.uint8, pin_memory=True)
    prev = L.pco_b200_zero_copy(C.c_int(7))
    try:
        _lib.check(L.pco_b200_compress_ex(C.c_void_p(src.data_ptr()), C.c_size_t(host.size), C.c_ubyte(2), C.byref(c), C.c_int(0), C.c_void_p(big.data_ptr()),
                                          C.c_size_t(big.numel()), C.byref(full), None, C.c_size_t(0), None, C.c_uint32(0), None))
        small = full.value // 2
        guard = torch.full((small + 64,), 0xAB, dtype=torch.uint8).pin_memory()
        nw = C.c_size_t()
        rc = L.pco_b200_compress_ex(C.c_void_p(src.data_ptr()), C.c_size_t(host.size), C.c_ubyte(2), C.byref(c), C.c_int(0), C.c_void_p(guard.data_ptr()),
                                    C.c_size_t(small), C.byref(nw), None, C.c_size_t(0), None, C.c_uint32(0), None)
        with pytest.raises(PcoError) as e:
            _lib.check(rc)
        assert e.value.kind == "Io"
        assert bool((guard[small:] == 0xAB).all())
    finally:
        L.pco_b200_zero_copy(C.c_int(prev))


def test_pageable_buffers_take_the_staged_path_whatever_the_mask():
    L = _lib.lib()
    from pcodec_b200 import standalone

    rng = np.random.default_rng(10)
    host = np.cumsum(rng.geometric(0.01, size=20000)).astype(np.uint64)
    cfg, ocfg = _cfgs(1, 4096)
    prev = L.pco_b200_zero_copy(C.c_int(7))
    try:
        data = standalone.simple_compress(host, cfg)
        assert data == oracle.simple_compress(host, ocfg)
        assert np.array_equal(standalone.simple_decompress(data, np.uint64), host)
    finally:
        L.pco_b200_zero_copy(C.c_int(prev))
