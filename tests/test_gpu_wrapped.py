"""GPU parity tests for the wrapped API (one page per chunk): bytes equal the oracle's ChunkCompressor, pages decode both ways."""
import numpy as np
import pytest

from tests.golden_generators import bits_view
from tests.test_gpu_encode import _cfgs, _walk

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,mode,order", [(np.uint64, "classic", 1), (np.int32, "classic", 0), (np.float64, "float_mult", 2),
                                              (np.float32, "float_quant", 0), (np.uint32, "int_mult", 1), (np.int16, "classic", 3)])
def test_wrapped_chunk_bytes_equal_oracle_and_pages_decode(oracle, dtype, mode, order):
    from pcodec_b200 import wrapped

    n = 20000
    x = _walk(dtype, n, seed=21)
    if mode == "float_mult":
        x = (np.round(np.cumsum(np.random.default_rng(1).normal(size=n)) * 100) * 0.01).astype(dtype)
    if mode == "int_mult":
        x = (x.astype(np.uint64) // 8 * 8 + 3).astype(dtype)
    ours_cfg, their_cfg = _cfgs(oracle, mode=mode, order=order, max_page_n=1 << 18)
    fc = wrapped.FileCompressor()
    assert fc.write_header() == bytes([4, 1])
    cc = fc.chunk_compressor(x, ours_cfg)
    occ = oracle.ChunkCompressor(x, their_cfg)
    assert cc.n_per_page() == occ.n_per_page() == [n]
    meta, page = cc.write_meta(), cc.write_page(0)
    assert meta == occ.write_meta()
    assert page == occ.write_page(0)
    # the oracle decodes our page; our page decoder decodes it too, through the reference-shaped handles
    got, mc, pc = oracle.wrapped_decompress_page(meta, page, dtype, n)
    assert mc == len(meta) and pc == len(page)
    np.testing.assert_array_equal(bits_view(got), bits_view(x))
    stream = fc.write_header() + meta + page
    fd, used = wrapped.FileDecompressor.new(stream)
    assert used == 2
    cd, mused = fd.chunk_decompressor(stream[used:], dtype)
    assert mused == len(meta)
    dst = np.zeros(n, dtype=dtype)
    prog, pused = cd.read_page_into(stream[used + mused:], n, dst)
    assert prog.finished and prog.n_processed == n and pused == len(page)
    np.testing.assert_array_equal(bits_view(dst), bits_view(x))


def test_wrapped_errors(oracle):
    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec, PcoError, wrapped

    x = np.arange(3000, dtype=np.uint32)
    fc = wrapped.FileCompressor()
    cfg = ChunkConfig(mode_spec=ModeSpec.try_int_mult(8), delta_spec=DeltaSpec.no_op(), paging_spec=PagingSpec.equal_pages_up_to(1000))
    with pytest.raises(PcoError) as e:
        fc.chunk_compressor(x, cfg)  # several pages sharing bins: classic mode only on the GPU path so far
    assert e.value.kind == "Unsupported"
    cc = fc.chunk_compressor(x, ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.no_op()))
    with pytest.raises(PcoError) as e:
        cc.write_page(1)  # chunk_compressor.rs:661-666
    assert e.value.kind == "InvalidArgument"
    with pytest.raises(PcoError) as e:
        wrapped.FileDecompressor.new(bytes([9, 0]))
    assert e.value.kind == "Corruption"
    # PageDecompressor::read's dst rule (page_decompressor.rs:200-206)
    meta, page = cc.write_meta(), cc.write_page(0)
    cd, _ = wrapped.FileDecompressor().chunk_decompressor(meta, np.uint32)
    with pytest.raises(PcoError) as e:
        cd.read_page_into(page, 3000, np.zeros(1000, dtype=np.uint32))
    assert e.value.kind == "InvalidArgument"
    dst = np.zeros(1024, dtype=np.uint32)
    prog, _ = cd.read_page_into(page, 3000, dst)
    assert prog.n_processed == 1024 and not prog.finished
    np.testing.assert_array_equal(dst, x[:1024])


@pytest.mark.parametrize("dtype,mode,order", [(np.uint64, "classic", 1), (np.float64, "float_mult", 2), (np.int32, "classic", 0)])
def test_pages_of_a_multi_page_chunk_decode(oracle, dtype, mode, order):
    """A chunk whose PagingSpec yields several pages (they share the chunk's bins, chunk_compressor.rs:129-140), written by
    the oracle: every page decodes through the wrapped handles from the one chunk meta."""
    from pcodec_b200 import wrapped

    n = 3 * 5000 + 123
    x = _walk(dtype, n, seed=33)
    if mode == "float_mult":
        x = (np.round(np.cumsum(np.random.default_rng(2).normal(size=n)) * 100) * 0.01).astype(dtype)
    _, their_cfg = _cfgs(oracle, mode=mode, order=order, max_page_n=5000)
    occ = oracle.ChunkCompressor(x, their_cfg)
    pages_n = occ.n_per_page()
    assert len(pages_n) == 4 and sum(pages_n) == n
    meta = occ.write_meta()
    cd, mused = wrapped.FileDecompressor().chunk_decompressor(meta, dtype)
    assert mused == len(meta)
    start = 0
    for i, pn in enumerate(pages_n):
        page = occ.write_page(i)
        dst = np.zeros(pn, dtype=dtype)
        prog, pused = cd.read_page_into(page, pn, dst)
        assert prog.finished and prog.n_processed == pn and pused == len(page), (i, prog, pused, len(page))
        np.testing.assert_array_equal(bits_view(dst), bits_view(x[start:start + pn]))
        start += pn


@pytest.mark.parametrize("dtype", [np.uint64, np.int32, np.float64, np.uint16])
@pytest.mark.parametrize("order", [0, 1, 2])
@pytest.mark.parametrize("exact", [None, [700, 1, 5000, 3000, 299]])
def test_multi_page_chunk_bytes_equal_oracle(oracle, dtype, order, exact):
    """Several pages per chunk: bins trained on all pages, deltas per page (chunk_compressor.rs:129-217).  The chunk meta and every
    page must equal the oracle's ChunkCompressor byte for byte, and decode page by page."""
    from pcodec_b200 import wrapped

    n = 9000
    x = _walk(dtype, n, seed=5 + order)
    ours_cfg, their_cfg = _cfgs(oracle, order=order, max_page_n=2500, exact=exact)
    cc = wrapped.FileCompressor().chunk_compressor(x, ours_cfg)
    occ = oracle.ChunkCompressor(x, their_cfg)
    assert cc.n_per_page() == occ.n_per_page() and len(cc.n_per_page()) > 1
    meta = cc.write_meta()
    assert meta == occ.write_meta()
    cd, _ = wrapped.FileDecompressor().chunk_decompressor(meta, dtype)
    start = 0
    for i, pn in enumerate(cc.n_per_page()):
        page = cc.write_page(i)
        assert page == occ.write_page(i), (i, len(page), len(occ.write_page(i)))
        dst = np.zeros(pn, dtype=dtype)
        prog, used = cd.read_page_into(page, pn, dst)
        assert prog.finished and used == len(page)
        np.testing.assert_array_equal(bits_view(dst), bits_view(x[start:start + pn]))
        start += pn


@pytest.mark.parametrize("dtype", ["f2", "f4", "f8", "i2", "i4", "i8", "u2", "u4", "u8"])
def test_reference_python_wrapped_case(dtype):
    """pco_python/test/test_wrapped.py::test_compress, against pcodec_b200.wrapped (default ChunkConfig, 2 pages)."""
    from pcodec_b200 import ChunkConfig, PagingSpec, PcoError, wrapped

    data = np.random.default_rng(12345).uniform(0, 1000, size=[10]).astype(dtype)
    page_sizes = [6, 4]
    fc = wrapped.FileCompressor()
    header = fc.write_header()
    cc = fc.chunk_compressor(data, ChunkConfig(paging_spec=PagingSpec.exact_page_sizes(page_sizes)))
    assert cc.n_per_page() == page_sizes
    chunk_meta, page0, page1 = cc.write_meta(), cc.write_page(0), cc.write_page(1)
    with pytest.raises(PcoError) as e:
        cc.write_page(2)
    assert e.value.kind == "InvalidArgument" and "page idx exceeds num pages" in str(e.value)
    fd, n_bytes_read = wrapped.FileDecompressor.new(header)
    assert n_bytes_read == len(header)
    _, n_bytes_read = wrapped.FileDecompressor.new(header + b"foo")  # undershooting is fine
    assert n_bytes_read == len(header)
    cd, n_bytes_read = fd.chunk_decompressor(chunk_meta, np.dtype(dtype))
    assert n_bytes_read == len(chunk_meta)
    dst1 = np.zeros(100).astype(dtype)  # page 1 holds elements 6..10
    _progress, n_bytes_read = cd.read_page_into(page1, 4, dst1)
    np.testing.assert_array_equal(dst1[4:], np.zeros(96))
    np.testing.assert_array_equal(dst1[:4], data[6:])
    assert n_bytes_read == len(page1)
    dst0 = np.zeros(6).astype(dtype)
    _progress, n_bytes_read = cd.read_page_into(page0, 6, dst0)
    np.testing.assert_array_equal(dst0, data[:6])
    assert n_bytes_read == len(page0)


def test_reference_low_level_wrapped_cases(oracle):
    """pco/src/tests/low_level.rs:98-131 (test_low_level_wrapped): chunks with several pages, a page shorter than the delta order,
    one-number pages - compressed through the wrapped handles, every page decoded back, and the bytes compared with the oracle's
    where the config is explicit; pco/src/wrapped/guarantee.rs:60-87: meta + pages never exceed chunk_size::<L>(n)."""
    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec, _lib, wrapped

    cases = [
        (np.arange(1700, dtype=np.int32), ChunkConfig(delta_spec=DeltaSpec.no_op(), paging_spec=PagingSpec.equal_pages_up_to(600)), None),
        (np.arange(500, dtype=np.int32), ChunkConfig(delta_spec=DeltaSpec.try_consecutive(2), paging_spec=PagingSpec.exact_page_sizes([1, 499])), None),
        (np.array([1, 2, 3], dtype=np.int32), ChunkConfig(), None),
        (np.array([1, 2, 3], dtype=np.int32), ChunkConfig(paging_spec=PagingSpec.equal_pages_up_to(1)), None),
        # explicit mode + delta: byte-exact against the oracle
        (np.arange(1700, dtype=np.int32), ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.no_op(), paging_spec=PagingSpec.equal_pages_up_to(600)),
         oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP, max_page_n=600)),
        (np.arange(500, dtype=np.int32), ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(2), paging_spec=PagingSpec.exact_page_sizes([1, 499])),
         oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=2, exact_pages=[1, 499])),
        (np.random.default_rng(0).integers(0, 2**32 - 1, size=100, dtype=np.uint64).astype(np.uint32),
         ChunkConfig(mode_spec=ModeSpec.classic(), delta_spec=DeltaSpec.try_consecutive(1), paging_spec=PagingSpec.equal_pages_up_to(10)),
         oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1, max_page_n=10)),
    ]
    L = _lib.lib()
    for nums, cfg, ocfg in cases:
        cc = wrapped.FileCompressor().chunk_compressor(nums, cfg)
        pages_n = cc.n_per_page()
        assert sum(pages_n) == nums.size
        meta = cc.write_meta()
        pages = [cc.write_page(i) for i in range(len(pages_n))]
        # chunk_size::<L>(n) = baseline meta + n * bits / 8 (wrapped/guarantee.rs:11-37) = the standalone chunk guarantee minus its 4-byte preamble
        bound = L.pco_standalone_guarantee_file_size(nums.size, _lib.dtype_byte(nums.dtype)) - L.pco_standalone_guarantee_file_size(0, _lib.dtype_byte(nums.dtype)) - 4
        if len(pages_n) == 1:
            assert len(meta) + sum(len(p) for p in pages) <= bound
        if ocfg is not None:
            occ = oracle.ChunkCompressor(nums, ocfg)
            assert occ.n_per_page() == pages_n and meta == occ.write_meta()
            assert pages == [occ.write_page(i) for i in range(len(pages_n))]
        cd, used = wrapped.FileDecompressor().chunk_decompressor(meta, nums.dtype)
        assert used == len(meta)
        start = 0
        for pn, page in zip(pages_n, pages):
            dst = np.zeros(pn, dtype=nums.dtype)
            prog, pused = cd.read_page_into(page, pn, dst)
            assert prog.finished and pused == len(page)
            np.testing.assert_array_equal(dst, nums[start:start + pn])
            start += pn
