"""GPU parity tests for the decode path (K6 walk, K7+K8+K9 fused decode) through the C-ABI.

The oracle (oracle/) is only the checker here: streams are produced by the oracle, decoded by
libcpcodec.so on the GPU, and compared bit-exactly with the original numbers and the oracle's decode.
"""
import numpy as np
import pytest

from tests.golden_generators import GENERATORS, bits_view, load_assets

pytestmark = pytest.mark.gpu

# all 13 reference assets (pco/assets/*.pco): the last three - Lookback delta, Dict mode, Conv1 delta - are served by the single-thread
# device decoder behind the fast kernels (pcodec_b200/csrc/decode_cold.cuh)
GPU_ASSETS = ["v0_0_0_classic", "v0_0_0_delta_float_mult", "v0_1_0_delta_int_mult", "v0_1_1_standalone_versioned", "v0_3_0_f16",
              "v0_3_0_float_quant", "v0_4_5_uniform_type", "v0_4_8_minor_version", "v1_0_0_u8", "v1_0_0_i8",
              "v0_4_0_lookback_delta", "v1_0_0_dict", "v1_0_0_conv1"]
OUT_OF_SCOPE_ASSETS = []


@pytest.fixture(scope="module")
def sa():
    from pcodec_b200 import standalone

    return standalone


def _walk(dtype, n, seed, scale=0.05):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        x = np.cumsum(rng.normal(size=n)).astype(dtype)
        if n > 5:
            x[3], x[4], x[5] = np.nan, -np.inf, -0.0
        return x
    steps = rng.geometric(scale, size=n).astype(np.int64) - int(1 / scale) // 2
    return np.cumsum(steps).astype(np.uint64).astype(np.dtype(dtype).str.replace("i", "u")).view(dtype)


@pytest.mark.parametrize("name", GPU_ASSETS)
def test_golden_assets_decode_on_gpu(sa, name):
    expected = GENERATORS[name]()
    got = sa.simple_decompress(load_assets()[name], expected.dtype)
    assert got.shape == expected.shape
    np.testing.assert_array_equal(bits_view(got), bits_view(expected))


def test_every_golden_asset_is_covered():
    assert sorted(GPU_ASSETS) == sorted(GENERATORS) and OUT_OF_SCOPE_ASSETS == []


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.uint16, np.int16, np.float16, np.uint32, np.int32, np.float32, np.uint64, np.int64, np.float64])
@pytest.mark.parametrize("order", [0, 1, 2, 7])
def test_classic_consecutive_vs_oracle(sa, oracle, dtype, order):
    for n in (1, 7, 255, 256, 257, 513, 5000, 70000):
        nums = _walk(dtype, n, seed=n + order)
        cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=order, max_page_n=1 << 14)
        data = oracle.simple_compress(nums, cfg)
        got = sa.simple_decompress(data, dtype)
        np.testing.assert_array_equal(bits_view(got), bits_view(nums))
        # the same through an explicit side index built on the device
        idx = sa.build_index(data, dtype)
        got2 = sa.simple_decompress(data, dtype, index=idx)
        np.testing.assert_array_equal(bits_view(got2), bits_view(nums))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", [0, 1, 2])
def test_float_mult_vs_oracle(sa, oracle, dtype, order):
    n = 40000
    x = (np.round(1e5 * np.cos(2 * np.pi * np.arange(n) / (n / 103))) * 0.01).astype(dtype)
    x[::97] = (x[::97] * (1 + 1e-6)).astype(dtype)
    x[5], x[6], x[7], x[8] = np.nan, np.inf, dtype(1e30), dtype(-0.0)
    cfg = oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=order,
                             max_page_n=1 << 13)
    data = oracle.simple_compress(x, cfg)
    assert oracle.inspect(data, dtype)["chunks"][0]["mode"] == 2
    got = sa.simple_decompress(data, dtype)
    np.testing.assert_array_equal(bits_view(got), bits_view(x))


def test_int_mult_and_float_quant_vs_oracle(sa, oracle):
    rng = np.random.default_rng(1)
    nums = (rng.integers(-1000, 1000, size=30000) * 8 - 1).astype(np.int64)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_INT_MULT, int_mult_base=8, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    np.testing.assert_array_equal(sa.simple_decompress(data, np.int64), nums)
    f = rng.normal(size=20000).astype(np.float16).astype(np.float32)
    data = oracle.simple_compress(f, oracle.make_config(mode=oracle.MODE_FLOAT_QUANT, float_quant_k=13, delta=oracle.DELTA_NOOP))
    np.testing.assert_array_equal(bits_view(sa.simple_decompress(data, np.float32)), bits_view(f))


@pytest.mark.parametrize("offset_bits", [1, 25, 26, 32, 33, 56, 57, 63, 64])
def test_wide_offsets(sa, oracle, offset_bits):  # recovery.rs:260-293
    nums = np.array([0, 1 << (offset_bits - 1)] * 700, dtype=np.uint64)
    data = oracle.simple_compress(nums, oracle.make_config(level=0, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP))
    np.testing.assert_array_equal(sa.simple_decompress(data, np.uint64), nums)


def test_trivial_and_constant_chunks(sa, oracle):
    for nums in (np.full(100000, 77, dtype=np.uint64), np.arange(100000, dtype=np.uint32) * 3 + 5, np.zeros(3, dtype=np.int16)):
        for order in (0, 1, 2):
            data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=order))
            np.testing.assert_array_equal(sa.simple_decompress(data, nums.dtype), nums)


def test_uniform_random_falls_back_to_one_wide_bin(sa, oracle):
    nums = np.random.default_rng(5).integers(0, 2**63, size=1 << 16, dtype=np.uint64) * 2 + 1
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    v = oracle.inspect(data, np.uint64)["chunks"][0]["vars"][0]
    assert len(v["bins"]) == 1 and v["bins"][0][2] == 64
    np.testing.assert_array_equal(sa.simple_decompress(data, np.uint64), nums)


def test_full_size_chunks_u64(sa, oracle):
    """BASELINE config 2 shape at small scale: 4 chunks x 2^18 u64, classic, consecutive order 1."""
    rng = np.random.default_rng(2)
    nums = np.cumsum(rng.geometric(0.001, size=4 << 18)).astype(np.uint64)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    info = oracle.inspect(data, np.uint64)
    assert len(info["chunks"]) == 4 and len(info["chunks"][0]["vars"][0]["bins"]) > 8
    got = sa.simple_decompress(data, np.uint64)
    np.testing.assert_array_equal(got, nums)
    idx = sa.build_index(data, np.uint64)
    np.testing.assert_array_equal(sa.simple_decompress(data, np.uint64, index=idx), nums)


def test_partial_destination_semantics(sa, oracle):  # pco/src/standalone/simple.rs:185-214
    nums = np.arange(600, dtype=np.int32)
    data = oracle.simple_compress(nums, oracle.make_config(level=0, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP, exact_pages=[300, 300]))
    for m in (0, 1, 256, 299, 300, 301, 556, 600, 601):
        dst = np.full(m, -1, dtype=np.int32)
        prog = sa.simple_decompress_into(data, dst)
        n = min(m, 600)
        assert prog.n_processed == n and prog.finished == (n >= 600), (m, prog)
        np.testing.assert_array_equal(dst[:n], nums[:n])


def test_reference_c_abi(oracle):
    """The three reference functions (pco_c/include/cpcodec_generated.h:33-64) on the decode side."""
    import ctypes as C

    from pcodec_b200 import _lib

    L = _lib.lib()
    nums = _walk(np.int64, 10000, 3)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1), uniform_type=True)
    dst = np.zeros(10000, dtype=np.int64)
    n = C.c_size_t()
    buf = np.frombuffer(data, dtype=np.uint8)
    assert L.pco_standalone_simple_decompress_into(buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), C.c_ubyte(4), dst.ctypes.data_as(C.c_void_p),
                                                   C.c_size_t(10000), C.byref(n)) == 0
    assert n.value == 10000
    np.testing.assert_array_equal(dst, nums)
    # too-small dst_cap -> PcoDecompressionError (pco_c/src/lib.rs:110-112); bad dtype -> PcoInvalidType; wrong dtype -> error
    assert L.pco_standalone_simple_decompress_into(buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), C.c_ubyte(4), dst.ctypes.data_as(C.c_void_p),
                                                   C.c_size_t(9999), C.byref(n)) == 3
    assert L.pco_standalone_simple_decompress_into(buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), C.c_ubyte(77), dst.ctypes.data_as(C.c_void_p),
                                                   C.c_size_t(10000), C.byref(n)) == 1
    assert L.pco_standalone_simple_decompress_into(buf.ctypes.data_as(C.c_void_p), C.c_size_t(len(data)), C.c_ubyte(2), dst.ctypes.data_as(C.c_void_p),
                                                   C.c_size_t(10000), C.byref(n)) == 3
    # a file that lost its terminator byte is an error even when its numbers fill dst_cap exactly (simple_decompress reads to the end:
    # standalone/simple.rs:93-113), and so is a file cut inside its last chunk
    for cut in (1, 40):
        short = np.frombuffer(data[:-cut], dtype=np.uint8)
        assert L.pco_standalone_simple_decompress_into(short.ctypes.data_as(C.c_void_p), C.c_size_t(len(data) - cut), C.c_ubyte(4),
                                                       dst.ctypes.data_as(C.c_void_p), C.c_size_t(10000), C.byref(n)) == 3


@pytest.mark.parametrize("case", ["short_bins", "multi_bin_delta"])
def test_truncation_is_insufficient_data(sa, oracle, case):  # stability.rs:8-34
    from pcodec_b200 import PcoError

    if case == "short_bins":
        nums = np.array([0] * 50 + [1000] * 50, dtype=np.uint32)
        cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP)
    else:
        nums = _walk(np.uint64, 2000, 9)
        cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=2)
    data = oracle.simple_compress(nums, cfg)
    step = 1 if len(data) < 300 else 13
    for i in range(0, len(data) - 1, step):
        with pytest.raises(PcoError) as e:
            sa.simple_decompress(data[:i], nums.dtype)
        assert e.value.kind == "InsufficientData", (i, len(data), e.value)


def test_bit_flips_never_crash_and_match_oracle_verdict(sa, oracle):  # corruption.rs:26-79 (sampled)
    from pcodec_b200 import PcoError

    nums = _walk(np.uint32, 3000, 4)
    data = bytearray(oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1)))
    rng = np.random.default_rng(0)
    for bit in rng.choice(len(data) * 8, size=150, replace=False):
        d = bytearray(data)
        d[bit // 8] ^= 1 << (bit % 8)
        try:
            ours = sa.simple_decompress(bytes(d), np.uint32)
        except PcoError:
            ours = None
        try:
            ref = oracle.simple_decompress(bytes(d), np.uint32)
        except oracle.OracleError:
            ref = None
        if ours is not None and ref is not None:
            np.testing.assert_array_equal(ours, ref)


@pytest.mark.parametrize("dtype", [np.uint64, np.uint32, np.uint16])
def test_chunks_starting_at_odd_element_offsets(sa, oracle, dtype):
    """Multi-chunk files whose chunk sizes are odd: a chunk's destination is then not 16-byte aligned (vector stores must not
    assume it), with and without the side index."""
    rng = np.random.default_rng(3)
    n = 4 * 7685 + 3
    x = np.cumsum(rng.geometric(0.01, size=n)).astype(dtype)
    for order in (0, 1):
        cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=order, max_page_n=7685)
        data = oracle.simple_compress(x, cfg)
        got = sa.simple_decompress(data, dtype)  # no index: device walk, then decode
        assert np.array_equal(bits_view(got), bits_view(x))
        idx = sa.build_index(data, dtype)
        assert np.array_equal(bits_view(sa.simple_decompress(data, dtype, index=idx)), bits_view(x))
        out = np.zeros(n, dtype=dtype)
        prog = sa.simple_decompress_into(data, out)
        assert prog.finished and prog.n_processed == n
        assert np.array_equal(bits_view(out), bits_view(x))


# ---- pco_b200_decompress_chunks: batched decompress at known chunk offsets, no side index ----
def _chunk_table(oracle, data, dtype):
    info = oracle.inspect(data, dtype)
    return [c["chunk_start"] for c in info["chunks"]], [c["n"] for c in info["chunks"]]


@pytest.mark.parametrize("dtype,mode,order", [(np.uint64, "classic", 1), (np.int32, "classic", 0), (np.float64, "float_mult", 2), (np.uint16, "classic", 2)])
def test_decompress_chunks_at_known_offsets(sa, oracle, dtype, mode, order):
    from pcodec_b200 import PcoError

    n = 9 * 3000 + 17
    if mode == "float_mult":
        nums = (np.round(np.cumsum(np.random.default_rng(1).normal(size=n)) * 100) * 0.01).astype(dtype)
        cfg = oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=order, max_page_n=3000)
    else:
        nums = _walk(dtype, n, seed=4)
        cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE if order else oracle.DELTA_NOOP, delta_order=order, max_page_n=3000)
    data = oracle.simple_compress(nums, cfg)
    offs, ns = _chunk_table(oracle, data, dtype)
    assert len(offs) > 5 and sum(ns) == n
    got = sa.decompress_chunks(data, dtype, offs, ns)
    assert np.array_equal(bits_view(got), bits_view(nums))
    # any subset, in any order, lands back to back
    pick = [4, 0, 2]
    got = sa.decompress_chunks(data, dtype, [offs[i] for i in pick], [ns[i] for i in pick])
    starts = np.concatenate([[0], np.cumsum(ns)])
    want = np.concatenate([nums[starts[i]:starts[i + 1]] for i in pick])
    assert np.array_equal(bits_view(got), bits_view(want))
    # bare chunks (no standalone header): the bytes from the first chunk on
    bare = data[offs[0]:]
    got = sa.decompress_chunks(bare, dtype, [o - offs[0] for o in offs], ns)
    assert np.array_equal(bits_view(got), bits_view(nums))
    # a count that contradicts the chunk's own header, an offset outside the buffer, a truncated last chunk
    bad = list(ns)
    bad[1] += 1
    with pytest.raises(PcoError) as e:
        sa.decompress_chunks(data, dtype, offs, bad)
    assert e.value.kind == "InvalidArgument"
    with pytest.raises(PcoError) as e:
        sa.decompress_chunks(data, dtype, offs[:-1] + [len(data) + 5], ns)
    assert e.value.kind == "InvalidArgument"
    with pytest.raises(PcoError) as e:
        sa.decompress_chunks(data[: offs[-1] + 40], dtype, offs, ns)
    assert e.value.kind in ("InsufficientData", "Corruption")


def test_decompress_chunks_full_size(sa, oracle):
    from pcodec_b200 import datagen

    nums = np.concatenate([datagen.c2_u64_cumsum_geometric(seed=s) for s in (21, 22, 23, 24)])
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    offs, ns = _chunk_table(oracle, data, np.uint64)
    assert np.array_equal(sa.decompress_chunks(data, np.uint64, offs, ns), nums)


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.float16])
@pytest.mark.parametrize("order", [0, 1])
def test_float_mult_nan_signs_and_payloads(sa, oracle, dtype, order):
    """FloatMult multiplies a NaN primary by the base (mode/float_mult.rs:17-36): the CPU keeps the operand NaN's sign and payload and
    quiets it; the join has to produce those bits (found by the property test: negative NaNs came back positive)."""
    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PcoError

    dt = np.dtype(dtype)
    u = {8: np.uint64, 4: np.uint32, 2: np.uint16}[dt.itemsize]
    ones = int(np.iinfo(u).max)
    exp_all = {8: 0x7ff0000000000000, 4: 0x7f800000, 2: 0x7c00}[dt.itemsize]
    sign = 1 << (8 * dt.itemsize - 1)
    bits = [ones, ones ^ sign, exp_all | 1, exp_all | 1 | sign, exp_all | (exp_all >> 3), 1, 0, sign, 1 | sign, exp_all, exp_all | sign]
    rng = np.random.default_rng(5)
    body = (rng.integers(-50, 50, size=300) * 0.5).astype(dt).view(u)
    nums = np.concatenate([np.array(bits, dtype=u), body, np.array(bits[::-1], dtype=u)]).view(dt)
    ocfg = oracle.make_config(level=4, mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.5, delta=oracle.DELTA_CONSECUTIVE if order else oracle.DELTA_NOOP, delta_order=order)
    want = oracle.simple_compress(nums, ocfg)
    back = sa.simple_decompress(want, dt)
    np.testing.assert_array_equal(back.view(u), nums.view(u))
    cfg = ChunkConfig(compression_level=4, mode_spec=ModeSpec.try_float_mult(0.5), delta_spec=DeltaSpec.try_consecutive(order) if order else DeltaSpec.no_op())
    try:
        got = sa.simple_compress(nums, cfg)
    except PcoError as e:  # f16 FloatMult is decoded on the device but not encoded
        assert dt.itemsize == 2 and e.kind == "Unsupported", e
    else:
        assert got == want
