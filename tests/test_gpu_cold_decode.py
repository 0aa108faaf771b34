"""The catch-all device decoder (pcodec_b200/csrc/decode_cold.cuh): valid pco the fast kernels decline - Dict mode, Lookback and Conv1
deltas, tANS tables beyond 2^10 states / 256 bins (compression levels 9..12), f16 FloatMult - is decoded on the GPU by one thread that
follows the reference's decompressor, never refused and never handed to the CPU.  Streams come from the oracle; the GPU's numbers must be
the oracle's and the original ones (pco/src/tests/recovery.rs is the reference's shape of these cases)."""
import numpy as np
import pytest

from tests.golden_generators import bits_view

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sa():
    from pcodec_b200 import standalone

    return standalone


def _roundtrip(sa, oracle, nums, cfg, expect_kernel="cold_decode_kernel"):
    import ctypes as C

    from pcodec_b200 import _lib

    data = oracle.simple_compress(nums, cfg)
    L = _lib.lib()
    L.pco_b200_profile_enable(1)
    got = sa.simple_decompress(data, nums.dtype)
    buf = C.create_string_buffer(4096)
    L.pco_b200_profile_last(buf, 4096)
    L.pco_b200_profile_enable(0)
    np.testing.assert_array_equal(bits_view(got), bits_view(nums))
    np.testing.assert_array_equal(bits_view(got), bits_view(oracle.simple_decompress(data, nums.dtype)))
    if expect_kernel:
        assert expect_kernel in buf.value.decode(), buf.value.decode()
    return data


@pytest.mark.parametrize("dtype", [np.uint32, np.int64, np.float32, np.uint16])
def test_lookback_delta(sa, oracle, dtype):
    rng = np.random.default_rng(1)
    motif = rng.integers(0, 1 << 14, size=97)
    nums = np.tile(motif, 60)[:5000].astype(dtype)
    nums[::211] = np.asarray(nums[::211] + 3, dtype=dtype)
    _roundtrip(sa, oracle, nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_LOOKBACK, max_page_n=1 << 12))


@pytest.mark.parametrize("dtype,order", [(np.int32, 1), (np.int32, 2), (np.float32, 3), (np.uint16, 2), (np.int16, 6)])
def test_conv1_delta(sa, oracle, dtype, order):
    i = np.arange(6000)
    x = 2000 * np.sin(i / 30.0) + 300 * np.cos(i / 7.0)
    nums = x.astype(dtype)
    _roundtrip(sa, oracle, nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONV1, delta_order=order, max_page_n=1 << 12))


@pytest.mark.parametrize("dtype", [np.uint64, np.float64, np.int32, np.uint8])
def test_dict_mode(sa, oracle, dtype):
    rng = np.random.default_rng(2)
    values = np.array([3, 77, 200, 5, 131], dtype=np.int64)
    nums = values[rng.integers(0, len(values), size=7000)].astype(dtype)
    _roundtrip(sa, oracle, nums, oracle.make_config(mode=oracle.MODE_DICT, delta=oracle.DELTA_NOOP, max_page_n=3000, enable_8_bit=True))


@pytest.mark.parametrize("level", [9, 10, 12])
def test_levels_above_8_decode(sa, oracle, level):
    """ans_size_log > 10 and more than 256 bins: files the reference writes at compression levels 9..12."""
    rng = np.random.default_rng(level)
    nums = (rng.lognormal(10, 2, size=1 << 16)).astype(np.uint64)
    data = _roundtrip(sa, oracle, nums, oracle.make_config(level=level, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP), expect_kernel="cold_decode_kernel" if level >= 12 else None)
    info = oracle.inspect(data, np.uint64)["chunks"][0]
    if level >= 12:
        assert max(len(v["bins"]) for v in info["vars"]) > 256 or max(v["ans_size_log"] for v in info["vars"]) > 10


def test_f16_float_mult(sa, oracle):
    nums = (np.round(np.random.default_rng(4).normal(size=4000) * 40) * 0.25).astype(np.float16)
    _roundtrip(sa, oracle, nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.25, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))


def test_lookback_with_a_two_var_mode_and_partial_destination(sa, oracle):
    from pcodec_b200 import _lib
    import ctypes as C

    rng = np.random.default_rng(6)
    motif = rng.integers(0, 5000, size=50)
    nums = (np.tile(motif, 80) * 8 + rng.integers(0, 3, size=4000)).astype(np.uint32)
    cfg = oracle.make_config(mode=oracle.MODE_INT_MULT, int_mult_base=8, delta=oracle.DELTA_LOOKBACK, max_page_n=1500)
    data = _roundtrip(sa, oracle, nums, cfg)
    # pco::standalone::simple_decompress_into semantics with a destination shorter than the file (standalone/simple.rs:100-143)
    L = _lib.lib()
    for cap in (0, 1, 255, 256, 1499, 1500, 1501, 3999):
        dst = np.zeros(cap + 8, dtype=np.uint32)
        prog = _lib._CProgress()
        rc = L.pco_b200_simple_decompress_into(data, C.c_size_t(len(data)), C.c_ubyte(1), dst.ctypes.data_as(C.c_void_p), C.c_size_t(cap), C.byref(prog))
        assert rc == 0, cap
        assert prog.n_processed == cap and not prog.finished
        np.testing.assert_array_equal(dst[:cap], nums[:cap])
        assert not dst[cap:].any()


def test_corrupt_cold_streams_fail_cleanly(sa, oracle):
    from pcodec_b200 import PcoError

    values = np.array([9, 4, 2], dtype=np.uint64)
    nums = values[np.random.default_rng(8).integers(0, 3, size=3000)]
    data = bytearray(oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_DICT, delta=oracle.DELTA_NOOP)))
    for cut in (len(data) - 1, len(data) // 2, 20):
        with pytest.raises(PcoError) as e:
            sa.simple_decompress(bytes(data[:cut]), np.uint64)
        assert e.value.kind in ("InsufficientData", "Corruption"), (cut, e.value.kind)
    rng = np.random.default_rng(9)
    for _ in range(30):
        bad = bytearray(data)
        bad[int(rng.integers(12, len(bad)))] ^= 1 << int(rng.integers(0, 8))
        try:
            out = sa.simple_decompress(bytes(bad), np.uint64)
            assert out.size <= nums.size + (1 << 24)
        except PcoError as e:
            assert e.kind in ("InsufficientData", "Corruption", "Unsupported", "InvalidArgument"), e.kind
