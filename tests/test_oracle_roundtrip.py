"""Round-trip / metadata assertions on the oracle, following pco/src/tests/recovery.rs and stability.rs."""
import numpy as np
import pytest

from tests.golden_generators import bits_view

DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float32, np.float64, np.float16]


def _data(dtype, n, seed):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        x = np.cumsum(rng.normal(size=n)).astype(dtype)
        if n > 5:
            x[3] = np.nan
            x[4] = -np.inf
            x[5] = -0.0
        return x
    info = np.iinfo(dtype)
    steps = rng.geometric(0.05, size=n).astype(np.int64) - 10
    vals = np.cumsum(steps)
    # wrap into the dtype's range (two's complement truncation of the int64 walk)
    return vals.astype(np.uint64).astype(np.dtype(dtype).str.replace("i", "u")).view(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("delta_order", [0, 1, 2, 7])
def test_recovers_consecutive(oracle, dtype, delta_order):  # recovery.rs:49-84
    for n in (1, 2, 255, 256, 257, 1000, 5000):
        nums = _data(dtype, n, n)
        cfg = oracle.make_config(level=8, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=delta_order)
        data = oracle.simple_compress(nums, cfg)
        got = oracle.simple_decompress(data, dtype)
        np.testing.assert_array_equal(bits_view(got), bits_view(nums))
        assert len(data) <= oracle.file_size_guarantee(n, dtype)


@pytest.mark.parametrize("dtype", [np.uint16, np.uint32, np.int64, np.float32])
def test_recovers_auto_delta_and_lookback(oracle, dtype):
    for n in (5, 100, 3000):
        nums = _data(dtype, n, 7 * n)
        for delta in (oracle.DELTA_AUTO, oracle.DELTA_LOOKBACK):
            cfg = oracle.make_config(level=8, mode=oracle.MODE_CLASSIC, delta=delta)
            got = oracle.simple_decompress(oracle.simple_compress(nums, cfg), dtype)
            np.testing.assert_array_equal(bits_view(got), bits_view(nums))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", [0, 2])
def test_recovers_float_mult(oracle, dtype, order):
    rng = np.random.default_rng(0)
    n = 3000
    x = (np.round(1e4 * np.cos(np.arange(n) / 50.0)) * 0.01).astype(dtype)
    x[::97] = (x[::97] * (1 + 1e-6)).astype(dtype)  # a few off-grid values -> nonzero ULP adjustments
    x[5] = np.nan
    x[6] = np.inf
    x[7] = dtype(1e30)
    cfg = oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=order)
    data = oracle.simple_compress(x, cfg)
    info = oracle.inspect(data, dtype)
    assert info["chunks"][0]["mode"] in (0, 2)  # FloatMult unless the size guarantee forced the Classic fallback
    got = oracle.simple_decompress(data, dtype)
    np.testing.assert_array_equal(bits_view(got), bits_view(x))


def test_recovers_int_mult_and_float_quant(oracle):
    rng = np.random.default_rng(1)
    nums = (rng.integers(-1000, 1000, size=300) * 8 - 1).astype(np.int32)
    cfg = oracle.make_config(mode=oracle.MODE_INT_MULT, int_mult_base=8, delta=oracle.DELTA_NOOP)
    data = oracle.simple_compress(nums, cfg)
    assert oracle.inspect(data, np.int32)["chunks"][0]["mode"] == 1
    np.testing.assert_array_equal(oracle.simple_decompress(data, np.int32), nums)
    f = rng.normal(size=1000).astype(np.float16).astype(np.float32)
    cfg = oracle.make_config(mode=oracle.MODE_FLOAT_QUANT, float_quant_k=13, delta=oracle.DELTA_NOOP)
    data = oracle.simple_compress(f, cfg)
    assert oracle.inspect(data, np.float32)["chunks"][0]["mode"] == 3
    np.testing.assert_array_equal(bits_view(oracle.simple_decompress(data, np.float32)), bits_view(f))


@pytest.mark.parametrize("offset_bits", [56, 57, 64])
def test_wide_offsets(oracle, offset_bits):  # recovery.rs:260-293
    nums = np.array([0, 1 << (offset_bits - 1)] * 50, dtype=np.uint64)
    cfg = oracle.make_config(level=0, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP)
    data = oracle.simple_compress(nums, cfg)
    var = oracle.inspect(data, np.uint64)["chunks"][0]["vars"][0]
    assert len(var["bins"]) == 1 and var["bins"][0][2] == offset_bits
    np.testing.assert_array_equal(oracle.simple_decompress(data, np.uint64), nums)


def test_empty_and_multichunk(oracle):  # recovery.rs:86-114, standalone/simple.rs:185-214
    cfg = oracle.make_config(level=0, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP, exact_pages=[300, 300])
    nums = np.arange(600, dtype=np.int32)
    data = oracle.simple_compress(nums, cfg)
    assert len(oracle.inspect(data, np.int32)["chunks"]) == 2
    for m in (0, 1, 256, 299, 300, 301, 556, 600, 601):
        dst = np.zeros(m, dtype=np.int32)
        n_proc, finished = oracle.simple_decompress_into(data, dst)
        n = min(m, 600)
        assert n_proc == n and finished == (n >= 600)
        np.testing.assert_array_equal(dst[:n], nums[:n])
    empty = oracle.simple_compress(np.zeros(0, dtype=np.uint32), oracle.make_config())
    assert len(empty) == 4 + 1 + 1 + 1 + 2 + 1  # magic, version, type, varint(0)=7 bits, format 4.1, terminator
    assert oracle.simple_decompress(empty, np.uint32).size == 0


@pytest.mark.parametrize("case", ["short_bins", "sparse", "long_offsets"])
def test_truncation_is_insufficient_data(oracle, case):  # stability.rs:8-105
    if case == "short_bins":
        nums, nb = np.array([0] * 50 + [1000] * 50, dtype=np.uint32), 2
    elif case == "sparse":
        nums, nb = np.array([0] + [1] * ((1 << 16) + 1), dtype=np.uint32), 2
    else:
        nums, nb = (np.arange(1000, dtype=np.uint64) * np.uint64((2**64 - 1) // 1000)), 1
    cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP)
    data = oracle.simple_compress(nums, cfg)
    assert len(oracle.inspect(data, nums.dtype)["chunks"][0]["vars"][0]["bins"]) == nb
    step = 1 if len(data) < 400 else 37
    for i in list(range(0, len(data) - 1, step)):
        with pytest.raises(oracle.OracleError) as e:
            oracle.simple_decompress(data[:i], nums.dtype)
        assert e.value.kind == "InsufficientData", (i, e.value)


def test_bit_flips_never_crash(oracle):  # corruption.rs:26-79 (sampled)
    nums = np.cumsum(np.random.default_rng(3).integers(0, 100, size=700)).astype(np.uint32)
    data = bytearray(oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1)))
    for i in range(0, len(data) * 8, 5):
        d = bytearray(data)
        d[i // 8] ^= 1 << (i % 8)
        try:
            oracle.simple_decompress(bytes(d), np.uint32)
        except oracle.OracleError:
            pass


def _assert_recovers(oracle, nums, level):  # recovery.rs:49-84: mode in {Classic, Auto} x the delta specs the oracle encodes
    deltas = [(oracle.DELTA_NOOP, 0), (oracle.DELTA_CONSECUTIVE, 0), (oracle.DELTA_CONSECUTIVE, 1), (oracle.DELTA_CONSECUTIVE, 7), (oracle.DELTA_LOOKBACK, 0),
              (oracle.DELTA_AUTO, 0)]
    if nums.dtype.itemsize <= 4:
        deltas += [(oracle.DELTA_CONV1, 2), (oracle.DELTA_CONV1, 6)]  # 6: the reference has a specialised decode path for it
    modes = [oracle.MODE_CLASSIC, oracle.MODE_AUTO]
    for mode in modes:
        for delta, order in deltas:
            cfg = oracle.make_config(level=level, mode=mode, delta=delta, delta_order=order, enable_8_bit=True)
            data = oracle.simple_compress(nums, cfg)
            got = oracle.simple_decompress(data, nums.dtype)
            np.testing.assert_array_equal(bits_view(got), bits_view(nums), err_msg=f"mode={mode} delta={delta}@{order}")
            assert len(data) <= oracle.file_size_guarantee(nums.size, nums.dtype)


def test_recovery_edge_cases(oracle):  # recovery.rs:86-114
    _assert_recovers(oracle, np.array([0, (1 << 64) - 1], dtype=np.uint64), 0)
    _assert_recovers(oracle, np.array([np.finfo(np.float64).min, np.finfo(np.float64).max]), 0)
    for level in (0, 1, 2):
        _assert_recovers(oracle, np.array([1.2], dtype=np.float32), level)
    for dtype, level in ((np.uint32, 6), (np.uint32, 0), (np.uint16, 6), (np.uint8, 6)):
        _assert_recovers(oracle, np.zeros(0, dtype=dtype), level)
    f16 = np.array([-np.inf, np.finfo(np.float16).min, -1.0, -0.0, np.nan, 0.0, 1.0, np.finfo(np.float16).max, np.inf], dtype=np.float16)
    _assert_recovers(oracle, f16, 5)


def test_recovery_moderate_and_sparse(oracle):  # recovery.rs:116-135, :318-330
    _assert_recovers(oracle, np.arange(-50000, 50000, dtype=np.int32), 3)
    _assert_recovers(oracle, np.array([1] * 10000 + [0, 0, 1], dtype=np.int32), 1)
    rng = np.random.default_rng(0)
    islands = np.concatenate([np.append(rng.integers(0, 8, size=99), rng.integers(1000, 1008)) for _ in range(20)]).astype(np.int32)
    _assert_recovers(oracle, islands, 4)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64, np.int8, np.int16, np.int32, np.int64, np.float16, np.float32, np.float64])
def test_recovery_codecs_with_auto_mode(oracle, dtype):  # recovery.rs:137-242: every type at several levels, extremes included
    dt = np.dtype(dtype)
    if dt.kind == "f":
        fi = np.finfo(dtype)
        nums = np.array([-np.inf, fi.min, -1.0, -0.0, np.nan, 0.0, fi.tiny / 2, 1.0, fi.max, np.inf] * 30, dtype=dtype)
    else:
        ii = np.iinfo(dtype)
        nums = np.array([ii.min, ii.min + 1, -1 if ii.min < 0 else 1, 0, 1, ii.max - 1, ii.max] * 40, dtype=dtype)
    for level in (0, 1, 5, 8):
        _assert_recovers(oracle, nums, level)
    _assert_recovers(oracle, _data(dtype, 3000, 4), 8)
