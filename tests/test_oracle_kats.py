"""Reproduces the reference's primitive-level known-answer tests on the oracle.

Each test cites the reference unit test whose expected values it copies.
"""
import ctypes as C

import numpy as np
import pytest

from tests.golden_generators import GENERATORS, load_assets


def u32arr(xs):
    return (C.c_uint32 * len(xs))(*xs)


def u64arr(xs):
    return (C.c_uint64 * len(xs))(*xs)


def test_spread_state_symbols(oracle):  # pco/src/ans/spec.rs:95-107
    L = oracle.lib()
    out = (C.c_uint32 * 16)()
    assert L.pco_oracle_kat_spread(4, u32arr([1, 1, 3, 11]), C.c_size_t(4), out) == 0
    assert list(out) == [0, 3, 2, 3, 2, 3, 3, 3, 3, 1, 3, 2, 3, 3, 3, 3]
    out = (C.c_uint32 * 2)()
    assert L.pco_oracle_kat_spread(1, u32arr([2]), C.c_size_t(1), out) == 0
    assert list(out) == [0, 0]
    # weights that do not sum to 2^size_log are a corruption (spec.rs:38-44)
    assert L.pco_oracle_kat_spread(4, u32arr([1, 1, 3, 10]), C.c_size_t(4), (C.c_uint32 * 16)()) == 1


@pytest.mark.parametrize("counts,total,size_log,expected", [
    ([777], 777, 0, [1]),
    ([777, 1], 778, 1, [1, 1]),
    ([777, 1], 778, 2, [3, 1]),
    ([2, 3, 6, 5, 1], 17, 3, [1, 1, 3, 2, 1]),
    ([1, 1], 2, 1, [1, 1]),
])
def test_quantize_weights_to(oracle, counts, total, size_log, expected):  # pco/src/ans/encoding.rs:181-197
    out = (C.c_uint32 * len(counts))()
    assert oracle.lib().pco_oracle_kat_quantize_weights_to(u32arr(counts), C.c_size_t(len(counts)), C.c_size_t(total), size_log, out) == 0
    assert list(out)[: len(expected)] == expected


@pytest.mark.parametrize("counts,total,max_log,exp_log,exp_w", [
    ([77, 100], 177, 4, 4, [7, 9]),
    ([77, 77], 154, 4, 1, [1, 1]),
])
def test_quantize_weights(oracle, counts, total, max_log, exp_log, exp_w):  # pco/src/ans/encoding.rs:199-206
    out = (C.c_uint32 * len(counts))()
    sl = C.c_uint32()
    assert oracle.lib().pco_oracle_kat_quantize_weights(u32arr(counts), C.c_size_t(len(counts)), C.c_size_t(total), max_log, C.byref(sl), out) == 0
    assert sl.value == exp_log and list(out) == exp_w


def _ans_roundtrip(oracle, size_log, state_symbols, weights, symbols):
    n = C.c_size_t()
    rc = oracle.lib().pco_oracle_kat_ans_roundtrip(size_log, u32arr(state_symbols), C.c_size_t(len(state_symbols)), u32arr(weights),
                                                   C.c_size_t(len(weights)), u32arr(symbols), C.c_size_t(len(symbols)), C.byref(n))
    assert rc == 0
    return n.value


def test_ans_encoder_decoder(oracle):  # pco/src/ans/mod.rs:67-115
    ss, w = [0, 1, 2, 0, 1, 2, 0, 1], [3, 3, 2]
    assert _ans_roundtrip(oracle, 3, ss, w, [2, 0, 1, 1, 1, 0, 0, 1, 2]) == 2
    assert _ans_roundtrip(oracle, 3, ss, w, [0, 1, 2] * 200) == 125
    assert _ans_roundtrip(oracle, 3, [0, 0, 0, 0, 0, 0, 0, 1], [7, 1], ([0] * 7 + [1]) * 100) == 50


def test_bit_writer_bytes(oracle):  # pco/src/bit_writer.rs:176-203
    vals = [(1 << 8) + 1, (1 << 16) + (1 << 5), 1 << 1, 1 << 1, (1 << 23) + (1 << 15)]
    bits = [9, 17, 17, 13, 24]
    out, n = C.c_void_p(), C.c_size_t()
    assert oracle.lib().pco_oracle_kat_bit_writer(u64arr(vals), u32arr(bits), C.c_size_t(5), C.byref(out), C.byref(n)) == 0
    data = C.string_at(out, n.value)
    oracle.lib().pco_oracle_free(out)
    assert list(data) == [1, 65, 0, 10, 0, 16, 0, 0, 128, 128]


def test_log2_approx(oracle):  # pco/src/bin_optimization.rs:277-316
    f = oracle.lib().pco_oracle_kat_log2_approx
    for e in range(32):
        assert f(float(1 << e)) == float(e)
    prev = -np.inf
    for i in range(1, 101):
        v = f(float(i))
        assert v >= prev
        assert abs(np.log2(np.float32(i)) - v) < 0.0076
        prev = v


def _hist(oracle, latents, n_bins_log, dtype=np.uint32):
    a = np.asarray(latents, dtype=dtype)
    out = (C.c_uint64 * (3 * max(1, 1 << n_bins_log)))()
    n = C.c_size_t()
    fn = oracle.lib().pco_oracle_kat_histogram_u32 if dtype == np.uint32 else oracle.lib().pco_oracle_kat_histogram_u64
    assert fn(a.ctypes.data_as(C.c_void_p), C.c_size_t(a.size), n_bins_log, out, C.byref(n)) == 0
    return [tuple(out[3 * i: 3 * i + 3]) for i in range(n.value)]


def _hist_sorted(oracle, slices, n, n_bins_log):
    flat = np.asarray([x for s in slices for x in s], dtype=np.uint32)
    lens = (C.c_size_t * len(slices))(*[len(s) for s in slices])
    out = (C.c_uint64 * (3 * (1 << n_bins_log)))()
    inc = (C.c_uint64 * 3)()
    n_, has = C.c_size_t(), C.c_int()
    rc = oracle.lib().pco_oracle_kat_histogram_sorted_u32(flat.ctypes.data_as(C.c_void_p), lens, C.c_size_t(len(slices)), C.c_size_t(n),
                                                          n_bins_log, out, C.byref(n_), inc, C.byref(has))
    assert rc == 0
    bins = [tuple(out[3 * i: 3 * i + 3]) for i in range(n_.value)]
    return bins, (tuple(inc) if has.value else None)


def test_histogram_sorted_simple(oracle):  # pco/src/histograms.rs:359-401
    assert _hist_sorted(oracle, [[8]], 1, 0) == ([(1, 8, 8)], None)
    assert _hist_sorted(oracle, [[1, 2, 3, 4, 5, 6, 7, 8, 9]], 9, 2) == ([(3, 1, 3), (2, 4, 5), (2, 6, 7), (2, 8, 9)], None)
    assert _hist_sorted(oracle, [[8] * 11], 11, 2) == ([(11, 8, 8)], None)
    assert _hist_sorted(oracle, [[0, 0, 0, 1, 2, 2, 2, 2]], 8, 3) == ([(3, 0, 0), (1, 1, 1), (4, 2, 2)], None)
    assert _hist_sorted(oracle, [[0, 0, 1, 2, 2, 2, 2, 2]], 8, 3) == ([(2, 0, 0), (1, 1, 1), (5, 2, 2)], None)


def test_histogram_sorted_complex(oracle):  # pco/src/histograms.rs:403-432
    assert _hist_sorted(oracle, [[1, 2], [3, 4, 5], [6, 7], [8]], 16, 3) == ([(2, 1, 2), (2, 3, 4), (2, 5, 6), (2, 7, 8)], None)
    assert _hist_sorted(oracle, [[1, 2, 3, 3, 3, 3, 3, 3, 3, 4], [5, 5, 5, 5]], 16, 2) == ([(2, 1, 2), (7, 3, 3), (1, 4, 4)], (4, 5, 5))
    assert _hist_sorted(oracle, [[1, 1, 2]], 16, 2) == ([], (3, 1, 2))


def test_histogram_quicksort(oracle):  # pco/src/histograms.rs:434-498 (shuffles via numpy instead of Xoroshiro)
    assert _hist(oracle, [8], 0) == [(1, 8, 8)]
    for seed in range(16):
        rng = np.random.default_rng(seed)
        assert _hist(oracle, rng.permutation(100), 2) == [(25, 0, 24), (25, 25, 49), (25, 50, 74), (25, 75, 99)]
        v = np.zeros(100, dtype=np.uint32); v[0] = 1
        assert _hist(oracle, rng.permutation(v), 2) == [(99, 0, 0), (1, 1, 1)]
        v = np.ones(100, dtype=np.uint32); v[0] = 0
        assert _hist(oracle, rng.permutation(v), 2) == [(1, 0, 0), (99, 1, 1)]
        v = np.full(100, 5, dtype=np.uint32); v[0] = 3; v[1:3] = 7
        p = rng.permutation(v)
        assert _hist(oracle, p, 2) == [(1, 3, 3), (97, 5, 5), (2, 7, 7)]
        assert _hist(oracle, p, 1) == [(98, 3, 5), (2, 7, 7)]
        v = np.full(100, 5, dtype=np.uint32); v[0:2] = 3; v[2] = 7
        assert _hist(oracle, rng.permutation(v), 1) == [(2, 3, 3), (98, 5, 7)]


def _optimize(oracle, bins, ans_size_log):
    flat = [x for b in bins for x in b]
    out = (C.c_uint64 * (5 * len(bins)))()
    n = C.c_size_t()
    assert oracle.lib().pco_oracle_kat_optimize_bins_u32(u64arr(flat), C.c_size_t(len(bins)), ans_size_log, out, C.byref(n)) == 0
    return [tuple(out[5 * i: 5 * i + 5]) for i in range(n.value)]


def test_bin_optimization(oracle):  # pco/src/bin_optimization.rs:215-275
    # (weight, lower, upper, offset_bits, symbol)
    assert _optimize(oracle, [(100, 1, 16), (100, 33, 48), (100, 49, 64), (100, 65, 74), (50, 75, 79)], 10) == [
        (100, 1, 16, 4, 0), (200, 33, 64, 5, 1), (150, 65, 79, 4, 2)]
    assert _optimize(oracle, [(1000, 0, 150), (1000, 200, 200)], 10) == [(1000, 0, 150, 8, 0), (1000, 200, 200, 0, 1)]


def test_consecutive_delta(oracle):  # pco/src/delta/consecutive.rs:57-78 and docs/format.md:236-238
    L = oracle.lib()
    orig = np.array([2, 2, 1, 0xFFFFFFFF, 0], dtype=np.uint32)
    d = orig.copy()
    mom = (C.c_uint32 * 2)()
    assert L.pco_oracle_kat_consecutive_encode_u32(d.ctypes.data_as(C.c_void_p), C.c_size_t(5), C.c_size_t(2), mom) == 0
    to_decode = np.concatenate([d[2:], np.array([1337, 1337], dtype=np.uint32)])
    assert L.pco_oracle_kat_consecutive_decode_u32(mom, C.c_size_t(2), to_decode[:3].ctypes.data_as(C.c_void_p), C.c_size_t(3)) == 0
    part1 = to_decode[:3].copy()
    rest = to_decode[3:].copy()
    assert L.pco_oracle_kat_consecutive_decode_u32(mom, C.c_size_t(2), rest.ctypes.data_as(C.c_void_p), C.c_size_t(2)) == 0
    np.testing.assert_array_equal(np.concatenate([part1, rest]), orig)
    # format.md: moments [1, 2], deltas [0, 10, 0] -> [1, 3, 5, 17, 29]
    mid = 1 << 31
    lat = np.array([(0 + mid) % 2**32, (10 + mid) % 2**32, (0 + mid) % 2**32, 0, 0], dtype=np.uint32)
    mom = u32arr([1, 2])
    assert L.pco_oracle_kat_consecutive_decode_u32(mom, C.c_size_t(2), lat.ctypes.data_as(C.c_void_p), C.c_size_t(5)) == 0
    assert list(lat) == [1, 3, 5, 17, 29]


def test_header_kat(oracle):  # SURVEY Appendix A KAT 1 (standalone/compressor.rs:12-16,85-105)
    nums = np.zeros(1 << 18, dtype=np.uint64)
    cfg = oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP)
    data = oracle.simple_compress(nums, cfg)
    assert data[:12] == bytes([0x70, 0x63, 0x6F, 0x21, 0x03, 0x00, 0x12, 0x00, 0x00, 0x01, 0x04, 0x01])
    assert data[12:16] == bytes([0x02, 0xFF, 0xFF, 0x03])
    data2 = oracle.simple_compress(nums, cfg, uniform_type=True)
    assert data2[5] == 2 and data2[:5] == data[:5] and data2[6:] == data[6:]


@pytest.mark.parametrize("name", ["v1_0_0_u8", "v1_0_0_i8"])
def test_reencode_8bit_assets(oracle, name):
    """Whole-pipeline encode KAT: default config (Auto delta picks Consecutive(1)) reproduces the
    asset written by pco 1.0.0 byte for byte (compatibility.rs:282-303).  ModeSpec::Auto resolves to
    Classic for this input (int_mult::choose_base finds no base), so Classic is passed explicitly."""
    nums = GENERATORS[name]()
    cfg = oracle.make_config(level=8, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_AUTO, enable_8_bit=True)
    assert oracle.simple_compress(nums, cfg) == load_assets()[name]


def test_reencode_uniform_type_chunk_bodies(oracle):
    """v0_4_5_uniform_type.pco (compatibility.rs:200-222): two chunks of u32 written with the default
    config; for n < 10 Auto delta is NoOp and Auto mode is Classic.  Chunk bytes are format-identical
    between wrapped major 3 and 4.1."""
    asset = load_assets()["v0_4_5_uniform_type"]
    info = oracle.inspect(asset, np.uint32)
    cfg = oracle.make_config(level=8, mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_AUTO)
    for chunk, nums in zip(info["chunks"], ([1, 2, 3], [4, 5])):
        ours = oracle.simple_compress(np.array(nums, dtype=np.uint32), cfg)
        oi = oracle.inspect(ours, np.uint32)["chunks"][0]
        assert ours[oi["chunk_start"]: oi["chunk_end"]] == asset[chunk["chunk_start"]: chunk["chunk_end"]]


def test_8bit_requires_opt_in(oracle):  # pco/src/chunk_config.rs:306-311
    cfg = oracle.make_config(enable_8_bit=False)
    with pytest.raises(oracle.OracleError) as e:
        oracle.simple_compress(np.zeros(10, dtype=np.uint8), cfg)
    assert e.value.kind == "InvalidArgument"


def test_file_size_guarantee(oracle):  # pco/src/standalone/guarantee.rs:11-38,53-64
    # header_size(): 4 + 1 + ceil((6+64+8)/8) + 2 = 17
    assert oracle.file_size_guarantee(0, np.int32) == 17 + 1
    n = 1 << 18
    # baseline meta: ceil((4 + 1102 + 4 + 15 + (0+64+7)) / 8) = 150 bytes for u64
    assert oracle.file_size_guarantee(n, np.uint64) == 17 + (4 + 150 + n * 8) + 1


def test_choose_mode_sample(oracle):  # pco/src/sampling.rs:186-201 (Xoroshiro128PlusPlus::seed_from_u64(0) + Floyd's algorithm)
    L = oracle.lib()
    nums = np.array([-float(i) for i in range(150)], dtype=np.float32)
    out = (C.c_uint64 * 64)()
    n_out = C.c_size_t()
    assert L.pco_oracle_kat_mode_sample_indices(C.c_size_t(150), out, C.byref(n_out)) == 0
    idx = list(out)[: n_out.value]
    assert len(set(idx)) == len(idx) == 13  # calc_sample_n(150) = 13, drawn without replacement
    sample = sorted(float(nums[i]) for i in idx if nums[i] != 0.0)  # the test's filter drops zeros
    assert len(sample) == 13
    assert sample[:3] == [-135.0, -131.0, -114.0]
    # calc_sample_n (sampling.rs:141-147)
    for n, want in ((9, 0), (10, 10), (100, 12), (1000010, 25010)):
        big = (C.c_uint64 * max(want, 1))()
        assert L.pco_oracle_kat_mode_sample_indices(C.c_size_t(n), big, C.byref(n_out)) == 0
        assert n_out.value == want
