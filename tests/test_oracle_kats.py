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


def test_int_mult_gcd_helpers(oracle):  # pco/src/mode/int_mult.rs:238-294
    L = oracle.lib()
    L.pco_oracle_kat_calc_gcd_u32.restype = C.c_uint32
    L.pco_oracle_kat_calc_triple_gcd_u32.restype = C.c_uint32
    for x, y, want in ((0, 0, 0), (0, 1, 1), (1, 0, 1), (2, 0, 2), (2, 3, 1), (6, 3, 3), (12, 30, 6)):
        assert L.pco_oracle_kat_calc_gcd_u32(C.c_uint32(x), C.c_uint32(y)) == want
    for triple, want in (((1, 5, 9), 4), ((8, 5, 2), 3), ((3, 3, 3), 0), ((5, 0, 10), 5)):
        assert L.pco_oracle_kat_calc_triple_gcd_u32((C.c_uint32 * 3)(*triple)) == want
    root = C.c_double()
    assert L.pco_oracle_kat_false_position(0, C.byref(root)) == 1 and abs(root.value - 1.0) < 1e-4
    assert L.pco_oracle_kat_false_position(1, C.byref(root)) == 0
    assert L.pco_oracle_kat_false_position(2, C.byref(root)) == 1 and root.value == 0.5


def _candidate_base(oracle, sample):
    a = np.asarray(sample, dtype=np.uint32)
    base, saved = C.c_uint32(), C.c_double()
    ok = oracle.lib().pco_oracle_kat_choose_candidate_base_u32(a.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(a.size), C.byref(base), C.byref(saved))
    return (base.value, saved.value) if ok else None


def test_int_mult_choose_candidate_base(oracle):  # pco/src/mode/int_mult.rs:296-320
    assert _candidate_base(oracle, [0, 4, 8]) is None  # not significant enough
    assert _candidate_base(oracle, [0, 4, 8, 10, 14, 18, 20, 24, 28])[0] == 4
    assert _candidate_base(oracle, [1, 11, 21, 31, 41, 51, 61, 71, 82])[0] == 10  # 2 of 3 triples congruent
    assert _candidate_base(oracle, [1, 11, 22, 31, 41, 51, 61, 71, 82]) is None  # 1 of 3
    # "even just evens can be useful if the signal is strong enough": 200 random evens below 2000 (the reference draws them with
    # rand 0.8's gen_range, which is not restated - same distribution, several seeds instead of its one draw)
    for seed in range(5):
        twos = np.random.default_rng(seed).integers(0, 1000, size=200).astype(np.uint32) * 2
        assert _candidate_base(oracle, twos)[0] == 2, seed


def _choose_base(oracle, latents):
    L = oracle.lib()
    a = np.ascontiguousarray(latents)
    if a.dtype == np.uint32:
        base = C.c_uint32()
        ok = L.pco_oracle_kat_int_mult_choose_base_u32(a.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(a.size), C.byref(base))
    else:
        base = C.c_uint64()
        ok = L.pco_oracle_kat_int_mult_choose_base_u64(a.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(a.size), C.byref(base))
    return base.value if ok else None


def test_int_mult_choose_base(oracle):  # int_mult.rs:216-230 + sampling.rs:105-141; cases after pco/src/tests/recovery.rs + data_types/unsigned.rs
    rng = np.random.default_rng(0)
    n = 20000
    # multiples of 77 with a wide spread of multipliers: int mult pays
    assert _choose_base(oracle, (rng.integers(0, 1 << 20, size=n) * 77).astype(np.uint32)) == 77
    assert _choose_base(oracle, (rng.integers(0, 1 << 40, size=n) * 1000 + 7).astype(np.uint64)) == 1000
    # a few outliers off the lattice do not break it
    lat = (rng.integers(0, 1 << 20, size=n) * 100).astype(np.uint32)
    lat[::200] += 1
    assert _choose_base(oracle, lat) == 100
    # plain uniform data: no gcd stands out
    assert _choose_base(oracle, rng.integers(0, 1 << 30, size=n).astype(np.uint32)) is None
    # multiples of 77 but only a handful of distinct multipliers: classic memorises them (est_bits_saved_per_num's frequent groups)
    assert _choose_base(oracle, (rng.integers(0, 8, size=n) * 77).astype(np.uint32)) is None
    # fewer than MIN_SAMPLE numbers: no sample, classic
    assert _choose_base(oracle, (np.arange(9) * 77).astype(np.uint32)) is None


@pytest.mark.parametrize("dtype", [np.uint16, np.uint32, np.int32, np.uint64, np.int64])
def test_auto_mode_for_ints_round_trips(oracle, dtype):  # data_types/unsigned.rs:28-35: Auto = int mult when choose_base finds a base, else classic
    rng = np.random.default_rng(3)
    n = 5000
    info = np.iinfo(dtype)
    span = min(int(info.max) // 50, 1 << 40)
    mult = (rng.integers(0, span, size=n) * 50).astype(dtype)
    if info.min < 0:
        mult = (mult - dtype(50 * (span // 2 // 50 * 1))).astype(dtype)
    plain = rng.integers(int(info.min), int(info.max), size=n, dtype=dtype)
    for nums, want_mode in ((mult, "IntMult"), (plain, "Classic")):
        cfg = oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_NOOP)
        data = oracle.simple_compress(nums, cfg)
        assert np.array_equal(oracle.simple_decompress(data, dtype), nums)
        if want_mode == "IntMult":
            explicit = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_INT_MULT, int_mult_base=50, delta=oracle.DELTA_NOOP))
            assert data == explicit  # Auto found base 50 and wrote exactly what TryIntMult(50) writes
        else:
            assert data == oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP))


class _Xoroshiro128PlusPlus:
    """Python twin of the oracle's generator (rand_xoshiro 0.6.0), used to redraw the reference's test inputs."""

    M = (1 << 64) - 1

    def __init__(self, seed):
        def splitmix():
            nonlocal seed
            seed = (seed + 0x9E3779B97F4A7C15) & self.M
            z = seed
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
            return z ^ (z >> 31)

        self.s0, self.s1 = splitmix(), splitmix()

    @classmethod
    def _rotl(cls, x, k):
        return ((x << k) | (x >> (64 - k))) & cls.M

    def next_u64(self):
        r = (self._rotl((self.s0 + self.s1) & self.M, 17) + self.s0) & self.M
        self.s1 ^= self.s0
        self.s0 = self._rotl(self.s0, 49) ^ self.s1 ^ ((self.s1 << 21) & self.M)
        self.s1 = self._rotl(self.s1, 28)
        return r


def _gen_range_i32(rng, low, high, upper_half):
    """rand 0.8.5 UniformInt<i32>::sample_single: widening multiply with a rejection zone over one u32 draw."""
    span = (high - low) & 0xFFFFFFFF
    zone = ((span << (32 - span.bit_length())) - 1) & 0xFFFFFFFF
    while True:
        v = rng.next_u64()
        v = (v >> 32) if upper_half else (v & 0xFFFFFFFF)
        m = v * span
        if (m & 0xFFFFFFFF) <= zone:
            return low + (m >> 32)


@pytest.mark.parametrize("upper_half", [False, True])
def test_recovery_with_int_mult(oracle, upper_half):  # pco/src/tests/recovery.rs:294-313: Auto mode + NoOp delta on 300 i32s = 8k - 1 -> IntMult(8)
    # both conventions for the generator's next_u32 reach the asserted mode here (test_recovery_decimals is the case that tells them apart)
    rng = _Xoroshiro128PlusPlus(0)
    nums = np.array([_gen_range_i32(rng, -1000, 1000, upper_half) * 8 - 1 for _ in range(300)], dtype=np.int32)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_NOOP))
    assert data == oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_INT_MULT, int_mult_base=8, delta=oracle.DELTA_NOOP))
    assert np.array_equal(oracle.simple_decompress(data, np.int32), nums)


# ---- ModeSpec::Auto for floats: the reference's unit tests of mode/float_mult.rs, mode/float_quant.rs and data_types/float.rs ----
F32_MAX = float(np.finfo(np.float32).max)
TAU32 = float(np.float32(6.2831855))


def _f32(xs):
    return np.ascontiguousarray(np.asarray(xs, dtype=np.float32))


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _plus_epsilons(a, eps):  # float_mult.rs:393-395 (f32)
    a = np.float32(a)
    bits = int(a.view(np.uint32))
    ordered = bits ^ 0x80000000 if not (bits >> 31) else (~bits) & 0xFFFFFFFF
    ordered = (ordered + eps) & 0xFFFFFFFF
    back = ordered ^ 0x80000000 if (ordered >> 31) else (~ordered) & 0xFFFFFFFF
    return float(np.uint32(back).view(np.float32))


def test_float_helpers(oracle):  # float_mult.rs:403-417, data_types/float.rs:528-547
    L = oracle.lib()
    L.pco_oracle_kat_insignificant_float_to_f32.restype = C.c_float
    L.pco_oracle_kat_insignificant_float_to_f32.argtypes = [C.c_float]
    L.pco_oracle_kat_insignificant_float_to_f64.restype = C.c_double
    L.pco_oracle_kat_insignificant_float_to_f64.argtypes = [C.c_double]
    L.pco_oracle_kat_float_exponent_f32.argtypes = [C.c_float]
    L.pco_oracle_kat_float_exp2_f32.restype = C.c_float
    assert L.pco_oracle_kat_insignificant_float_to_f64(1.0) == 1.0 / (1 << 46)
    assert L.pco_oracle_kat_insignificant_float_to_f32(1.0) == 1.0 / (1 << 17)
    assert L.pco_oracle_kat_insignificant_float_to_f32(32.0) == 1.0 / (1 << 12)
    for x, want in ((1.0, 0), (2.0, 1), (3.3333, 1), (0.3333, -2), (31.0, 4)):
        assert L.pco_oracle_kat_float_exponent_f32(x) == want
    for p, want in ((0, 1.0), (1, 2.0), (-1, 0.5), (2, 4.0)):
        assert L.pco_oracle_kat_float_exp2_f32(p) == want


def test_approx_pair_gcd(oracle):  # float_mult.rs:427-456
    L = oracle.lib()
    L.pco_oracle_kat_approx_pair_gcd_f32.argtypes = [C.c_float, C.c_float, C.POINTER(C.c_float)]
    L.pco_oracle_kat_approx_pair_gcd_f64.argtypes = [C.c_double, C.c_double, C.POINTER(C.c_double)]

    def g32(a, b):
        out = C.c_float()
        return out.value if L.pco_oracle_kat_approx_pair_gcd_f32(a, b, C.byref(out)) else None

    def g64(a, b):
        out = C.c_double()
        return out.value if L.pco_oracle_kat_approx_pair_gcd_f64(a, b, C.byref(out)) else None

    assert g32(0.0, 0.0) is None
    assert g32(1.0, 0.0) is None
    assert g32(1.0, 1.0) is None
    assert g32(1.0, 2.0) == 1.0
    assert g32(6.0, 3.0) == 3.0
    assert g64(10.01, 0.009999999999999787) == 0.009999999999999787
    assert g32(2.0**100, 3.0) is None
    mfs = float(np.float32(F32_MAX) * np.float32(0.5))
    assert g32(mfs, float(np.float32(mfs) * np.float32(0.6))) == float(np.float32(mfs) * np.float32(0.2))
    assert g32(mfs, 0.0000000000001) is None
    got = np.float32(g32(float(np.float32(1.0) / np.float32(3.0)), 0.25))
    want = np.float32(1.0) / np.float32(12.0)
    assert abs(int(got.view(np.uint32)) - int(want.view(np.uint32))) <= 1


def test_float_mult_config_candidates(oracle):  # float_mult.rs:419-425, :458-505
    L = oracle.lib()
    base, inv = C.c_float(), C.c_float()
    s = _f32([0.0, 3.0, 6.0, 21.0, 2.0**100] * 5)
    assert L.pco_oracle_kat_config_by_trailing_zeros_f32(_fptr(s), C.c_size_t(s.size), C.byref(base), C.byref(inv)) == 1
    assert base.value == 3.0 and np.float32(inv.value) == np.float32(1.0) / np.float32(3.0)
    s = _f32([0.0, 2.0**-100, 0.0037, 1.0001] * 5 + [F32_MAX])
    assert L.pco_oracle_kat_config_by_euclidean_f32(_fptr(s), C.c_size_t(s.size), C.byref(base), C.byref(inv)) == 1
    assert abs(base.value - 1.0e-4) <= 1.0e-6
    out = C.c_float()
    s = _f32([0.0, 2.0**-100, 0.0037, 1.0001, F32_MAX] * 5)
    assert L.pco_oracle_kat_sample_gcd_euclidean_f32(_fptr(s), C.c_size_t(s.size), C.byref(out)) == 1 and abs(out.value - 1.0e-4) <= 1.0e-6
    s = _f32([0.0, 2.0**-100, 0.0037, 0.0049, 1.0001, F32_MAX] * 5)
    assert L.pco_oracle_kat_sample_gcd_euclidean_f32(_fptr(s), C.c_size_t(s.size), C.byref(out)) == 1 and abs(out.value - 1.0e-4) <= 1.0e-9
    # 25 uniform draws from [0, 1) have no common gcd (the reference draws them with rand's gen_range - same distribution here)
    for seed in range(5):
        s = np.random.default_rng(seed).random(25, dtype=np.float32)
        assert L.pco_oracle_kat_sample_gcd_euclidean_f32(_fptr(s), C.c_size_t(s.size), C.byref(out)) == 0, seed
    # test_center_gcd
    L.pco_oracle_kat_center_sample_base_f32.restype = C.c_float
    L.pco_oracle_kat_center_sample_base_f32.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_size_t]
    s = _f32([6.0 / 7.0 - 1e-4, 16.0 / 7.0 + 1e-4, 18.0 / 7.0 - 1e-4])
    assert abs(L.pco_oracle_kat_center_sample_base_f32(0.28, _fptr(s), s.size) - 2.0 / 7.0) <= 1e-4


def test_float_mult_snap(oracle):  # float_mult.rs:507-538
    L = oracle.lib()
    L.pco_oracle_kat_snap_to_int_reciprocal_f32.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]

    def snap(x):
        b, i = C.c_float(), C.c_float()
        L.pco_oracle_kat_snap_to_int_reciprocal_f32(x, C.byref(b), C.byref(i))
        return np.float32(b.value), np.float32(i.value)

    one = np.float32(1.0)
    assert snap(0.01000333) == (np.float32(0.01), np.float32(100.0))
    assert snap(0.009999666) == (np.float32(0.01), np.float32(100.0))
    assert snap(0.143) == (one / np.float32(7.0), np.float32(7.0))
    assert snap(0.0105) == (np.float32(0.0105), one / np.float32(0.0105))
    assert snap(TAU32)[0] == np.float32(TAU32)


def _fm_better(oracle, inv_base, sample):  # float_mult.rs:397-401
    s = _f32(sample)
    out = C.c_double()
    L = oracle.lib()
    L.pco_oracle_kat_float_mult_bits_saved_f32.argtypes = [C.c_float, C.POINTER(C.c_float), C.c_size_t, C.POINTER(C.c_double)]
    return bool(L.pco_oracle_kat_float_mult_bits_saved_f32(inv_base, _fptr(s), s.size, C.byref(out))) and out.value >= 0.5


def test_float_mult_better_or_worse_than_classic(oracle):  # float_mult.rs:540-604
    nums = [-np.inf, -np.nan, -999.0, -0.3, 0.0, 0.1, 0.2, 0.3, 0.3, 0.4, 0.5, 0.6, 0.7, np.nan, np.inf]
    assert _fm_better(oracle, 10.0, nums)
    for n in (10, 1000):
        assert _fm_better(oracle, 10.0, [_plus_epsilons(np.float32(x) * np.float32(0.1), x % 2) for x in range(n)]), n
        assert not _fm_better(oracle, 10.0, [0.1] * n), n
        assert not _fm_better(oracle, 10.0, [(np.float32(x) + np.float32(1.0)) * np.float32(TAU32) for x in range(n)]), n
        # at this magnitude each increment of base is only ~2 bits
        assert not _fm_better(oracle, 10.0, [np.float32(x + 5_000_000) * np.float32(0.1) for x in range(n)]), n
    for seed in range(3):  # test_float_mult_worse_than_classic_zeros: 1000 zeros + 1000 uniform [0, 1) against base 1e-7
        nums = np.concatenate([np.zeros(1000, np.float32), np.random.default_rng(seed).random(1000, dtype=np.float32)])
        assert not _fm_better(oracle, 1e7, nums), seed


def test_float_mult_compute_bid(oracle):  # float_mult.rs:606-653
    L = oracle.lib()

    def bid(sample):
        s = _f32(sample)
        base, saved = C.c_float(), C.c_double()
        return np.float32(base.value) if L.pco_oracle_kat_float_mult_compute_bid_f32(_fptr(s), C.c_size_t(s.size), C.byref(base), C.byref(saved)) else None

    sevenths = [np.float32((i % 50) - 20) * (np.float32(1.0) / np.float32(7.0)) for i in range(1000)]
    ones = [1.0] * 1000
    noisy_decimals = [_plus_epsilons(np.float32(0.1) * np.float32(i - 100), -7 + i % 15) for i in range(1000)]
    junk = np.sin(np.arange(1000, dtype=np.float32)).astype(np.float32)
    assert bid(sevenths[:50]) == np.float32(1.0) / np.float32(7.0)
    assert bid(sevenths) is None  # (not in the reference's test) 50 multipliers x 20 occurrences each: classic memorises them
    assert bid(noisy_decimals) == np.float32(1.0) / np.float32(10.0)
    bid([F32_MAX] * 10 + [float(np.float32(F32_MAX) * np.float32(0.6))] * 10)  # just check this terminates
    assert bid(ones) is None  # not enough distinct mults
    assert bid(junk) is None


def test_float_quant_estimates(oracle):  # float_quant.rs:155-173, :263-289
    L = oracle.lib()
    k, saved = C.c_uint32(), C.c_double()
    # all but the last of these have 21 of 23 mantissa bits zeroed
    s = _f32([1.0, 1.25, -1.5, 1.75, -0.875, 0.75, 0.625] * 3 + [float(np.uint32(0x3F800001).view(np.float32))])
    L.pco_oracle_kat_float_quant_best_k_f32(_fptr(s), C.c_size_t(s.size), C.byref(k), C.byref(saved))
    assert k.value == 21 and 10.0 < saved.value < 21.0
    d = np.ones(20, dtype=np.float64)
    L.pco_oracle_kat_float_quant_best_k_f64(d.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(d.size), C.byref(k), C.byref(saved))
    assert k.value == 52 and saved.value == 52.0

    def qbid(sample):
        s = _f32(sample)
        ok = L.pco_oracle_kat_float_quant_compute_bid_f32(_fptr(s), C.c_size_t(s.size), C.byref(k), C.byref(saved))
        return (k.value, saved.value) if ok else None

    # the larger numbers in this sample have 23 - 6 = 17 bits of quantization
    assert qbid(np.arange(100)) == (17, 17.0)
    s = np.arange(100, dtype=np.float32)
    s[0] += np.float32(0.1)
    s[37] -= np.float32(0.1)
    kk, sv = qbid(s)
    assert kk == 17 and 15.0 < sv < 17.0
    assert qbid([0.0, 1.0] * 50) is None  # too few primary values: memorizable


def _choose_float_mode(oracle, nums):
    L = oracle.lib()
    kind, base, k = C.c_int(), C.c_double(), C.c_uint32()
    a = np.ascontiguousarray(nums)
    if a.dtype == np.float32:
        L.pco_oracle_kat_choose_float_mode_f32(_fptr(a), C.c_size_t(a.size), C.byref(kind), C.byref(base), C.byref(k))
    else:
        L.pco_oracle_kat_choose_float_mode_f64(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_size_t(a.size), C.byref(kind), C.byref(base), C.byref(k))
    return ("Classic", "IntMult", "FloatMult", "FloatQuant", "Dict")[kind.value], base.value, k.value


def test_choose_float_mode(oracle):  # data_types/float.rs:459-466, :511-521; tests/recovery.rs:388-402
    assert _choose_float_mode(oracle, np.arange(2000, dtype=np.float64) * 1.5)[:2] == ("FloatMult", 1.5)
    lowest = int(np.float64(1.0).view(np.uint64))
    nums = (np.uint64(lowest) + (np.arange(1000, dtype=np.uint64) << np.uint64(20))).view(np.float64)
    mode = _choose_float_mode(oracle, nums)
    assert (mode[0], mode[2]) == ("FloatQuant", 20)
    trivial = np.arange(100, dtype=np.float32)
    trivial[77] += np.float32(0.0001)
    assert _choose_float_mode(oracle, trivial)[:2] == ("FloatMult", 1.0)
    assert _choose_float_mode(oracle, np.random.default_rng(0).standard_normal(5000))[0] == "Classic"
    assert _choose_float_mode(oracle, np.arange(9, dtype=np.float64) * 1.5)[0] == "Classic"  # fewer than MIN_SAMPLE numbers


def _f64_plus_epsilons(a, eps):  # tests/recovery.rs:338-340
    bits = int(np.float64(a).view(np.uint64))
    m = (1 << 64) - 1
    ordered = bits ^ (1 << 63) if not (bits >> 63) else (~bits) & m
    ordered = (ordered + eps) & m
    back = ordered ^ (1 << 63) if (ordered >> 63) else (~ordered) & m
    return float(np.uint64(back).view(np.float64))


def test_recovery_decimals(oracle):  # tests/recovery.rs:332-358: ChunkConfig::default() on noisy hundredths -> FloatMult(1/100), small file
    # This case also pins the generator's next_u32: with the upper half of next_u64 the redrawn input gives the mode the reference
    # asserts; with the lower half the 10-element sample it leaves is a different one and the search lands on Classic.
    upper_half = True
    rng = _Xoroshiro128PlusPlus(0)
    n = 300
    nums = []
    for _ in range(n):
        unadjusted = float(_gen_range_i32(rng, -1, 100, upper_half)) * 0.01
        nums.append(_f64_plus_epsilons(unadjusted, _gen_range_i32(rng, -1, 2, upper_half)))
    nums = np.array(nums + [np.inf] * n, dtype=np.float64)
    assert _choose_float_mode(oracle, nums)[:2] == ("FloatMult", 1.0 / 100.0)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO))
    assert np.array_equal(oracle.simple_decompress(data, np.float64).view(np.uint64), nums.view(np.uint64))
    overhead_bytes = 100
    assert len(data) < (9 * n + 3 * n) // 8 + overhead_bytes


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_auto_mode_for_floats_round_trips(oracle, dtype):  # data_types/float.rs:82-98 through the whole chunk compressor
    rng = np.random.default_rng(11)
    n = 6000
    u = np.dtype(f"u{np.dtype(dtype).itemsize}")
    k = 9 if dtype == np.float32 else 30
    cases = {
        "FloatMult": (rng.integers(-20000, 20000, size=n).astype(dtype) * dtype(1.5)).astype(dtype),
        "FloatQuant": ((rng.standard_normal(n).astype(dtype).view(u) >> u.type(k)) << u.type(k)).view(dtype),
        "Classic": rng.standard_normal(n).astype(dtype),
    }
    for want, nums in cases.items():
        kind, base, kk = _choose_float_mode(oracle, nums)
        assert kind == want, (want, kind, base, kk)
        data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_NOOP))
        assert np.array_equal(oracle.simple_decompress(data, dtype).view(u), nums.view(u)), want
        if want == "FloatMult":
            assert base == 1.5
            assert data == oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=1.5, delta=oracle.DELTA_NOOP))
        elif want == "FloatQuant":
            assert kk == k
            assert data == oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_QUANT, float_quant_k=k, delta=oracle.DELTA_NOOP))
        else:
            assert data == oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_NOOP))
    # decimals: the snapped config has inv_base = 100 exactly and base = 1/100
    nums = (rng.integers(-5000, 5000, size=n).astype(np.float64) / 100.0).astype(dtype)
    kind, base, _ = _choose_float_mode(oracle, nums)
    assert kind == "FloatMult" and dtype(base) == dtype(1.0) / dtype(100.0)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO))
    assert np.array_equal(oracle.simple_decompress(data, dtype).view(u), nums.view(u))
    assert len(data) < n * 2 + 200  # ~13.3 bits per number for the multiplier, adjustments almost free


def test_conv1_matrix_kats(oracle):  # pco/src/delta/conv1.rs:503-583
    L = oracle.lib()
    dp = C.POINTER(C.c_double)

    def arr(x):
        return np.ascontiguousarray(np.asarray(x, dtype=np.float64))

    v = arr([1.0, 2.0, -1.0, 5.0, -3.0])
    xtx, xty = np.zeros(9), np.zeros(3)
    L.pco_oracle_kat_conv_autocov_mats(v.ctypes.data_as(dp), C.c_size_t(5), C.c_size_t(2), C.c_double(0.7), xtx.ctypes.data_as(dp), xty.ctypes.data_as(dp))
    assert xtx.tolist() == [6.7, -5.0, 2.0, -5.0, 30.7, 6.0, 2.0, 6.0, 3.7]  # exact, as the reference asserts
    assert xty.tolist() == [12.0, -22.0, 1.0]
    a = arr([[0.01, -0.2, -0.4], [-0.2, 13.0, 23.0], [-0.4, 23.0, 77.0]])
    want = arr([[0.1, 0.0, 0.0], [-2.0, 3.0, 0.0], [-4.0, 5.0, 6.0]])
    out = np.zeros(9)
    L.pco_oracle_kat_conv_cholesky(a.ctypes.data_as(dp), C.c_size_t(3), out.ctypes.data_as(dp))
    assert out.tolist() == want.T.reshape(-1).tolist()  # column-major data, exact
    low, y, x = arr([[2.0, 0.0], [3.0, -4.0]]), arr([1.0, 2.0]), np.zeros(2)
    L.pco_oracle_kat_conv_sub(0, low.ctypes.data_as(dp), C.c_size_t(2), y.ctypes.data_as(dp), x.ctypes.data_as(dp))
    assert np.allclose(x, [0.5, -0.125], atol=1e-6)
    L.pco_oracle_kat_conv_sub(1, low.ctypes.data_as(dp), C.c_size_t(2), y.ctypes.data_as(dp), x.ctypes.data_as(dp))
    assert np.allclose(x, [1.25, -0.5], atol=1e-6)


def test_conv1_asset_reencodes_byte_for_byte(oracle):
    """pco/assets/v1_0_0_conv1.pco (2000 i32, ChunkConfig::default() + TryConv1(2), pco/src/tests/compatibility.rs:261-279) written by
    pco 1.0.0.  Between 1.0.0 and 1.0.3 the fit gained an L2 term (L2_REGULARIZATION = 0.1) and lost one bit of quantization (the
    `- 1` at delta/conv1.rs:403): with those two parameters set back, the oracle reproduces the asset's 1668 bytes exactly - the Auto
    mode search, the f64 least-squares fit (bias equal in all 17 digits), the residuals, histogram, bin optimisation, tANS and
    bit packing are then all pinned against real output of the Rust crate.  With 1.0.3's parameters the file differs as expected."""
    from tests.golden_generators import GENERATORS, load_assets

    asset, nums = load_assets()["v1_0_0_conv1"], GENERATORS["v1_0_0_conv1"]()
    cfg = oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_CONV1, delta_order=2)
    L = oracle.lib()
    try:
        L.pco_oracle_kat_conv1_v1_0_0_parameters(1)
        assert oracle.simple_compress(nums, cfg) == asset
    finally:
        L.pco_oracle_kat_conv1_v1_0_0_parameters(0)
    current = oracle.simple_compress(nums, cfg)
    assert current != asset and abs(len(current) - len(asset)) < 16
    assert np.array_equal(oracle.simple_decompress(current, np.int32), nums)
    info = oracle.inspect(current, np.int32)["chunks"][0]
    assert info["delta"] == 3  # Conv1


def test_conv1_recovery_cases(oracle):  # pco/src/tests/recovery.rs:454-540
    x0, x1, x2 = 31, 77, -54
    nums = [x0, x1, x2]
    for _ in range(2000):
        x = x2 - x1 + int(np.float32(0.99) * np.float32(x0)) + 3
        nums.append(x)
        x0, x1, x2 = x1, x2, x
    nums = np.array(nums, dtype=np.int32)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_CONV1, delta_order=3))
    assert oracle.inspect(data, np.int32)["chunks"][0]["delta"] == 3  # test_conv1_nominal: compressed with conv1
    assert np.array_equal(oracle.simple_decompress(data, np.int32), nums)
    # test_conv1_degenerate
    for arr in (np.array([3], dtype=np.uint16), np.zeros(100, dtype=np.uint32), np.random.default_rng(0).integers(0, 1000, size=1000).astype(np.uint32)):
        data = oracle.simple_compress(arr, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_CONV1, delta_order=2))
        assert np.array_equal(oracle.simple_decompress(data, arr.dtype), arr)
    # test_conv1_actually_applied
    i = np.arange(1000, dtype=np.int64)
    parabola = np.maximum(998 * 998 - i * i, 0).astype(np.uint32)
    for order in (3, 6):
        data = oracle.simple_compress(parabola, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_CONV1, delta_order=order))
        from pcodec_b200 import inspect as insp

        c = insp.inspect(data)["chunk"][0]
        assert c["delta_encoding"].startswith("Conv1(") and c["delta_encoding"].count(",") == 1 + order  # quantization, bias, `order` weights
        assert np.array_equal(oracle.simple_decompress(data, np.uint32), parabola)
    # 64-bit latents are refused (delta/mod.rs:53-59, chunk_config.rs:285-295)
    with pytest.raises(oracle.OracleError):
        oracle.simple_compress(np.arange(100, dtype=np.uint64), oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONV1, delta_order=2))


@pytest.mark.parametrize("name,kw", [
    ("v0_0_0_classic", dict(mode="MODE_AUTO", delta="DELTA_NOOP")),               # compatibility.rs:70-82: 2000 i32, 2751-byte page
    ("v0_3_0_f16", dict(mode="MODE_AUTO", delta="DELTA_AUTO")),                   # :145-155: 2000 f16, ChunkConfig::default()
    ("v0_4_0_lookback_delta", dict(mode="MODE_AUTO", delta="DELTA_LOOKBACK")),   # :182-197: lookback choice + encode
    ("v0_4_8_minor_version", dict(mode="MODE_AUTO", delta="DELTA_AUTO")),        # :225-245
    ("v0_3_0_float_quant", dict(mode="MODE_FLOAT_QUANT", float_quant_k=13, delta="DELTA_AUTO")),  # :157-178: two latent vars, 2905-byte page
])
def test_older_assets_reencode_to_the_same_bins_and_page(oracle, name, kw):
    """Assets written by older pco versions carry older header / metadata layouts, so whole files cannot match - but their bins and
    their PAGE bytes (page meta + every batch) do: the planner and the encoder of 1.0.3, restated here, still produce exactly what
    those versions wrote."""
    from pcodec_b200 import inspect as insp
    from tests.golden_generators import GENERATORS, load_assets

    asset, nums = load_assets()[name], GENERATORS[name]()
    data = oracle.simple_compress(nums, oracle.make_config(**{k: getattr(oracle, v) if isinstance(v, str) else v for k, v in kw.items()}))

    def first_chunk(buf):
        c = insp.inspect(buf)["chunk"][0]
        start = c["byte_offset"] + c["meta_size"]
        return c, buf[start:start + c["page_size"]]

    ca, page_a = first_chunk(asset)
    cd, page_d = first_chunk(data)
    assert (ca["mode"], ca["delta_encoding"]) == (cd["mode"], cd["delta_encoding"])
    assert {k: (v["ans_size_log"], v["bins"]) for k, v in ca["latent_var"].items()} == {k: (v["ans_size_log"], v["bins"]) for k, v in cd["latent_var"].items()}
    assert page_a == page_d and len(page_a) > 0


def test_f16_arithmetic_and_mode_search(oracle):  # data_types/float.rs:254-366 (half 2.7.1), tests/recovery.rs:360-374, compatibility.rs:145-155
    L = oracle.lib()
    L.pco_oracle_kat_f64_to_f16_bits.restype = C.c_uint16
    L.pco_oracle_kat_f64_to_f16_bits.argtypes = [C.c_double]
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-9, 6, size=2000), [0.0, -0.0, 65504.0, 65519.99, 65520.0, 1e9, -1e9, 2.0**-24, 2.0**-25, 1.5 * 2.0**-25,
                                                                                         2.0**-14, 2.0**-14 - 2.0**-25, np.inf, -np.inf, 0.1, 100.0, 0.01]])
    for x in xs:  # numpy converts double -> half directly, rounding once to nearest even
        with np.errstate(over="ignore"):
            want = int(np.float64(x).astype(np.float16).view(np.uint16))
        assert L.pco_oracle_kat_f64_to_f16_bits(float(x)) == want, x

    def mode(nums):
        a = np.ascontiguousarray(nums, dtype=np.float16)
        kind, base, k = C.c_int(), C.c_double(), C.c_uint32()
        L.pco_oracle_kat_choose_float_mode_f16(a.view(np.uint16).ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(a.size), C.byref(kind), C.byref(base), C.byref(k))
        return ("Classic", "IntMult", "FloatMult", "FloatQuant", "Dict")[kind.value], base.value, k.value

    # the f16 golden asset was written with ChunkConfig::default(): Auto mode settled on Classic there
    from tests.golden_generators import GENERATORS, load_assets

    assert mode(GENERATORS["v0_3_0_f16"]())[0] == "Classic"
    assert mode(np.arange(300) % 200)[0] in ("FloatMult", "FloatQuant")  # small integers leave mantissa bits unused: a mult or a quant mode, not Classic
    # whole files with the default config: round trip, and the asset's page is reproduced from Auto / Auto
    nums = GENERATORS["v0_3_0_f16"]()
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO))
    assert np.array_equal(oracle.simple_decompress(data, np.float16).view(np.uint16), nums.view(np.uint16))
    from pcodec_b200 import inspect as insp

    def page(buf):
        c = insp.inspect(buf)["chunk"][0]
        s0 = c["byte_offset"] + c["meta_size"]
        return buf[s0:s0 + c["page_size"]]

    assert page(data) == page(load_assets()["v0_3_0_f16"])
    # test_f16_mult: explicit TryFloatMult(100) on f16
    nums = np.array([100.1, 299.9, 200.0] * 100, dtype=np.float16)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=100.0, delta=oracle.DELTA_AUTO))
    assert insp.inspect(data)["chunk"][0]["mode"] == "FloatMult(100.0)"
    assert np.array_equal(oracle.simple_decompress(data, np.float16).view(np.uint16), nums.view(np.uint16))
    for arr in (rng.integers(-500, 500, size=3000) * 0.5, rng.standard_normal(3000), rng.integers(0, 60, size=3000) / 8.0):
        arr = arr.astype(np.float16)
        data = oracle.simple_compress(arr, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO))
        assert np.array_equal(oracle.simple_decompress(data, np.float16).view(np.uint16), arr.view(np.uint16))


def test_dict_mode_encode(oracle):  # pco/src/mode/dict.rs:10-68, tests/recovery.rs:426-451, compatibility.rs:247-259
    from pcodec_b200 import inspect as insp
    from tests.golden_generators import GENERATORS, load_assets

    # test_dict: 2000 distinct squares, five of each
    nums = np.repeat(np.arange(2000, dtype=np.int64) ** 2, 5)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_DICT, delta=oracle.DELTA_NOOP))
    c = insp.inspect(data)["chunk"][0]
    assert c["mode"] == "Dict(2000 values)" and set(c["latent_var"]) == {"primary"} and c["latent_var"]["primary"]["latent_type"] == "U32"
    assert np.array_equal(oracle.simple_decompress(data, np.int64), nums)
    # the golden asset: three u64 values x 1000, equally frequent - the dictionary order is whatever the writer's HashMap gave,
    # so the asset's own order is passed as the tie order; then the whole 664-byte file is reproduced
    asset, anums = load_assets()["v1_0_0_dict"], GENERATORS["v1_0_0_dict"]()
    meta, _ = insp.read_chunk_meta(asset, insp.inspect(asset)["chunk"][0]["byte_offset"] + 4, 64)
    order = (C.c_uint64 * 3)(*meta["mode"]["dict"])
    L = oracle.lib()
    try:
        L.pco_oracle_kat_dict_tie_order(order, C.c_size_t(3))
        assert oracle.simple_compress(anums, oracle.make_config(mode=oracle.MODE_DICT, delta=oracle.DELTA_NOOP)) == asset
    finally:
        L.pco_oracle_kat_dict_tie_order(order, C.c_size_t(0))
    # other types, deltas on the indices, the fallback for incompressible input, several chunks
    rng = np.random.default_rng(2)
    for dtype in (np.uint8, np.int16, np.uint32, np.float32, np.float64):
        vals = rng.integers(0, 40, size=5000).astype(dtype) * dtype(3)
        for delta, order_ in ((oracle.DELTA_NOOP, 0), (oracle.DELTA_CONSECUTIVE, 1), (oracle.DELTA_AUTO, 0), (oracle.DELTA_LOOKBACK, 0)):
            data = oracle.simple_compress(vals, oracle.make_config(mode=oracle.MODE_DICT, delta=delta, delta_order=order_, max_page_n=2000))
            assert np.array_equal(oracle.simple_decompress(data, dtype).view(np.uint8), vals.view(np.uint8)), (dtype, delta)
            if np.dtype(dtype).itemsize >= 4:  # narrow types can hit the worst-case size fallback (Classic), which is legitimate
                assert all(ch["mode"].startswith("Dict(") for ch in insp.inspect(data)["chunk"]), (dtype, delta)
    noise = rng.integers(0, 1 << 62, size=3000).astype(np.uint64)  # every value distinct: the dictionary costs more than it saves
    data = oracle.simple_compress(noise, oracle.make_config(mode=oracle.MODE_DICT, delta=oracle.DELTA_NOOP))
    assert insp.inspect(data)["chunk"][0]["mode"] == "Classic" and len(data) <= oracle.file_size_guarantee(noise.size, np.uint64)
    assert np.array_equal(oracle.simple_decompress(data, np.uint64), noise)


def test_minimal_file_of_the_reference_docs(oracle):  # pco/src/standalone/decompressor.rs:45: "the minimal .pco file"
    from pcodec_b200 import inspect as insp

    minimal = bytes([112, 99, 111, 33, 0, 0])
    assert oracle.simple_decompress(minimal, np.int64).size == 0
    dst = np.zeros(256, dtype=np.int64)
    assert oracle.simple_decompress_into(minimal, dst) == (0, True)
    s = insp.inspect(minimal)
    assert (s["n"], s["n_chunks"], s["compressed"]["total_size"], s["compressed"]["unknown_trailing_bytes"]) == (0, 0, 6, 0)
