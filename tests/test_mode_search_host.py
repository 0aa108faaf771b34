"""pco_b200_choose_mode (pcodec_b200/csrc/mode_search.hpp, host-only planner logic of libcpcodec.so: what ModeSpec::Auto resolves
to for one chunk) against the oracle's restatement of the same search and the cases the reference's own tests assert
(pco/src/tests/recovery.rs:294-313, :332-358, :388-402; pco/src/data_types/float.rs:459-466, :511-521).  No device needed."""
import numpy as np
import pytest

from tests.test_oracle_kats import _choose_base, _choose_float_mode, _f64_plus_epsilons, _gen_range_i32, _Xoroshiro128PlusPlus

KINDS = {1: "Classic", 2: "FloatMult", 3: "FloatQuant", 4: "IntMult"}


@pytest.fixture(scope="module")
def sa():
    from pcodec_b200 import standalone

    return standalone


def _ordered(nums):
    a = np.ascontiguousarray(nums)
    u = a.view(np.dtype(f"u{a.dtype.itemsize}"))
    return u ^ u.dtype.type(1 << (8 * a.dtype.itemsize - 1)) if a.dtype.kind == "i" else u


def _agree(oracle, sa, nums):
    got = sa.choose_mode(nums)
    a = np.ascontiguousarray(nums)
    if a.dtype.kind == "f":
        kind, base, k = _choose_float_mode(oracle, a)
        assert KINDS[got.kind] == kind, (KINDS[got.kind], kind)
        if kind == "FloatMult":
            assert got.base == base
        if kind == "FloatQuant":
            assert got.k == k
    else:
        base = _choose_base(oracle, _ordered(a))
        assert (got.int_base if got.kind == 4 else None) == base
    return got


def test_reference_recovery_cases(oracle, sa):
    rng = _Xoroshiro128PlusPlus(0)
    nums = np.array([_gen_range_i32(rng, -1000, 1000, True) * 8 - 1 for _ in range(300)], dtype=np.int32)
    got = _agree(oracle, sa, nums)
    assert (got.kind, got.int_base) == (4, 8)  # recovery.rs:294-313
    rng = _Xoroshiro128PlusPlus(0)
    dec = []
    for _ in range(300):
        unadjusted = float(_gen_range_i32(rng, -1, 100, True)) * 0.01
        dec.append(_f64_plus_epsilons(unadjusted, _gen_range_i32(rng, -1, 2, True)))
    got = _agree(oracle, sa, np.array(dec + [np.inf] * 300))
    assert (got.kind, got.base, got.inv_base) == (2, 1.0 / 100.0, 100.0)  # recovery.rs:332-358
    trivial = np.arange(100, dtype=np.float32)
    trivial[77] += np.float32(0.0001)
    got = _agree(oracle, sa, trivial)
    assert (got.kind, got.base) == (2, 1.0)  # recovery.rs:388-402
    assert _agree(oracle, sa, np.arange(2000, dtype=np.float64) * 1.5).base == 1.5  # data_types/float.rs:459-466
    lowest = int(np.float64(1.0).view(np.uint64))
    quant = (np.uint64(lowest) + (np.arange(1000, dtype=np.uint64) << np.uint64(20))).view(np.float64)
    got = _agree(oracle, sa, quant)
    assert (got.kind, got.k) == (3, 20)  # data_types/float.rs:511-521


@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.uint64, np.int64])
def test_ints_agree_with_the_oracle(oracle, sa, dtype):
    rng = np.random.default_rng(5)
    info = np.iinfo(dtype)
    for trial in range(40):
        n = int(rng.choice([9, 10, 50, 300, 5000, 1 << 16]))
        base = int(rng.choice([1, 2, 3, 7, 10, 50, 1000, 65536, 10**6]))
        span = max(2, min(int(info.max) // max(base, 1) // 2, 1 << int(rng.integers(3, 40))))
        nums = (rng.integers(0, span, size=n) * base + int(rng.integers(0, base))).astype(dtype)
        if rng.random() < 0.3:  # off-lattice noise
            nums[:: int(rng.integers(2, 50))] += dtype(1)
        if info.min < 0 and rng.random() < 0.5:
            nums = (nums - dtype(span // 2 * base)).astype(dtype)
        _agree(oracle, sa, nums)
    assert sa.choose_mode(np.arange(9, dtype=dtype) * dtype(77)).kind == 1  # fewer than MIN_SAMPLE numbers


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.uint16])
def test_small_ints(sa, dtype):
    rng = np.random.default_rng(6)
    assert sa.choose_mode((rng.integers(0, 25, size=5000) * 5).astype(dtype)).kind in (1, 4)  # a 25-value alphabet is memorizable: either is defensible
    assert sa.choose_mode(rng.integers(0, 100, size=5000).astype(dtype)).kind == 1
    if np.dtype(dtype).itemsize == 2:
        got = sa.choose_mode((rng.integers(0, 3000, size=5000) * 10).astype(dtype))
        assert (got.kind, got.int_base) == (4, 10)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_floats_agree_with_the_oracle(oracle, sa, dtype):
    rng = np.random.default_rng(8)
    u = np.dtype(f"u{np.dtype(dtype).itemsize}")
    for trial in range(40):
        n = int(rng.choice([10, 64, 300, 5000, 1 << 16]))
        pick = trial % 5
        if pick == 0:  # multiples of a base, some snapping to an integer reciprocal, some not
            base = float(rng.choice([0.01, 0.1, 0.125, 1.0 / 7.0, 0.0105, 1.5, 3.0, 1e-4]))
            nums = (rng.integers(-20000, 20000, size=n).astype(np.float64) * base).astype(dtype)
        elif pick == 1:  # quantised mantissas
            k = int(rng.integers(2, 12 if dtype == np.float32 else 40))
            nums = ((rng.standard_normal(n).astype(dtype).view(u) >> u.type(k)) << u.type(k)).view(dtype)
        elif pick == 2:  # nothing to find
            nums = rng.standard_normal(n).astype(dtype)
        elif pick == 3:  # decimals with noise in the last bits, infinities and NaNs in between
            nums = (rng.integers(0, 10000, size=n).astype(np.float64) / 100.0).astype(dtype)
            bits = nums.view(u).copy()
            bits += rng.integers(0, 3, size=n).astype(u)
            nums = bits.view(dtype)
            nums[:: 17] = np.inf
            nums[5:: 29] = np.nan
        else:  # integers stored as floats
            nums = rng.integers(0, 1 << 20, size=n).astype(dtype)
        _agree(oracle, sa, nums)
    assert sa.choose_mode(np.zeros(1000, dtype=dtype)).kind == 1  # zeros are not normal: nothing to sample


def test_argument_checks(sa):
    import ctypes as C

    from pcodec_b200 import _lib

    L = _lib.lib()
    out = sa._CModeChoice()
    assert L.pco_b200_choose_mode(None, C.c_size_t(5), C.c_ubyte(1), C.byref(out)) != 0
    assert L.pco_b200_choose_mode(None, C.c_size_t(0), C.c_ubyte(1), C.byref(out)) == 0 and out.mode_spec == 1
    a = np.zeros(4, dtype=np.uint32)
    assert L.pco_b200_choose_mode(C.c_void_p(a.ctypes.data), C.c_size_t(4), C.c_ubyte(99), C.byref(out)) != 0
    assert L.pco_b200_choose_mode(C.c_void_p(a.ctypes.data), C.c_size_t(4), C.c_ubyte(1), None) != 0


def test_f16_agrees_with_the_oracle(oracle, sa):  # the half crate's widen-to-f32 arithmetic, op by op (data_types/float.rs:254-366)
    import ctypes as C

    def oracle_mode(a):
        kind, base, k = C.c_int(), C.c_double(), C.c_uint32()
        oracle.lib().pco_oracle_kat_choose_float_mode_f16(a.view(np.uint16).ctypes.data_as(C.POINTER(C.c_uint16)), C.c_size_t(a.size), C.byref(kind), C.byref(base), C.byref(k))
        return {0: 1, 2: 2, 3: 3}[kind.value], base.value, k.value

    rng = np.random.default_rng(21)
    seen = set()
    for trial in range(120):
        n = int(rng.choice([10, 64, 300, 5000, 1 << 15]))
        pick = trial % 6
        if pick == 0:
            nums = rng.integers(-300, 300, size=n) * float(rng.choice([0.5, 0.25, 1.0, 2.0, 0.125, 3.0]))
        elif pick == 1:
            nums = rng.integers(0, 60, size=n) / 8.0
        elif pick == 2:
            nums = rng.standard_normal(n)
        elif pick == 3:
            k = int(rng.integers(1, 9))
            nums = ((rng.standard_normal(n).astype(np.float16).view(np.uint16) >> np.uint16(k)) << np.uint16(k)).view(np.float16)
        elif pick == 4:
            nums = rng.integers(0, 2000, size=n).astype(np.float64)
        else:
            nums = np.round(rng.standard_normal(n) * 20.0, 1)
        a = np.ascontiguousarray(nums, dtype=np.float16)
        if pick == 5:
            a[::13] = np.inf
            a[3::31] = np.nan
        got = sa.choose_mode(a)
        want = oracle_mode(a)
        assert got.kind == want[0], (trial, got.kind, want)
        if want[0] == 2:
            assert got.base == want[1]
        if want[0] == 3:
            assert got.k == want[2]
        seen.add(want[0])
    assert seen == {1, 2, 3}  # all three outcomes occurred
    from tests.golden_generators import GENERATORS

    assert sa.choose_mode(GENERATORS["v0_3_0_f16"]()).kind == 1  # the f16 golden asset: its Auto config resolved to Classic
