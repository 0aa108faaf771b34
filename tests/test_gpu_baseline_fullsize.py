"""Byte parity at BASELINE.json's FULL chunk size (2^18 numbers per chunk) for every BASELINE config, through the C-ABI:
  C1  u32 lomax, Classic, NoOp                                  (configs[0])
  C2  u64 cumsum-geometric, Classic, Consecutive(1)             (configs[1], several chunks)
  C3  f64 decimal sinusoid, FloatMult(0.01), Consecutive(2)     (configs[2])
  C5  {u8,u16,i32,i64,f32,f64} x consecutive orders 0..7        (configs[4])
GPU bytes must equal the oracle's bytes for the same ChunkConfig, and the GPU must decode them back to the input bit for
bit (pco/src/tests/recovery.rs:49-84 is the reference's shape of this test).  Data: pcodec_b200/datagen.py.
"""
import numpy as np
import pytest

from tests.golden_generators import bits_view

pytestmark = pytest.mark.gpu

CHUNK_N = 1 << 18


@pytest.fixture(scope="module")
def sa():
    from pcodec_b200 import standalone

    return standalone


def _pair(oracle, mode="classic", order=0, base=0.01):
    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec

    ms = ModeSpec.classic() if mode == "classic" else ModeSpec.try_float_mult(base)
    om = oracle.MODE_CLASSIC if mode == "classic" else oracle.MODE_FLOAT_MULT
    ours = ChunkConfig(mode_spec=ms, delta_spec=DeltaSpec.try_consecutive(order) if order else DeltaSpec.no_op(), enable_8_bit=True)
    theirs = oracle.make_config(mode=om, float_mult_base=base, delta=oracle.DELTA_CONSECUTIVE if order else oracle.DELTA_NOOP, delta_order=order,
                                enable_8_bit=True)
    return ours, theirs


def _check(sa, oracle, nums, ours_cfg, their_cfg):
    want = oracle.simple_compress(nums, their_cfg)
    got = sa.simple_compress(nums, ours_cfg)
    assert len(got) == len(want), (len(got), len(want))
    assert got == want
    back = sa.simple_decompress(got, nums.dtype)
    np.testing.assert_array_equal(bits_view(back), bits_view(nums))


def test_c1_u32_lomax_classic_noop(sa, oracle):
    from pcodec_b200 import datagen

    for seed in (0, 1):
        ours, theirs = _pair(oracle)
        _check(sa, oracle, datagen.c1_u32_lomax(seed=seed), ours, theirs)


def test_c2_u64_order1_several_chunks(sa, oracle):
    from pcodec_b200 import datagen

    nums = np.concatenate([datagen.c2_u64_cumsum_geometric(seed=s) for s in range(6)])
    ours, theirs = _pair(oracle, order=1)
    _check(sa, oracle, nums, ours, theirs)


def test_c2_secondary_shapes(sa, oracle):
    """C2(ii) uniform random u64 (trips should_fallback: one 64-bit bin) and C2(iii) an arithmetic sequence (trivial bins)."""
    rng = np.random.default_rng(5)
    ours, theirs = _pair(oracle, order=1)
    _check(sa, oracle, rng.integers(0, 2**64, size=CHUNK_N, dtype=np.uint64), ours, theirs)
    _check(sa, oracle, (np.arange(CHUNK_N, dtype=np.uint64) * np.uint64(77) + np.uint64(12345)), ours, theirs)


def test_c3_f64_float_mult_order2(sa, oracle):
    from pcodec_b200 import datagen

    nums = np.concatenate([datagen.c3_f64_decimal_sinusoid(seed=s) for s in range(2)])
    ours, theirs = _pair(oracle, mode="float_mult", order=2)
    _check(sa, oracle, nums, ours, theirs)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("order", list(range(8)))
def test_c5_dtype_sweep(sa, oracle, dtype, order):
    from pcodec_b200 import datagen

    ours, theirs = _pair(oracle, order=order)
    _check(sa, oracle, datagen.c5_sweep(dtype, seed=order), ours, theirs)
