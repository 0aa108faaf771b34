"""ModeSpec::Auto and DeltaSpec::Auto on the GPU path, PER CHUNK like the reference (pco/src/wrapped/chunk_compressor.rs:310-360,396-440;
pco/src/data_types/unsigned.rs:28-35, float.rs:70-98): the bytes of ChunkConfig::default() - what every reference binding uses
(pco_c/src/lib.rs:43-55) - must equal the oracle's Auto bytes, chunk by chunk, also when the chunks of one array make different choices.
The oracle's Auto also weighs the Lookback candidate; these inputs are ones where a consecutive order (or none) wins."""
import ctypes as C

import numpy as np
import pytest

from tests.golden_generators import bits_view

pytestmark = pytest.mark.gpu

CH = 1 << 18


@pytest.fixture(scope="module")
def sa():
    from pcodec_b200 import standalone

    return standalone


def _auto_cfgs(oracle, max_page_n=0):
    from pcodec_b200 import ChunkConfig, DeltaSpec, ModeSpec, PagingSpec

    ours = ChunkConfig(mode_spec=ModeSpec.auto(), delta_spec=DeltaSpec.auto(), paging_spec=PagingSpec.equal_pages_up_to(max_page_n or CH), enable_8_bit=True)
    theirs = oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO, max_page_n=max_page_n, enable_8_bit=True)
    return ours, theirs


def _check(sa, oracle, nums, max_page_n=0):
    ours_cfg, their_cfg = _auto_cfgs(oracle, max_page_n)
    want = oracle.simple_compress(nums, their_cfg)
    info = oracle.inspect(want, nums.dtype)
    assert all(c["delta"] in (0, 1) for c in info["chunks"]), "test input on which the oracle picks Lookback / Conv1"
    got = sa.simple_compress(nums, ours_cfg)
    if got != want:
        mine = oracle.inspect(got, nums.dtype)
        keys = ("mode", "mode_base_latent", "mode_k", "delta", "delta_order")
        diff = [(i, {k: (a[k], b[k]) for k in keys if a[k] != b[k]}) for i, (a, b) in enumerate(zip(mine["chunks"], info["chunks"])) if any(a[k] != b[k] for k in keys)]
        raise AssertionError(f"Auto bytes differ from the oracle's: ours {len(got)} theirs {len(want)}; choices that differ (ours, oracle): {diff[:4]}")
    back = sa.simple_decompress(got, nums.dtype)
    np.testing.assert_array_equal(bits_view(back), bits_view(nums))
    return info


def test_auto_on_the_baseline_configs(sa, oracle):
    from pcodec_b200 import datagen

    _check(sa, oracle, datagen.c1_u32_lomax(seed=0))
    _check(sa, oracle, np.concatenate([datagen.c2_u64_cumsum_geometric(seed=s) for s in range(3)]))
    _check(sa, oracle, np.concatenate([datagen.c3_f64_decimal_sinusoid(seed=s) for s in range(2)]))
    for dtype in (np.uint8, np.uint16, np.int32, np.int64, np.float32, np.float64):
        _check(sa, oracle, datagen.c5_sweep(dtype, seed=1))


def test_auto_per_chunk_choices_differ_within_one_array(sa, oracle):
    """Three chunks, three answers: IntMult(1000) + a delta order, Classic with order 1, Classic without delta - several runs of the pipeline."""
    rng = np.random.default_rng(3)
    n = 30000
    a = (np.cumsum(rng.integers(-50, 50, size=n)) * 1000 + 10**9).astype(np.int64)
    b = np.cumsum(rng.geometric(0.01, size=n)).astype(np.int64)
    c = rng.integers(0, 1 << 40, size=n).astype(np.int64)
    nums = np.concatenate([a, b, c, b + 7, a])
    info = _check(sa, oracle, nums, max_page_n=n)
    kinds = [(ch["mode"], ch.get("delta_order", 0)) for ch in info["chunks"]]
    assert len(set(kinds)) >= 3, kinds


def test_auto_float_modes_per_chunk(sa, oracle):
    rng = np.random.default_rng(5)
    n = 20000
    dec = rng.integers(-5000, 5000, size=n).astype(np.float64) / 100.0                                     # FloatMult(0.01)
    quant = ((rng.standard_normal(n).view(np.uint64) >> np.uint64(30)) << np.uint64(30)).view(np.float64)  # FloatQuant(30)
    plain = rng.standard_normal(n)                                                                         # Classic
    info = _check(sa, oracle, np.concatenate([dec, quant, plain, dec]), max_page_n=n)
    assert len({ch["mode"] for ch in info["chunks"]}) == 3, [ch["mode"] for ch in info["chunks"]]
    _check(sa, oracle, np.concatenate([dec, plain]).astype(np.float32), max_page_n=n)


@pytest.mark.parametrize("dtype_byte,dtype", [(1, np.uint32), (2, np.uint64), (4, np.int64), (6, np.float64)])
def test_reference_abi_with_null_config_is_byte_identical(oracle, dtype_byte, dtype):
    """pco_standalone_simple_compress_into(config = NULL) == the oracle's simple_compress_into with ChunkConfig::default() (uniform-type header)."""
    from pcodec_b200 import _lib, datagen

    L = _lib.lib()
    nums = {np.uint32: datagen.c1_u32_lomax(n=100000, seed=2), np.uint64: datagen.c2_u64_cumsum_geometric(n=CH + 1000, seed=3),
            np.int64: datagen.c5_sweep(np.int64, n=70000, seed=4), np.float64: datagen.c3_f64_decimal_sinusoid(n=90000, seed=1)}[dtype]
    cap = L.pco_standalone_guarantee_file_size(C.c_size_t(nums.size), C.c_ubyte(dtype_byte))
    dst = np.zeros(cap, dtype=np.uint8)
    nw = C.c_size_t()
    rc = L.pco_standalone_simple_compress_into(nums.ctypes.data_as(C.c_void_p), C.c_size_t(nums.size), C.c_ubyte(dtype_byte), None, dst.ctypes.data_as(C.c_void_p),
                                               C.c_size_t(cap), C.byref(nw))
    assert rc == 0
    want = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO, enable_8_bit=True), uniform_type=True)
    assert dst[: nw.value].tobytes() == want


def test_auto_refuses_loudly_where_lookback_wins(sa, oracle):
    """The reference's Auto delta also weighs Lookback (chunk_compressor.rs:326-338).  This path weighs it the same way but does not encode
    it: on a chunk where it wins the call fails with Unsupported instead of writing bytes the reference would not write."""
    from pcodec_b200 import PcoError

    rng = np.random.default_rng(1)
    motif = rng.integers(0, 1 << 30, size=97)
    nums = np.tile(motif, 700)[:60000].astype(np.uint32)
    ours_cfg, their_cfg = _auto_cfgs(oracle)
    info = oracle.inspect(oracle.simple_compress(nums, their_cfg), np.uint32)
    assert info["chunks"][0]["delta"] == 2, "the oracle is expected to pick Lookback on a repeating motif"
    with pytest.raises(PcoError) as e:
        sa.simple_compress(nums, ours_cfg)
    assert e.value.kind == "Unsupported"
    # a chunk before it that picks a consecutive order does not change that
    walk = np.cumsum(rng.geometric(0.01, size=30000)).astype(np.uint32)
    with pytest.raises(PcoError) as e:
        sa.simple_compress(np.concatenate([walk, nums[:30000]]), _auto_cfgs(oracle, 30000)[0])
    assert e.value.kind == "Unsupported"
