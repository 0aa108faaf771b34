"""Property tests of the checker itself (hypothesis): whatever the number type, the data, the mode / delta spec, the level and
the paging, the oracle's compressor output decodes back bit for bit, stays within the reference's size guarantee
(pco/src/standalone/guarantee.rs) and parses with the independent metadata reader (pcodec_b200.inspect).  An oracle that is wrong in
some corner would mislead every GPU parity test, so the corners are searched, not just listed."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from pcodec_b200 import inspect as insp
from tests.golden_generators import bits_view

DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float16, np.float32, np.float64]


@st.composite
def arrays(draw):
    dtype = np.dtype(draw(st.sampled_from(DTYPES)))
    n = draw(st.one_of(st.integers(0, 12), st.integers(250, 260), st.integers(0, 700)))
    seed = draw(st.integers(0, 2**31 - 1))
    shape = draw(st.sampled_from(["walk", "small_range", "extremes", "constant", "multiples", "noise", "two_clusters"]))
    rng = np.random.default_rng(seed)
    u = np.dtype(f"u{dtype.itemsize}")
    if shape == "noise":
        a = rng.integers(0, 1 << (8 * dtype.itemsize), size=n, dtype=np.uint64 if dtype.itemsize == 8 else np.int64).astype(u).view(dtype)
    elif shape == "extremes":
        pool = np.array([0, 1, (1 << (8 * dtype.itemsize)) - 1, 1 << (8 * dtype.itemsize - 1), (1 << (8 * dtype.itemsize - 1)) - 1], dtype=np.uint64).astype(u)
        a = pool[rng.integers(0, len(pool), size=n)].view(dtype)
    elif dtype.kind == "f":
        base = {"walk": np.cumsum(rng.normal(size=n)), "small_range": rng.integers(0, 9, size=n) * 0.5, "constant": np.full(n, 1.25),
                "multiples": rng.integers(-500, 500, size=n) * 0.01, "two_clusters": np.where(rng.random(n) < 0.5, 1.0, 1e4) + rng.integers(0, 4, size=n)}[shape]
        a = np.asarray(base).astype(dtype)
    else:
        span = 1 << (8 * dtype.itemsize - 2)
        base = {"walk": np.cumsum(rng.integers(-3, 9, size=n)), "small_range": rng.integers(0, 9, size=n), "constant": np.full(n, 7),
                "multiples": rng.integers(0, 100, size=n) * 7 + 3, "two_clusters": np.where(rng.random(n) < 0.5, 5, span) + rng.integers(0, 4, size=n)}[shape]
        a = np.asarray(base).astype(np.int64).astype(np.uint64).astype(u).view(dtype)
    return np.ascontiguousarray(a)


@st.composite
def configs(draw, dtype):
    dtype = np.dtype(dtype)
    modes = ["MODE_CLASSIC", "MODE_AUTO", "MODE_DICT"] + (["MODE_FLOAT_MULT", "MODE_FLOAT_QUANT"] if dtype.kind == "f" else ["MODE_INT_MULT"])
    mode = draw(st.sampled_from(modes))
    deltas = ["DELTA_NOOP", "DELTA_CONSECUTIVE", "DELTA_LOOKBACK", "DELTA_AUTO"] + (["DELTA_CONV1"] if dtype.itemsize <= 4 else [])
    delta = draw(st.sampled_from(deltas))
    kw = dict(mode=mode, delta=delta, level=draw(st.integers(0, 12)), enable_8_bit=True)
    if delta == "DELTA_CONSECUTIVE":
        kw["delta_order"] = draw(st.integers(0, 7))
    if delta == "DELTA_CONV1":
        kw["delta_order"] = draw(st.integers(0, 8))
    if mode == "MODE_FLOAT_MULT":
        kw["float_mult_base"] = draw(st.sampled_from([0.01, 0.5, 1.0, 3.0, 1.0 / 7.0]))
    if mode == "MODE_FLOAT_QUANT":
        kw["float_quant_k"] = draw(st.integers(1, {2: 10, 4: 23, 8: 52}[dtype.itemsize]))
    if mode == "MODE_INT_MULT":
        kw["int_mult_base"] = draw(st.sampled_from([1, 2, 7, 10, 255]))
    kw["max_page_n"] = draw(st.sampled_from([0, 100, 256, 300]))
    return kw


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])
@given(data=st.data())
def test_compress_then_decompress_is_the_identity(oracle, data):
    nums = data.draw(arrays())
    kw = data.draw(configs(nums.dtype))
    cfg = oracle.make_config(**{k: getattr(oracle, v) if isinstance(v, str) else v for k, v in kw.items()})
    blob = oracle.simple_compress(nums, cfg)
    back = oracle.simple_decompress(blob, nums.dtype)
    assert back.size == nums.size and np.array_equal(bits_view(back), bits_view(nums)), kw
    page_n = kw["max_page_n"] or (1 << 18)
    n_chunks = -(-nums.size // page_n) if nums.size else 0
    per_chunk_bound = sum(oracle.file_size_guarantee(min(page_n, nums.size - i * page_n), nums.dtype) for i in range(n_chunks))
    assert len(blob) <= max(per_chunk_bound, oracle.file_size_guarantee(nums.size, nums.dtype)), kw
    summary = insp.inspect(blob)
    assert summary["n"] == nums.size and summary["compressed"]["total_size"] == len(blob) and summary["compressed"]["unknown_trailing_bytes"] == 0
    # partial destinations (standalone/simple.rs:100-143)
    for m in {0, min(3, nums.size), min(256, nums.size), nums.size}:
        dst = np.zeros(m, dtype=nums.dtype)
        n_proc, finished = oracle.simple_decompress_into(blob, dst)
        assert n_proc == min(m, nums.size) and finished == (m >= nums.size) and np.array_equal(bits_view(dst[:n_proc]), bits_view(nums[:n_proc]))
