"""The reference's own Python test cases (pco_python/test/test_standalone.py), run against pcodec_b200.standalone where the
GPU path implements the feature; the cases the path refuses (TryDict, TryLookback, TryConv1 compress) are asserted to fail
loudly with kind "Unsupported" instead of silently falling back.  Error texts differ from the Rust crate's; kinds are compared.
"""
import numpy as np
import pytest

from tests.golden_generators import load_assets

pytestmark = pytest.mark.gpu

ALL_LENGTHS = (0, 900)
ALL_DTYPES = ("f2", "f4", "f8", "i2", "i4", "i8", "u2", "u4", "u8")


@pytest.fixture(scope="module")
def api():
    import pcodec as p  # the reference's import name, served by pcodec_b200 (pcodec/__init__.py)

    return p


@pytest.mark.parametrize("length", ALL_LENGTHS)
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_round_trip_decompress_into(api, length, dtype):  # test_standalone.py:22-33
    rng = np.random.default_rng(12345)
    data = rng.uniform(0, 1000, size=length).astype(dtype)
    compressed = api.standalone.simple_compress(data, api.ChunkConfig())
    out = np.empty_like(data)
    progress = api.standalone.simple_decompress_into(compressed, out)
    np.testing.assert_array_equal(data, out)
    assert progress.n_processed == data.size
    assert progress.finished


@pytest.mark.parametrize("length", ALL_LENGTHS)
@pytest.mark.parametrize("dtype", ALL_DTYPES)
def test_round_trip_simple_decompress(api, length, dtype):  # :36-44
    rng = np.random.default_rng(12345)
    data = rng.uniform(0, 1000, size=length).astype(dtype)
    compressed = api.standalone.simple_compress(data, api.ChunkConfig(paging_spec=api.PagingSpec.equal_pages_up_to(300)))
    out = api.standalone.simple_decompress(compressed, dtype)
    np.testing.assert_array_equal(data, out)


def test_inexact_decompression(api):  # :47-66
    data = np.random.default_rng(1).uniform(size=300)
    compressed = api.standalone.simple_compress(data, api.ChunkConfig())
    out = np.zeros(3)
    progress = api.standalone.simple_decompress_into(compressed, out)
    np.testing.assert_array_equal(out, data[:3])
    assert progress.n_processed == 3
    assert not progress.finished
    out = np.zeros(600)
    progress = api.standalone.simple_decompress_into(compressed, out)
    np.testing.assert_array_equal(out[:300], data)
    np.testing.assert_array_equal(out[300:], np.zeros(300))
    assert progress.n_processed == 300
    assert progress.finished


def test_simple_decompress_into_errors(api):  # :69-77: dtype of dst does not match the chunk's number type
    data = np.random.default_rng(2).uniform(size=100).astype(np.float32)
    compressed = api.standalone.simple_compress(data, api.ChunkConfig())
    with pytest.raises(RuntimeError) as e:
        api.standalone.simple_decompress_into(compressed, np.zeros(100, dtype=np.float64))
    assert e.value.kind == "Corruption"  # pco/src/standalone/decompressor.rs:203-219


def test_simple_decompress_errors(api):  # :80-108, on the golden asset v0_4_5_uniform_type.pco
    compressed = bytearray(load_assets()["v0_4_5_uniform_type"])
    with pytest.raises(RuntimeError) as e:
        api.standalone.simple_decompress(bytes(compressed[:8]), np.uint32)
    assert e.value.kind == "InsufficientData"
    compressed[8] = 99  # byte 8 is the first chunk's number type, byte 5 the file's uniform type
    with pytest.raises(RuntimeError) as e:
        api.standalone.simple_decompress(bytes(compressed), np.uint32)
    assert e.value.kind == "Corruption"
    compressed[8] = 0  # a file with no chunks: an empty array of the uniform type
    got = api.standalone.simple_decompress(bytes(compressed))
    assert got.dtype == np.uint32 and got.size == 0
    compressed[5] = 0  # no uniform type and no chunk: nothing to infer a dtype from
    assert api.standalone.simple_decompress(bytes(compressed)) is None


@pytest.mark.parametrize("delta", ["no_op", "consecutive1", "lookback", "conv1"])
def test_compression_options(api, delta):  # :111-136
    data = np.random.default_rng(3).normal(size=100).astype(np.float32)
    default_size = len(api.standalone.simple_compress(data, api.ChunkConfig()))
    spec = {"no_op": api.DeltaSpec.no_op(), "consecutive1": api.DeltaSpec.try_consecutive(1), "lookback": api.DeltaSpec.try_lookback(),
            "conv1": api.DeltaSpec.try_conv1(1)}[delta]
    cfg = api.ChunkConfig(compression_level=0, delta_spec=spec, mode_spec=api.ModeSpec.classic(), paging_spec=api.PagingSpec.equal_pages_up_to(77))
    if delta in ("lookback", "conv1"):
        with pytest.raises(RuntimeError) as e:  # outside the GPU hot path: refused, never a fallback
            api.standalone.simple_compress(data, cfg)
        assert e.value.kind == "Unsupported"
        return
    compressed = api.standalone.simple_compress(data, cfg)
    np.testing.assert_array_equal(data, api.standalone.simple_decompress(compressed, np.float32))
    assert len(compressed) >= default_size


@pytest.mark.parametrize("mode", ["auto", "classic", "int_mult", "dict"])
def test_compression_int_mode_spec_options(api, mode):  # :139-159
    data = (np.random.default_rng(4).normal(size=100) * 1000).astype(np.int32)
    spec = {"auto": api.ModeSpec.auto(), "classic": api.ModeSpec.classic(), "int_mult": api.ModeSpec.try_int_mult(10), "dict": api.ModeSpec.try_dict()}[mode]
    if mode == "dict":
        with pytest.raises(RuntimeError) as e:
            api.standalone.simple_compress(data, api.ChunkConfig(mode_spec=spec))
        assert e.value.kind == "Unsupported"
        return
    compressed = api.standalone.simple_compress(data, api.ChunkConfig(mode_spec=spec))
    np.testing.assert_array_equal(data, api.standalone.simple_decompress(compressed, np.int32))


@pytest.mark.parametrize("mode", ["auto", "classic", "float_mult", "float_quant"])
def test_compression_float_mode_spec_options(api, mode):  # :162-182
    data = (np.random.default_rng(5).normal(size=100) * 1000).astype(np.int32) * np.pi
    spec = {"auto": api.ModeSpec.auto(), "classic": api.ModeSpec.classic(), "float_mult": api.ModeSpec.try_float_mult(10.0),
            "float_quant": api.ModeSpec.try_float_quant(4)}[mode]
    compressed = api.standalone.simple_compress(data, api.ChunkConfig(mode_spec=spec))
    np.testing.assert_array_equal(data, api.standalone.simple_decompress(compressed, np.float64))


def test_decompress_without_n_hint(api):  # :185-190: old files have no n_hint
    assert len(api.standalone.simple_decompress(load_assets()["v0_0_0_classic"], np.int32)) == 2000
