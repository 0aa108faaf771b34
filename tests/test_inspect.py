"""pcodec_b200.inspect (host-side `pco inspect`, pco_cli/src/inspect/mod.rs) against the oracle's own parse of the same bytes:
golden assets of every format version, oracle-written files of every mode, and the planner parameters SURVEY.md Appendix D
derives for the BASELINE configs.  Metadata reading only - runs on CPU."""
import numpy as np
import pytest

from pcodec_b200 import inspect as insp
from tests.golden_generators import GENERATORS, load_assets

ASSETS = load_assets()


def _same_as_oracle(oracle, data, dtype, summary):
    ref = oracle.inspect(data, dtype)
    assert summary["n_chunks"] == len(ref["chunks"])
    assert summary["format_version"] == f"{ref['format'][0]}.{ref['format'][1]}"
    for c, rc in zip(summary["chunk"], ref["chunks"]):
        assert (c["n"], c["byte_offset"], c["byte_offset"] + c["meta_size"], c["byte_offset"] + c["meta_size"] + c["page_size"]) == \
               (rc["n"], rc["chunk_start"], rc["page_start"], rc["chunk_end"])
        assert insp.MODE_NAMES.index(c["mode"].split("(")[0]) == rc["mode"]
        assert insp.DELTA_NAMES.index(c["delta_encoding"].split("(")[0]) == rc["delta"]
        got_vars = list(c["latent_var"].values())
        assert len(got_vars) == len(rc["vars"])
        for v, rv in zip(got_vars, rc["vars"]):
            assert (v["ans_size_log"], v["n_bins"]) == (rv["ans_size_log"], len(rv["bins"]))
            assert [list(b) for b in v["bins"]] == rv["bins"]
    assert summary["compressed"]["total_size"] == ref["end"]


@pytest.mark.parametrize("name", sorted(GENERATORS))
def test_golden_assets(oracle, name):
    data, expected = ASSETS[name], GENERATORS[name]()
    s = insp.inspect(data)
    assert s["n"] == expected.size and s["compressed"]["unknown_trailing_bytes"] == 0 and s["compressed"]["total_size"] == len(data)
    if expected.size:
        assert np.dtype(expected.dtype) == np.dtype({"f16": "f2", "f32": "f4", "f64": "f8"}.get(s["number_type"], s["number_type"][0] + str(int(s["number_type"][1:]) // 8)))
    _same_as_oracle(oracle, data, expected.dtype, s)


def test_known_modes_of_the_assets():  # pco/src/tests/compatibility.rs:85-114, :157-178, :248-259
    assert insp.inspect(ASSETS["v0_0_0_delta_float_mult"])["chunk"][0]["mode"] == "FloatMult(1.0)"
    assert insp.inspect(ASSETS["v0_1_0_delta_int_mult"])["chunk"][0]["mode"] == "IntMult(1000)"
    assert insp.inspect(ASSETS["v0_3_0_float_quant"])["chunk"][0]["mode"].startswith("FloatQuant(")
    assert insp.inspect(ASSETS["v1_0_0_dict"])["chunk"][0]["mode"] == "Dict(3 values)"
    assert insp.inspect(ASSETS["v0_4_0_lookback_delta"])["chunk"][0]["delta_encoding"].startswith("Lookback(window_n_log=")
    assert insp.inspect(ASSETS["v1_0_0_conv1"])["chunk"][0]["delta_encoding"].startswith("Conv1(quantization=")
    assert set(insp.inspect(ASSETS["v0_4_0_lookback_delta"])["chunk"][0]["latent_var"]) == {"delta", "primary"}


CASES = {
    "classic_order1": (np.uint64, dict(mode="MODE_CLASSIC", delta="DELTA_CONSECUTIVE", delta_order=1)),
    "classic_order7": (np.int32, dict(mode="MODE_CLASSIC", delta="DELTA_CONSECUTIVE", delta_order=7)),
    "float_mult_order2": (np.float64, dict(mode="MODE_FLOAT_MULT", float_mult_base=0.01, delta="DELTA_CONSECUTIVE", delta_order=2)),
    "int_mult": (np.uint32, dict(mode="MODE_INT_MULT", int_mult_base=50, delta="DELTA_NOOP")),
    "float_quant": (np.float32, dict(mode="MODE_FLOAT_QUANT", float_quant_k=9, delta="DELTA_NOOP")),
    "auto_u16": (np.uint16, dict(mode="MODE_AUTO", delta="DELTA_AUTO")),
}


def _data(dtype, n, seed):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "f":
        return (np.round(1e3 * np.cos(np.arange(n) / 50.0) + rng.integers(0, 5, size=n)) * 0.01).astype(dtype)
    vals = np.cumsum(rng.geometric(0.01, size=n)) * (50 if np.dtype(dtype).itemsize >= 4 else 1)
    return (vals % (1 << min(62, 8 * np.dtype(dtype).itemsize - 1))).astype(dtype)


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("n", [1, 255, 256, 257, 3000])
def test_oracle_written_files(oracle, case, n):
    dtype, kw = CASES[case]
    kw = {k: getattr(oracle, v) if isinstance(v, str) else v for k, v in kw.items()}
    nums = _data(dtype, n, 7)
    data = oracle.simple_compress(nums, oracle.make_config(max_page_n=1000, **kw))  # 3000 numbers -> 3 chunks
    s = insp.inspect(data)
    assert s["n"] == n and s["uncompressed_size"] == nums.nbytes and s["compressed"]["total_size"] == len(data)
    _same_as_oracle(oracle, data, dtype, s)
    # the same with the chunk offsets given (no page walk)
    offs = [c["byte_offset"] for c in s["chunk"]]
    s2 = insp.inspect(data, chunk_offsets=offs)
    assert s2["chunk"] == s["chunk"] and s2["compressed"] == s["compressed"]
    if len(offs) > 1:
        with pytest.raises(insp.InspectError):
            insp.inspect(data, chunk_offsets=[offs[0], offs[1] + 1] + offs[2:])


def test_truncated_and_foreign_bytes(oracle):
    nums = _data(np.uint64, 3000, 1)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    with pytest.raises(insp.InspectError):
        insp.inspect(b"nope" + data[4:])
    for cut in (3, 6, 20, len(data) // 2, len(data) - 1):
        with pytest.raises(insp.InspectError):
            insp.inspect(data[:cut])
    assert insp.inspect(data + b"\x01\x02")["compressed"]["unknown_trailing_bytes"] == 2


def test_baseline_config_2_planner_parameters(oracle):  # SURVEY.md Appendix D, BASELINE.json configs[1]
    from pcodec_b200 import datagen

    nums = datagen.c2_u64_cumsum_geometric(seed=0)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_CONSECUTIVE, delta_order=1))
    s = insp.inspect(data)  # one page walk over 2^18 symbols
    assert s["n"] == 1 << 18 and s["n_chunks"] == 1 and s["compressed"]["total_size"] == len(data)
    v = s["chunk"][0]["latent_var"]["primary"]
    assert s["chunk"][0]["delta_encoding"].startswith("Consecutive(order=1")
    assert 1 < v["n_bins"] <= 256 and v["ans_size_log"] <= 10
    assert s["chunk"][0]["meta_size"] <= 4 + 2700  # chunk preamble + <= ~2.6 KiB of bins
    # the bins' approximate cost is the page: within 2 % (the tANS estimate is an entropy bound, the offsets are exact)
    approx_bytes = v["approx_avg_bits"] * ((1 << 18) - 1) / 8
    assert abs(approx_bytes - s["chunk"][0]["page_size"]) / s["chunk"][0]["page_size"] < 0.02


def test_baseline_config_3_planner_parameters(oracle):  # FloatMult(0.01) + order 2: the secondary var is limited to 2^6 bins / size_log 8
    from pcodec_b200 import datagen

    nums = datagen.c3_f64_decimal_sinusoid(seed=0)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=2))
    s = insp.inspect(data)
    c = s["chunk"][0]
    assert c["mode"] == "FloatMult(0.01)" and c["delta_encoding"].startswith("Consecutive(order=2")
    assert c["latent_var"]["primary"]["n_bins"] <= 256 and c["latent_var"]["primary"]["ans_size_log"] <= 10
    assert c["latent_var"]["secondary"]["n_bins"] <= 64 and c["latent_var"]["secondary"]["ans_size_log"] <= 8


def test_auto_config_on_the_baseline_data(oracle):
    """ChunkConfig::default() (Auto mode, Auto delta) on the BASELINE data: the reference's search settles on the explicit configs the
    BASELINE names - Classic + order 1 for C2, FloatMult(0.01) + order 2 for C3, Classic + no delta for C1 - so the explicit-config
    bench measures what a default-config user would get; C3 is the case where the mode search matters (21x smaller than Classic)."""
    from pcodec_b200 import datagen

    def chosen(nums):
        data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_AUTO, delta=oracle.DELTA_AUTO))
        c = insp.inspect(data)["chunk"][0]
        return c["mode"], c["delta_encoding"].split(",")[0].rstrip(")") + ")" if "(" in c["delta_encoding"] else c["delta_encoding"], len(data)

    assert chosen(datagen.c1_u32_lomax(seed=0))[:2] == ("Classic", "NoOp")
    assert chosen(datagen.c2_u64_cumsum_geometric(seed=0))[:2] == ("Classic", "Consecutive(order=1)")
    mode, delta, size = chosen(datagen.c3_f64_decimal_sinusoid(seed=0))
    assert (mode, delta) == ("FloatMult(0.01)", "Consecutive(order=2)")
    nums = datagen.c3_f64_decimal_sinusoid(seed=0)
    assert size == len(oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=2)))
    classic = len(oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_CLASSIC, delta=oracle.DELTA_AUTO)))
    assert classic > 15 * size


def test_bit_flips_only_raise_inspect_errors(oracle):
    """Corrupt input is reported (InspectError) or read as whatever valid file it now is - never an IndexError / ZeroDivisionError."""
    rng = np.random.default_rng(4)
    nums = _data(np.float32, 3000, 2)
    data = oracle.simple_compress(nums, oracle.make_config(mode=oracle.MODE_FLOAT_MULT, float_mult_base=0.01, delta=oracle.DELTA_CONSECUTIVE, delta_order=1, max_page_n=1000))
    for bit in rng.choice(min(len(data), 400) * 8, size=300, replace=False):  # the metadata-heavy front of the file
        d = bytearray(data)
        d[bit // 8] ^= 1 << (bit % 8)
        try:
            insp.inspect(bytes(d))
        except insp.InspectError:
            pass
    for name in ("v0_4_0_lookback_delta", "v1_0_0_conv1", "v1_0_0_dict"):
        base = ASSETS[name]
        for bit in rng.choice(len(base) * 8, size=150, replace=False):
            d = bytearray(base)
            d[bit // 8] ^= 1 << (bit % 8)
            try:
                insp.inspect(bytes(d))
            except insp.InspectError:
                pass
