"""Codec strings and the results CSV of the reference's bench tool (pco_cli/src/bench/codecs/{mod,pco}.rs, parse.rs, bench/mod.rs)
as mirrored by pcodec_b200/benchfmt.py.  Text handling only: runs on CPU."""
import pytest

from pcodec_b200 import benchfmt as bf


def test_default_codec_prints_bare_name():  # docs/benchmark_results/*.csv call the default configuration just "pco"
    assert str(bf.PcoCodec()) == "pco"
    assert bf.PcoCodec().name(explicit=True) == "pco:level=8:delta=Auto:mode=Auto:chunk-n=262144"
    assert str(bf.parse_codec("pco")) == "pco" and str(bf.parse_codec("pcodec:level=8")) == "pco"


@pytest.mark.parametrize("s", [
    "pco:level=12", "pco:delta=Consecutive@1", "pco:level=8:delta=Consecutive@2:mode=FloatMult@0.01", "pco:mode=Classic", "pco:mode=IntMult@77",
    "pco:mode=FloatQuant@20", "pco:delta=NoOp:mode=Dict", "pco:delta=Lookback", "pco:delta=Conv1@3:chunk-n=65536", "pco:mode=FloatMult@100",
    "pco:mode=FloatMult@1.5:chunk-n=1000",
])
def test_codec_round_trip(s):
    c = bf.parse_codec(s)
    want = s.replace("pco:level=8:", "pco:")  # defaults are dropped when printing
    assert str(c) == want
    assert str(bf.parse_codec(c.name(explicit=True))) == want


def test_spec_grammar_is_case_insensitive_and_strict():  # pco_cli/src/parse.rs:8-48
    assert bf.parse_delta_spec("consecutive@3").order == 3 and bf.parse_delta_spec("NOOP").kind == 1
    assert bf.parse_mode_spec("floatmult@0.25").base == 0.25 and bf.parse_mode_spec("AUTO").kind == 0
    for bad in ("Consecutive", "Delta@1", "conv1@x"):
        with pytest.raises(ValueError):
            bf.parse_delta_spec(bad)
    for bad in ("FloatMult", "Mult@3", "intmult@1.5"):
        with pytest.raises(ValueError):
            bf.parse_mode_spec(bad)
    for bad in ("zstd:level=3", "pco:level", "pco:foo=1", "pco:level=8=9"):
        with pytest.raises(ValueError):
            bf.parse_codec(bad)


def test_chunk_config_of_a_codec():  # pco_cli/src/chunk_config_opt.rs:28-36
    cfg = bf.parse_codec("pco:level=5:delta=Consecutive@1:mode=FloatMult@0.01:chunk-n=4096").chunk_config()
    assert cfg.compression_level == 5 and cfg.delta_spec.kind == 2 and cfg.delta_spec.order == 1
    assert cfg.mode_spec.kind == 2 and cfg.mode_spec.base == 0.01
    assert cfg.paging_spec.kind == 0 and cfg.paging_spec.n == 4096 and cfg.enable_8_bit
    c = cfg._to_c()
    assert (c.compression_level, c.mode_spec, c.float_mult_base, c.delta_spec, c.delta_order, c.max_page_n, c.enable_8_bit) == (5, 2, 0.01, 2, 1, 4096, 1)


def test_results_csv_merge(tmp_path):  # pco_cli/src/bench/mod.rs:325-372
    p = tmp_path / "results.csv"
    p.write_text(bf.CSV_HEADER + "\nair_quality,pco,0.09211633,0.018496584,4283268,42834636\nr_place,zstd,1.5,0.5,10,20")
    n = bf.merge_results_csv(str(p), [
        dict(input="c2_u64_cumsum_geometric", codec=bf.parse_codec("pco:delta=Consecutive@1:mode=Classic"), compress_dt=0.00204, decompress_dt=0.00075,
             compressed_size=381234567, uncompressed_size=2147483648),
        dict(input="air_quality", codec="pco", compress_dt=0.01, decompress_dt=0.002, compressed_size=4283268, uncompressed_size=42834636),
    ])
    assert n == 3
    lines = p.read_text().split("\n")
    assert lines[0] == bf.CSV_HEADER
    assert lines[1] == "air_quality,pco,0.01,0.002,4283268,42834636"  # replaced
    assert lines[2].startswith("c2_u64_cumsum_geometric,pco:delta=Consecutive@1:mode=Classic,0.00204,0.00075,381234567,2147483648")
    assert lines[3] == "r_place,zstd,1.5,0.5,10,20"  # untouched
    # the file it writes parses back to the same thing
    assert bf.merge_results_csv(str(p), []) == 3 and p.read_text().split("\n") == lines
