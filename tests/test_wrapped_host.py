"""Device-free parts of the wrapped-format C-ABI (pco/src/wrapped/): the 2-byte header and the host-side walk over a chunk's
metadata that tells a caller where the page starts (pco_b200_chunk_meta_size) - checked against pcodec_b200.inspect / the oracle's
parse on every mode and delta encoding and on the golden assets.  CPU test."""
import ctypes as C

import numpy as np
import pytest

from pcodec_b200 import _lib
from pcodec_b200 import inspect as insp
from tests.golden_generators import GENERATORS, load_assets


def _meta_size(buf, off, dtype):
    L = _lib.lib()
    src = (C.c_uint8 * (len(buf) - off)).from_buffer_copy(buf[off:])
    out = C.c_size_t()
    rc = L.pco_b200_chunk_meta_size(src, C.c_size_t(len(buf) - off), C.c_ubyte(_lib.dtype_byte(dtype)), C.byref(out))
    return rc, out.value


def test_wrapped_header_bytes():  # metadata/format_version.rs:30-34, :60-91
    L = _lib.lib()
    buf = (C.c_uint8 * 8)()
    n = C.c_size_t()
    assert L.pco_b200_file_compressor_write_header(buf, C.c_size_t(8), C.byref(n)) == 0 and n.value == 2 and bytes(buf[:2]) == b"\x04\x01"
    assert L.pco_b200_file_compressor_write_header(buf, C.c_size_t(1), C.byref(n)) != 0
    assert L.pco_b200_file_decompressor_read_header(buf, C.c_size_t(2), C.byref(n)) == 0 and n.value == 2
    assert L.pco_b200_file_decompressor_read_header(buf, C.c_size_t(1), C.byref(n)) != 0  # cut short
    assert L.pco_b200_file_decompressor_read_header(buf, C.c_size_t(0), C.byref(n)) != 0
    buf[0] = 5
    assert L.pco_b200_file_decompressor_read_header(buf, C.c_size_t(2), C.byref(n)) != 0  # a newer major version


CASES = [
    (np.uint64, dict(mode="MODE_CLASSIC", delta="DELTA_CONSECUTIVE", delta_order=1)),
    (np.int32, dict(mode="MODE_CLASSIC", delta="DELTA_CONSECUTIVE", delta_order=7)),
    (np.uint8, dict(mode="MODE_CLASSIC", delta="DELTA_NOOP")),
    (np.float64, dict(mode="MODE_FLOAT_MULT", float_mult_base=0.01, delta="DELTA_CONSECUTIVE", delta_order=2)),
    (np.uint32, dict(mode="MODE_INT_MULT", int_mult_base=50, delta="DELTA_NOOP")),
    (np.float32, dict(mode="MODE_FLOAT_QUANT", float_quant_k=9, delta="DELTA_NOOP")),
    (np.uint16, dict(mode="MODE_CLASSIC", delta="DELTA_LOOKBACK")),
    (np.int32, dict(mode="MODE_CLASSIC", delta="DELTA_CONV1", delta_order=3)),
    (np.int64, dict(mode="MODE_DICT", delta="DELTA_NOOP")),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_chunk_meta_size_matches_the_parse(oracle, case):
    dtype, kw = CASES[case]
    kw = {k: getattr(oracle, v) if isinstance(v, str) else v for k, v in kw.items()}
    rng = np.random.default_rng(case)
    if np.dtype(dtype).kind == "f":
        nums = (np.round(1e3 * np.cos(np.arange(4000) / 40.0) + rng.integers(0, 7, size=4000)) * 0.01).astype(dtype)
    else:
        nums = ((np.cumsum(rng.integers(0, 9, size=4000)) * 50) % (1 << min(40, 8 * np.dtype(dtype).itemsize - 1))).astype(dtype)
    data = oracle.simple_compress(nums, oracle.make_config(max_page_n=1500, **kw))
    for ch in insp.inspect(data)["chunk"]:
        rc, got = _meta_size(data, ch["byte_offset"] + 4, dtype)  # a standalone chunk = 4-byte preamble + the wrapped chunk meta + page
        if ch["mode"].startswith("Dict") or ch["delta_encoding"].startswith(("Lookback", "Conv1")):
            assert rc == 7  # outside the GPU hot path: refused as Unsupported, never mis-measured
            continue
        assert rc == 0 and got == ch["meta_size"] - 4, (ch["mode"], ch["delta_encoding"], got, ch["meta_size"] - 4)
        # cut anywhere inside the metadata: insufficient data, never a wrong length
        for cut in (0, 1, got // 2, got - 1):
            rc2, _ = _meta_size(data[: ch["byte_offset"] + 4 + cut], ch["byte_offset"] + 4, dtype) if cut else (1, 0)
            assert rc2 != 0


def test_chunk_meta_size_on_current_format_assets():
    assets = load_assets()
    for name in ("v1_0_0_u8", "v1_0_0_i8", "v0_4_8_minor_version"):  # format 4.x, in-scope encodings
        ch = insp.inspect(assets[name])["chunk"][0]
        rc, got = _meta_size(assets[name], ch["byte_offset"] + 4, GENERATORS[name]().dtype)
        assert rc == 0 and got == ch["meta_size"] - 4, name
    for name in ("v1_0_0_conv1", "v1_0_0_dict"):
        ch = insp.inspect(assets[name])["chunk"][0]
        assert _meta_size(assets[name], ch["byte_offset"] + 4, GENERATORS[name]().dtype)[0] == 7, name
