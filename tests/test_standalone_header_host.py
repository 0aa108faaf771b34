"""Host-side parsing in the Python mirror (pcodec_b200/standalone.py: _header, peek_dtype, n_hint) on the reference's golden assets -
the part of pcodec.standalone.simple_decompress that infers the dtype and sizes the destination before any device work
(docs/format.md:173-192; pco/src/standalone/decompressor.rs:85-148, :203-231)."""
import numpy as np
import pytest

from tests.golden_generators import GENERATORS, load_assets


@pytest.mark.parametrize("name", sorted(GENERATORS))
def test_peek_dtype_and_n_hint_on_golden_assets(name):
    from pcodec_b200 import standalone

    data = load_assets()[name]
    want = GENERATORS[name]()
    dt = standalone.peek_dtype(data)
    if want.size == 0:
        # an empty file: the dtype is known only when the header carries a uniform type (standalone version >= 3)
        assert dt is None or np.dtype(dt) == want.dtype
    else:
        assert np.dtype(dt) == want.dtype
    hint = standalone.n_hint(data)
    # n_hint exists from standalone version 2 on; older files report 0 and the mirror grows the destination as it goes
    assert hint in (0, want.size)


def test_header_rejects_foreign_bytes():
    from pcodec_b200 import PcoError, standalone

    with pytest.raises(PcoError) as e:
        standalone.peek_dtype(b"not a pco file at all")
    assert e.value.kind == "Corruption"
    assert standalone.n_hint(b"xx") == 0
