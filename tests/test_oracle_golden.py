"""Pins the oracle against the reference's golden assets (decode KATs)."""
import numpy as np
import pytest

from tests.golden_generators import GENERATORS, bits_view, load_assets

ASSETS = load_assets()


@pytest.mark.parametrize("name", sorted(GENERATORS))
def test_asset_decodes_to_generator(oracle, name):
    expected = GENERATORS[name]()
    got = oracle.simple_decompress(ASSETS[name], expected.dtype)
    assert got.shape == expected.shape
    np.testing.assert_array_equal(bits_view(got), bits_view(expected))


@pytest.mark.parametrize("name", sorted(GENERATORS))
def test_asset_consumed_exactly(oracle, name):
    expected = GENERATORS[name]()
    info = oracle.inspect(ASSETS[name], expected.dtype)
    # every asset ends with the terminator byte directly after the last chunk
    assert info["end"] == len(ASSETS[name])
