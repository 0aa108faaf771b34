"""GPU encoder against bytes the Rust crate itself wrote: the golden assets whose configs the GPU path implements re-encode, on the
GPU, to the assets' own bins and PAGE bytes (headers / metadata layouts of those older format versions differ, the page does not).
The oracle does the same on CPU (tests/test_oracle_kats.py::test_older_assets_reencode_to_the_same_bins_and_page); this closes the
chain GPU == Rust output directly, for a one-var page, an f16 page and a two-var FloatQuant page.  (Sorted last on purpose.)"""
import numpy as np
import pytest

from pcodec_b200 import inspect as insp
from tests.golden_generators import GENERATORS, load_assets

pytestmark = pytest.mark.gpu

CASES = {
    "v0_0_0_classic": lambda p: p.ChunkConfig(mode_spec=p.ModeSpec.classic(), delta_spec=p.DeltaSpec.no_op()),            # compatibility.rs:70-82
    "v0_3_0_f16": lambda p: p.ChunkConfig(mode_spec=p.ModeSpec.classic(), delta_spec=p.DeltaSpec.no_op()),                # :145-155
    "v0_3_0_float_quant": lambda p: p.ChunkConfig(mode_spec=p.ModeSpec.try_float_quant(13), delta_spec=p.DeltaSpec.no_op()),  # :157-178
    "v0_4_8_minor_version": lambda p: p.ChunkConfig(mode_spec=p.ModeSpec.classic(), delta_spec=p.DeltaSpec.no_op()),      # :225-245
}


def _first_chunk(buf):
    c = insp.inspect(buf)["chunk"][0]
    start = c["byte_offset"] + c["meta_size"]
    return c, bytes(buf[start:start + c["page_size"]])


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reencodes_the_assets_page(name):
    import pcodec_b200 as p

    asset, nums = load_assets()[name], GENERATORS[name]()
    data = p.standalone.simple_compress(nums, CASES[name](p))
    back = p.standalone.simple_decompress(data, nums.dtype)
    assert np.array_equal(back.view(np.uint8), nums.view(np.uint8))
    ca, page_a = _first_chunk(asset)
    cg, page_g = _first_chunk(data)
    assert (ca["mode"], ca["delta_encoding"]) == (cg["mode"], cg["delta_encoding"])
    assert {k: (v["ans_size_log"], v["bins"]) for k, v in ca["latent_var"].items()} == {k: (v["ans_size_log"], v["bins"]) for k, v in cg["latent_var"].items()}
    assert page_g == page_a


@pytest.mark.parametrize("name", ["v1_0_0_u8", "v1_0_0_i8"])
def test_gpu_reencodes_whole_current_format_assets(name):  # compatibility.rs:281-303: format 4.1 files, the whole file must match
    import pcodec_b200 as p

    asset, nums = load_assets()[name], GENERATORS[name]()
    cfg = p.ChunkConfig(mode_spec=p.ModeSpec.classic(), delta_spec=p.DeltaSpec.try_consecutive(1), enable_8_bit=True)  # what the asset's Auto config resolved to
    assert p.standalone.simple_compress(nums, cfg) == asset
