"""A plain C program against the reference's C ABI (tests/c/dropin_roundtrip.c, modelled on pco_c/test/test_cpcodec.c), compiled with
gcc against include/cpcodec.h and linked with pcodec_b200/libcpcodec.so - the drop-in boundary exercised the way a C user would."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmpdir):
    from pcodec_b200 import _build as b

    lib = b.build()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = os.path.join(tmpdir, "dropin_roundtrip")
    libdir = os.path.dirname(lib)
    res = subprocess.run([gcc, "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "dropin_roundtrip.c"), "-o", exe, "-L", libdir,
                          "-lcpcodec", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr  # every symbol of the three-function ABI resolves
    return exe


def test_c_program_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch

    exe = _build(str(tmp_path))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert res.returncode == 0, res.stdout + res.stderr
    else:
        assert res.returncode == 3, res.stdout + res.stderr  # PcoCompressionError from the first call: no CPU fallback
        assert "compress_into error 2" in res.stdout


@pytest.mark.gpu
def test_c_program_round_trips_on_the_gpu(tmp_path):
    exe = _build(str(tmp_path))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "all ok" in res.stdout, res.stdout + res.stderr
