"""include/*.h are valid C99 and the library links and runs from plain C - for the entry points that need no device (CPU test;
the full C drop-in round trip is tests/test_c_dropin.py on the GPU)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


@pytest.mark.skipif(shutil.which("gcc") is None, reason="no gcc")
def test_headers_are_c99_and_host_entry_points_run(tmp_path):
    from pcodec_b200 import _lib

    _lib.lib()  # builds libcpcodec.so if it is not there yet
    libdir = os.path.join(ROOT, "pcodec_b200")
    exe = str(tmp_path / "host_only")
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "host_only.c"), "-o", exe,
                            "-L", libdir, "-l:libcpcodec.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "HOST_ONLY_OK" in res.stdout, (res.returncode, res.stdout, res.stderr)
