"""The drop-in boundary consumed from plain C: tests/csrc/abi_smoke.c is compiled with gcc against
include/varpro_b200.h and linked to libvarpro_b200.so (no Python, no ctypes in between)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "csrc", "abi_smoke.c")
EXE = os.path.join(ROOT, "tests", "csrc", "abi_smoke")


def _build():
    from varpro_b200 import _lib
    _lib.load()  # the library must exist
    libdir = os.path.join(ROOT, "varpro_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-o", EXE, SRC, f"-L{libdir}",
                    "-lvarpro_b200", "-lm", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    return EXE


def test_header_compiles_as_c99_and_links():
    """-Wall -Wextra -Werror: the header is valid C (not only C++), every call in the C program resolves."""
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True)
    # without a GPU the program must stop at vp_ctx_create with VP_ERR_CUDA (exit 77): no CPU fallback
    assert r.returncode in (0, 77), (r.returncode, r.stdout, r.stderr)
    if r.returncode == 77:
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_c_program_fits_the_reference_mrhs_problem():
    exe = _build()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "worst error" in r.stdout
