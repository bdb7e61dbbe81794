"""The drop-in boundary from plain C: tests/c_abi/dropin_smoke.c includes only include/cuda_bulletproof.h
(reference structs, reference prototypes), is compiled with gcc as C99 and linked against the shared library.
CPU: it compiles and links (the header is valid C, every used symbol resolves).  GPU: it runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "dropin_smoke.c")
OUT = os.path.join(ROOT, "tests", "c_abi", "dropin_smoke")


def build_binary():
    import cudabulletproof_b200.build as b
    lib = b.build()
    libdir = os.path.dirname(lib)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", SRC, "-o", OUT, "-L" + libdir, "-lcudabulletproof_b200",
           "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return OUT


def test_c_caller_compiles_and_links():
    assert os.path.exists(build_binary())


@pytest.mark.gpu
def test_c_caller_runs_on_gpu():
    exe = build_binary() if not os.path.exists(OUT) else OUT
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "dropin_smoke ok" in r.stdout
    assert "Vector lengths must match" in r.stderr  # the reference's message for the mismatched call
