"""CPU-only: the C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("bpk.h", "cuda_bulletproof.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b((?:bpk|cuda)_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    assert lib._missing == []
    decl = declared_symbols()
    assert decl == set(cbp.SIGNATURES), decl ^ set(cbp.SIGNATURES)
    for name in decl:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bpk_version()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import cudabulletproof_b200 as cbp
    monkeypatch.setattr(cbp, "_lib", None)
    monkeypatch.setattr(cbp, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(cbp.BpkError):
        cbp.load()


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "cudabulletproof_b200")
    pat = re.compile(r'(#include\s+"[^"]*oracle|import\s+oracle|from\s+oracle|liboracle|libref_verbatim|oracle\.binding)')
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not pat.search(text), f
