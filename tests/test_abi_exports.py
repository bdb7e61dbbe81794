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


def test_msm_plan_is_host_only_and_consistent():
    """bpk_msm_window_bits / bpk_msm_workspace_bytes are pure host logic (the plan): the measured window table,
    the workspace growing with the input, the slotted front end's extra room from 2^19 points, argument checks."""
    import ctypes as C
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    OPT_SLOTS = 0  # include/bpk.h BPK_OPT_MSM_SLOTS: -1 auto, 0 off, 1 on (no entry point reads the environment)
    assert lib.bpk_debug_set_option(OPT_SLOTS, -1) == 0
    assert [lib.bpk_msm_window_bits(n) for n in (1, 1 << 10, 1 << 11, 1 << 18, 1 << 19, 1 << 24)] == [13, 13, 15, 15, 16, 16]

    def ws(n, c=0):
        out = C.c_size_t(0)
        assert lib.bpk_msm_workspace_bytes(n, c, C.byref(out)) == 0
        return out.value

    sizes = [ws(1 << lg) for lg in range(14, 25, 2)]  # Pippenger sizes (up to 2^13 points the Straus tables are added)
    assert sizes == sorted(sizes) and sizes[0] > 0
    small = [ws(1 << lg) for lg in range(0, 14)]
    assert small == sorted(small) and small[0] > 0
    assert ws(1 << 13) - ws(1 << 13, 15) >= (1 << 13) * (8 * 128 + 64)  # the multiples 1..8 and 64 digits per point
    slotted = ws(1 << 20)
    lib.bpk_debug_set_option(OPT_SLOTS, 0)
    compact = ws(1 << 20)
    # slots: 15 windows x 2^15 buckets x cap entries (mean 32 + 8 sigma + 8, a multiple of 8 = 88) + the top window's
    # compact run and ranks, against 16 x 2^20 entries
    cap = 88
    assert slotted - compact == (15 * 32768 * cap + (1 << 20) - 16 * (1 << 20)) * 4 + 8 * (1 << 20)
    lib.bpk_debug_set_option(OPT_SLOTS, 1)
    assert ws(1 << 12, 11) > 0
    lib.bpk_debug_set_option(OPT_SLOTS, -1)
    out = C.c_size_t(0)
    for bad in (3, 18, -1):
        assert lib.bpk_msm_workspace_bytes(1 << 12, bad, C.byref(out)) != 0
    assert lib.bpk_msm_workspace_bytes(1 << 31, 0, C.byref(out)) != 0
    # entry offsets are 32-bit prefix sums over n * W entries: n * ceil(256 / c) must stay below 2^32
    assert lib.bpk_msm_workspace_bytes((1 << 28) - 1, 16, C.byref(out)) == 0
    assert lib.bpk_msm_workspace_bytes(1 << 28, 16, C.byref(out)) != 0
    assert lib.bpk_msm_workspace_bytes(1 << 28, 0, C.byref(out)) != 0
    assert lib.bpk_msm_workspace_bytes(1 << 27, 4, C.byref(out)) != 0
    assert lib.bpk_debug_set_option(99, 0) != 0
    lib.bpk_last_error()  # clear
