"""Pins oracle/ref_corrected.c (the parity oracle) against known answers that do NOT come from
this repo: RFC 8032 §7.1 public keys, [l]B = O, and the independent Python big-int model
oracle/pyref.py.  CPU only."""
import ctypes as C
import random

import numpy as np

from oracle import binding as ob
from oracle import pyref

P, L = pyref.P, pyref.L
EDGE = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**255, 2**256 - 1, 2**256 - 38, 2**255 - 20, 2**128, L, L - 1]


def fe_op(lib, name, *vals):
    out = np.zeros(4, dtype=np.uint64)
    args = [ob.int_to_fe(v) for v in vals]
    getattr(lib, name)(ob.ptr(out), *[ob.ptr(a) for a in args])
    return ob.fe_to_int(out)


def test_fe_ops_match_bigint(oracle):
    rng = random.Random(1)
    vals = EDGE + [rng.getrandbits(256) for _ in range(200)]
    for a in vals:
        assert fe_op(oracle, "fe25519_sq", a) == a * a % P
        assert fe_op(oracle, "fe25519_neg", a) == (-a) % P
        for b in rng.sample(vals, 6):
            assert fe_op(oracle, "fe25519_add", a, b) == (a + b) % P
            assert fe_op(oracle, "fe25519_sub", a, b) == (a - b) % P
            assert fe_op(oracle, "fe25519_mul", a, b) == (a * b) % P


def test_fe_invert_pow_tobytes(oracle):
    rng = random.Random(2)
    for a in [1, 2, P - 1, 2**255 - 20, 2**256 - 1] + [rng.getrandbits(255) for _ in range(20)]:
        assert fe_op(oracle, "fe25519_invert", a) == pow(a, P - 2, P)
        assert fe_op(oracle, "fe25519_pow2523", a) == pow(a, (P - 5) // 8, P)
        out = (C.c_uint8 * 32)()
        oracle.fe25519_tobytes(out, ob.ptr(ob.int_to_fe(a)))
        assert bytes(out) == (a % P).to_bytes(32, "little")
    assert fe_op(oracle, "fe25519_invert", 0) == 0


def test_fe_batch_invert(oracle):
    rng = random.Random(3)
    vals = [rng.getrandbits(256) for _ in range(50)]
    vals[7] = 0
    vals[20] = P
    a = ob.ints_to_fe(vals)
    out = np.zeros_like(a)
    oracle.fe25519_batch_invert(ob.ptr(out), ob.ptr(a), len(vals))
    for i, v in enumerate(vals):
        assert ob.fe_to_int(out[i]) == (pow(v, P - 2, P) if v % P else 0)


def test_sc_ops_match_bigint(oracle):
    rng = random.Random(4)
    vals = EDGE + [rng.getrandbits(256) for _ in range(60)]
    for a in vals:
        assert fe_op(oracle, "sc25519_reduce", a) == a % L
        assert fe_op(oracle, "sc25519_neg", a) == (-a) % L
        for b in rng.sample(vals, 4):
            assert fe_op(oracle, "sc25519_add", a, b) == (a + b) % L
            assert fe_op(oracle, "sc25519_sub", a, b) == (a - b) % L
            assert fe_op(oracle, "sc25519_mul", a, b) == (a * b) % L
    for a in [1, 2, L - 1, rng.getrandbits(252)]:
        assert fe_op(oracle, "sc25519_invert", a) == pow(a, L - 2, L)


def _scalarmult(oracle, k, pt):
    g = ob.affine_to_ge(*pt)
    out = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult(ob.ptr(out), int(k).to_bytes(32, "little"), ob.ptr(g))
    return out


def _pack(oracle, g):
    out = (C.c_uint8 * 32)()
    oracle.ge25519_pack(out, ob.ptr(g))
    return bytes(out)


def test_rfc8032_public_keys(oracle):
    """k*B for the RFC 8032 §7.1 secret scalars must encode to the published public keys."""
    for sk, pk in pyref.RFC8032_VECTORS:
        k = pyref.rfc8032_scalar(bytes.fromhex(sk))
        out = np.zeros(16, dtype=np.uint64)
        oracle.ge25519_scalarmult_base(ob.ptr(out), k.to_bytes(32, "little"))
        assert _pack(oracle, out).hex() == pk
        assert oracle.ge25519_is_on_curve(ob.ptr(out)) == 1


def test_group_order_and_torsion(oracle):
    lB = _scalarmult(oracle, L, pyref.B)
    assert oracle.ge25519_is_identity(ob.ptr(lB)) == 1
    # a point of order 8: decode of y = 0x7a03... is not needed; use (L * Q) for a random curve point Q
    y = 3
    while pyref.recover_x(y, 0) is None:
        y += 1
    Q = (pyref.recover_x(y, 0), y)
    t = _scalarmult(oracle, L, Q)  # lies in the 8-torsion subgroup
    t8 = _scalarmult(oracle, 8, ob.ge_to_affine(t))
    assert oracle.ge25519_is_identity(ob.ptr(t8)) == 1


def test_scalarmult_add_pack_unpack_match_bigint(oracle):
    rng = random.Random(5)
    pts = [pyref.B, pyref.pt_mul(7, pyref.B), pyref.pt_mul(rng.getrandbits(200), pyref.B)]
    for pt in pts:
        for k in [0, 1, 2, L - 1, L, 2**255 - 20, 2**256 - 1, rng.getrandbits(256)]:
            got = _scalarmult(oracle, k, pt)
            assert ob.ge_to_affine(got) == pyref.pt_mul(k, pt)
            enc = _pack(oracle, got)
            assert enc == pyref.encode(pyref.pt_mul(k, pt))
            back = np.zeros(16, dtype=np.uint64)
            assert oracle.ge25519_unpack(ob.ptr(back), enc) == 1
            assert ob.ge_to_affine(back) == pyref.pt_mul(k, pt)
    a, b = ob.affine_to_ge(*pts[1]), ob.affine_to_ge(*pts[2])
    out = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_add(ob.ptr(out), ob.ptr(a), ob.ptr(b))
    assert ob.ge_to_affine(out) == pyref.pt_add(pts[1], pts[2])
    oracle.ge25519_add(ob.ptr(out), ob.ptr(a), ob.ptr(a))  # unified: doubling through add
    assert ob.ge_to_affine(out) == pyref.pt_add(pts[1], pts[1])


def test_unpack_rejects_invalid(oracle):
    back = np.zeros(16, dtype=np.uint64)
    bad = 0
    for y in range(2, 40):
        if pyref.recover_x(y, 0) is None:
            assert oracle.ge25519_unpack(ob.ptr(back), y.to_bytes(32, "little")) == 0
            bad += 1
    assert bad > 5
    assert oracle.ge25519_unpack(ob.ptr(back), (P + 3).to_bytes(32, "little")) == 0  # non-canonical y


def test_naive_msm_matches_bigint(oracle):
    rng = random.Random(6)
    n = 12
    ks = [rng.getrandbits(256) for _ in range(n)]
    pts = [pyref.pt_mul(rng.getrandbits(128) + 1, pyref.B) for _ in range(n)]
    sc = ob.ints_to_fe(ks)
    pv = np.stack([ob.affine_to_ge(*p) for p in pts])
    out = np.zeros(16, dtype=np.uint64)
    fv, pvv = ob.field_vector(sc), ob.point_vector(pv)
    oracle.point_vector_multi_scalar_mul(ob.ptr(out), C.byref(fv), C.byref(pvv))
    # the reference's scalar convention: k = canonical(fe25519_tobytes(scalar)), all 256 bits used
    want = pyref.msm([k % P for k in ks], pts)
    assert ob.ge_to_affine(out) == want
    assert ob.fe_to_int(out[8:12]) == 1  # returned normalised, as the reference's CPU MSM does


def test_sha256_and_challenge(oracle):
    import hashlib
    for msg in [b"", b"abc", b"a" * 55, b"b" * 56, b"c" * 64, b"d" * 200]:
        out = (C.c_uint8 * 32)()
        oracle.oracle_sha256(out, msg, len(msg))
        assert bytes(out) == hashlib.sha256(msg).digest()
    out = (C.c_uint8 * 32)()
    oracle.generate_challenge(out, b"xyz", 3, b"BulletproofYChal")
    want = bytearray(hashlib.sha256(b"BulletproofYChalxyz").digest())
    want[31] &= 0x7F
    assert bytes(out) == bytes(want)
