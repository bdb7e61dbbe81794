"""Protocol-level checks of the corrected oracle: configs C1 (16-bit, v=42, v=65536) and C2
(64-bit) of BASELINE.json — honest proofs accept, every tampered field rejects.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from tests.helpers import Gens, oracle_prove, oracle_verify


@pytest.fixture(scope="module")
def gens16(oracle):
    return Gens(oracle, 16)


def test_generators_on_curve_prime_order(oracle, gens16):
    from oracle import pyref
    for g in list(gens16.G) + list(gens16.H) + [gens16.g, gens16.h]:
        assert oracle.ge25519_is_on_curve(ob.ptr(g)) == 1
        out = np.zeros(16, dtype=np.uint64)
        oracle.ge25519_scalarmult(ob.ptr(out), pyref.L.to_bytes(32, "little"), ob.ptr(g))
        assert oracle.ge25519_is_identity(ob.ptr(out)) == 1


def test_c1_16bit_value_42_accepts(oracle, gens16):
    proof, V = oracle_prove(oracle, gens16, 42, seed=1)
    assert oracle_verify(oracle, gens16, proof, V)
    oracle.range_proof_free(C.byref(proof))


def test_c1_out_of_range_65536_rejected(oracle, gens16):
    proof, V = oracle_prove(oracle, gens16, 65536, seed=2)
    assert not oracle_verify(oracle, gens16, proof, V)
    oracle.range_proof_free(C.byref(proof))  # D20: initialised, safe to free


def test_tampering_any_field_rejects(oracle, gens16):
    proof, V = oracle_prove(oracle, gens16, 4242, seed=3)
    assert oracle_verify(oracle, gens16, proof, V)

    def flip(obj, field=None):
        tgt = getattr(obj, field) if field else obj
        limbs = tgt.limbs if hasattr(tgt, "limbs") else tgt.X.limbs
        limbs[0] ^= 4
        ok = oracle_verify(oracle, gens16, proof, V)
        limbs[0] ^= 4
        return ok

    for f in ["A", "S", "T1", "T2", "taux", "mu", "t", "V"]:
        assert not flip(proof, f), f
    ip = proof.ip_proof
    assert not flip(ip, "c")
    assert not flip(ip, "x")
    assert not flip(ip.a.elements[0])
    assert not flip(ip.b.elements[0])
    for j in range(ip.L_len):
        assert not flip(ip.L.elements[j])
        assert not flip(ip.R.elements[j])
    V2 = V.copy()
    V2[0] ^= 1
    assert not oracle_verify(oracle, gens16, proof, V2)
    assert oracle_verify(oracle, gens16, proof, V)
    oracle.range_proof_free(C.byref(proof))


@pytest.mark.parametrize("value", [0, 2**64 - 1])
def test_c2_64bit_roundtrip(oracle, value):
    gens = Gens(oracle, 64)
    proof, V = oracle_prove(oracle, gens, value, seed=0xB0070002)
    assert oracle_verify(oracle, gens, proof, V)
    proof.t.limbs[1] ^= 1
    assert not oracle_verify(oracle, gens, proof, V)
    oracle.range_proof_free(C.byref(proof))


def test_standalone_ipa_prove_verify(oracle, gens16):
    """inner_product_prove / inner_product_verify on their own (zero initial transcript, :589)."""
    import random
    from oracle import pyref
    rng = random.Random(9)
    n = 8
    a = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    b = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    G, H = gens16.G[:n].copy(), gens16.H[:n].copy()
    Q = gens16.h
    c = np.zeros(4, dtype=np.uint64)
    av, bv, Gv, Hv = ob.field_vector(a), ob.field_vector(b), ob.point_vector(G), ob.point_vector(H)
    oracle.field_vector_inner_product(ob.ptr(c), C.byref(av), C.byref(bv))
    assert ob.fe_to_int(c) == sum(ob.fe_to_int(x) * ob.fe_to_int(y) for x, y in zip(a, b)) % pyref.L
    # P = <a,G> + <b,H> + c Q
    P1, P2, P3, P = (np.zeros(16, dtype=np.uint64) for _ in range(4))
    oracle.point_vector_multi_scalar_mul(ob.ptr(P1), C.byref(av), C.byref(Gv))
    oracle.point_vector_multi_scalar_mul(ob.ptr(P2), C.byref(bv), C.byref(Hv))
    oracle.ge25519_scalarmult(ob.ptr(P3), ob.fe_to_int(c).to_bytes(32, "little"), ob.ptr(Q))
    oracle.ge25519_add(ob.ptr(P), ob.ptr(P1), ob.ptr(P2))
    oracle.ge25519_add(ob.ptr(P), ob.ptr(P), ob.ptr(P3))
    proof = ob.InnerProductProof()
    oracle.inner_product_prove(C.byref(proof), C.byref(av), C.byref(bv), C.byref(Gv), C.byref(Hv), ob.ptr(Q),
                               ob.ptr(c), bytes(32))
    assert oracle.inner_product_verify(C.byref(proof), ob.ptr(P), C.byref(Gv), C.byref(Hv), ob.ptr(Q))
    proof.a.elements[0].limbs[0] ^= 1
    assert not oracle.inner_product_verify(C.byref(proof), ob.ptr(P), C.byref(Gv), C.byref(Hv), ob.ptr(Q))
    oracle.inner_product_proof_free(C.byref(proof))
