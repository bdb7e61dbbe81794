"""The protocol-level anchor outside the C oracle: tests/pyverify.py (plain Python integers, textbook protocol with
literal generator folding) against the C oracle's prover and verifier (CPU) and the device prover and verifiers (GPU)."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binding as ob
from tests import pyverify
from tests.helpers import Gens, flatten_proof, oracle_prove, oracle_verify


def _aff(g):
    return ob.ge_to_affine(g)


def _gens_py(g):
    return [_aff(p) for p in g.G], [_aff(p) for p in g.H], _aff(g.g), _aff(g.h)


@pytest.fixture(scope="module")
def gens8(oracle):
    return Gens(oracle, 8)


@pytest.mark.parametrize("n", [8, 16, 64])
def test_oracle_prover_accepted_by_the_independent_verifier(oracle, n):
    """BASELINE configs 0-1 (16-bit value 42 / out-of-range 65536, one 64-bit proof) and an 8-bit one"""
    gens = Gens(oracle, n)
    G, H, g, h = _gens_py(gens)
    for value, seed in ((42, 1), (0, 2), (2**n - 1, 3)):
        proof, V = oracle_prove(oracle, gens, value, seed)
        rec = flatten_proof(proof, n)
        assert oracle_verify(oracle, gens, proof, V)
        assert pyverify.verify(rec, n, G, H, g, h, V_ext=V)
        oracle.range_proof_free(C.byref(proof))
    if n < 64:
        proof, V = oracle_prove(oracle, gens, 2**n, 4)  # out of range (65536 at n = 16): must not verify anywhere
        assert not oracle_verify(oracle, gens, proof, V)
        assert not pyverify.verify(flatten_proof(proof, n), n, G, H, g, h)
        oracle.range_proof_free(C.byref(proof))


def test_tamper_matrix_same_decisions_as_the_c_oracle(oracle, gens8):
    """one bit flipped in every field of the record: the independent verifier and the C oracle reject all of them"""
    G, H, g, h = _gens_py(gens8)
    proof, V = oracle_prove(oracle, gens8, 0xA5, 9)
    rec = flatten_proof(proof, 8)
    oracle.range_proof_free(C.byref(proof))
    rng = random.Random(8)
    k = 3
    fields = [0, 16, 32, 48, 64, 80, 84, 88, 92, 96, 100, 104] + [108 + 16 * j for j in range(2 * k)]
    for off in fields:
        bad = rec.copy()
        bad[off + rng.randrange(2)] ^= np.uint64(1 << rng.randrange(40))
        assert not pyverify.verify(bad, 8, G, H, g, h), off


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 64])
def test_device_prover_and_verifiers_agree_with_the_independent_verifier(oracle, n):
    import cudabulletproof_b200 as cbp
    g = Gens(oracle, n)
    G, H, gg, hh = _gens_py(g)
    dg = cbp.Generators(g.G, g.H, g.g, g.h)
    rng = random.Random(0x1D + n)
    m = 6
    vals = [rng.getrandbits(n) for _ in range(m)]
    gam = ob.ints_to_fe([rng.getrandbits(250) for _ in range(m)])
    keys = np.frombuffer(bytes(rng.getrandbits(8) for _ in range(32 * m)), dtype=np.uint8).reshape(m, 32).copy()
    recs = cbp.range_prove_batch(dg, vals, gam, keys=keys).cpu().numpy()
    words = recs.view(np.uint64).reshape(m, -1).copy()
    # tamper half of them, one bit each, anywhere in the record
    for i in range(0, m, 2):
        words[i, rng.randrange(words.shape[1])] ^= np.uint64(1 << rng.randrange(64))
    import torch
    acc = cbp.RangeVerifier(dg, m)(torch.from_numpy(words.view(np.uint8).reshape(m, -1)).cuda()).cpu().numpy().astype(bool)
    want = [pyverify.verify(words[i], n, G, H, gg, hh) for i in range(m)]
    assert acc.tolist() == want
    assert want[1] and want[3] and want[5] and not any(want[0::2])
    # the host-pointer drop-in on one honest and one tampered record
    for i in (0, 1):
        proof, V, keep = cbp.record_to_range_proof(words[i].view(np.uint8), n)
        assert cbp.cuda_range_proof_verify(proof, V, n, g.G, g.H, g.g, g.h) == want[i]
