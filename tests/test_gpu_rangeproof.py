"""GPU parity for the range-proof path (BASELINE.json configs 1, 2, 5): batched prover and verifier
vs the CPU oracle (oracle/ref_corrected.c) — proofs byte-identical, accept/reject identical, every
tampered field rejected — plus the host-pointer drop-ins cuda_range_proof_verify /
cuda_inner_product_verify and the IPA folding kernels."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyref
from tests.helpers import Gens, flatten_proof, oracle_prove, oracle_verify

pytestmark = pytest.mark.gpu

L = pyref.L


@pytest.fixture(scope="module")
def gens16(oracle):
    return Gens(oracle, 16)


@pytest.fixture(scope="module")
def gens64(oracle):
    return Gens(oracle, 64)


_DEV_GENS = {}


def dev_gens(g, window_bits=8):
    """device generator tables, built once per (n, window width): the 16-bit tables are 6.5 GB at n = 64"""
    import cudabulletproof_b200 as cbp
    key = (len(g.G), window_bits)
    if key not in _DEV_GENS:
        _DEV_GENS[key] = cbp.Generators(g.G, g.H, g.g, g.h, window_bits=window_bits)
    return _DEV_GENS[key]


def gamma_for(seed):
    return (0x1234567 + seed * 7919) % (2**252)


@pytest.mark.parametrize("window_bits", [8, 16])
@pytest.mark.parametrize("n", [16, 64])
def test_gpu_prover_is_byte_identical_to_oracle(oracle, gens16, gens64, n, window_bits):
    import cudabulletproof_b200 as cbp
    g = gens16 if n == 16 else gens64
    dg = dev_gens(g, window_bits)
    cases = [(42, 1), (0, 2), (2**n - 1, 3), (0xBEEF & (2**n - 1), 4)]
    vals = [v for v, _ in cases]
    seeds = [s for _, s in cases]
    gam = ob.ints_to_fe([gamma_for(s) for s in seeds])
    got = cbp.range_prove_batch(dg, vals, gam, seeds).cpu().numpy()
    for i, (v, s) in enumerate(cases):
        proof, V = oracle_prove(oracle, g, v, s)
        want = flatten_proof(proof, n).view(np.uint8)
        assert np.array_equal(got[i], want), (n, v, s, np.nonzero(got[i] != want)[0][:8])
        oracle.range_proof_free(C.byref(proof))


def test_gpu_prover_out_of_range_value(oracle, gens16):
    """config 1: value 65536 with n = 16 must not yield a valid proof (D20: initialised, invalid)."""
    import cudabulletproof_b200 as cbp
    dg = dev_gens(gens16)
    gam = ob.ints_to_fe([gamma_for(2)])
    got = cbp.range_prove_batch(dg, [65536], gam, [2])
    proof, V = oracle_prove(oracle, gens16, 65536, 2)
    assert np.array_equal(got[0].cpu().numpy(), flatten_proof(proof, 16).view(np.uint8))
    ver = cbp.RangeVerifier(dg, 1)
    assert int(ver(got)[0]) == 0
    assert not oracle_verify(oracle, gens16, proof, V)
    oracle.range_proof_free(C.byref(proof))


def tamper_cases(rec_bytes, k, rng):
    """byte offsets covering every field of the record"""
    offs = [0, 32 + 5, 64, 96 + 31, 128 + 3, 256 + 40, 384 + 70, 512 + 100, 640, 672 + 8, 704 + 16, 736, 768 + 9, 800, 832 + 1]
    for j in range(2 * k):
        offs.append(864 + 128 * j + rng.randrange(0, 64))
    return offs


@pytest.mark.parametrize("window_bits", [8, 16])
@pytest.mark.parametrize("n", [16, 64])
def test_batch_verify_matches_oracle_honest_and_tampered(oracle, gens16, gens64, n, window_bits):
    import torch
    import cudabulletproof_b200 as cbp
    g = gens16 if n == 16 else gens64
    dg = dev_gens(g, window_bits)
    k = n.bit_length() - 1
    rng = random.Random(n)
    recs, expect, structs = [], [], []
    for s, v in [(11, 42), (12, 2**n - 1)]:
        proof, V = oracle_prove(oracle, g, v, s)
        base = flatten_proof(proof, n).view(np.uint8).copy()
        recs.append(base)
        expect.append(oracle_verify(oracle, g, proof, V))
        oracle.range_proof_free(C.byref(proof))
    assert expect == [True, True]
    base = recs[0]
    for off in tamper_cases(len(base), k, rng):
        bad = base.copy()
        bad[off] ^= 1 << rng.randrange(8)
        recs.append(bad)
        expect.append(False)
    d = torch.from_numpy(np.stack(recs)).cuda()
    ver = cbp.RangeVerifier(dg, len(recs))
    got = ver(d).cpu().numpy().astype(bool).tolist()
    assert got == expect
    # spot-check a few tampered records against the oracle itself through the drop-in struct path
    from tests.test_gpu_rangeproof import record_to_struct  # noqa
    for idx in [2, 5, 9, 12, len(recs) - 1]:
        proof, V, keep = record_to_struct(recs[idx], n)
        assert oracle_verify(oracle, g, proof, V) == got[idx]


@pytest.mark.parametrize("group,window_bits", [(2, 16), (3, 8), (8, 16), (12, 8), (64, 16)])
def test_grouped_verification_same_decisions_as_one_by_one(oracle, gens16, group, window_bits):
    """Grouped verification (one combined identity per `group` proofs, hash-derived weights, members of failed groups
    verified again one by one; BPK_OPT_VERIFY_GROUP) on a batch that mixes honest proofs, one bit flipped in every field
    of the record, and runs of consecutive honest proofs long enough to fill whole groups: decisions equal the oracle's
    and those of the one-by-one path, for groups that are all honest, all bad, mixed, and ragged at the end."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    n, g = 16, gens16
    dg = dev_gens(g, window_bits)
    k = n.bit_length() - 1
    rng = random.Random(1000 + group)
    honest = []
    for s, v in [(21, 7), (22, 2**n - 1), (23, 0), (24, 31337)]:
        proof, V = oracle_prove(oracle, g, v, s)
        assert oracle_verify(oracle, g, proof, V)
        honest.append(flatten_proof(proof, n).view(np.uint8).copy())
        oracle.range_proof_free(C.byref(proof))
    recs, expect = [], []
    for i in range(20):  # whole groups of honest proofs first
        recs.append(honest[i % 4])
        expect.append(True)
    for off in tamper_cases(len(honest[0]), k, rng):
        bad = honest[0].copy()
        bad[off] ^= 1 << rng.randrange(8)
        recs.append(bad)
        expect.append(False)
        if rng.random() < 0.5:
            recs.append(honest[rng.randrange(4)])
            expect.append(True)
    d = torch.from_numpy(np.stack(recs)).cuda()
    ver = cbp.RangeVerifier(dg, len(recs))
    try:
        cbp.check(lib.bpk_debug_set_option(13, 0), "set_option")  # BPK_OPT_VERIFY_GROUP: one by one
        plain = ver(d).cpu().numpy().astype(bool).tolist()
        cbp.check(lib.bpk_debug_set_option(13, group), "set_option")
        grouped = ver(d).cpu().numpy().astype(bool).tolist()
    finally:
        lib.bpk_debug_set_option(13, -1)
    assert plain == expect
    assert grouped == expect


def test_grouped_verification_across_the_pass_boundary(gens16):
    """More proofs than one pass of the batch verifier takes (2^14): the second pass starts a fresh set of groups; a ragged
    last group, tampered proofs on both sides of the boundary and in the last group.  Decisions equal the one-by-one path."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    dg = dev_gens(gens16, 8)
    m = (1 << 14) + 101
    rng = random.Random(77)
    vals = [rng.getrandbits(16) for _ in range(m)]
    seeds = list(range(5000, 5000 + m))
    gam = ob.ints_to_fe([gamma_for(s) for s in seeds])
    proofs = cbp.range_prove_batch(dg, vals, gam, seeds)
    bad = sorted({3, 16383, 16384, 16390, m - 1, *rng.sample(range(m), 40)})
    h = proofs.cpu().numpy()
    for i in bad:
        h[i, rng.randrange(h.shape[1])] ^= 1 << rng.randrange(8)
    d = torch.from_numpy(h).cuda()
    ver = cbp.RangeVerifier(dg, m)
    try:
        grouped = ver(d).cpu().numpy().astype(bool)
        cbp.check(lib.bpk_debug_set_option(13, 0), "set_option")  # BPK_OPT_VERIFY_GROUP: one by one
        plain = ver(d).cpu().numpy().astype(bool)
    finally:
        lib.bpk_debug_set_option(13, -1)
    assert [i for i in range(m) if not plain[i]] == bad
    assert np.array_equal(grouped, plain)


def record_to_struct(rec, n):
    """flat record -> ctypes RangeProof (oracle layout) + V, keeping the backing arrays alive"""
    k = n.bit_length() - 1
    w = np.ascontiguousarray(rec).view(np.uint64)
    proof = ob.RangeProof()
    C.memmove(C.byref(proof), w[0:92].tobytes(), 92 * 8)
    a, b = w[92:96].copy().reshape(1, 4), w[96:100].copy().reshape(1, 4)
    Ls = w[108:108 + 16 * k].copy().reshape(k, 16)
    Rs = w[108 + 16 * k:108 + 32 * k].copy().reshape(k, 16)
    ip = proof.ip_proof
    ip.n = n
    ip.a, ip.b = ob.field_vector(a), ob.field_vector(b)
    C.memmove(C.byref(ip.c), w[100:104].tobytes(), 32)
    C.memmove(C.byref(ip.x), w[104:108].tobytes(), 32)
    ip.L, ip.R = ob.point_vector(Ls), ob.point_vector(Rs)
    ip.L_len = k
    V = w[0:16].copy()
    return proof, V, (a, b, Ls, Rs)


def test_dropin_cuda_range_proof_verify_config1(oracle, gens16):
    """complete_bulletproof_test.cu:153,247: 16-bit proof of 42 accepts; the 65536 attempt rejects;
    a wrong external V rejects; wrong generator vector length rejects (nb:6669-6672)."""
    import cudabulletproof_b200 as cbp
    g = gens16
    proof, V = oracle_prove(oracle, g, 42, seed=1)
    assert cbp.cuda_range_proof_verify(proof, V, 16, g.G, g.H, g.g, g.h) is True
    assert cbp.cuda_range_proof_verify(proof, V, 16, g.G, g.H, g.g, g.h) is True  # cached generator tables
    V2 = V.copy()
    V2[1] ^= 2
    assert cbp.cuda_range_proof_verify(proof, V2, 16, g.G, g.H, g.g, g.h) is False
    assert cbp.cuda_range_proof_verify(proof, V, 16, g.G[:8], g.H, g.g, g.h) is False
    proof.mu.limbs[0] ^= 1
    assert cbp.cuda_range_proof_verify(proof, V, 16, g.G, g.H, g.g, g.h) is False
    oracle.range_proof_free(C.byref(proof))
    bad, Vb = oracle_prove(oracle, g, 65536, seed=2)
    assert cbp.cuda_range_proof_verify(bad, Vb, 16, g.G, g.H, g.g, g.h) is False
    oracle.range_proof_free(C.byref(bad))


@pytest.mark.parametrize("n", [2, 8, 64])
def test_dropin_cuda_inner_product_verify(oracle, gens64, n):
    import cudabulletproof_b200 as cbp
    rng = random.Random(90 + n)
    a = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    b = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    G, H, Q = gens64.G[:n].copy(), gens64.H[:n].copy(), gens64.h
    av, bv, Gv, Hv = ob.field_vector(a), ob.field_vector(b), ob.point_vector(G), ob.point_vector(H)
    c = np.zeros(4, dtype=np.uint64)
    oracle.field_vector_inner_product(ob.ptr(c), C.byref(av), C.byref(bv))
    P1, P2, P3, P = (np.zeros(16, dtype=np.uint64) for _ in range(4))
    oracle.point_vector_multi_scalar_mul(ob.ptr(P1), C.byref(av), C.byref(Gv))
    oracle.point_vector_multi_scalar_mul(ob.ptr(P2), C.byref(bv), C.byref(Hv))
    oracle.ge25519_scalarmult(ob.ptr(P3), ob.fe_to_int(c).to_bytes(32, "little"), ob.ptr(Q))
    oracle.ge25519_add(ob.ptr(P), ob.ptr(P1), ob.ptr(P2))
    oracle.ge25519_add(ob.ptr(P), ob.ptr(P), ob.ptr(P3))
    proof = ob.InnerProductProof()
    oracle.inner_product_prove(C.byref(proof), C.byref(av), C.byref(bv), C.byref(Gv), C.byref(Hv), ob.ptr(Q), ob.ptr(c),
                               bytes(32))
    assert oracle.inner_product_verify(C.byref(proof), ob.ptr(P), C.byref(Gv), C.byref(Hv), ob.ptr(Q))
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is True
    proof.b.elements[0].limbs[2] ^= 4
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is False
    proof.b.elements[0].limbs[2] ^= 4
    proof.L.elements[0].Y.limbs[0] ^= 1  # off-curve L
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is False
    proof.L.elements[0].Y.limbs[0] ^= 1
    P[5] ^= 1
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is False
    oracle.inner_product_proof_free(C.byref(proof))


@pytest.mark.parametrize("n_half", [1, 8, 32])
def test_ipa_fold_kernels_match_oracle(oracle, gens64, n_half):
    import torch
    import cudabulletproof_b200 as cbp
    rng = random.Random(400 + n_half)
    n = 2 * n_half
    a = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])  # unreduced inputs allowed
    b = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    u_int = rng.getrandbits(252) % L
    u, ui = ob.int_to_fe(u_int), ob.int_to_fe(pow(u_int, L - 2, L))
    G, H = gens64.G[:n].copy(), gens64.H[:n].copy()
    wa, wb = np.zeros((n_half, 4), np.uint64), np.zeros((n_half, 4), np.uint64)
    wG, wH = np.zeros((n_half, 16), np.uint64), np.zeros((n_half, 16), np.uint64)
    oracle.ipa_fold_scalars(ob.ptr(wa), ob.ptr(wb), ob.ptr(a), ob.ptr(b), n_half, ob.ptr(u), ob.ptr(ui))
    oracle.ipa_fold_points(ob.ptr(wG), ob.ptr(wH), ob.ptr(G), ob.ptr(H), n_half, ob.ptr(u), ob.ptr(ui))

    def dev(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).cuda()

    ga, gb = cbp.ipa_fold_scalars(dev(a), dev(b), dev(u), dev(ui))
    gG, gH = cbp.ipa_fold_points(dev(G), dev(H), dev(u), dev(ui))
    assert np.array_equal(ga.cpu().numpy().view(np.uint64).reshape(n_half, 4), wa)
    assert np.array_equal(gb.cpu().numpy().view(np.uint64).reshape(n_half, 4), wb)
    assert np.array_equal(gG.cpu().numpy().view(np.uint64).reshape(n_half, 16), wG)
    assert np.array_equal(gH.cpu().numpy().view(np.uint64).reshape(n_half, 16), wH)


def test_batch_verify_larger_batch_with_one_percent_tampered(oracle, gens64):
    """config 5 in miniature: 256 GPU-proved 64-bit proofs, a few tampered; decisions must equal the
    oracle's on a sample and all honest proofs must accept."""
    import torch
    import cudabulletproof_b200 as cbp
    dg = dev_gens(gens64, 16)
    m = 256
    rng = random.Random(5)
    vals = [rng.getrandbits(64) for _ in range(m)]
    seeds = list(range(1000, 1000 + m))
    gam = ob.ints_to_fe([gamma_for(s) for s in seeds])
    proofs = cbp.range_prove_batch(dg, vals, gam, seeds)
    bad_idx = sorted(rng.sample(range(m), 5))
    h = proofs.cpu().numpy()
    for i in bad_idx:
        h[i, rng.randrange(h.shape[1])] ^= 1 << rng.randrange(8)
    d = torch.from_numpy(h).cuda()
    ver = cbp.RangeVerifier(dg, m)
    got = ver(d).cpu().numpy().astype(bool)
    assert [i for i in range(m) if not got[i]] == bad_idx
    for i in bad_idx[:2] + [0, 7]:
        proof, V, keep = record_to_struct(h[i], 64)
        assert oracle_verify(oracle, gens64, proof, V) == bool(got[i])


def test_ipa_n4096_config4(oracle):
    """BASELINE config 4 (aggregated m = 64 x 64-bit: IPA over n = 4096, 12 rounds): one folding round of the
    a/b and G/H vectors at full size against the oracle, and stand-alone verification of an oracle-made
    4096-wide inner-product proof (accept, then reject after tampering)."""
    import torch
    import cudabulletproof_b200 as cbp
    from tests.helpers import gen_points
    rng = random.Random(0xA66E0040)
    n = 4096
    # generators: additive walk (cheap for the CPU oracle), all in the prime-order subgroup
    base, step = pyref.pt_mul(rng.getrandbits(128) | 1, pyref.B), pyref.pt_mul(rng.getrandbits(128) | 1, pyref.B)
    pts, cur = [], base
    for _ in range(2 * n + 1):
        pts.append(ob.affine_to_ge(*cur))
        cur = pyref.pt_add(cur, step)
    G, H, Q = np.stack(pts[:n]), np.stack(pts[n:2 * n]), pts[2 * n]
    a = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])
    b = ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)])

    def dev(x):
        return torch.from_numpy(x.view(np.uint8).reshape(-1)).cuda()

    # one folding round at full width
    u_int = rng.getrandbits(252) % L
    u, ui = ob.int_to_fe(u_int), ob.int_to_fe(pow(u_int, L - 2, L))
    nh = n // 2
    wa, wb = np.zeros((nh, 4), np.uint64), np.zeros((nh, 4), np.uint64)
    oracle.ipa_fold_scalars(ob.ptr(wa), ob.ptr(wb), ob.ptr(a), ob.ptr(b), nh, ob.ptr(u), ob.ptr(ui))
    ga, gb = cbp.ipa_fold_scalars(dev(a), dev(b), dev(u), dev(ui))
    assert np.array_equal(ga.cpu().numpy().view(np.uint64).reshape(nh, 4), wa)
    assert np.array_equal(gb.cpu().numpy().view(np.uint64).reshape(nh, 4), wb)
    sub = 64  # the point fold is 4 scalar multiplications per output on the CPU: check a 64-wide slice of it
    idx = np.r_[0:sub // 2, nh:nh + sub // 2]
    Gs, Hs = G[idx].copy(), H[idx].copy()
    wG, wH = np.zeros((sub // 2, 16), np.uint64), np.zeros((sub // 2, 16), np.uint64)
    oracle.ipa_fold_points(ob.ptr(wG), ob.ptr(wH), ob.ptr(Gs), ob.ptr(Hs), sub // 2, ob.ptr(u), ob.ptr(ui))
    gG, gH = cbp.ipa_fold_points(dev(G), dev(H), dev(u), dev(ui))
    assert np.array_equal(gG.cpu().numpy().view(np.uint64).reshape(nh, 16)[:sub // 2], wG)
    assert np.array_equal(gH.cpu().numpy().view(np.uint64).reshape(nh, 16)[:sub // 2], wH)

    # full 12-round proof from the oracle, verified by the single-MSM device verifier
    av, bv, Gv, Hv = ob.field_vector(a), ob.field_vector(b), ob.point_vector(G), ob.point_vector(H)
    c = np.zeros(4, dtype=np.uint64)
    oracle.field_vector_inner_product(ob.ptr(c), C.byref(av), C.byref(bv))
    assert np.array_equal(cbp.cuda_field_vector_inner_product(a, b), c)
    P1, P2, P3, P = (np.zeros(16, dtype=np.uint64) for _ in range(4))
    sc_all = np.concatenate([a, b, c.reshape(1, 4)])
    pt_all = np.concatenate([G, H, Q.reshape(1, 16)])
    P = cbp.cuda_point_vector_multi_scalar_mul(sc_all, pt_all)  # P = <a,G> + <b,H> + c Q (GPU MSM, 8193 points)
    s64, p64 = sc_all[:64].copy(), pt_all[:64].copy()  # keep alive: the vector structs hold raw pointers
    fv, pv = ob.field_vector(s64), ob.point_vector(p64)
    oracle.point_vector_multi_scalar_mul(ob.ptr(P1), C.byref(fv), C.byref(pv))
    assert np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(s64, p64), P1)
    proof = ob.InnerProductProof()
    oracle.inner_product_prove(C.byref(proof), C.byref(av), C.byref(bv), C.byref(Gv), C.byref(Hv), ob.ptr(Q), ob.ptr(c),
                               bytes(32))
    assert proof.L_len == 12
    # the same 12-round argument proved on the device: every L_j, R_j, the final a, b and the stored challenge
    dL, dR, fa, fb, fx = cbp.ipa_prove(G, H, Q, a, b)
    wantL = np.stack([np.frombuffer(bytes(proof.L.elements[j]), dtype=np.uint64) for j in range(12)])
    wantR = np.stack([np.frombuffer(bytes(proof.R.elements[j]), dtype=np.uint64) for j in range(12)])
    assert np.array_equal(dL.cpu().numpy().view(np.uint64).reshape(12, 16), wantL)
    assert np.array_equal(dR.cpu().numpy().view(np.uint64).reshape(12, 16), wantR)
    assert fa.cpu().numpy().tobytes() == bytes(proof.a.elements[0])
    assert fb.cpu().numpy().tobytes() == bytes(proof.b.elements[0])
    assert fx.cpu().numpy().tobytes() == bytes(proof.x)
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is True
    proof.R.elements[7].X.limbs[0] ^= 1
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is False
    proof.R.elements[7].X.limbs[0] ^= 1
    proof.a.elements[0].limbs[0] ^= 1
    assert cbp.cuda_inner_product_verify(proof, P, G, H, Q) is False
    oracle.inner_product_proof_free(C.byref(proof))


@pytest.mark.parametrize("n,window_bits", [(16, 8), (64, 16), (32, 16), (8, 8), (2, 16), (1, 8)])
def test_batched_prover_is_byte_identical(oracle, gens16, gens64, n, window_bits):
    """Batches of 64+ proofs take the phase-split prover (batch inversions across proofs, one window per lane);
    its records must equal the one-CTA-per-proof kernel's byte for byte (which is pinned to the oracle above),
    and a few are compared with the oracle directly.  n = 16 includes an out-of-range value."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    g = gens16 if n == 16 else gens64 if n == 64 else Gens(oracle, n)
    dg = dev_gens(g, window_bits)
    m = 150
    rng = random.Random(0xBA7C4 + n)
    vals = [rng.getrandbits(n) for _ in range(m)]
    if n == 16:
        vals[5] = 65536  # config 1: must not yield a valid proof
    seeds = list(range(7000, 7000 + m))
    gam = ob.ints_to_fe([gamma_for(s) for s in seeds])
    got = cbp.range_prove_batch(dg, vals, gam, seeds)  # workspace -> batched path
    # the same inputs through the single-kernel path (no workspace)
    d_v = torch.from_numpy(np.asarray(vals, dtype=np.uint64).view(np.int64)).cuda()
    d_s = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64)).cuda()
    d_g = torch.from_numpy(np.ascontiguousarray(gam, dtype=np.uint64).view(np.uint8).reshape(-1)).cuda()
    ref = torch.zeros_like(got)
    assert lib.bpk_range_prove_batch_device(dg.workspace.data_ptr(), d_v.data_ptr(), d_g.data_ptr(), d_s.data_ptr(), n, m,
                                            ref.data_ptr(), None, 0, None) == 0
    torch.cuda.synchronize()
    diff = (got != ref).any(dim=1).nonzero().flatten().tolist()
    if diff:
        i = diff[0]
        off = int((got[i] != ref[i]).nonzero()[0])
        raise AssertionError(f"{len(diff)} records differ; first: proof {i}, byte offset {off}")
    h = got.cpu().numpy()
    for i in (0, 77, m - 1):
        proof, V = oracle_prove(oracle, g, vals[i], seeds[i])
        want = flatten_proof(proof, n).view(np.uint8)
        assert np.array_equal(h[i], want)
        oracle.range_proof_free(C.byref(proof))
    acc = cbp.RangeVerifier(dg, m)(got).cpu().numpy().astype(bool)
    assert acc.tolist() == [not (n == 16 and i == 5) for i in range(m)]


@pytest.mark.parametrize("n", [2, 8, 64])
def test_device_ipa_prover_matches_oracle_small(oracle, gens64, n):
    """bpk_ipa_prove_device against inner_product_prove (bulletproof_vectors.cu:375-509 restated) at small widths,
    with a non-zero initial transcript"""
    import cudabulletproof_b200 as cbp
    rng = random.Random(0x1BA + n)
    G, H, Q = gens64.G[:n].copy(), gens64.H[:n].copy(), gens64.g.copy()
    a = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])  # unreduced inputs are reduced mod l first
    b = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])
    tr0 = bytes(rng.getrandbits(8) for _ in range(32))
    av, bv, Gv, Hv = ob.field_vector(a), ob.field_vector(b), ob.point_vector(G), ob.point_vector(H)
    c = np.zeros(4, dtype=np.uint64)
    oracle.field_vector_inner_product(ob.ptr(c), C.byref(av), C.byref(bv))
    proof = ob.InnerProductProof()
    oracle.inner_product_prove(C.byref(proof), C.byref(av), C.byref(bv), C.byref(Gv), C.byref(Hv), ob.ptr(Q), ob.ptr(c), tr0)
    k = n.bit_length() - 1
    dL, dR, fa, fb, fx = cbp.ipa_prove(G, H, Q, a, b, transcript0=tr0)
    for j in range(k):
        assert dL[j].cpu().numpy().tobytes() == bytes(proof.L.elements[j]), j
        assert dR[j].cpu().numpy().tobytes() == bytes(proof.R.elements[j]), j
    assert fa.cpu().numpy().tobytes() == bytes(proof.a.elements[0])
    assert fb.cpu().numpy().tobytes() == bytes(proof.b.elements[0])
    assert fx.cpu().numpy().tobytes() == bytes(proof.x)


def test_device_ipa_prover_hybrid_fold_then_unfolded_rounds(oracle, gens64):
    """Vectors longer than 4096 are first folded like the reference (bulletproof_vectors.cu:641-663, Pippenger MSMs for
    L and R), then the remaining rounds run unfolded over the folded generators (csrc/ipa_straus.cu).  Forced here at
    n = 64 by lowering the switch-over (BPK_OPT_IPA_COMPOSITE_MAX) to 16, 8 and 2: every split point gives the oracle's
    bytes."""
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    n = 64
    rng = random.Random(0x4B1D)
    G, H, Q = gens64.G[:n].copy(), gens64.H[:n].copy(), gens64.h.copy()
    a = ob.ints_to_fe([rng.getrandbits(253) for _ in range(n)])
    b = ob.ints_to_fe([rng.getrandbits(253) for _ in range(n)])
    tr0 = bytes(rng.getrandbits(8) for _ in range(32))
    av, bv, Gv, Hv = ob.field_vector(a), ob.field_vector(b), ob.point_vector(G), ob.point_vector(H)
    c = np.zeros(4, dtype=np.uint64)
    oracle.field_vector_inner_product(ob.ptr(c), C.byref(av), C.byref(bv))
    proof = ob.InnerProductProof()
    oracle.inner_product_prove(C.byref(proof), C.byref(av), C.byref(bv), C.byref(Gv), C.byref(Hv), ob.ptr(Q), ob.ptr(c), tr0)
    try:
        for comp_max in (16, 8, 2, -1):
            cbp.check(lib.bpk_debug_set_option(7, comp_max), "set_option")
            dL, dR, fa, fb, fx = cbp.ipa_prove(G, H, Q, a, b, transcript0=tr0)
            for j in range(6):
                assert dL[j].cpu().numpy().tobytes() == bytes(proof.L.elements[j]), (comp_max, j)
                assert dR[j].cpu().numpy().tobytes() == bytes(proof.R.elements[j]), (comp_max, j)
            assert fa.cpu().numpy().tobytes() == bytes(proof.a.elements[0])
            assert fb.cpu().numpy().tobytes() == bytes(proof.b.elements[0])
            assert fx.cpu().numpy().tobytes() == bytes(proof.x)
    finally:
        lib.bpk_debug_set_option(7, -1)


@pytest.mark.parametrize("n,m", [(16, 5), (64, 150)])
def test_keyed_prover_csprng_nonces(oracle, gens16, gens64, n, m):
    """bpk_range_prove_batch_keyed_device: blinding values and nonces are SHA-256("cbp-bp-nonce" || key || j) of a
    32-byte per-proof secret (the reference draws them from RAND_bytes, bulletproof_range_proof.cu:153), not the
    oracle's 64-bit SplitMix64 test stream.  Proofs must verify on the GPU and with the CPU oracle's
    range_proof_verify, be a function of the key, and be the same bytes through the one-CTA (m < 64) and the
    phase-split (m >= 64, workspace) provers."""
    import hashlib
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    g = gens16 if n == 16 else gens64
    dg = dev_gens(g, 8)
    rng = random.Random(0x5EC + n)
    vals = [rng.getrandbits(n) for _ in range(m)]
    gam = ob.ints_to_fe([rng.getrandbits(250) for _ in range(m)])
    keys = np.frombuffer(b"".join(hashlib.sha256(b"test key %d" % i).digest() for i in range(m)), dtype=np.uint8).reshape(m, 32).copy()
    got = cbp.range_prove_batch(dg, vals, gam, keys=keys)
    assert cbp.RangeVerifier(dg, m)(got).cpu().numpy().all()
    h = got.cpu().numpy()
    for i in sorted({0, m // 2, m - 1}):
        proof, V, keep = cbp.record_to_range_proof(h[i], n)
        assert oracle_verify(oracle, g, proof, V), i
    # alpha = draw 2n: A's blinding.  Recompute it from the key on the host and check it is NOT the seeded stream
    again = cbp.range_prove_batch(dg, vals, gam, keys=keys)
    assert torch.equal(got, again)
    keys2 = keys.copy()
    keys2[0, 31] ^= 1
    other = cbp.range_prove_batch(dg, vals, gam, keys=keys2).cpu().numpy()
    assert np.array_equal(other[0, :128], h[0, :128])          # V = v g + gamma h does not depend on the key
    assert not np.array_equal(other[0, 128:256], h[0, 128:256])  # A does
    assert np.array_equal(other[1:], h[1:])
    # no workspace -> one CTA per proof: same bytes
    d_v = torch.from_numpy(np.asarray(vals, dtype=np.uint64).view(np.int64)).cuda()
    d_g = torch.from_numpy(np.ascontiguousarray(gam, dtype=np.uint64).view(np.uint8).reshape(-1)).cuda()
    d_k = torch.from_numpy(keys).cuda()
    ref = torch.zeros_like(got)
    assert lib.bpk_range_prove_batch_keyed_device(dg.workspace.data_ptr(), d_v.data_ptr(), d_g.data_ptr(), d_k.data_ptr(), n, m,
                                                  ref.data_ptr(), None, 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    assert lib.bpk_range_prove_batch_keyed_device(dg.workspace.data_ptr(), d_v.data_ptr(), d_g.data_ptr(), None, n, m,
                                                  ref.data_ptr(), None, 0, None) != 0
    lib.bpk_clear_last_error()
