"""Shared test helpers (CPU side): deterministic generators, proofs via the oracle prover."""
import ctypes as C

import numpy as np

from oracle import binding as ob


def seed32(b):
    s = bytearray(32)
    s[0] = b
    return bytes(s)


def gen_points(oracle, seed_byte, n):
    """On-curve prime-order generators with the reference's derivation labels
    (complete_bulletproof_test.cu:33-41,79-88 — seed byte + big-endian index), SURVEY.md §8d C1."""
    out = np.zeros((n, 16), dtype=np.uint64)
    for i in range(n):
        oracle.oracle_hash_to_point(ob.ptr(out[i]), seed32(seed_byte), i)
    return out


class Gens:
    def __init__(self, oracle, n):
        self.n = n
        self.G = gen_points(oracle, 0x01, n)
        self.H = gen_points(oracle, 0x02, n)
        self.g = gen_points(oracle, 0x03, 1)[0].copy()
        self.h = gen_points(oracle, 0x04, 1)[0].copy()
        self.Gv = ob.point_vector(self.G)
        self.Hv = ob.point_vector(self.H)


def oracle_prove(oracle, gens, value, seed, gamma=None):
    """Returns (RangeProof ctypes struct, V ndarray). Caller frees with oracle.range_proof_free."""
    oracle.oracle_seed_rng(seed)
    proof = ob.RangeProof()
    v = ob.int_to_fe(value)
    gam = ob.int_to_fe(gamma if gamma is not None else (0x1234567 + seed * 7919) % (2**252))
    oracle.generate_range_proof(C.byref(proof), ob.ptr(v), ob.ptr(gam), gens.n, C.byref(gens.Gv), C.byref(gens.Hv),
                                ob.ptr(gens.g), ob.ptr(gens.h))
    V = np.frombuffer(bytes(proof.V), dtype=np.uint64).copy()
    return proof, V


def oracle_verify(oracle, gens, proof, V):
    return bool(oracle.range_proof_verify(C.byref(proof), ob.ptr(V), gens.n, C.byref(gens.Gv), C.byref(gens.Hv),
                                          ob.ptr(gens.g), ob.ptr(gens.h)))


def flatten_proof(proof, n):
    """RangeProof -> flat uint64 record used by the device batch API (include/bpk.h, bpk_proof layout):
    V,A,S,T1,T2 (5x16) | taux,mu,t (3x4) | a,b,c,x (4x4) | L[k] (k x16) | R[k] (k x16)."""
    k = n.bit_length() - 1
    rec = np.zeros(5 * 16 + 7 * 4 + 2 * k * 16, dtype=np.uint64)
    raw = np.frombuffer(bytes(proof), dtype=np.uint64)
    rec[0:80] = raw[0:80]
    rec[80:92] = raw[80:92]
    ip = proof.ip_proof
    rec[92:96] = np.ctypeslib.as_array(ip.a.elements[0].limbs)
    rec[96:100] = np.ctypeslib.as_array(ip.b.elements[0].limbs)
    rec[100:104] = np.ctypeslib.as_array(ip.c.limbs)
    rec[104:108] = np.ctypeslib.as_array(ip.x.limbs)
    for j in range(k):
        rec[108 + 16 * j:108 + 16 * (j + 1)] = np.frombuffer(bytes(ip.L.elements[j]), dtype=np.uint64)
        rec[108 + 16 * (k + j):108 + 16 * (k + j + 1)] = np.frombuffer(bytes(ip.R.elements[j]), dtype=np.uint64)
    return rec


def dot_mod_l(sc_h, ks_h):
    """sum_i s_i * k_i mod l for (n, 4) uint64 little-endian 256-bit s_i and (n,) uint64 k_i, vectorised:
    32 x 32-bit partial products are exact in uint64; their low and high halves are summed separately so that
    no sum over up to 2^24 terms overflows."""
    from oracle import pyref
    s32 = np.ascontiguousarray(sc_h, dtype=np.uint64).view(np.uint32).reshape(-1, 8).astype(np.uint64)
    k32 = np.ascontiguousarray(ks_h, dtype=np.uint64).view(np.uint32).reshape(-1, 2).astype(np.uint64)
    mask = np.uint64(0xFFFFFFFF)
    total = 0
    for a in range(8):
        for b in range(2):
            prod = s32[:, a] * k32[:, b]
            lo, hi = int((prod & mask).sum(dtype=np.uint64)), int((prod >> np.uint64(32)).sum(dtype=np.uint64))
            total += (lo + (hi << 32)) << (32 * (a + b))
    return total % pyref.L


def base_multiple(oracle, k):
    """(k mod l) * B normalised, as a (16,) uint64 ge25519 (one CPU scalar multiplication by the oracle)"""
    from oracle import pyref
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), (k % pyref.L).to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    return want
