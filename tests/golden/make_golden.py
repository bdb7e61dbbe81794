"""Generates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

msm_vectors.json   — small MSM instances with results computed by the independent Python big-int model
                     (oracle/pyref.py, affine arithmetic): pins BOTH the C oracle and the CUDA path.
proof_n16.json     — one 16-bit range proof (value 42, seed 1) as the flat record of include/bpk.h, produced
                     by the C oracle's prover; the device prover must reproduce these bytes and both verifiers
                     must accept them.
codec_vectors.json — generator derivation (hash -> decode -> x8) and point encodings computed by the independent
                     Python model (hashlib + big integers): pins the C oracle's oracle_hash_to_point /
                     ge25519_pack / ge25519_unpack and the device codec.
"""
import ctypes as C
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from oracle import pyref  # noqa: E402
from tests.helpers import Gens, flatten_proof, oracle_prove  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = random.Random(20261018)
    cases = []
    L, P = pyref.L, pyref.P
    for n, kind in [(1, "random"), (2, "random"), (5, "edge"), (17, "random"), (33, "random")]:
        pts = [pyref.pt_mul(rng.getrandbits(96) | 1, pyref.B) for _ in range(n)]
        if kind == "edge":
            ks = [0, 1, L, P - 1, 2**256 - 1]
        else:
            ks = [rng.getrandbits(256) for _ in range(n)]
        want = pyref.msm([k % P for k in ks], pts)  # the reference's scalar convention: k = s mod p, all bits
        cases.append({"n": n, "scalars": [hex(k) for k in ks], "points": [[hex(x), hex(y)] for x, y in pts],
                      "result_affine": [hex(want[0]), hex(want[1])], "result_encoding": pyref.encode(want).hex()})
    with open(os.path.join(HERE, "msm_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/pyref.py big-int model)", "cases": cases}, f, indent=1)

    oracle = ob.load_oracle()
    g = Gens(oracle, 16)
    proof, V = oracle_prove(oracle, g, 42, 1)
    rec = flatten_proof(proof, 16)
    oracle.range_proof_free(C.byref(proof))
    with open(os.path.join(HERE, "proof_n16.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (oracle/ref_corrected.c prover, value 42, seed 1, "
                                "gamma (0x1234567 + 7919) mod 2^252, generators tests/helpers.Gens(16))",
                   "n": 16, "value": 42, "seed": 1, "record_hex": rec.tobytes().hex()}, f, indent=1)
    import hashlib
    gens = []
    for seed_byte in (1, 2, 3, 4):
        seed = bytes([seed_byte]) + bytes(31)
        for idx in range(3):
            ctr = 0
            while True:
                msg = seed + idx.to_bytes(4, "big") + (ctr.to_bytes(4, "big") if ctr else b"")
                h = hashlib.sha256(msg).digest()
                yv = int.from_bytes(h, "little")
                sign = yv >> 255
                yv &= (1 << 255) - 1
                x = pyref.recover_x(yv, sign) if yv < P else None
                if x is not None:
                    pt = pyref.pt_mul(8, (x, yv))
                    if pt != (0, 1):
                        break
                ctr += 1
            gens.append({"seed_byte": seed_byte, "index": idx, "counter": ctr, "affine": [hex(pt[0]), hex(pt[1])],
                         "encoding": pyref.encode(pt).hex()})
    invalid = [P.to_bytes(32, "little").hex(), (1 | 1 << 255).to_bytes(32, "little").hex(), (2).to_bytes(32, "little").hex()]
    with open(os.path.join(HERE, "codec_vectors.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py (hashlib + oracle/pyref.py)", "generators": gens,
                   "invalid_encodings": invalid}, f, indent=1)
    print("golden files written")


if __name__ == "__main__":
    main()
