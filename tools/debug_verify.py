"""Debug aid: run batched verification on one oracle-made proof and compare every intermediate the
kernels leave in the workspace (VScal block, generator sums, window sums) with Python big-ints."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import cudabulletproof_b200 as cbp  # noqa: E402
from oracle import binding as ob  # noqa: E402
from oracle import pyref  # noqa: E402
from tests.helpers import Gens, flatten_proof, oracle_prove, oracle_verify  # noqa: E402

L, P = pyref.L, pyref.P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = n.bit_length() - 1
oracle = ob.load_oracle()
g = Gens(oracle, n)
proof, V = oracle_prove(oracle, g, 42, 11)
print("oracle verify:", oracle_verify(oracle, g, proof, V))
rec = flatten_proof(proof, n)
dg = cbp.Generators(g.G, g.H, g.g, g.h)
ver = cbp.RangeVerifier(dg, 1)
d = torch.from_numpy(rec.view(np.uint8).copy()).cuda().reshape(1, -1)
acc = ver(d).cpu().numpy()
print("gpu accept:", acc)
ws = ver.workspace.cpu().numpy()
vs = ws[:960].view(np.uint32)
print("valid flag:", vs[0])


def sc_at(i):
    return int.from_bytes(vs[8 + 8 * i: 16 + 8 * i].tobytes(), "little")


names = ["z", "z2", "x", "x2", "a", "b", "g1", "h1", "h2", "s0"] + [f"ypow{m}" for m in range(7)] + \
        [f"usq{j}" for j in range(6)] + [f"uinvsq{j}" for j in range(6)]
dev = {nm: sc_at(i) for i, nm in enumerate(names)}


def chal(fn, *args):
    out = (C.c_uint8 * 32)()
    fn(out, *args)
    return bytes(out)


w = rec
pt = lambda i: w[16 * i:16 * i + 16].copy()
yb = chal(oracle.generate_challenge_y, ob.ptr(pt(0)), ob.ptr(pt(1)), ob.ptr(pt(2)))
zb = chal(oracle.generate_challenge_z, yb)
xb = chal(oracle.generate_challenge_x, ob.ptr(pt(3)), ob.ptr(pt(4)))
y, z, x = (int.from_bytes(b, "little") % L for b in (yb, zb, xb))
taux, mu, t = (ob.fe_to_int(w[80 + 4 * i:84 + 4 * i]) % L for i in range(3))
a, b = ob.fe_to_int(w[92:96]) % L, ob.fe_to_int(w[96:100]) % L
fin = t.to_bytes(32, "little") + taux.to_bytes(32, "little") + mu.to_bytes(32, "little")
tr = chal(oracle.generate_challenge, fin, 96, b"BulletproofIP")
us = []
for j in range(k):
    Lx = ob.ge_to_affine(w[108 + 16 * j:124 + 16 * j])[0]
    Rx = ob.ge_to_affine(w[108 + 16 * (k + j):124 + 16 * (k + j)])[0]
    tr = chal(oracle.generate_challenge, tr + Lx.to_bytes(32, "little") + Rx.to_bytes(32, "little"), 96, b"InnerProductChal")
    us.append(int.from_bytes(tr, "little") % L)
inv = lambda v: pow(v, L - 2, L)
delta = ((z - z * z) * sum(pow(y, i, L) for i in range(n)) - pow(z, 3, L) * (2**n - 1)) % L
want = {"z": z, "z2": z * z % L, "x": x, "x2": x * x % L, "a": a, "b": b, "g1": (t - delta) % L, "h1": taux,
        "h2": (mu + a * b - t) % L, "s0": inv(np.prod([1]) * 1) if False else 0}
s0 = 1
for u in us:
    s0 = s0 * inv(u) % L
want["s0"] = s0
for m in range(k + 1):
    want[f"ypow{m}"] = pow(inv(y), 2**m, L)
for j in range(k):
    want[f"usq{j}"] = us[j] * us[j] % L
    want[f"uinvsq{j}"] = inv(us[j]) ** 2 % L
for nm, v in want.items():
    print(f"{nm:8s}", "OK" if dev[nm] == v else f"MISMATCH dev={dev[nm]:x} want={v:x}")

yi_ = inv(y)
dbg = {"ypow5": sum(pow(y, i, L) for i in range(n)) % L, "ypow6": delta, "usq4": (z - z * z) % L, "usq5": pow(z, 3, L),
       "uinvsq4": 2**n - 1, "uinvsq5": inv(us[0]) * inv(us[1]) % L}
for nm, v in dbg.items():
    print("DBG", nm, "OK" if dev[nm] == v else f"MISMATCH dev={dev[nm]:x} want={v:x}")
# generator sums
off_f = (960 + 255) // 256 * 256
fs = ws[off_f:off_f + 256].view(np.uint64)
gp, hp = ob.ge_to_affine(g.g), ob.ge_to_affine(g.h)
F1 = pyref.pt_add(pyref.pt_mul((t - delta) % L, gp), pyref.pt_mul(taux, hp))
print("F1", "OK" if ob.ge_to_affine(fs[0:16]) == F1 else "MISMATCH")
s = []
for i in range(n):
    v = 1
    for j in range(k):
        v = v * (us[j] if (i >> (k - 1 - j)) & 1 else inv(us[j])) % L
    s.append(v)
F2 = pyref.pt_mul(want["h2"], hp)
yi = inv(y)
for i in range(n):
    F2 = pyref.pt_add(F2, pyref.pt_mul((a * s[i] + z) % L, ob.ge_to_affine(g.G[i])))
    ch = ((b * s[n - 1 - i] - z * z * 2**i) * pow(yi, i, L) - z) % L
    F2 = pyref.pt_add(F2, pyref.pt_mul(ch, ob.ge_to_affine(g.H[i])))
print("F2", "OK" if ob.ge_to_affine(fs[16:32]) == F2 else "MISMATCH")
Vp, A, S, T1, T2 = (ob.ge_to_affine(pt(i)) for i in range(5))
Var1 = pyref.pt_add(pyref.pt_add(pyref.pt_mul(z * z % L, Vp), pyref.pt_mul(x, T1)), pyref.pt_mul(x * x % L, T2))
print("identity1 holds (python):", Var1 == F1)
Var2 = pyref.pt_add(A, pyref.pt_mul(x, S))
for j in range(k):
    Lp = ob.ge_to_affine(w[108 + 16 * j:124 + 16 * j])
    Rp = ob.ge_to_affine(w[108 + 16 * (k + j):124 + 16 * (k + j)])
    Var2 = pyref.pt_add(Var2, pyref.pt_add(pyref.pt_mul(us[j] ** 2 % L, Lp), pyref.pt_mul(inv(us[j]) ** 2 % L, Rp)))
print("identity2 holds (python):", Var2 == F2)
off_w = off_f + 256
wsums = ws[off_w:off_w + 2 * 64 * 128].view(np.uint64).reshape(2, 64, 16)
for idx, Var in ((0, Var1), (1, Var2)):
    accp = pyref.IDENT
    for wdx in range(63, -1, -1):
        for _ in range(4):
            accp = pyref.pt_add(accp, accp)
        accp = pyref.pt_add(accp, ob.ge_to_affine(wsums[idx, wdx]))
    print(f"window sums identity {idx + 1}:", "OK" if accp == Var else "MISMATCH")
flags = ws[off_w + 2 * 64 * 128: off_w + 2 * 64 * 128 + 2]
print("flags", flags)
