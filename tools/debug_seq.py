import ctypes as C, random, sys
import numpy as np
sys.path.insert(0, ".")
import cudabulletproof_b200 as cbp
from oracle import binding as ob, pyref
oracle = ob.load_oracle()
lib = cbp.load()
rng = random.Random(5)
def mk(n):
    base, step = pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B), pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B)
    pts, cur = [], base
    for _ in range(n):
        pts.append(ob.affine_to_ge(*cur)); cur = pyref.pt_add(cur, step)
    return ob.ints_to_fe([rng.getrandbits(252) for _ in range(n)]), np.stack(pts)
def omsm(sc, pts):
    out = np.zeros(16, dtype=np.uint64); fv, pv = ob.field_vector(sc), ob.point_vector(pts)
    oracle.point_vector_multi_scalar_mul(ob.ptr(out), C.byref(fv), C.byref(pv)); return out
for seq in ([8193, 64], [64, 8193, 64], [300, 64], [2000, 64], [4097,64]):
    for n in seq:
        sc, pts = mk(n)
        got = cbp.cuda_point_vector_multi_scalar_mul(sc, pts)
        ok = True
        if n <= 300:
            ok = np.array_equal(got, omsm(sc, pts))
        print(seq, n, "ok" if ok else "MISMATCH", "err", lib.bpk_last_error(), lib.bpk_last_cuda_error(), "zero" if not got.any() else "", flush=True)
