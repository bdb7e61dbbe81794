"""GPU parity for the octet-form arithmetic (csrc/fe8.cuh: one 32-bit word per lane, a warp shares each point
operation; used by the latency-bound chains — window combine, small MSMs, point folds) through the
bpk_debug_fe8_op_device hook, against Python integers / the affine group law of oracle/pyref.py."""
import random

import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyref

pytestmark = pytest.mark.gpu
P, L = pyref.P, pyref.L
MUL, ADD, SUB, DBL, PADD, PADD_CACHED, DBL_CHAIN, NORMALIZE = range(8)
EDGE = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**255, 2**256 - 1, 2**256 - 38, 2**255 - 20, 2**224 - 1,
        2**255 - 1, int("ffffffff" * 7, 16), (2**256 - 1) ^ (2**32 - 1), 2**128, 2**128 - 1]


def run(op, a, b, width):
    import torch
    import cudabulletproof_b200 as cbp
    n = a.shape[0]
    da = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(n, width * 8)).cuda()
    db = torch.from_numpy(np.ascontiguousarray(b).view(np.uint8).reshape(n, width * 8)).cuda() if b is not None else None
    out = torch.zeros_like(da)
    cbp.check(cbp.load().bpk_debug_fe8_op_device(op, da.data_ptr(), db.data_ptr() if db is not None else None, out.data_ptr(),
                                                 n, None), "bpk_debug_fe8_op_device")
    return out.cpu().numpy().view(np.uint64).reshape(n, width)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 4099])
def test_fe8_field_ops_bit_exact(n):
    rng = random.Random(80 + n)
    va = [rng.getrandbits(256) for _ in range(n)]
    vb = [rng.getrandbits(256) for _ in range(n)]
    for i in range(min(n, len(EDGE))):
        va[i] = EDGE[i]
        vb[i] = EDGE[-1 - i]
    for i in range(len(EDGE), min(n, 400)):  # runs of all-ones words: the cross-lane carry ripple
        lo, hi = sorted((rng.randrange(9), rng.randrange(9)))
        for j in range(lo, hi):
            va[i] |= 0xFFFFFFFF << (32 * j)
    a, b = ob.ints_to_fe(va), ob.ints_to_fe(vb)
    for op, f in ((MUL, lambda x, y: x * y), (ADD, lambda x, y: x + y), (SUB, lambda x, y: x - y)):
        got = run(op, a, b, 4)
        for i in range(n):
            assert ob.fe_to_int(got[i]) == f(va[i], vb[i]) % P, (op, i, hex(va[i]), hex(vb[i]))


def _points(rng, n):
    base = pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B)
    step = pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B)
    pts, cur = [], base
    for _ in range(n):
        pts.append(cur)
        cur = pyref.pt_add(cur, step)
    return pts


def test_ge8_point_ops_match_group_law():
    import torch
    import cudabulletproof_b200 as cbp
    rng = random.Random(88)
    n = 37
    pa, pb = _points(rng, n), _points(rng, n)
    pa[0], pb[1] = (0, 1), (0, 1)            # identity operands
    pb[2] = pa[2]                             # p + p through the unified addition
    pb[3] = pyref.pt_neg(pa[3])               # p + (-p) = identity
    a = np.stack([ob.affine_to_ge(*p) for p in pa])
    b = np.stack([ob.affine_to_ge(*p) for p in pb])
    # projective representatives: scale every coordinate by a random z (the debug hook of capi_msm.cu)
    da = torch.from_numpy(a.view(np.uint8).reshape(n, 128)).cuda()
    db = torch.from_numpy(b.view(np.uint8).reshape(n, 128)).cuda()
    lib = cbp.load()
    cbp.check(lib.bpk_debug_projectivize_device(da.data_ptr(), n, 11, None, 0, None), "projectivize")
    cbp.check(lib.bpk_debug_projectivize_device(db.data_ptr(), n, 12, None, 0, None), "projectivize")
    ap = da.cpu().numpy().view(np.uint64).reshape(n, 16)
    bp = db.cpu().numpy().view(np.uint64).reshape(n, 16)
    for A, B in ((a, b), (ap, bp)):
        for op, want in ((DBL, [pyref.pt_add(p, p) for p in pa]), (PADD, [pyref.pt_add(p, q) for p, q in zip(pa, pb)]),
                         (PADD_CACHED, [pyref.pt_add(p, q) for p, q in zip(pa, pb)]),
                         (DBL_CHAIN, [pyref.pt_mul(2**64, p) for p in pa]), (NORMALIZE, pa)):
            got = run(op, A, B, 16)
            for i in range(n):
                x, y = want[i]
                assert np.array_equal(got[i], ob.affine_to_ge(x, y)), (op, i)  # canonical X, Y, Z = 1, T = xy
