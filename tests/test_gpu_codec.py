"""GPU parity: batched point encoding / decoding and generator derivation against the
CPU oracle's ge25519_pack / ge25519_unpack / oracle_hash_to_point (curve25519_ops.cu:449-531 and
complete_bulletproof_test.cu:33-63 restated; the oracle itself is pinned to RFC 8032 in test_oracle_kat.py)."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyref
from tests.helpers import gen_points, seed32

pytestmark = pytest.mark.gpu

P = pyref.P


def oracle_pack(oracle, pts):
    out = np.zeros((len(pts), 32), dtype=np.uint8)
    for i in range(len(pts)):
        oracle.ge25519_pack(ob.ptr(out[i]), ob.ptr(pts[i]))
    return out


def oracle_unpack(oracle, enc):
    pts = np.zeros((len(enc), 16), dtype=np.uint64)
    ok = np.zeros(len(enc), dtype=np.uint8)
    for i in range(len(enc)):
        e = np.ascontiguousarray(enc[i])
        ok[i] = oracle.ge25519_unpack(ob.ptr(pts[i]), ob.ptr(e))
    return pts, ok


def projective(rng, x, y):
    z = rng.randrange(2, P)
    return np.concatenate([ob.int_to_fe(x * z % P), ob.int_to_fe(y * z % P), ob.int_to_fe(z),
                           ob.int_to_fe(x * y % P * z % P)])


def test_pack_matches_oracle_affine_projective_and_special(oracle):
    import cudabulletproof_b200 as cbp
    rng = random.Random(0xC0DEC)
    rows = []
    cur = pyref.B
    for i in range(300):  # more than one 256-thread CTA, Z = 1 and random Z mixed
        cur = pyref.pt_add(cur, pyref.pt_mul(rng.getrandbits(32) | 1, pyref.B))
        rows.append(ob.affine_to_ge(*cur) if i % 3 == 0 else projective(rng, *cur))
    rows.append(ob.affine_to_ge(0, 1))                       # identity
    rows.append(ob.affine_to_ge(0, P - 1))                   # order 2
    rows.append(np.zeros(16, dtype=np.uint64))               # Z = 0 garbage: the oracle encodes zeros
    pts = np.stack(rows)
    want = oracle_pack(oracle, pts)
    got = cbp.point_pack(pts).cpu().numpy()
    assert np.array_equal(got, want)


def test_unpack_matches_oracle_valid_and_invalid(oracle):
    import cudabulletproof_b200 as cbp
    rng = random.Random(0xDEC0DE)
    enc = []
    cur = pyref.B
    for _ in range(200):
        cur = pyref.pt_add(cur, pyref.pt_mul(rng.getrandbits(32) | 1, pyref.B))
        x, y = cur
        enc.append((y | ((x & 1) << 255)).to_bytes(32, "little"))
    for _ in range(200):  # random 32 bytes: about half are not on the curve
        enc.append(rng.getrandbits(256).to_bytes(32, "little"))
    enc.append((1).to_bytes(32, "little"))                    # identity, sign 0
    enc.append((1 | 1 << 255).to_bytes(32, "little"))         # x = 0 with sign 1: invalid
    enc.append((P - 1).to_bytes(32, "little"))                # (0, -1)
    enc.append(P.to_bytes(32, "little"))                      # non-canonical y = p
    enc.append((P + 1).to_bytes(32, "little"))                # non-canonical y = p + 1
    enc.append((2**255 - 1).to_bytes(32, "little"))           # non-canonical
    enc.append((0).to_bytes(32, "little"))                    # y = 0: x^2 = -1 ... valid (order 4)
    arr = np.frombuffer(b"".join(enc), dtype=np.uint8).reshape(-1, 32).copy()
    want_pts, want_ok = oracle_unpack(oracle, arr)
    got_pts, got_ok = cbp.point_unpack(arr)
    got_pts = got_pts.cpu().numpy().view(np.uint64).reshape(-1, 16)
    got_ok = got_ok.cpu().numpy()
    assert np.array_equal(got_ok, want_ok)
    assert 0 < int(want_ok.sum()) < len(want_ok)
    good = want_ok.astype(bool)
    assert np.array_equal(got_pts[good], want_pts[good])
    ident = ob.affine_to_ge(0, 1)
    assert all(np.array_equal(r, ident) for r in got_pts[~good])


def test_pack_unpack_round_trip_large():
    """size-independent property at 2^18 points: unpack(pack(P)) == normalised P"""
    import torch
    import cudabulletproof_b200 as cbp
    n = 1 << 18
    pts, _ = cbp.synth_points(n, seed=0xAB)  # normalised, canonical
    enc = cbp.point_pack(pts)
    back, ok = cbp.point_unpack(enc)
    assert bool(ok.all().item())
    assert torch.equal(back, pts)


@pytest.mark.parametrize("seed_byte,count", [(1, 64), (2, 64), (3, 1), (4, 1)])
def test_generator_derivation_matches_oracle(oracle, seed_byte, count):
    import cudabulletproof_b200 as cbp
    want = gen_points(oracle, seed_byte, count)
    got = cbp.derive_generators(seed32(seed_byte), count).cpu().numpy().view(np.uint64).reshape(count, 16)
    assert np.array_equal(got, want)
    # a window of indices starting elsewhere
    got2 = cbp.derive_generators(seed32(seed_byte), 3, first_index=count - 1).cpu().numpy().view(np.uint64).reshape(3, 16)
    assert np.array_equal(got2[0], want[count - 1])


def test_group_law_hook_matches_python():
    """ge_dbl / ge_add / 8P through the test hook against the big-integer model, on hash-derived and walk points"""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    rng = random.Random(3)
    pts, cur = [], pyref.B
    for _ in range(150):
        cur = pyref.pt_add(cur, pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B))
        pts.append(cur)
    pts += [(0, 1), (0, P - 1)]
    arr = np.stack([ob.affine_to_ge(*p) for p in pts])
    d = torch.from_numpy(arr.view(np.uint8).reshape(-1, 128)).cuda()
    for op, k in ((0, 2), (1, 2), (2, 8)):
        out = torch.empty_like(d)
        assert lib.bpk_debug_ge_op_device(op, d.data_ptr(), None, out.data_ptr(), len(pts), None) == 0
        o = out.cpu().numpy().view(np.uint64).reshape(-1, 16)
        for i, p in enumerate(pts):
            assert ob.ge_to_affine(o[i]) == pyref.pt_mul(k, p), (op, i)


def test_field_ops_on_compile_time_constants():
    """Regression: read-write inline-asm operands must be early-clobber.  With plain "+r" the squaring of the
    constant 1 returned 39 * (1 + 2^64 + 2^128 + 2^192): the compiler had given the zero limbs of the input and
    the zero accumulators one register (found through the generator derivation, where Z = 1 is a constant)."""
    import torch
    import cudabulletproof_b200 as cbp
    out = torch.zeros(48, dtype=torch.int32, device="cuda")
    assert cbp.load().bpk_debug_const_operands_device(out.data_ptr(), None) == 0
    w = out.cpu().numpy().view(np.uint32).reshape(6, 8)
    vals = [sum(int(w[r][i]) << (32 * i) for i in range(8)) for r in range(6)]
    assert vals == [1, 1, 2, 4, P - 1, (2 * 2 * pyref.D) % P]
