"""GPU parity: Pippenger MSM vs the CPU oracle's naive MSM (bulletproof_vectors.cu:189-224 restated),
through the host-pointer drop-in cuda_point_vector_multi_scalar_mul, plus size-independent checks
at BASELINE.json's full size (2^20) through the device-resident API."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyref

pytestmark = pytest.mark.gpu

P, L = pyref.P, pyref.L


@pytest.fixture
def slot_option():
    """forces the slotted first digit pass off / on (bpk_debug_set_option; the library never reads the environment
    on a call path), restored to automatic afterwards"""
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    yield lambda v: cbp.check(lib.bpk_debug_set_option(0, v), "bpk_debug_set_option")
    lib.bpk_debug_set_option(0, -1)


def oracle_msm(oracle, sc, pts):
    out = np.zeros(16, dtype=np.uint64)
    fv, pv = ob.field_vector(sc), ob.point_vector(pts)
    oracle.point_vector_multi_scalar_mul(ob.ptr(out), C.byref(fv), C.byref(pv))
    return out


def curve_points(rng, n):
    base = pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B)
    step = pyref.pt_mul(rng.getrandbits(64) | 1, pyref.B)
    pts, cur = [], base
    for _ in range(n):
        pts.append(cur)
        cur = pyref.pt_add(cur, step)
    return pts


@pytest.mark.parametrize("n", [1, 2, 3, 16, 64, 147, 300])
def test_msm_matches_oracle_random(oracle, n):
    import cudabulletproof_b200 as cbp
    rng = random.Random(1000 + n)
    sc = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])
    pts = np.stack([ob.affine_to_ge(*p) for p in curve_points(rng, n)])
    for shared in (False, True):
        got = cbp.cuda_point_vector_multi_scalar_mul(sc, pts, shared=shared)
        assert np.array_equal(got, oracle_msm(oracle, sc, pts))


def test_msm_edge_scalars_and_points(oracle):
    """SURVEY.md §8d C3 edge suite: special scalars, all-equal scalars/points, (P,-P), identity, torsion."""
    import cudabulletproof_b200 as cbp
    rng = random.Random(7)
    special = [0, 1, L - 1, L, 2**252, 2**255 - 20, 2**255 - 19, 2**256 - 1, 2**255, 2**16, 2**16 - 1, 2**15, 2**15 + 1]
    pts = curve_points(rng, len(special))
    sc = ob.ints_to_fe(special)
    pv = np.stack([ob.affine_to_ge(*p) for p in pts])
    assert np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(sc, pv), oracle_msm(oracle, sc, pv))
    # all-equal scalars, all-equal points (one bucket per window gets everything)
    k = rng.getrandbits(255)
    sc = ob.ints_to_fe([k] * 40)
    pv = np.stack([ob.affine_to_ge(*pts[0])] * 40)
    assert np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(sc, pv), oracle_msm(oracle, sc, pv))
    # (P, -P) pairs cancel to the identity; identity inputs
    pv = np.stack([ob.affine_to_ge(*pts[1]), ob.affine_to_ge(*pyref.pt_neg(pts[1])), ob.affine_to_ge(0, 1)])
    sc = ob.ints_to_fe([k, k, 12345])
    got = cbp.cuda_point_vector_multi_scalar_mul(sc, pv)
    assert np.array_equal(got, oracle_msm(oracle, sc, pv))
    assert ob.ge_to_affine(got) == (0, 1)
    # non-normalised (Z != 1) inputs
    z = 0x1234567890ABCDEF1234567890ABCDEF
    proj = []
    for (x, y) in pts[:5]:
        proj.append(np.concatenate([ob.int_to_fe(x * z % P), ob.int_to_fe(y * z % P), ob.int_to_fe(z),
                                    ob.int_to_fe(x * y % P * z % P)]))
    pv = np.stack(proj)
    sc = ob.ints_to_fe([rng.getrandbits(256) for _ in range(5)])
    assert np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(sc, pv), oracle_msm(oracle, sc, pv))
    # a point with an 8-torsion component: k is NOT reduced mod l, so this must still match
    y = 3
    while pyref.recover_x(y, 0) is None:
        y += 1
    Q = (pyref.recover_x(y, 0), y)
    pv = np.stack([ob.affine_to_ge(*Q), ob.affine_to_ge(*pyref.pt_mul(L, Q))])
    sc = ob.ints_to_fe([L + 5, 7])
    got = cbp.cuda_point_vector_multi_scalar_mul(sc, pv)
    assert np.array_equal(got, oracle_msm(oracle, sc, pv))
    assert ob.ge_to_affine(got) == pyref.pt_add(pyref.pt_mul(L + 5, Q), pyref.pt_mul(7, pyref.pt_mul(L, Q)))


def test_msm_empty_and_mismatch(capfd):
    import cudabulletproof_b200 as cbp
    e_s, e_p = np.zeros((0, 4), np.uint64), np.zeros((0, 16), np.uint64)
    got = cbp.cuda_point_vector_multi_scalar_mul(e_s, e_p)
    assert ob.ge_to_affine(got) == (0, 1)
    res = np.full(16, 0xDEADBEEF, dtype=np.uint64)
    cbp.cuda_point_vector_multi_scalar_mul(ob.ints_to_fe([1, 2]), np.zeros((3, 16), np.uint64), result=res)
    assert (res == 0xDEADBEEF).all()  # cuda_bulletproof_kernels.cu:65-68
    assert "Vector lengths must match" in capfd.readouterr().err


@pytest.mark.parametrize("window_bits", [4, 7, 11, 13, 15, 16, 17])
def test_msm_all_window_sizes_agree(oracle, window_bits):
    import torch
    import cudabulletproof_b200 as cbp
    rng = random.Random(50 + window_bits)
    n = 200
    sc = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])
    pts = np.stack([ob.affine_to_ge(*p) for p in curve_points(rng, n)])
    want = oracle_msm(oracle, sc, pts)
    d_s = torch.from_numpy(sc.view(np.uint8).reshape(n, 32)).cuda()
    d_p = torch.from_numpy(pts.view(np.uint8).reshape(n, 128)).cuda()
    msm = cbp.Msm(n, window_bits=window_bits)
    got = msm(d_s, d_p).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("log_n", [12, 16, 20])
def test_msm_full_size_scalar_identity(oracle, log_n):
    """Size-independent check at full size: points P_i = k_i*B, so MSM(s, P) = (sum s_i k_i mod l) * B,
    one CPU scalar multiplication."""
    import torch
    import cudabulletproof_b200 as cbp
    n = 1 << log_n
    pts, ks = cbp.synth_points(n, seed=0xC3 + log_n)
    sc = cbp.synth_scalars(n, seed=0x5CA1A000 + log_n, bits=252)
    msm = cbp.Msm(n)
    got = msm(sc, pts).cpu().numpy().view(np.uint64).copy()
    torch.cuda.synchronize()
    ks_h = ks.cpu().numpy().astype(np.uint64)
    sc_h = sc.cpu().numpy().view(np.uint64).reshape(n, 4)
    acc = 0
    for i in range(n):
        acc += (int(sc_h[i, 0]) | int(sc_h[i, 1]) << 64 | int(sc_h[i, 2]) << 128 | int(sc_h[i, 3]) << 192) * int(ks_h[i])
    acc %= L
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), acc.to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    assert np.array_equal(got, want)
    # atomics reorder bucket contents between runs and the window groups are reduced on concurrent streams;
    # the group element must not change (this caught a buffer-sharing race between group tails once)
    for _ in range(6):
        again = msm(sc, pts).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, again)


def test_msm_adversarial_equal_scalars_large(oracle):
    """All scalars equal: every window has ONE bucket holding all N entries.  Segment splitting bounds the
    serial work per thread; the result must equal k * (sum of points) = k * (sum k_i) * B."""
    import torch
    import cudabulletproof_b200 as cbp
    n = 1 << 16
    pts, ks = cbp.synth_points(n, seed=77)
    k = 0x0123456789ABCDEF_FEDCBA9876543210_0F1E2D3C4B5A6978_13579BDF02468ACE % (2**253)
    one = torch.from_numpy(np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint8).copy()).cuda()
    sc = one.repeat(n, 1).contiguous()
    got = cbp.Msm(n)(sc, pts).cpu().numpy().view(np.uint64).copy()
    total = int(ks.cpu().numpy().astype(np.uint64).astype(object).sum()) % L
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), ((k % L) * total % L).to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    assert np.array_equal(got, want)


def test_msm_host_pointer_chunked_upload(oracle):
    """The host-pointer drop-in cuts inputs of >= 2^20 points into chunks that are multiplied while later
    chunks are still being uploaded; the sum of the partial MSMs must be the same group element
    (checked against one CPU scalar multiplication, ragged last chunk included)."""
    import cudabulletproof_b200 as cbp
    n = (1 << 20) + 12345
    pts, ks = cbp.synth_points(n, seed=0xE2E)
    sc = cbp.synth_scalars(n, seed=0x5EED, bits=253)
    h_pts = pts.cpu().numpy().view(np.uint64).reshape(n, 16)
    h_sc = sc.cpu().numpy().view(np.uint64).reshape(n, 4)
    got = cbp.cuda_point_vector_multi_scalar_mul(h_sc, h_pts)
    ks_h = ks.cpu().numpy().astype(np.uint64)
    acc = 0
    for i in range(n):
        acc += (int(h_sc[i, 0]) | int(h_sc[i, 1]) << 64 | int(h_sc[i, 2]) << 128 | int(h_sc[i, 3]) << 192) * int(ks_h[i])
    acc %= L
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), acc.to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    assert np.array_equal(got, want)
    # and it agrees with the device-resident single MSM
    dev = cbp.Msm(n)(sc, pts).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, dev)


@pytest.mark.parametrize("chunk_log2,taper_log2", [(0, 0), (16, 0), (17, 13), (18, 15)])
def test_msm_host_pointer_chunk_layouts_pinned(chunk_log2, taper_log2):
    """The chunked host path sorts every chunk's scalars ahead of its points (msm_run front / back split, one front
    workspace per chunk, shared buckets).  Pinned inputs (every copy queued before the first kernel), several chunk
    sizes, a tapered and ragged last chunk: the result is the device-resident MSM's, byte for byte, call after call."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    n = (1 << 20) + 4097
    pts, _ = cbp.synth_points(n, seed=0xE30)
    sc = cbp.synth_scalars(n, seed=0x5EEE, bits=253)
    want = cbp.Msm(n)(sc, pts).cpu().numpy().view(np.uint64)
    h_sc = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    h_sc.copy_(sc)
    h_pts = torch.empty((n, 128), dtype=torch.uint8).pin_memory()
    h_pts.copy_(pts)
    fv, pv = cbp.FieldVector(h_sc.data_ptr(), n), cbp.PointVector(h_pts.data_ptr(), n)
    out = np.zeros(16, dtype=np.uint64)
    try:
        cbp.check(lib.bpk_debug_set_option(2, chunk_log2), "set_option")   # BPK_OPT_HOST_CHUNK_LOG2
        cbp.check(lib.bpk_debug_set_option(10, taper_log2), "set_option")  # BPK_OPT_HOST_TAPER_LOG2
        for _ in range(2):  # the second call reuses the front workspaces of the first
            out[:] = 0
            lib.cuda_point_vector_multi_scalar_mul(out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
            assert np.array_equal(out, want)
    finally:
        lib.bpk_debug_set_option(2, 0)
        lib.bpk_debug_set_option(10, 0)


@pytest.mark.parametrize("n", [1, 300, (1 << 16) + 5, (1 << 20) + 4097])
def test_msm_affine_input_extension(n):
    """bpk_msm_device_affine / bpk_msm_host_affine (64-byte x || y points, an extension without a counterpart in the
    reference): the same bytes as the reference-layout MSM over the same points, on the device and through the
    host-pointer path (one piece below 2^20 pairs, chunked above), also when some coordinates are given as x + p."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    pts, _ = cbp.synth_points(n, seed=0xAFF + n)
    sc = cbp.synth_scalars(n, seed=0xAFE + n, bits=253)
    msm = cbp.Msm(n)
    want = msm(sc, pts).cpu().numpy().view(np.uint64).copy()
    xy = pts[:, :64].contiguous()
    h_xy = xy.cpu().numpy().copy()
    P = 2**255 - 19
    for i in range(0, n, max(1, n // 7)):  # weakly reduced containers: x + p, y + p
        for off in (0, 32):
            v = int.from_bytes(h_xy[i, off:off + 32].tobytes(), "little") + P
            h_xy[i, off:off + 32] = np.frombuffer(v.to_bytes(32, "little"), dtype=np.uint8)
    xy = torch.from_numpy(h_xy).cuda()
    got = msm.affine(sc, xy).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)
    h_sc = sc.cpu().numpy().copy()
    out = np.zeros(16, dtype=np.uint64)
    cbp.check(lib.bpk_msm_host_affine(out.ctypes.data_as(C.c_void_p), h_sc.ctypes.data_as(C.c_void_p),
                                      h_xy.ctypes.data_as(C.c_void_p), n), "bpk_msm_host_affine")
    assert np.array_equal(out, want)


def test_msm_host_pointer_chunked_equal_scalars(oracle):
    """chunks add into shared buckets (carry-in): with all scalars equal every chunk hits the same, split
    ("heavy") buckets, so the carry path of the split-bucket kernels is exercised"""
    import cudabulletproof_b200 as cbp
    n = (1 << 20) + 777
    pts, ks = cbp.synth_points(n, seed=0xE2F)
    k = 0x0F1E2D3C4B5A69788796A5B4C3D2E1F00123456789ABCDEFFEDCBA9876543210 % (2**253)
    h_pts = pts.cpu().numpy().view(np.uint64).reshape(n, 16)
    h_sc = np.tile(np.frombuffer(k.to_bytes(32, "little"), dtype=np.uint64), (n, 1))
    got = cbp.cuda_point_vector_multi_scalar_mul(h_sc, h_pts)
    total = int(ks.cpu().numpy().astype(np.uint64).astype(object).sum()) % L
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), ((k % L) * total % L).to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    assert np.array_equal(got, want)


@pytest.mark.parametrize("slots", ["0", "1"])
@pytest.mark.parametrize("window_bits", [7, 11])
def test_msm_slotted_front_end_small(oracle, slot_option, slots, window_bits):
    """The first digit pass places every window but the top one into fixed bucket slots (msm.cu,
    msm_digits_kernel<0>; default from 2^19 points, forced here).  Random scalars stay inside the slots;
    repeated scalars overflow them and take the exact two-pass placement: same bytes either way."""
    import torch
    import cudabulletproof_b200 as cbp
    slot_option(int(slots))
    rng = random.Random(900 + window_bits)
    n = 600
    pts = np.stack([ob.affine_to_ge(*p) for p in curve_points(rng, n)])
    d_p = torch.from_numpy(pts.view(np.uint8).reshape(n, 128)).cuda()
    k = rng.getrandbits(255)
    cases = {
        "random": [rng.getrandbits(256) for _ in range(n)],
        "mod_l": [rng.getrandbits(256) % L for _ in range(n)],
        "equal (every slot of the run overflows)": [k] * n,
        "half equal": [k if i % 2 else rng.getrandbits(253) for i in range(n)],
        "one overflowing low window": [(rng.getrandbits(240) << 16) | 0x1234 for _ in range(n)],
    }
    msm = cbp.Msm(n, window_bits=window_bits)
    for name, ints in cases.items():
        sc = ob.ints_to_fe(ints)
        d_s = torch.from_numpy(sc.view(np.uint8).reshape(n, 32)).cuda()
        got = msm(d_s, d_p).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, oracle_msm(oracle, sc, pts)), name


@pytest.mark.parametrize("slots", [0, 1])
def test_msm_fused_front_end_matches_separate_kernels(slot_option, slots):
    """The scans / segment build / histogram / scatter of the front end run as one cooperative launch
    (msm_front_tail_kernel) or as the eleven separate kernels it replaced (BPK_OPT_MSM_FUSED_FRONT = 0): same bytes, for
    random scalars, for repeated scalars (split buckets; with slots forced on, the overflow fallback) and for a size
    whose bucket count is not a multiple of the tile."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    slot_option(slots)
    n = (1 << 16) + 123
    pts, _ = cbp.synth_points(n, seed=0xF05ED)
    sc = cbp.synth_scalars(n, seed=0xF05EE, bits=253)
    adv = sc.clone()
    adv[: n // 3] = adv[1]
    out = {}
    try:
        for fused in (1, 0):
            cbp.check(lib.bpk_debug_set_option(14, fused), "set_option")  # BPK_OPT_MSM_FUSED_FRONT
            for wb in (0, 11):
                msm = cbp.Msm(n, window_bits=wb)
                out[(fused, wb)] = (msm(sc, pts).cpu().numpy().copy(), msm(adv, pts).cpu().numpy().copy())
                torch.cuda.synchronize()
    finally:
        lib.bpk_debug_set_option(14, 1)
    for wb in (0, 11):
        assert np.array_equal(out[(1, wb)][0], out[(0, wb)][0])
        assert np.array_equal(out[(1, wb)][1], out[(0, wb)][1])
    assert np.array_equal(out[(1, 0)][0], out[(1, 11)][0])


def test_msm_cached_launch_graph_replays(oracle):
    """Device MSMs of 2^13 < n < 2^19 pairs replay a CUDA graph of their launch DAG, cached by (buffers, n, window width)
    (msm_run_cached; BPK_OPT_MSM_GRAPH = 0: plain launches).  Same bytes as plain launches: on first use (capture), on
    replays after the buffers' CONTENTS changed, with more live (workspace, buffer) combinations than the cache holds
    (eviction), and after an option change that alters the DAG (stale graphs are not replayed)."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    n = (1 << 14) + 77
    pts, _ = cbp.synth_points(n, seed=0x6A)
    scs = [cbp.synth_scalars(n, seed=0x6B + i, bits=253) for i in range(3)]
    try:
        cbp.check(lib.bpk_debug_set_option(15, 0), "set_option")
        plain_msm = cbp.Msm(n)
        want = [plain_msm(s, pts).cpu().numpy().copy() for s in scs]
        cbp.check(lib.bpk_debug_set_option(15, 1), "set_option")
        msms = [cbp.Msm(n) for _ in range(10)]  # ten workspaces: more entries than the cache keeps
        buf = scs[0].clone()
        for rnd in range(3):
            buf.copy_(scs[rnd])  # same buffer, new contents: the replayed graph must read them
            for m in msms:
                got = m(buf, pts).cpu().numpy()
                assert np.array_equal(got, want[rnd])
        cbp.check(lib.bpk_debug_set_option(4, 0x666), "set_option")  # other window groups: a different DAG
        assert np.array_equal(msms[0](buf, pts).cpu().numpy(), want[2])
    finally:
        lib.bpk_debug_set_option(4, 0)
        lib.bpk_debug_set_option(15, 1)
    torch.cuda.synchronize()


def test_msm_slotted_equals_two_pass_at_size(oracle, slot_option):
    """2^18 points, both forced: the slotted and the two-pass front end give identical bytes, for
    252-bit random scalars and for an adversarial input that overflows the slots."""
    import torch
    import cudabulletproof_b200 as cbp
    n = 1 << 18
    pts, ks = cbp.synth_points(n, seed=0x51075)
    sc = cbp.synth_scalars(n, seed=0x51076, bits=252)
    adv = sc.clone()
    adv[: n // 2] = adv[0]  # half of the scalars equal: 2^17 entries in one bucket per window
    out = {}
    for slots in ("1", "0"):
        slot_option(int(slots))
        msm = cbp.Msm(n)
        out[slots] = (msm(sc, pts).cpu().numpy().copy(), msm(adv, pts).cpu().numpy().copy())
        torch.cuda.synchronize()
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.array_equal(out["0"][1], out["1"][1])
    # anchor on the oracle: sum s_i k_i * B
    ks_h = ks.cpu().numpy().astype(np.uint64)
    sc_h = sc.cpu().numpy().view(np.uint64).reshape(n, 4)
    acc = 0
    for i in range(n):
        acc += (int(sc_h[i, 0]) | int(sc_h[i, 1]) << 64 | int(sc_h[i, 2]) << 128 | int(sc_h[i, 3]) << 192) * int(ks_h[i])
    want = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_scalarmult_base(ob.ptr(want), (acc % L).to_bytes(32, "little"))
    oracle.ge25519_normalize(ob.ptr(want))
    assert np.array_equal(out["1"][0].view(np.uint64), want)


def test_msm_running_sum_reduction_cross_check(oracle):
    """The bucket reduction has two implementations: the shallow 2-D one (default for c >= 9) and the
    work-efficient running-sum levels (narrow windows; CBP_MSM_NO2D=1 forces them, read once per process).
    Both must give the oracle's bytes — single group (n < 2^15) and the window-group pipeline (2^16)."""
    import os
    import subprocess
    import sys
    import cudabulletproof_b200 as cbp
    script = (
        "import sys, numpy as np, torch; sys.path.insert(0, %r); import cudabulletproof_b200 as cbp\n"
        "for n, seed in ((3000, 5), (1 << 16, 6)):\n"
        "    pts, _ = cbp.synth_points(n, seed=seed); sc = cbp.synth_scalars(n, seed=seed + 100, bits=253)\n"
        "    print(cbp.Msm(n)(sc, pts).cpu().numpy().tobytes().hex())\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, CBP_MSM_NO2D="1")
    out = subprocess.run([sys.executable, "-c", script], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    forced = out.stdout.split()
    for (n, seed), hexed in zip(((3000, 5), (1 << 16, 6)), forced):
        pts, ks = cbp.synth_points(n, seed=seed)
        sc = cbp.synth_scalars(n, seed=seed + 100, bits=253)
        got = cbp.Msm(n)(sc, pts).cpu().numpy()
        assert got.tobytes().hex() == hexed
        ks_h = ks.cpu().numpy().astype(np.uint64)
        sc_h = sc.cpu().numpy().view(np.uint64).reshape(n, 4)
        acc = 0
        for i in range(n):
            acc += (int(sc_h[i, 0]) | int(sc_h[i, 1]) << 64 | int(sc_h[i, 2]) << 128 | int(sc_h[i, 3]) << 192) * int(ks_h[i])
        want = np.zeros(16, dtype=np.uint64)
        oracle.ge25519_scalarmult_base(ob.ptr(want), (acc % L).to_bytes(32, "little"))
        oracle.ge25519_normalize(ob.ptr(want))
        assert np.array_equal(got.view(np.uint64), want)


def test_msm_2_22_scalar_identity(oracle):
    """BASELINE's largest sweep size (2^22 points): c = 16, window groups 8, 4, 4 (the n >= 2^21 pipeline), slotted
    front end.  Checked against (sum s_i k_i mod l) * B — one CPU scalar multiplication."""
    import torch
    import cudabulletproof_b200 as cbp
    from tests.helpers import base_multiple, dot_mod_l
    n = 1 << 22
    pts, ks = cbp.synth_points(n, seed=0xC3 + 22)
    sc = cbp.synth_scalars(n, seed=0x5CA1A000 + 22, bits=252)
    msm = cbp.Msm(n)
    assert msm.window_bits == 16
    got = msm(sc, pts).cpu().numpy().view(np.uint64).copy()
    torch.cuda.synchronize()
    want = base_multiple(oracle, dot_mod_l(sc.cpu().numpy().view(np.uint64).reshape(n, 4), ks.cpu().numpy()))
    assert np.array_equal(got, want)
    assert np.array_equal(msm(sc, pts).cpu().numpy().view(np.uint64), want)  # and again (atomics reorder buckets)


def _torsion_point():
    """a point of order 8 (l * Q for the first curve point Q with a full torsion component), affine"""
    y = 3
    while True:
        x = pyref.recover_x(y, 0)
        if x is not None:
            t = pyref.pt_mul(L, (x, y))
            if pyref.pt_mul(4, t) != (0, 1):
                return t
        y += 1


@pytest.mark.parametrize("log_n", [16, 18, 20])
def test_msm_projective_and_torsion_inputs_at_size(oracle, log_n):
    """Full-size inputs OUTSIDE the Z = 1 fast path of msm_precompute_kernel: every point re-randomised to a
    pseudo-random Z (the Montgomery-trick branch), every 1000th point shifted by a point of order 8, scalars with
    all 255 bits.  k = s mod p is used as an integer, never reduced mod l (cuda_bulletproof_kernels.cu:33-37), so
    the expected value is (sum s_i k_i mod l) * B + (sum over shifted points of s_i mod 8) * T8."""
    import torch
    import cudabulletproof_b200 as cbp
    from tests.helpers import base_multiple, dot_mod_l
    n, stride = 1 << log_n, 1000
    pts, ks = cbp.synth_points(n, seed=0x7035 + log_n)
    sc = cbp.synth_scalars(n, seed=0x7036 + log_n, bits=255)
    t8 = _torsion_point()
    d_t8 = torch.from_numpy(ob.affine_to_ge(*t8).view(np.uint8).copy()).cuda()
    plain = cbp.Msm(n)(sc, pts).cpu().numpy().view(np.uint64).copy()
    cbp.check(cbp.load().bpk_debug_projectivize_device(pts.data_ptr(), n, 0xABCD + log_n, d_t8.data_ptr(), stride, None),
              "bpk_debug_projectivize_device")
    got = cbp.Msm(n)(sc, pts).cpu().numpy().view(np.uint64).copy()
    torch.cuda.synchronize()
    sc_h = sc.cpu().numpy().view(np.uint64).reshape(n, 4)
    assert (sc_h[:, 3] >> np.uint64(62)).max() > 0  # the top window is really populated
    base = base_multiple(oracle, dot_mod_l(sc_h, ks.cpu().numpy()))
    assert np.array_equal(plain, base)
    m8 = int((sc_h[::stride, 0] & np.uint64(7)).sum()) % 8  # s < p for these inputs, so s mod p = s
    want = pyref.pt_add(ob.ge_to_affine(base), pyref.pt_mul(m8, t8))
    assert ob.ge_to_affine(got) == want
    assert int(got[8]) == 1 and not got[9:12].any()  # normalised: Z = 1
    # the host-pointer drop-in sees the same projective inputs
    if log_n == 16:
        h = cbp.cuda_point_vector_multi_scalar_mul(sc_h, pts.cpu().numpy().view(np.uint64).reshape(n, 16))
        assert np.array_equal(h, got)


def test_msm_concurrent_host_threads_and_streams(oracle):
    """Four host threads, each with its own stream, inputs and workspace, call bpk_msm_device (2^16 points: the
    window-group pipeline with its per-device side streams and events) while a fifth calls the host-pointer
    drop-in.  The enqueue is serialised per device inside the library, so an event recorded by one call can never
    pair with another call's wait: every result must be its own expected point, every time."""
    import threading
    import torch
    import cudabulletproof_b200 as cbp
    from tests.helpers import base_multiple, dot_mod_l
    n, nthreads, reps = 1 << 16, 4, 12
    jobs = []
    for t in range(nthreads):
        pts, ks = cbp.synth_points(n, seed=0x7E4D + t)
        sc = cbp.synth_scalars(n, seed=0x7E5D + t, bits=253)
        want = base_multiple(oracle, dot_mod_l(sc.cpu().numpy().view(np.uint64).reshape(n, 4), ks.cpu().numpy()))
        jobs.append((pts, sc, want, cbp.Msm(n), torch.cuda.Stream()))
    hn = 3000
    h_pts_d, h_ks = cbp.synth_points(hn, seed=0x7E6D)
    h_sc_d = cbp.synth_scalars(hn, seed=0x7E7D, bits=253)
    h_pts = h_pts_d.cpu().numpy().view(np.uint64).reshape(hn, 16)
    h_sc = h_sc_d.cpu().numpy().view(np.uint64).reshape(hn, 4)
    h_want = base_multiple(oracle, dot_mod_l(h_sc, h_ks.cpu().numpy()))
    torch.cuda.synchronize()
    errors = []

    def device_worker(t):
        pts, sc, want, msm, stream = jobs[t]
        try:
            for _ in range(reps):
                out = torch.zeros(128, dtype=torch.uint8, device="cuda")
                with torch.cuda.stream(stream):
                    msm(sc, pts, out=out, stream=stream)
                stream.synchronize()
                if not np.array_equal(out.cpu().numpy().view(np.uint64), want):
                    errors.append(("device", t))
        except Exception as ex:  # noqa: BLE001
            errors.append(("device", t, repr(ex)))

    def host_worker():
        try:
            for _ in range(reps):
                if not np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(h_sc, h_pts), h_want):
                    errors.append(("host",))
        except Exception as ex:  # noqa: BLE001
            errors.append(("host", repr(ex)))

    threads = [threading.Thread(target=device_worker, args=(t,)) for t in range(nthreads)] + [threading.Thread(target=host_worker)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


@pytest.mark.parametrize("n", [1, 2, 31, 33, 64, 147, 257, 513, 1024, 1025])
def test_msm_small_path_matches_oracle_and_pippenger(oracle, n):
    """Up to 1024 points an MSM with the automatic window width is Straus in three launches (msm.cu section 8: the
    reference's production shape, cuda_bulletproof_kernels.cu:119-207); above, Pippenger.  Same bytes as the oracle's
    naive MSM and as the Pippenger path forced onto the same inputs (BPK_OPT_MSM_SMALL_MAX = 0), edge scalars and a
    torsion-carrying, projective point included, with and without normalisation."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    rng = random.Random(4000 + n)
    special = [0, 1, 8, 9, 2**255 - 20, 2**256 - 1, L, L - 1, 0x8888888888888888, 0x7777777777777777 << 192, 2**252, 16**63 * 8]
    ints = [rng.getrandbits(256) for _ in range(n)]
    for i, v in enumerate(special[:n]):
        ints[-1 - i] = v
    sc = ob.ints_to_fe(ints)
    pl = curve_points(rng, n)
    if n > 2:  # a point with a torsion component, given projectively
        y = 3
        while pyref.recover_x(y, 0) is None:
            y += 1
        pl[1] = (pyref.recover_x(y, 0), y)
    pts = np.stack([ob.affine_to_ge(*p) for p in pl])
    if n > 2:
        z = 0x1234567890ABCDEF1234567890ABCDEF1234567
        x, y = pl[1]
        pts[1] = np.concatenate([ob.int_to_fe(x * z % P), ob.int_to_fe(y * z % P), ob.int_to_fe(z), ob.int_to_fe(x * y % P * z % P)])
    want = oracle_msm(oracle, sc, pts)
    d_s = torch.from_numpy(sc.view(np.uint8).reshape(n, 32)).cuda()
    d_p = torch.from_numpy(pts.view(np.uint8).reshape(n, 128)).cuda()
    got = cbp.Msm(n)(d_s, d_p).cpu().numpy().view(np.uint64).copy()
    assert np.array_equal(got, want)
    raw = cbp.Msm(n)(d_s, d_p, normalize=False).cpu().numpy().view(np.uint64).copy()
    assert ob.ge_to_affine(raw) == ob.ge_to_affine(want)
    try:
        cbp.check(lib.bpk_debug_set_option(5, 0), "set_option")  # BPK_OPT_MSM_SMALL_MAX = 0: Pippenger for every n
        pip = cbp.Msm(n)(d_s, d_p).cpu().numpy().view(np.uint64).copy()
    finally:
        lib.bpk_debug_set_option(5, -1)
    assert np.array_equal(pip, want)
    assert np.array_equal(cbp.cuda_point_vector_multi_scalar_mul(sc, pts, shared=True), want)
