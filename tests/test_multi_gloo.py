"""CPU, world_size 2 over gloo: the N > 1 host logic — contiguous point-range sharding, the all-gather
of 128-byte partial points (same order on every rank) and the claim the multi-GPU path rests on:
the sum of the per-shard MSMs is the full MSM, bit-identical as a canonical encoding.  The per-shard
arithmetic is done by the CPU oracle here (no GPU in this container); on the GPU box the same plumbing
carries kernel results (bench.py --gpus N, tests/test_gpu_multi.py)."""
import ctypes as C
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cudabulletproof_b200.multi import all_gather_bytes, shard_range


def test_shard_range_partitions_exactly():
    for n in [0, 1, 5, 16, 1000, 2**20 + 3]:
        for world in [1, 2, 3, 4, 8]:
            cover = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                cover.append((lo, hi))
            assert cover[0][0] == 0 and cover[-1][1] == n
            assert all(cover[i][1] == cover[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cover]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import binding as ob
    from oracle import pyref
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    oracle = ob.load_oracle()
    rng = random.Random(99)  # same inputs on every rank
    sc = ob.ints_to_fe([rng.getrandbits(256) for _ in range(n)])
    pts, cur, step = [], pyref.pt_mul(5, pyref.B), pyref.pt_mul(11, pyref.B)
    for _ in range(n):
        pts.append(ob.affine_to_ge(*cur))
        cur = pyref.pt_add(cur, step)
    pts = np.stack(pts)
    lo, hi = shard_range(n, rank, world)
    part = np.zeros(16, dtype=np.uint64)
    if hi > lo:
        s_sh, p_sh = sc[lo:hi].copy(), pts[lo:hi].copy()
        fv, pv = ob.field_vector(s_sh), ob.point_vector(p_sh)
        oracle.point_vector_multi_scalar_mul(ob.ptr(part), C.byref(fv), C.byref(pv))
    else:
        oracle.ge25519_0(ob.ptr(part))
    gathered = all_gather_bytes(torch.from_numpy(part.view(np.uint8).copy()), world, dist=dist).numpy()
    total = np.zeros(16, dtype=np.uint64)
    oracle.ge25519_0(ob.ptr(total))
    for r in range(world):
        p_r = gathered[r].view(np.uint64).copy()
        oracle.ge25519_add(ob.ptr(total), ob.ptr(total), ob.ptr(p_r))
    oracle.ge25519_normalize(ob.ptr(total))
    full = np.zeros(16, dtype=np.uint64)
    fv, pv = ob.field_vector(sc), ob.point_vector(pts)
    oracle.point_vector_multi_scalar_mul(ob.ptr(full), C.byref(fv), C.byref(pv))
    q.put((rank, bytes(total.tobytes()), bytes(full.tobytes()), gathered.tobytes()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 24])
def test_point_range_sharding_world2_gloo(n):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[0][2], "sum of shard MSMs != full MSM"
    assert res[0][1] == res[1][1] and res[0][3] == res[1][3], "ranks disagree"
