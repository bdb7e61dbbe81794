"""torchrun helper: every rank computes the same global MSM (a) sharded by point range across ranks and
(b) alone on its own GPU, and checks the two canonical results are identical bytes."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cudabulletproof_b200 as cbp  # noqa: E402
from cudabulletproof_b200.multi import ShardedMsm, shard_range  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (1) point-range-sharded MSM: Pippenger shards (2^14 and 2^17 per rank) and small shards on the Straus path
    for n_local in (1 << 14, 1 << 17, 300, 7):
        n = n_local * world
        pts, _ = cbp.synth_points(n, seed=4242 + n_local, device=dev)      # identical global inputs on every rank
        sc = cbp.synth_scalars(n, seed=2424 + n_local, bits=252, device=dev)
        lo, hi = shard_range(n, rank, world)
        sharded = ShardedMsm(n_local, world, dev)(sc[lo:hi].contiguous(), pts[lo:hi].contiguous()).cpu()
        single = cbp.Msm(n, device=dev)(sc, pts).cpu()
        torch.cuda.synchronize()
        ok = ok and torch.equal(sharded, single)
    # (2) proof-sharded batch verification: every rank proves the same global batch (deterministic prover), verifies
    # its contiguous shard, the accept masks meet through an all_gather; compared with one GPU verifying everything
    import numpy as np
    nb, per = 64, 96
    m = per * world
    gp, _ = cbp.synth_points(2 * nb + 2, seed=0xB0070002, device=dev)
    gens = cbp.Generators(gp[:nb], gp[nb:2 * nb], gp[2 * nb], gp[2 * nb + 1], device=dev)
    rng = np.random.default_rng(99)
    vals = rng.integers(0, 2**63, size=m, dtype=np.uint64)
    gam = rng.integers(0, 2**63, size=(m, 4), dtype=np.uint64)
    gam[:, 3] &= np.uint64((1 << 59) - 1)
    proofs = cbp.range_prove_batch(gens, vals, gam, np.arange(m, dtype=np.uint64))
    bad = rng.choice(m, size=max(2, m // 20), replace=False)
    hb = proofs.cpu().numpy()
    for i in bad:
        hb[i, rng.integers(0, hb.shape[1])] ^= np.uint8(1 << rng.integers(0, 8))
    proofs = torch.from_numpy(hb).to(dev)
    whole = cbp.RangeVerifier(gens, m)(proofs).clone()
    mine = cbp.RangeVerifier(gens, per)(proofs[rank * per:(rank + 1) * per].contiguous()).clone()
    gathered = torch.zeros((world, per), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered.view(-1), mine.contiguous())
    torch.cuda.synchronize()
    ok = ok and torch.equal(gathered.view(-1), whole) and int((whole == 0).sum()) == len(bad)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_OK" if int(flag.item()) == 1 else "MULTI_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
