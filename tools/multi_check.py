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
    n_local = 1 << 14
    n = n_local * world
    pts, _ = cbp.synth_points(n, seed=4242, device=dev)      # identical global inputs on every rank
    sc = cbp.synth_scalars(n, seed=2424, bits=252, device=dev)
    lo, hi = shard_range(n, rank, world)
    sharded = ShardedMsm(n_local, world, dev)(sc[lo:hi].contiguous(), pts[lo:hi].contiguous()).cpu()
    single = cbp.Msm(n, device=dev)(sc, pts).cpu()
    torch.cuda.synchronize()
    ok = torch.equal(sharded, single)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_OK" if int(flag.item()) == 1 else "MULTI_MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
