"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink) for the exchange.

The path shards naturally: a large MSM is partitioned by contiguous POINT RANGE, batch
verification by PROOF.  The only exchange is tiny — one 128-byte partial point (or one accept mask) per
rank — and NCCL has no elliptic-curve reduction, so it is an all-gather followed by a point-sum kernel.
Results are bit-identical for any world size because the group law is exact and the final point is
normalised to its canonical coordinates."""


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of rank's items; sizes differ by at most one."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_bytes(local, world, dist=None, out=None):
    """all-gather of equal-sized uint8 tensors -> (world, nbytes) tensor (same order on every rank)."""
    import torch
    if world == 1:
        return local.reshape(1, -1)
    if dist is None:
        import torch.distributed as dist
    if out is None:
        out = torch.empty((world, local.numel()), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.reshape(-1).contiguous())
    return out


class ShardedMsm:
    """MSM over world * n_local pairs, this rank holding pairs [rank*n_local, (rank+1)*n_local)."""

    def __init__(self, n_local, world, device, window_bits=0):
        import torch
        from .host import Msm
        self.world, self.device = world, device
        self.msm = Msm(n_local, device=device, window_bits=window_bits)
        self.partial = torch.zeros(128, dtype=torch.uint8, device=device)
        self.gathered = torch.zeros((world, 128), dtype=torch.uint8, device=device)

    def __call__(self, scalars, points):
        from .host import point_sum
        if self.world == 1:
            return self.msm(scalars, points, normalize=True)
        self.msm(scalars, points, normalize=False, out=self.partial)
        all_gather_bytes(self.partial, self.world, out=self.gathered)
        return point_sum(self.gathered, normalize=True)
