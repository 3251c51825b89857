"""Multi-GPU plumbing (one process per GPU).  The path shards by independent objects — LPs of a
batch, IP / knapsack instances — so there is no data-path collective: each rank solves a
contiguous slice.  torch.distributed (NCCL on the GPU box, gloo in the CPU tests) only carries the
barrier, the max-over-ranks of the timings, the gather of per-instance results and the incumbent
max-reduction of a node-sharded tree."""
import numpy as np


def shard_range(count, rank, world):
    """Contiguous slice [lo, hi) of `count` units for `rank`; sizes differ by at most one."""
    base, extra = divmod(count, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def all_max(value, device="cpu"):
    """Max over ranks of a scalar (timings: the slowest rank defines the step)."""
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(value, device="cpu"):
    dist = _dist()
    if dist is None:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def share_incumbent(best, device="cpu"):
    """Element-wise max of the ranks' incumbent objectives (the reference maximises,
    R/Models/Branch&Bound.cs:182): every rank prunes against the global best."""
    dist = _dist()
    best = np.asarray(best, dtype=np.float64)
    if dist is None:
        return best.copy()
    import torch
    t = torch.from_numpy(best.copy()).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def gather_shards(local, count, device="cpu"):
    """Reassemble per-unit results (first axis = units of this rank's slice) on every rank."""
    dist = _dist()
    local = np.ascontiguousarray(local)
    if dist is None:
        return local
    import torch
    world = dist.get_world_size()
    sizes = [shard_range(count, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((width,) + local.shape[1:], dtype=local.dtype)
    pad[: local.shape[0]] = local
    mine = torch.from_numpy(pad).to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([parts[r].cpu().numpy()[: hi - lo] for r, (lo, hi) in enumerate(sizes)], axis=0)
