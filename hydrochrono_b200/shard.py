"""Ensemble partitioning across GPUs.

Ensemble members are independent (own eta, own velocity history, own state) and the hydro tables are small and
read-only, so the instances are split into contiguous blocks, one block per GPU / process, tables replicated, and
there is no collective on the step path.  The only exchanges are a final gather of per-instance results and the
max-over-ranks of the timing in bench.py; both go through torch.distributed (NCCL on GPUs, gloo in CPU tests).
"""
import numpy as np


def shard_range(total, world, rank):
    """Contiguous block [lo, hi) of `total` instances owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def instance_seeds(lo, hi, first_seed=1):
    """std::mt19937 seed of every global instance index in [lo, hi): first_seed + index."""
    return (first_seed + np.arange(lo, hi)).astype(np.int32)


def gather_results(local, total, world, rank, dist=None, device=None):
    """Final result gather: `local` is [n_local, ...] for this rank's block; rank 0 gets [total, ...] in global
    instance order (other ranks get None).  Uneven blocks are padded to the largest block for the collective."""
    import torch
    if world == 1:
        return np.asarray(local)
    local_t = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64)
    sizes = [shard_range(total, world, r)[1] - shard_range(total, world, r)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local_t.shape[1:]), dtype=torch.float64)
    buf[:local_t.shape[0]] = local_t
    if device is not None:
        buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o[:n].cpu().numpy() for o, n in zip(out, sizes)], axis=0)


def max_over_ranks(x, world, dist=None, device=None):
    import torch
    if world == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
