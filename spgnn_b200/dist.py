"""Sharding helpers for data-parallel-by-graph runs.  Trees are independent components, so a global
batch is split into contiguous blocks of tree indices; no data-path collective is needed for inference."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of rank `rank`; sizes differ by at most one and cover [0, n_items) exactly."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_nodes(n_nodes, world_size: int):
    """Ragged batches: contiguous blocks of trees balanced by NODE count (prefix-scan split).  Returns the
    world_size+1 tree boundaries."""
    n_nodes = torch.as_tensor(n_nodes, dtype=torch.int64)
    csum = torch.cumsum(n_nodes, 0)
    total = int(csum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        bounds.append(int(torch.searchsorted(csum, torch.tensor(target), right=True)))
    bounds.append(int(n_nodes.numel()))
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def allreduce_sum_(t: torch.Tensor, group=None):
    """In-place SUM all-reduce (the one collective of a training step: the flat gradient bucket, and the two loss
    sums); a no-op in single-process runs."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def global_weighted_ce(local_nll_sum: torch.Tensor, local_w_sum: torch.Tensor, group=None):
    """F.cross_entropy(..., weight=) over the GLOBAL batch from per-rank (Σ w·nll, Σ w): each rank's loss term is
    divided by the global Σw, so summed gradients equal the single-process gradient."""
    sums = torch.stack([local_nll_sum.detach(), local_w_sum.detach()]).double()
    allreduce_sum_(sums, group)
    return local_nll_sum / sums[1].to(local_nll_sum.dtype), sums[0] / sums[1]
