"""Barcode-group sharding over the GPUs of one box.

Barcode groups are independent units of work (reference: inference/lariat.go:375,436-437 hands one group to each worker
goroutine), so the multi-GPU form has no data-path collective: every rank owns a contiguous range of groups balanced by
read count, keeps its own replica of the index in HBM and returns results in barcode order.  torch.distributed is used for
the rendezvous, the barrier around timed regions and the max-over-ranks reduction of device times only.
"""
from __future__ import annotations

import numpy as np


def shard_groups(group_sizes, world: int):
    """Contiguous ranges [lo, hi) of barcode groups per rank, balanced by the number of read pairs.

    group_sizes: pairs per barcode group, in barcode order.  Returns a list of `world` (lo, hi) tuples that cover
    range(len(group_sizes)) exactly once, in order.  Greedy prefix split at multiples of total/world."""
    sizes = np.asarray(group_sizes, dtype=np.int64)
    n = len(sizes)
    if world <= 0:
        raise ValueError("world must be positive")
    cum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the boundary closest to the target, never going backwards
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        cuts.append(max(cuts[-1], min(k, n)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def pair_range(group_sizes, lo: int, hi: int):
    """Read-pair index range [p0, p1) of barcode groups [lo, hi)."""
    sizes = np.asarray(group_sizes, dtype=np.int64)
    p0 = int(sizes[:lo].sum())
    return p0, p0 + int(sizes[lo:hi].sum())


def max_over_ranks(seconds: float, device=None) -> float:
    """Device/wall time of the slowest rank (all_reduce MAX); identity without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_ranges(lo: int, hi: int):
    """All ranks' (lo, hi) in rank order (used by tests and by an optional result gather)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [(lo, hi)]
    mine = torch.tensor([lo, hi], dtype=torch.int64)
    out = [torch.zeros(2, dtype=torch.int64) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [(int(t[0]), int(t[1])) for t in out]
