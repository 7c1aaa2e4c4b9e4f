"""Multi-process sharding of a query batch: one process per GPU, model replicated, no data-path collective.

The path is embarrassingly parallel: rank ``r`` of ``W`` evaluates the contiguous slice
``rank_range(n, r, W)`` of the batch on its own replica of the CPT arena; the only exchange is the gather of the
fp32 results (4 B per query) on the host, which is outside the hot path.  ``torch.distributed`` supplies the
plumbing (NCCL on the GPU box for barriers, gloo for the host gather and for the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def rank_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice of rank ``rank``: equal parts, remainder to the last rank (== ShardedModel.split)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base = n // world
    return rank * base, (n if rank == world - 1 else (rank + 1) * base)


def csr_slice(row_off: np.ndarray, entries: np.ndarray, a: int, b: int) -> Tuple[np.ndarray, np.ndarray]:
    """Rows ``[a, b)`` of a SPARSE (CSR) batch, re-based to start at entry 0."""
    row_off = np.asarray(row_off, dtype=np.uint32)
    e0, e1 = int(row_off[a]), int(row_off[b])
    return (row_off[a:b + 1] - np.uint32(e0)).astype(np.uint32), np.asarray(entries, dtype=np.uint32)[e0:e1]


def evaluate_sharded(evaluate: Callable[[int, int], np.ndarray], n: int, group=None,
                     gather_to: Optional[int] = None) -> Optional[np.ndarray]:
    """Run ``evaluate(a, b) -> fp32[b-a]`` on this rank's slice and gather all slices on the host.

    ``gather_to=None``: every rank gets the full result (all_gather); otherwise only that rank does.
    Uses the given process group (any backend that moves CPU tensors, e.g. gloo); with no initialised
    process group it is the single-process identity.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(evaluate(0, n), dtype=np.float32)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = rank_range(n, rank, world)
    local = np.ascontiguousarray(evaluate(a, b), dtype=np.float32)
    if local.shape != (b - a,):
        raise ValueError(f"evaluate returned shape {local.shape} for a slice of {b - a} queries")
    sizes = [rank_range(n, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, dtype=torch.float32)
    pad[: b - a] = torch.from_numpy(local)
    if gather_to is None:
        parts = [torch.empty(width, dtype=torch.float32) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    else:
        parts = [torch.empty(width, dtype=torch.float32) for _ in range(world)] if rank == gather_to else None
        dist.gather(pad, parts, dst=gather_to, group=group)
        if rank != gather_to:
            return None
    out = np.empty(n, dtype=np.float32)
    for (lo, hi), p in zip(sizes, parts):
        out[lo:hi] = p[: hi - lo].numpy()
    return out


def max_over_ranks(value: float, group=None, device=None) -> float:
    """Timing rule of bench.py: a multi-GPU step takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
