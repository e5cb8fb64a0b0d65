"""Data-parallel plumbing for the hot path (one process per GPU, torch.distributed).

Volumes are independent end to end: the path shards over the batch with NO data-path
collective. What remains is bookkeeping - which samples a rank owns, and device-time reductions for reporting
(max over ranks). Backend-agnostic so the host logic is covered on CPU with gloo (tests/test_parallel.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> range:
    """Contiguous balanced block of `n_items` owned by `rank` (sizes differ by at most one; rank r gets samples
    r*n/W .. (r+1)*n/W as in SURVEY.md section 8e)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    lo = n_items * rank // world_size
    hi = n_items * (rank + 1) // world_size
    return range(lo, hi)


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world_size: int) -> List[torch.Tensor]:
    """Slice every tensor's leading (batch) dim to this rank's block."""
    out = []
    for t in tensors:
        r = shard_range(t.shape[0], rank, world_size)
        out.append(t[r.start:r.stop])
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device-side elapsed ms): the number a multi-GPU benchmark must report."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_ids(ids: torch.Tensor) -> List[torch.Tensor]:
    """All ranks' generated ids (ragged lengths allowed), for reporting only; CPU tensors."""
    rank, ws = world()
    if ws == 1:
        return [ids.cpu()]
    objs = [None] * ws
    dist.all_gather_object(objs, ids.cpu())
    return objs
