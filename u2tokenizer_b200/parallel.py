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


# ------------------------------------------------------------------------------------------------
# Gradient exchange of the training step: ZeRO-1 form of data-parallel training -
# reduce-scatter the gradients, update the local 1/W shard of the parameters, all-gather the updated parameters.
# The reference gets this from DeepSpeed ZeRO-1 (config/ds_config.json:33-38: 2e8-element buckets, overlap_comm). Host
# logic only: layouts, buckets and collectives; the shard update itself is a callable (the fused AdamW kernel on the GPU,
# a plain restatement in the CPU tests). Backend-agnostic: NCCL uses reduce_scatter_tensor / all_gather_into_tensor, gloo
# (CPU tests) falls back to all_reduce + slice and list all_gather.
# ------------------------------------------------------------------------------------------------
class FlatLayout:
    """Placement of a list of parameters in one flat vector whose length is a multiple of world_size * align, so that
    every rank owns one contiguous, equally sized, `align`-element-aligned shard."""

    def __init__(self, numels: Sequence[int], world_size: int, align: int = 8):
        if world_size < 1 or align < 1:
            raise ValueError("world_size and align must be >= 1")
        self.numels = [int(n) for n in numels]
        self.world_size = world_size
        self.offsets, off = [], 0
        for n in self.numels:
            self.offsets.append(off)
            off += n
        self.used = off
        q = world_size * align
        self.total = (off + q - 1) // q * q if off else q
        self.shard = self.total // world_size

    def shard_bounds(self, rank: int) -> Tuple[int, int]:
        if not (0 <= rank < self.world_size):
            raise ValueError(f"bad rank {rank}")
        return rank * self.shard, (rank + 1) * self.shard

    def flatten(self, tensors: Sequence, out: torch.Tensor) -> torch.Tensor:
        """Copy tensors (None = an unused parameter: zeros, the reference relies on find_unused_parameters for
        `layer_linagg.linear_aggregator.{wv,dense}`, train_stage1.py:21) into `out` [total]; the padding tail is zeroed."""
        if out.numel() != self.total:
            raise ValueError(f"flat buffer has {out.numel()} elements, layout needs {self.total}")
        for t, off, n in zip(tensors, self.offsets, self.numels):
            if t is None:
                out[off:off + n].zero_()
            else:
                if t.numel() != n:
                    raise ValueError("tensor size does not match the layout")
                out[off:off + n].copy_(t.reshape(-1))
        out[self.used:].zero_()
        return out

    def unflatten(self, flat: torch.Tensor, tensors: Sequence[torch.Tensor]) -> None:
        for t, off, n in zip(tensors, self.offsets, self.numels):
            t.copy_(flat[off:off + n].view_as(t))


def plan_buckets(numels: Sequence[int], bucket_elems: int) -> List[List[int]]:
    """Indices of the parameters grouped into buckets of at most `bucket_elems` elements, walking the list BACKWARDS
    (gradients become ready in reverse order of use, so the first bucket to fill is the last layers': its collective can
    start while the rest of backward still runs). A parameter larger than the bucket gets a bucket of its own."""
    if bucket_elems < 1:
        raise ValueError("bucket_elems must be >= 1")
    buckets, cur, size = [], [], 0
    for i in range(len(numels) - 1, -1, -1):
        n = int(numels[i])
        if cur and size + n > bucket_elems:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(i)
        size += n
    if cur:
        buckets.append(cur)
    return buckets


def _group_world(group=None) -> Tuple[int, int]:
    """(rank, world size) INSIDE `group` (the default group when None)."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def reduce_scatter_mean(flat: torch.Tensor, layout: FlatLayout, group=None) -> torch.Tensor:
    """Mean over ranks of `flat` [total], returning this rank's shard [total / W] (a new tensor)."""
    rank, ws = _group_world(group)
    lo, hi = layout.shard_bounds(rank)
    if ws == 1:
        return flat[lo:hi].clone()
    if flat.is_cuda:  # NCCL path chosen from where the data lives, not from the backend's name
        out = torch.empty(layout.shard, dtype=flat.dtype, device=flat.device)
        dist.reduce_scatter_tensor(out, flat, op=dist.ReduceOp.SUM, group=group)
        return out.div_(ws)
    buf = flat.clone()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[lo:hi].div_(ws)


def all_gather_flat(shard: torch.Tensor, layout: FlatLayout, out: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate every rank's shard into `out` [total]."""
    rank, ws = _group_world(group)
    if ws == 1:
        out.copy_(shard)
        return out
    if shard.is_cuda:
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
        return out
    parts = [torch.empty_like(shard) for _ in range(ws)]
    dist.all_gather(parts, shard.contiguous(), group=group)
    for r, p in enumerate(parts):
        lo, hi = layout.shard_bounds(r)
        out[lo:hi].copy_(p)
    return out


class Zero1Step:
    """One optimiser step of ZeRO-1 data parallelism over `params`:
        flat gradient (unused parameters contribute zeros)  -> reduce-scatter (mean over ranks)
        -> update_fn(param_shard, grad_shard, state, step)  on this rank's 1/W shard, in place
        -> all-gather of the updated shards -> parameters.
    `update_fn` owns the optimiser arithmetic and its state (`state` is a dict this object keeps per rank; on the GPU it is
    the fused AdamW kernel with fp32 moments - optimiser state per rank drops to 1/W, which is what lets cfg 5 fit:
    SURVEY.md section 8e). The master copy of the shard is kept in `master_dtype` (fp32 by default).
    This class is the backend-agnostic reference of the exchange (gloo tests on CPU); the product's step -
    bucket-interleaved ownership, reduce-scatter overlapped with the backward, all-gather overlapped with the next
    forward - is `train.TrainEngine.optimizer_step`."""

    def __init__(self, params: Sequence[torch.Tensor], update_fn, master_dtype=torch.float32, align: int = 8, group=None):
        self.params = list(params)
        self.update_fn = update_fn
        self.group = group
        self.rank, ws = _group_world(group)
        self.layout = FlatLayout([p.numel() for p in self.params], ws, align)
        dev = self.params[0].device
        self.flat_grad = torch.zeros(self.layout.total, dtype=self.params[0].dtype, device=dev)
        self.flat_param = torch.zeros(self.layout.total, dtype=self.params[0].dtype, device=dev)
        self.layout.flatten([p.detach() for p in self.params], self.flat_param)
        lo, hi = self.layout.shard_bounds(self.rank)
        self.master = self.flat_param[lo:hi].to(master_dtype).clone()
        self.state, self.steps = {}, 0

    @torch.no_grad()
    def step(self) -> None:
        self.layout.flatten([None if p.grad is None else p.grad for p in self.params], self.flat_grad)
        g = reduce_scatter_mean(self.flat_grad, self.layout, self.group).to(self.master.dtype)
        self.steps += 1
        self.update_fn(self.master, g, self.state, self.steps)
        all_gather_flat(self.master.to(self.flat_param.dtype), self.layout, self.flat_param, self.group)
        self.layout.unflatten(self.flat_param, [p.data for p in self.params])
