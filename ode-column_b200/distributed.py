"""Trial-parallel sharding: trials never interact, so each rank integrates a contiguous block
of trials with a replicated W_aug; training adds ONE all-reduce(sum) of the parameter gradients per optimizer step.
One process per GPU, ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_trials: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the trial axis owned by `rank`; blocks differ by at most one trial."""
    base, rem = divmod(num_trials, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_trials(x: torch.Tensor, rank: int, world_size: int, dim: int = 0) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[dim], rank, world_size)
    return x.narrow(dim, lo, hi - lo)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Sum .grad of every parameter over the ranks with ONE flattened all-reduce; returns the number of floats sent.
    The reference's grad masks (scripts/wta_ode.py:182-183, xor_ode.py:179-183, parity_ode.py:185-197) are applied
    afterwards, identically on every rank."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    at = 0
    for g in grads:
        g.copy_(flat[at:at + g.numel()].view_as(g))
        at += g.numel()
    return flat.numel()


def allreduce_tensor_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
