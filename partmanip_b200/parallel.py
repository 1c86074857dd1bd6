"""Env-sharded data parallelism — the host-side collective contract.

The reference has no distributed code at all; this module is the single place where ranks exchange data.  Every rank
owns `num_envs` envs, its rollout buffer and a replica of both networks; per optimiser step the ranks exchange

  * the flat gradient buffer            (sum; the loss kernels already scale by 1/(B*world) so the sum IS the global mean),
  * [sum surrogate, sum KL]             (sum; makes the KL-skip decision identical on every rank),
  * per-feature observation statistics  (sum of column sums, then sum of squared deviations about the GLOBAL mean).

Backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests); tensors are reduced in place on whatever device they live.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks (no-op for a single process)."""
    if world() > 1:
        dist.all_reduce(t)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if world() > 1:
        dist.broadcast(t, src)
    return t


def inv_global_batch(local_batch: int) -> float:
    """1 / (B_local * world): the scale the loss kernels apply so that summed gradients equal the global-mean gradient."""
    return 1.0 / (local_batch * world())


def global_count(local_rows: int) -> float:
    """Rows behind an all-reduced column sum (RMS.py:14 `x.mean(dim=0)` over the GLOBAL env batch)."""
    return float(local_rows * world())
