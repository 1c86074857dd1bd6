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


class SymmetricGrad:
    """A gradient buffer every rank can read with plain loads (NVLink peer memory), for the fused all-reduce + clip + Adam kernel
    (csrc/fused_step.cu).  Built on torch.distributed._symmetric_memory: the same allocation is mapped into every rank's
    address space; `grad_ptrs` / `flag_ptrs` are DEVICE arrays of the `world` peer pointers.  `create()` returns None when the
    process group cannot provide it (one process, gloo, no P2P): the caller then keeps the NCCL all-reduce path."""

    def __init__(self, buf, flags, hdl, fhdl):
        self.buf, self.flags, self._hdl, self._fhdl = buf, flags, hdl, fhdl
        self.grad_ptrs, self.flag_ptrs = int(hdl.buffer_ptrs_dev), int(fhdl.buffer_ptrs_dev)
        self.rank, self.world = int(hdl.rank), int(hdl.world_size)

    @staticmethod
    def create(numel: int, device):
        if world() == 1 or dist.get_backend() != "nccl" or not str(device).startswith("cuda"):
            return None
        try:
            import torch.distributed._symmetric_memory as symm
            group = dist.group.WORLD
            if hasattr(symm, "enable_symm_mem_for_group") and not symm.is_symm_mem_enabled_for_group(group.group_name):
                symm.enable_symm_mem_for_group(group.group_name)
            buf = symm.empty(int(numel), dtype=torch.float32, device=device)
            flags = symm.empty(64, dtype=torch.int32, device=device)
            buf.zero_()
            flags.zero_()
            hdl, fhdl = symm.rendezvous(buf, group), symm.rendezvous(flags, group)
            torch.cuda.synchronize()
            dist.barrier()                      # every rank's flags are zero before anybody's first fused step
            return SymmetricGrad(buf, flags, hdl, fhdl)
        except Exception as e:  # pragma: no cover - depends on the platform
            print(f"[partmanip_b200] symmetric memory unavailable ({type(e).__name__}: {e}); gradients go through NCCL all-reduce")
            return None

    def peers(self):
        return (self.grad_ptrs, self.flag_ptrs, self.flags, self.rank, self.world)
