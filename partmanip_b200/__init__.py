"""partmanip_b200 — B200-native (sm_100a) implementation of PartManip's vision-RL training hot path.

Scope: the PointNet/MLP actor-critic forward/backward, Gaussian policy head, PPO losses,
GAE, running mean/std observation normaliser, grad-clip + Adam and the rollout buffer — as hand-written
CUDA kernels behind a C-ABI (include/partmanip_b200.h), with `algorithms.{ppo}` / `algorithms.algo_utils`
mirroring the reference's plugin surface (train.py:68-72).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401  (fails loudly if the CUDA extension is missing)

__all__ = ["ops", "algorithms", "envs"]
