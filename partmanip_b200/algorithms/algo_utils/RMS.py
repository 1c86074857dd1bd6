"""Running mean/std observation normaliser — mirror of algorithms/algo_utils/RMS.py:3-45 on the K6 kernels.

The update formula is the reference's non-standard one; there is no epsilon in the divide.
With torch.distributed initialised the batch statistics are all-reduced so every rank holds the same
running stats; `count` is then the global env count.
"""
from __future__ import annotations

import torch

from ... import ops, parallel


class RunningMeanStd:
    def __init__(self, shape, device):
        self.n = 0
        self.mean = torch.zeros((1, shape), device=device)
        self.S = torch.ones((1, shape), device=device) * 1e-4
        self.std = torch.sqrt(self.S)
        self._colsum = torch.empty(shape, device=device)
        self._sqdev = torch.empty(shape, device=device)

    def update(self, x):
        """RMS.py:10-18."""
        self.n += 1
        count = parallel.global_count(x.shape[0])
        ops.rms_colsum(x, self._colsum)
        parallel.all_reduce_sum_(self._colsum)
        ops.rms_colsqdev(x, self._colsum, count, self._sqdev)
        parallel.all_reduce_sum_(self._sqdev)
        ops.rms_update(self.mean, self.S, self.std, self._colsum, self._sqdev, count, self.n)

    def load(self, load_dict):
        dev = self.mean.device
        self.mean = load_dict['mean'].to(dev).float().contiguous()
        self.std = load_dict['std'].to(dev).float().contiguous()
        self.S = load_dict['S'].to(dev).float().contiguous()
        self.n = load_dict['n']

    def save(self):
        return {'mean': self.mean, 'std': self.std, 'S': self.S, 'n': self.n}


class Normalization:
    def __init__(self, shape, device):
        self.running_ms = RunningMeanStd(shape=shape, device=device)

    def __call__(self, x, update=True, out=None):
        """RMS.py:40-45.  `out` (optional) receives the result — the PPO runner passes the rollout-buffer slot so
        the normalised observation is written exactly once."""
        if update:
            self.running_ms.update(x)
        if out is None:
            out = torch.empty_like(x)
        ops.rms_normalize(x, out, self.running_ms.mean, self.running_ms.std)
        return out
