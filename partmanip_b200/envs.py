"""Synthetic zero-physics vec-env with the attribute/method surface the algo classes consume
(reference tasks/hand_base.py:252-290, 363-402; SURVEY.md §1 'Env' row and §8(d) 'Synthetic inputs').

It stands in for the closed Isaac Gym stepper in tests and in bench.py: observations come from a
pre-generated pool cycled per step so RNG cost is excluded from timings.  `host=True` keeps the pool in
pinned host memory and copies each step's observation host->device inside step()/reset() (the e2e
measurement); otherwise the pool lives in HBM.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class FakeVecEnv:
    def __init__(self, num_envs: int, obs_dim: int, num_actions: int, device, *, obs_mode: str = "obs",
                 cloud: bool = True, channels: int = 3, pool: int = 16, seed: int = 1234, succ_p: float = 0.0,
                 done_p: float = 0.05, host: bool = False, extra_obs: Optional[Dict[str, int]] = None,
                 max_episode_length: int = 200, shard=None):
        self.num_envs, self.num_actions, self.max_episode_length = num_envs, num_actions, max_episode_length
        self.device = torch.device(device)
        self.obs_mode = obs_mode
        self.num_obs = {obs_mode: obs_dim, "proprio_state": 0}
        self.extra = dict(extra_obs or {})
        self.num_obs.update(self.extra)
        self.train_test_flag = "train"
        self.host = host
        g = torch.Generator().manual_seed(seed)
        # shard=(rank, world): generate the GLOBAL env batch of num_envs*world envs and keep this rank's slice, so that
        # R ranks see exactly the data one process with R*num_envs envs sees (multi-GPU parity runs)
        srank, sworld = shard if shard is not None else (0, 1)
        E, D = num_envs * sworld, obs_dim

        def make_obs():
            if cloud:
                # xyz ~ U over the open_drawer TSDF box x,y in (-1,1), z in (0.05,2.05) (cfg/tasks/open_drawer.yaml:12-14);
                # 10 % of the points are exactly (0,0,0) like the env's invalid-point padding (utils/depth2tsdf.py:159)
                n = D // channels
                pc = torch.rand(E, n, channels, generator=g)
                pc[..., :2] = pc[..., :2] * 2 - 1
                pc[..., 2] = pc[..., 2] * 2 + 0.05
                pc[torch.rand(E, n, generator=g) < 0.1] = 0.0
                o = pc.reshape(E, n * channels)
                if o.shape[1] < D:
                    o = torch.cat([o, torch.randn(E, D - o.shape[1], generator=g)], dim=1)
                return o
            return torch.randn(E, D, generator=g)

        self._pool = [make_obs() for _ in range(pool)]
        self._extra_pool = {k: [torch.randn(E, d, generator=g) for _ in range(pool)] for k, d in self.extra.items()}
        self._rew = [torch.randn(E, generator=g) for _ in range(pool)]
        self._done = [torch.rand(E, generator=g) < done_p for _ in range(pool)]
        self._succ = [torch.rand(E, generator=g) < succ_p for _ in range(pool)]
        if sworld > 1:
            cut = lambda t: t[srank * num_envs:(srank + 1) * num_envs].contiguous()
            self._pool = [cut(t) for t in self._pool]
            self._extra_pool = {k: [cut(t) for t in v] for k, v in self._extra_pool.items()}
            self._rew, self._done, self._succ = [cut(t) for t in self._rew], [cut(t) for t in self._done], [cut(t) for t in self._succ]
        E = num_envs
        if host:
            self._pool = [t.pin_memory() for t in self._pool]
            # two device buffers + a copy stream: the H2D copy of the NEXT pool entry runs while the learner computes on the current
            # one.  Every copy still happens inside the caller's timed region and is counted in h2d_bytes; what the
            # overlap models is a simulator that renders step t+1 concurrently with the learner — the pool does not depend on the
            # actions, so the prefetch is exact here (a real env can only do this with one-step-stale actions).
            self._obs_dev = [torch.empty(E, D, device=self.device) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._copied = [None, None]                # event: buffer i holds its pool entry
            self._consumed = [None, None]              # event: the compute stream is done with buffer i
            self._slot = 0
            self._act_host = torch.empty(E, num_actions).pin_memory()
        else:
            self._pool = [t.to(self.device) for t in self._pool]
        self._extra_pool = {k: [t.to(self.device) for t in v] for k, v in self._extra_pool.items()}
        self._rew = [t.to(self.device) for t in self._rew]
        self._done = [t.to(self.device) for t in self._done]
        self._succ = [t.to(self.device) for t in self._succ]
        self._k = 0
        self.reset_succ = torch.zeros(E, dtype=torch.bool, device=self.device)
        self.rew_buf = torch.zeros(E, device=self.device)
        self.success = torch.zeros(1, device=self.device)
        self.progress_buf = torch.zeros(E, dtype=torch.long, device=self.device)
        self._zero = torch.zeros(1, device=self.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _obs(self):
        k = self._k % len(self._pool)
        self._k += 1
        if self.host:
            cur = torch.cuda.current_stream(self.device)
            i = self._slot
            if self._copied[i] is None:                 # first call: nothing prefetched yet
                self._issue_copy(i, k, cur)
            cur.wait_event(self._copied[i])             # the compute stream may read buffer i from here on
            o = self._obs_dev[i]
            # the OTHER buffer was handed out one call ago; everything the learner enqueued on it is ordered before this point
            j = i ^ 1
            self._consumed[j] = torch.cuda.Event()
            self._consumed[j].record(cur)
            self._issue_copy(j, (k + 1) % len(self._pool), cur)
            self._slot = j
        else:
            o = self._pool[k]
        d = {self.obs_mode: o}
        for name, v in self._extra_pool.items():
            d[name] = v[k]
        return d

    def _issue_copy(self, i, k, cur):
        """H2D copy of pool entry k into device buffer i on the copy stream, after the compute stream has finished with buffer i."""
        cs = self._copy_stream
        if self._consumed[i] is not None:
            cs.wait_event(self._consumed[i])
        with torch.cuda.stream(cs):
            self._obs_dev[i].copy_(self._pool[k], non_blocking=True)
            self._copied[i] = torch.cuda.Event()
            self._copied[i].record(cs)
        self.h2d_bytes += self._pool[k].numel() * 4

    def reset(self):
        return self._obs()

    def step(self, actions, save_image_path=None):
        k = self._k % len(self._pool)
        if self.host:   # the simulator consumes the actions on the host side of the boundary
            self._act_host.copy_(actions, non_blocking=True)
            self.d2h_bytes += actions.numel() * 4
        self.rew_buf = self._rew[k]
        self.reset_succ = self._succ[k]
        return self._obs(), self.rew_buf, self._done[k], {"succ_rate": self._zero}
