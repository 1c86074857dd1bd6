"""ActorCritic — mirror of the reference class (algorithms/algo_utils/actor_critic.py:8-100): same constructor,
same method names / return tuples, same state_dict keys (`actor.*`, `critic.*`, `log_std`), with every method
body running in libpartmanip_b200.so kernels.

Reference quirks kept on purpose: Q1 sampling std = exp(log_std)^2, Q2 "sigma" = log_std repeated,
Q4 log-prob of atanh(clamp(stored action)) without a tanh-Jacobian term.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from ... import ops
from .network import MLP, Conv3DNet, PointNet, PoolConv3DNet  # noqa: F401  (resolved by name, like the reference's eval())

_NETWORKS = {"MLP": MLP, "PointNet": PointNet, "Conv3DNet": Conv3DNet, "PoolConv3DNet": PoolConv3DNet}


class ActorCritic(nn.Module):

    def __init__(self, obs_shape, actions_shape, model_cfg, proprio_shape=0):
        super().__init__()
        net_cfg = model_cfg['network']
        if net_cfg['name'] not in _NETWORKS:
            # the reference eval()s the name (actor_critic.py:16); ResNet / depthResNet are image students outside the hot path and its
            # next rows
            raise NotImplementedError(f"network {net_cfg['name']!r} is outside the B200 hot path (MLP, PointNet, Conv3DNet, PoolConv3DNet)")
        cls = _NETWORKS[net_cfg['name']]
        self.actor = cls(obs_shape, actions_shape, net_cfg, proprio_shape=proprio_shape)     # policy
        self.critic = cls(obs_shape, 1, net_cfg, proprio_shape=proprio_shape)                # value function
        with np.errstate(divide='ignore'):      # bc.yaml ships action_std 0.0: log_std = -inf, exactly as in the reference (actor_critic.py:21)
            self.log_std = nn.Parameter(np.log(model_cfg['action_std']) * torch.ones(actions_shape))
        self.max_action = model_cfg['clipAction']
        assert self.max_action > 0
        self.action_activate = model_cfg['action_activate']
        if self.action_activate not in ('tanh', None):
            raise NotImplementedError
        self.num_actions = actions_shape
        # counter-based RNG state for pm_randn (Philox): (seed, offset).  The rank is mixed into the key so that env shards on
        # different GPUs draw independent exploration noise when every rank calls torch.manual_seed(seed) (the usual set_seed)
        from ... import parallel
        self._seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * parallel.rank()) & (2 ** 63 - 1)
        self._offset = 0
        # flat buffers (set by flatten_())
        self.actor_flat: Optional[torch.Tensor] = None
        self.critic_flat: Optional[torch.Tensor] = None

    def rng_state(self) -> dict:
        """Philox (seed, offset) — saved with checkpoints so that a resumed run continues the noise sequence."""
        return {'seed': self._seed, 'offset': self._offset}

    def set_rng_state(self, st: dict):
        self._seed, self._offset = int(st['seed']), int(st['offset'])

    # ------------------------------------------------------------------ flat parameter storage
    def flatten_(self):
        """Re-home the parameters into two contiguous fp32 buffers — [actor params | log_std] and [critic params]
        — matching the reference's two optimisers (ppo.py:73-74).  Parameters become views, so state_dict /
        load_state_dict keep working and the optimiser kernel sees one flat tensor."""
        dev = self.log_std.device

        def pack(params: List[nn.Parameter]):
            n = sum(p.numel() for p in params)
            # 16-byte align every tensor so float4 paths stay legal
            offs, off = [], 0
            for p in params:
                offs.append(off)
                off += (p.numel() + 3) // 4 * 4
            flat = torch.zeros(off, device=dev, dtype=torch.float32)
            for p, o in zip(params, offs):
                flat[o:o + p.numel()].copy_(p.data.reshape(-1))
                p.data = flat[o:o + p.numel()].view(p.shape)
            return flat, offs, n

        a_params = list(self.actor.parameters())
        self.actor_flat, self.actor_offs, _ = pack(a_params + [self.log_std])
        self.actor_n_clip = self.actor_offs[-1]                  # everything before log_std is clipped (Q8)
        self.critic_flat, self.critic_offs, _ = pack(list(self.critic.parameters()))
        return self

    def grad_views(self, flat_grad: torch.Tensor, which: str) -> List[torch.Tensor]:
        params = (list(self.actor.parameters()) + [self.log_std]) if which == "actor" else list(self.critic.parameters())
        offs = self.actor_offs if which == "actor" else self.critic_offs
        return [flat_grad[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)]

    # ------------------------------------------------------------------ reference API
    def forward(self):
        raise NotImplementedError

    @property
    def _squash(self) -> bool:
        return self.action_activate == 'tanh'

    def _eps(self, like: torch.Tensor, eps: Optional[torch.Tensor]) -> torch.Tensor:
        if eps is not None:
            return eps.contiguous()
        out = torch.empty_like(like)
        ops.randn(out, self._seed, self._offset)
        self._offset += (out.numel() + 3) // 4
        return out

    @torch.no_grad()
    def cri(self, observations):
        return self.critic.runner.forward(observations, need_backward=False).clone()

    @torch.no_grad()
    def random_act_cri(self, observations, eps: Optional[torch.Tensor] = None):
        """actor_critic.py:36-47.  `eps` (E,A) optionally injects the standard-normal draw (parity tests)."""
        mu = self.actor.runner.forward(observations, need_backward=False)
        actions, logp, sigma = ops.policy_sample(mu, self.log_std.data, self._eps(mu, eps), self.max_action, self._squash)
        value = self.critic.runner.forward(observations, need_backward=False)
        return actions, logp, value.clone(), mu.clone(), sigma

    @torch.no_grad()
    def random_act(self, observations, eps: Optional[torch.Tensor] = None):
        mu = self.actor.runner.forward(observations, need_backward=False)
        actions, _, _ = ops.policy_sample(mu, self.log_std.data, self._eps(mu, eps), self.max_action, self._squash)
        return actions

    @torch.no_grad()
    def act(self, observations):
        mu = self.actor.runner.forward(observations, need_backward=False)
        return ops.action_activation(mu, self.max_action, self._squash)

    @torch.no_grad()
    def act_cri(self, observations):
        mu = self.actor.runner.forward(observations, need_backward=False)
        value = self.critic.runner.forward(observations, need_backward=False)
        return ops.action_activation(mu, self.max_action, self._squash), value.clone()

    @torch.no_grad()
    def update_act_cri(self, observations, actions):
        """actor_critic.py:71-82, values only.  (The PPO engine fuses log-prob, losses and their gradients in
        pm_ppo_actor_loss / pm_value_loss; this method exists for callers that inspect the distribution.)"""
        mu = self.actor.runner.forward(observations)
        logp, ent = ops.policy_logprob(mu, self.log_std.data, actions.contiguous(), self.max_action, self._squash)
        value = self.critic.runner.forward(observations)
        sigma = self.log_std.data.repeat(mu.shape[0], 1)
        return logp, ent, value.clone(), mu.clone(), sigma

    def update_act(self, observations):
        """actor_critic.py:67-69 — differentiable (DAgger / BC students)."""
        mu = self.actor(observations)
        return _Squash.apply(mu, self.max_action) if self._squash else mu

    def action_activation(self, action):
        return ops.action_activation(action.contiguous(), self.max_action, self._squash)


class _Squash(torch.autograd.Function):
    """tanh(mu)*max_action with its derivative, through pm_action_activation."""

    @staticmethod
    def forward(ctx, mu, max_action):
        y = ops.action_activation(mu.contiguous(), max_action, True)
        ctx.save_for_backward(y)
        ctx.max_action = max_action
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        t = y / ctx.max_action
        return dy * ctx.max_action * (1 - t * t), None
