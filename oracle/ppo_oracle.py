"""TEST INFRASTRUCTURE ONLY — CPU oracle (torch-on-CPU restatement) of PartManip's PPO hot path.

This file restates, function by function, what the reference computes on the path
named by BASELINE.json:north_star.  It is the *checker* for the CUDA kernels in
partmanip_b200/csrc and the CPU baseline timed by bench.py; it is never imported by
the product package.

Parity pinning: the reference ships no tests/golden vectors for this path
(SURVEY.md H3).  The oracle is therefore pinned against outputs of the UNMODIFIED
reference modules executed in the build container (tests/golden/make_golden.py
imports /root/reference, writes tests/golden/*.npz) and against the shipped
checkpoint KAT.  tests/test_oracle_golden.py replays them.

Third-party arithmetic restated here (reference calls PyTorch, unpinned; oracle
validated against torch 2.11): torch.optim.Adam (single-tensor, amsgrad off),
nn.utils.clip_grad_norm_, MultivariateNormal(scale_tril=diag) log_prob/sample,
nn.Linear default init and nn.init.orthogonal_.

All file:line citations are relative to the reference root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# activations  (algorithms/algo_utils/network.py:7-24)
# --------------------------------------------------------------------------------------
_ACT = {
    "elu": torch.nn.functional.elu,
    "selu": torch.selu,
    "relu": torch.relu,
    "crelu": torch.relu,  # network.py:13-14 maps "crelu" to plain ReLU
    "lrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.01),
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
}


def activation(name: str):
    if name not in _ACT:
        raise NotImplementedError(f"invalid activation function {name!r}")
    return _ACT[name]


# --------------------------------------------------------------------------------------
# parameter construction
# --------------------------------------------------------------------------------------
def _linear_default_init(out_f: int, in_f: int, gen: Optional[torch.Generator]) -> Tuple[Tensor, Tensor]:
    """nn.Linear.reset_parameters: kaiming_uniform(a=sqrt 5) == U(-1/sqrt(in), 1/sqrt(in)) for W and b."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def mlp_layer_dims(in_dim: int, out_dim: int, hid: Sequence[int]) -> List[Tuple[int, int]]:
    """network.py:33-40 — Linear(in,h0), Linear(h0,h1) ..., Linear(h_last,out)."""
    dims = [in_dim, *hid, out_dim]
    return [(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]


def mlp_init(in_dim: int, out_dim: int, hid: Sequence[int], gen: Optional[torch.Generator] = None) -> Params:
    """network.py:27-51.  Keys follow the reference state_dict: model.{0,2,4,..}.{weight,bias}.
    Orthogonal init with gains sqrt2,...,sqrt2, then 1 (critic, out_dim==1) or 0.01 (actor)."""
    p: Params = {}
    dims = mlp_layer_dims(in_dim, out_dim, hid)
    gains = [math.sqrt(2)] * len(hid) + [1.0 if out_dim == 1 else 0.01]
    for i, (fi, fo) in enumerate(dims):
        w, b = _linear_default_init(fo, fi, gen)
        torch.nn.init.orthogonal_(w, gain=gains[i], generator=gen)
        p[f"model.{2 * i}.weight"] = w
        p[f"model.{2 * i}.bias"] = b
    return p


def mlp_forward(p: Params, x: Tensor, act: str = "tanh") -> Tensor:
    """network.py:53 — Sequential(Linear, act, ..., Linear); no activation after the last layer."""
    f = activation(act)
    n = len([k for k in p if k.endswith(".weight")])
    h = x
    for i in range(n):
        h = torch.nn.functional.linear(h, p[f"model.{2 * i}.weight"], p[f"model.{2 * i}.bias"])
        if i != n - 1:
            h = f(h)
    return h


def pointnet_init(input_dim: int, out_dim: int, *, point_num: int = 1024, proprio: int = 0,
                  max_mean: bool = False, gen: Optional[torch.Generator] = None) -> Params:
    """network.py:141-163.  Per-point Linear(C,128)-act-Linear(128,256)-act-Linear(256,512);
    head Linear(512*(1+max_mean)+proprio,128)-act-Linear(128,32)-act-Linear(32,out).
    PyTorch default Linear init (no custom init).
    NB the reference computes C as input_dim // point_num with point_num hard-coded 1024
    (network.py:146-148); `point_num` is the build's parameterisation of that constant."""
    c = input_dim // point_num
    p: Params = {}
    for name, (fi, fo) in {
        "mlp.0": (c, 128), "mlp.2": (128, 256), "mlp.4": (256, 512),
        "final_mlp.0": (512 * (1 + int(max_mean)) + proprio, 128),
        "final_mlp.2": (128, 32), "final_mlp.4": (32, out_dim),
    }.items():
        w, b = _linear_default_init(fo, fi, gen)
        p[name + ".weight"], p[name + ".bias"] = w, b
    return p


def pointnet_encode(p: Params, pc: Tensor, act: str = "tanh") -> Tensor:
    """network.py:175 — the per-point MLP; pc (b, N, C) -> (b, N, 512)."""
    f = activation(act)
    h = f(torch.nn.functional.linear(pc, p["mlp.0.weight"], p["mlp.0.bias"]))
    h = f(torch.nn.functional.linear(h, p["mlp.2.weight"], p["mlp.2.bias"]))
    return torch.nn.functional.linear(h, p["mlp.4.weight"], p["mlp.4.bias"])


def pointnet_forward(p: Params, x: Tensor, *, point_num: int = 1024, proprio: int = 0,
                     max_mean: bool = False, sub_mean: bool = False, act: str = "tanh",
                     return_feat: bool = False):
    """network.py:165-198.  With sub_mean the xyz centring is written THROUGH A VIEW of the
    caller's tensor (network.py:172-173) — replicated here on purpose."""
    b = x.shape[0]
    if proprio != 0:
        tail = x[:, -proprio:]
        pc = x[:, :-proprio].reshape(b, point_num, -1)
    else:
        tail = None
        pc = x.reshape(b, point_num, -1)
    if sub_mean:
        pc[..., :3] = pc[..., :3] - pc[..., :3].mean(dim=1, keepdim=True)
    h = pointnet_encode(p, pc, act)
    if max_mean:
        feat = torch.cat((h.max(dim=1)[0], h.mean(dim=1)), dim=-1)
    else:
        feat = h.max(dim=1)[0]
    if tail is not None:
        feat = torch.cat((feat, tail), dim=-1)
    f = activation(act)
    o = f(torch.nn.functional.linear(feat, p["final_mlp.0.weight"], p["final_mlp.0.bias"]))
    o = f(torch.nn.functional.linear(o, p["final_mlp.2.weight"], p["final_mlp.2.bias"]))
    o = torch.nn.functional.linear(o, p["final_mlp.4.weight"], p["final_mlp.4.bias"])
    return (o, feat) if return_feat else o


def net_forward(kind: str, p: Params, x: Tensor, net_cfg: dict, proprio: int = 0) -> Tensor:
    """actor_critic.py:16,19 — eval(net_cfg['name']) dispatch restricted to the two hot-path classes."""
    if kind == "MLP":
        return mlp_forward(p, x, net_cfg["activation"])
    if kind == "PointNet":
        return pointnet_forward(p, x, point_num=net_cfg.get("point_num", 1024), proprio=proprio,
                                max_mean=net_cfg["max_mean"], sub_mean=net_cfg["sub_mean"],
                                act=net_cfg["activation"])
    raise NotImplementedError(kind)


# --------------------------------------------------------------------------------------
# Gaussian policy head  (algorithms/algo_utils/actor_critic.py:36-100)
# --------------------------------------------------------------------------------------
LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def policy_std(log_std: Tensor) -> Tensor:
    """actor_critic.py:39-40: scale_tril = diag(exp(ls)*exp(ls))  => std = exp(2*log_std)  (SURVEY Q1)."""
    return log_std.exp() * log_std.exp()


def gaussian_logp(mu: Tensor, log_std: Tensor, a: Tensor) -> Tensor:
    """MultivariateNormal(mu, scale_tril=diag(s)).log_prob(a) = sum_a[-0.5((a-mu)/s)^2 - ln s - 0.5 ln 2pi]."""
    s = policy_std(log_std)
    z = (a - mu) / s
    return (-0.5 * z * z - torch.log(s) - LOG_SQRT_2PI).sum(-1)


def gaussian_entropy(log_std: Tensor, batch: int) -> Tensor:
    s = policy_std(log_std)
    h = (0.5 + LOG_SQRT_2PI) * s.numel() + torch.log(s).sum()
    return h.expand(batch)


def action_activation(a: Tensor, max_action: float, mode: Optional[str] = "tanh") -> Tensor:
    """actor_critic.py:84-91."""
    if mode == "tanh":
        return torch.tanh(a) * max_action
    if mode is None:
        return a
    raise NotImplementedError


def action_deactivation(a: Tensor, max_action: float, mode: Optional[str] = "tanh") -> Tensor:
    """actor_critic.py:93-100 — atanh(clamp(a/max_a, +-(1-1e-5)))  (SURVEY Q4)."""
    if mode == "tanh":
        return torch.atanh(torch.clamp(a / max_action, max=1 - 1e-5, min=-1 + 1e-5))
    if mode is None:
        return a
    raise NotImplementedError


def policy_sample(mu: Tensor, log_std: Tensor, eps: Tensor, max_action: float, mode: Optional[str] = "tanh"):
    """actor_critic.py:36-47 with the standard-normal draw `eps` made explicit:
    sample = mu + s*eps (rsample of MultivariateNormal with diagonal scale_tril);
    returns (activated action, log_prob of the raw sample)."""
    s = policy_std(log_std)
    raw = mu + s * eps
    return action_activation(raw, max_action, mode), gaussian_logp(mu, log_std, raw)


# --------------------------------------------------------------------------------------
# rollout buffer math  (algorithms/algo_utils/storage.py:96-138)
# --------------------------------------------------------------------------------------
def gae(rewards: Tensor, values: Tensor, dones: Tensor, succs: Tensor, last_values: Tensor,
        gamma: float, lam: float, succ_value: Optional[float], whole_adv_norm: bool = False):
    """storage.py:96-114.  All inputs (T, E, 1) (last_values (E, 1)); dones/succs bool.
    Returns (returns, advantages)."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        not_terminal = ~dones[step]
        delta = rewards[step] + gamma * next_values - values[step]
        advantage = not_terminal * (delta + gamma * lam * advantage)
        if succ_value is not None:
            returns[step] = (~succs[step]) * (advantage + values[step]) + succs[step] * succ_value
        else:
            returns[step] = advantage + values[step]
    advantages = returns - values
    if whole_adv_norm:
        advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def minibatch_geometry(buf_size: int, num_mini_batches: int) -> Tuple[int, int]:
    """storage.py:125-138: size = min(buf // nmb, 2048); BatchSampler(drop_last=True) => count = buf // size."""
    size = min(int(buf_size // num_mini_batches), 2048)
    return size, buf_size // size


def minibatch_indices(buf_size: int, num_mini_batches: int, sampler: str = "sequential",
                      gen: Optional[torch.Generator] = None) -> List[Tensor]:
    size, count = minibatch_geometry(buf_size, num_mini_batches)
    if sampler == "sequential":
        order = torch.arange(buf_size)
    elif sampler == "random":
        order = torch.randperm(buf_size, generator=gen)
    else:
        raise NotImplementedError(sampler)
    return [order[k * size:(k + 1) * size] for k in range(count)]


# --------------------------------------------------------------------------------------
# running mean / std  (algorithms/algo_utils/RMS.py:3-45)
# --------------------------------------------------------------------------------------
class RunningStats:
    """RMS.py:3-34 (non-standard update)."""

    def __init__(self, dim: int):
        self.n = 0
        self.mean = torch.zeros(1, dim)
        self.S = torch.ones(1, dim) * 1e-4
        self.std = torch.sqrt(self.S)

    def update(self, x: Tensor):
        self.n += 1
        old = self.mean.clone()
        new = x.mean(dim=0, keepdim=True)
        self.mean = old + (new - old) / self.n
        self.S = self.S + (x - new).pow(2).mean(dim=0, keepdim=True) + (old - new).pow(2) * (self.n - 1) / self.n
        self.std = torch.sqrt(self.S / self.n)

    def normalize(self, x: Tensor, update: bool = True) -> Tensor:
        """RMS.py:40-45 (no epsilon in the divide)."""
        if update:
            self.update(x)
        return (x - self.mean) / self.std


# --------------------------------------------------------------------------------------
# PPO losses  (algorithms/ppo.py:326-374)
# --------------------------------------------------------------------------------------
def kl_old_new(mu: Tensor, log_std_b: Tensor, mu_old: Tensor, log_std_old_b: Tensor) -> Tensor:
    """ppo.py:332-333 — uses exp(log_std) (NOT the exp(2 log_std) the sampler uses)."""
    return torch.sum(log_std_b - log_std_old_b
                     + (torch.square(log_std_old_b.exp()) + torch.square(mu_old - mu))
                     / (2.0 * torch.square(log_std_b.exp())) - 0.5, dim=-1)


def surrogate_loss(logp: Tensor, logp_old: Tensor, adv: Tensor, eps_clip: float) -> Tensor:
    """ppo.py:341-344."""
    ratio = torch.exp(logp - logp_old)
    s1 = -adv * ratio
    s2 = -adv * torch.clamp(ratio, 1.0 - eps_clip, 1.0 + eps_clip)
    return torch.max(s1, s2).mean()


def value_loss(value: Tensor, returns: Tensor, old_values: Tensor, eps_clip: float, clipped: bool) -> Tensor:
    """ppo.py:368-374."""
    if clipped:
        with torch.no_grad():
            d = (eps_clip * old_values).abs().mean()
            target = old_values + (returns - old_values).clamp(-d, d)
        return (value - target).pow(2).mean()
    return (returns - value).pow(2).mean()


def mini_adv_norm(adv: Tensor) -> Tensor:
    """ppo.py:328-329 (unbiased std)."""
    return (adv - adv.mean()) / (adv.std() + 1e-8)


# --------------------------------------------------------------------------------------
# optimiser restatement  (torch.optim.Adam defaults + nn.utils.clip_grad_norm_; ppo.py:73-74,351-353)
# --------------------------------------------------------------------------------------
def clip_coef(grads: Sequence[Tensor], max_norm: float) -> Tuple[Tensor, Tensor]:
    """clip_grad_norm_: total = || (||g_i||_2)_i ||_2 ; coef = min(1, max_norm / (total + 1e-6))."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    return total, torch.clamp(max_norm / (total + 1e-6), max=1.0)


def adam_step(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.adam._single_tensor_adam (amsgrad=False, weight_decay=0, maximize=False); in place.
    `step` is the 1-based step count AFTER the increment."""
    m.lerp_(grad, 1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-(lr / bc1))


class AdamState:
    """Flat-free Adam over a dict of named tensors; mirrors two reference optimisers (ppo.py:73-74)."""

    def __init__(self, params: Params, lr: float):
        self.params = params
        self.lr = lr
        self.step = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply(self, grads: Params):
        self.step += 1
        for k, p in self.params.items():
            adam_step(p, grads[k], self.m[k], self.v[k], self.step, self.lr)


# --------------------------------------------------------------------------------------
# one PPO update over a filled buffer  (algorithms/ppo.py:307-411), autograd for the gradients
# --------------------------------------------------------------------------------------
def _leaf(p: Params) -> Params:
    return {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}


def ppo_update(actor: Params, critic: Params, log_std: Tensor, opt_a: AdamState, opt_c: AdamState,
               buf: Dict[str, Tensor], cfg: dict, net_kind: str, net_cfg: dict, proprio: int = 0,
               gen: Optional[torch.Generator] = None) -> Dict[str, float]:
    """Restates ppo.update.  `buf` holds flattened (T*E, .) tensors: obs, actions, values, returns,
    logp, adv, mu, sigma.  actor/critic/log_std are updated in place; opt_a holds actor params +
    'log_std' (ppo.py:73), opt_c the critic's.  Only the network each phase needs is evaluated —
    the reference's extra forward (actor_critic.py:71-82 runs both) has no effect on any output."""
    n = buf["obs"].shape[0]
    eps_clip = cfg["epsilon_clip"]
    tricks = cfg["tricks"]
    mean_v = mean_s = mean_kl = 0.0
    kl_max = 0.0
    count = 0
    nb = 0
    for _ in range(cfg["n_updates"]):
        for idx in minibatch_indices(n, cfg["n_minibatches"], cfg["sampler"], gen):
            obs = buf["obs"][idx].clone()  # reference gathers a copy: x[list] (ppo.py:317)
            a_leaf = _leaf(actor)
            ls_leaf = log_std.detach().clone().requires_grad_(True)
            mu = net_forward(net_kind, a_leaf, obs, net_cfg, proprio)
            raw = action_deactivation(buf["actions"][idx], cfg["model"]["clipAction"], cfg["model"]["action_activate"])
            logp = gaussian_logp(mu, ls_leaf, raw)
            adv = buf["adv"][idx].squeeze(-1)
            if tricks["mini_adv_norm"]:
                adv = mini_adv_norm(buf["adv"][idx]).squeeze(-1)
            sig_b = ls_leaf.repeat(mu.shape[0], 1)
            kl = kl_old_new(mu, sig_b, buf["mu"][idx], buf["sigma"][idx]).mean()
            kl_max = max(kl_max, float(kl.detach()))
            if float(kl.detach()) > cfg["desired_kl"]:
                continue
            loss = surrogate_loss(logp, buf["logp"][idx].squeeze(-1), adv, eps_clip)
            loss.backward()
            grads = {k: v.grad for k, v in a_leaf.items()}
            if tricks["use_grad_clip"]:
                _, coef = clip_coef(list(grads.values()), tricks["max_grad_norm"])  # log_std excluded (Q8)
                for g in grads.values():
                    g.mul_(coef)
            grads["log_std"] = ls_leaf.grad
            opt_a.apply(grads)
            mean_s += float(loss.detach())
            mean_kl += float(kl.detach())
            count += 1
    for _ in range(cfg["n_updates"]):
        for idx in minibatch_indices(n, cfg["n_minibatches"], cfg["sampler"], gen):
            obs = buf["obs"][idx].clone()
            c_leaf = _leaf(critic)
            val = net_forward(net_kind, c_leaf, obs, net_cfg, proprio)
            loss = value_loss(val, buf["returns"][idx], buf["values"][idx], eps_clip, tricks["use_clipped_value_loss"])
            loss.backward()
            grads = {k: v.grad for k, v in c_leaf.items()}
            if tricks["use_grad_clip"]:
                _, coef = clip_coef(list(grads.values()), tricks["max_grad_norm"])
                for g in grads.values():
                    g.mul_(coef)
            opt_c.apply(grads)
            mean_v += float(loss.detach())
            nb += 1
    return {
        "value_loss": mean_v / max(nb, 1),
        "surrogate_loss": mean_s / count if count else float("nan"),
        "kl": mean_kl / count if count else float("nan"),
        "kl_max": kl_max,
        "count": count,
    }


def dagger_update_step(student_actor: Params, opt: AdamState, stu_obs: Tensor, tea_act: Tensor,
                       net_kind: str, net_cfg: dict, max_action: float, proprio: int = 0) -> float:
    """dagger.py:310-319 — loss = mean((tanh(mu_teacher) - tanh(mu_student))^2), one Adam over the
    student (critic/log_std receive no gradient)."""
    leaf = _leaf(student_actor)
    stu = action_activation(net_forward(net_kind, leaf, stu_obs.clone(), net_cfg, proprio), max_action)
    loss = (tea_act - stu).pow(2).mean()
    loss.backward()
    opt.apply({k: v.grad for k, v in leaf.items()})
    return float(loss)
