"""Tensor-level wrappers over the C-ABI (include/partmanip_b200.h).

Each function validates device/dtype/contiguity, passes raw device pointers + the current CUDA stream
to libpartmanip_b200.so and returns torch tensors that merely OWN the memory (torch is plumbing here:
allocation, streams, torch.distributed).  No function in this module computes anything with torch ops,
and none falls back to CPU.
"""
from __future__ import annotations

import ctypes as ct
from typing import Optional, Sequence, Tuple

import torch

from ._lib import EncoderParams, LAUNCHES, PM_ACT, PM_PREC, check, lib

Tensor = torch.Tensor


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32(t: Tensor, name: str, last_contig: bool = True) -> Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (the product path has no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if last_contig and t.dim() >= 1 and t.numel() > 0 and t.stride(-1) != 1:
        raise ValueError(f"{name}: last dimension must be contiguous")
    return t


def _rows(t: Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a 2-D row-major view (row stride may exceed cols)."""
    if t.dim() != 2:
        raise ValueError(f"{name}: expected 2-D, got {tuple(t.shape)}")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t.shape[0], t.shape[1], ld


def launch_count() -> int:
    """Kernels launched through the C-ABI so far (per-call multiplicities in _lib.KERNELS_PER_CALL)."""
    return LAUNCHES[0]


def count_launches(n: int):
    """Account for kernels launched by a CUDA-graph replay of previously counted C-ABI calls."""
    LAUNCHES[0] += int(n)


_scratch = {}
_scratch_gen = [0]


def scratch_generation() -> int:
    """Bumped whenever an EXISTING workspace is replaced by a larger one.  A captured CUDA graph bakes workspace pointers in;
    its owner compares this counter with the value at capture time and re-captures when it moved (ppo.update)."""
    return _scratch_gen[0]


_scratch_ns = [""]


class scratch_ns:
    """`with ops.scratch_ns("critic"):` — workspaces requested inside get their own copies, so that two kernel sequences that run
    CONCURRENTLY on different streams (the actor and the critic phase of the PPO update) never share a scratch buffer."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.prev, _scratch_ns[0] = _scratch_ns[0], self.name
        return self

    def __exit__(self, *a):
        _scratch_ns[0] = self.prev


def scratch(nbytes: int, device, tag: str = "default") -> Tensor:
    """Grow-only per-(device, namespace + tag) byte workspace (never shrinks, never shared across tags)."""
    key = (str(device), _scratch_ns[0] + tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError(f"workspace {tag!r} would be (re)allocated inside a CUDA-graph capture; run the same call "
                               "eagerly once first")
        if buf is not None:
            _scratch_gen[0] += 1
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


# ------------------------------------------------------------------------------------------- K6 RMS
def rms_forward(x: Tensor, out: Tensor, mean: Tensor, S: Tensor, std: Tensor, n_after: int, update: bool) -> Tensor:
    """RMS.py:40-45 (+10-18 when update): out = (x-mean)/std after the optional running-stat update."""
    _f32(x, "x"); _f32(out, "out")
    E, D, ldx = _rows(x, "x")
    _, _, ldo = _rows(out, "out")
    for t, n in ((mean, "mean"), (S, "S"), (std, "std")):
        _f32(t, n)
        assert t.numel() == D and t.is_contiguous(), n
    ws = scratch(lib.pm_rms_forward_ws_bytes(E, D), x.device, "rms")
    check(lib.pm_rms_forward(_p(x), ldx, _p(out), ldo, E, D, _p(mean), _p(S), _p(std), int(n_after), int(bool(update)),
                             _p(ws), _stream()), "pm_rms_forward")
    return out


def rms_colsum(x: Tensor, colsum: Tensor):
    E, D, ldx = _rows(_f32(x, "x"), "x")
    ws = scratch(lib.pm_colreduce_ws_bytes(E, D), x.device, "rms")
    check(lib.pm_rms_colsum(_p(x), ldx, E, D, _p(_f32(colsum, "colsum")), _p(ws), _stream()), "pm_rms_colsum")


def rms_colsqdev(x: Tensor, colsum: Tensor, count: float, sqdev: Tensor):
    E, D, ldx = _rows(_f32(x, "x"), "x")
    ws = scratch(lib.pm_colreduce_ws_bytes(E, D), x.device, "rms")
    check(lib.pm_rms_colsqdev(_p(x), ldx, E, D, _p(colsum), float(count), _p(_f32(sqdev, "sqdev")), _p(ws), _stream()),
          "pm_rms_colsqdev")


def rms_update(mean: Tensor, S: Tensor, std: Tensor, colsum: Tensor, sqdev: Tensor, count: float, n_after: int):
    check(lib.pm_rms_update(_p(mean), _p(S), _p(std), _p(colsum), _p(sqdev), float(count), int(n_after), mean.numel(),
                            _stream()), "pm_rms_update")


def rms_normalize(x: Tensor, out: Tensor, mean: Tensor, std: Tensor):
    E, D, ldx = _rows(_f32(x, "x"), "x")
    _, _, ldo = _rows(_f32(out, "out"), "out")
    check(lib.pm_rms_normalize(_p(x), ldx, _p(out), ldo, E, D, _p(mean), _p(std), _stream()), "pm_rms_normalize")


# ------------------------------------------------------------------------------------------- K5 GAE
def gae(rewards: Tensor, values: Tensor, dones: Tensor, succs: Optional[Tensor], last_values: Tensor, gamma: float,
        lam: float, succ_value: Optional[float], returns: Optional[Tensor] = None,
        advantages: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """storage.py:96-112.  Shapes (T,E[,1]); dones/succs torch.bool or uint8."""
    T, E = rewards.shape[0], rewards.shape[1]
    for t, n in ((rewards, "rewards"), (values, "values"), (last_values, "last_values")):
        _f32(t, n)
        assert t.is_contiguous(), n
    d8 = dones.view(torch.uint8) if dones.dtype == torch.bool else dones
    s8 = None if succs is None else (succs.view(torch.uint8) if succs.dtype == torch.bool else succs)
    assert d8.dtype == torch.uint8 and d8.is_contiguous()
    if returns is None:
        returns = torch.empty_like(rewards)
    if advantages is None:
        advantages = torch.empty_like(rewards)
    use = succ_value is not None
    # gamma*lam is a python double product rounded once to fp32, as `gamma * lam * advantage` evaluates it
    check(lib.pm_gae(_p(rewards), _p(values), _p(d8), _p(s8), _p(last_values), _p(returns), _p(advantages), T, E,
                     float(gamma), float(gamma * lam), int(use), float(succ_value if use else 0.0), _stream()), "pm_gae")
    return returns, advantages


def normalize_(x: Tensor) -> Tensor:
    """(x-mean)/(std_unbiased+1e-8) in place (storage.py:113-114)."""
    assert x.is_contiguous()
    ws = scratch(256, x.device, "norm")
    check(lib.pm_normalize_inplace(_p(_f32(x, "x")), x.numel(), _p(ws), _stream()), "pm_normalize_inplace")
    return x


def normalize_stats(x: Tensor, stats: Tensor) -> Tensor:
    """stats <- {mean, std_unbiased+1e-8} of x without touching x (mini_adv_norm, ppo.py:328-329)."""
    assert x.is_contiguous() and stats.numel() >= 2
    check(lib.pm_normalize(_p(_f32(x, "x")), None, x.numel(), _p(_f32(stats, "stats")), _stream()), "pm_normalize")
    return stats


# ------------------------------------------------------------------------------------------- K4 policy
def randn(out: Tensor, seed: int, offset: int) -> Tensor:
    assert out.is_contiguous()
    check(lib.pm_randn(_p(_f32(out, "out")), out.numel(), int(seed) & (2 ** 64 - 1), int(offset), _stream()), "pm_randn")
    return out


def policy_sample(mu: Tensor, log_std: Tensor, eps: Tensor, max_action: float, squash: bool,
                  actions: Optional[Tensor] = None, logp: Optional[Tensor] = None, sigma: Optional[Tensor] = None):
    """actor_critic.py:39-47 after the actor forward."""
    E, A = mu.shape
    for t, n in ((mu, "mu"), (log_std, "log_std"), (eps, "eps")):
        _f32(t, n)
        assert t.is_contiguous(), n
    if actions is None:
        actions = torch.empty_like(mu)
    if logp is None:
        logp = torch.empty(E, device=mu.device, dtype=torch.float32)
    if sigma is None:
        sigma = torch.empty_like(mu)
    check(lib.pm_policy_sample(_p(mu), _p(log_std), _p(eps), E, A, float(max_action), int(bool(squash)), _p(actions),
                               _p(logp), _p(sigma), _stream()), "pm_policy_sample")
    return actions, logp, sigma


def action_activation(mu: Tensor, max_action: float, squash: bool, out: Optional[Tensor] = None) -> Tensor:
    assert mu.is_contiguous()
    if out is None:
        out = torch.empty_like(mu)
    check(lib.pm_action_activation(_p(_f32(mu, "mu")), _p(out), mu.numel(), float(max_action), int(bool(squash)),
                                   _stream()), "pm_action_activation")
    return out


def policy_logprob(mu: Tensor, log_std: Tensor, actions: Tensor, max_action: float, squash: bool):
    """actor_critic.py:71-82 without gradient: (logp, entropy) of stored squashed actions."""
    B, A, ldmu = _rows(_f32(mu, "mu"), "mu")
    assert actions.is_contiguous()
    logp = torch.empty(B, device=mu.device, dtype=torch.float32)
    ent = torch.empty(B, device=mu.device, dtype=torch.float32)
    check(lib.pm_policy_logprob(_p(mu), ldmu, _p(log_std), _p(_f32(actions, "actions")), B, A, float(max_action),
                                int(bool(squash)), _p(logp), _p(ent), _stream()), "pm_policy_logprob")
    return logp, ent


def ppo_actor_loss(mu: Tensor, log_std: Tensor, actions: Tensor, logp_old: Tensor, mu_old: Tensor, sigma_old: Tensor,
                   adv: Tensor, adv_stats: Optional[Tensor], inv_batch: float, eps_clip: float, max_action: float,
                   squash: bool, stats: Tensor, dmu: Tensor, dlog_std: Tensor, logp_out: Optional[Tensor] = None):
    """ppo.py:326-344 forward + analytic backward; see include/partmanip_b200.h."""
    B, A, ldmu = _rows(_f32(mu, "mu"), "mu")
    _, _, lddmu = _rows(_f32(dmu, "dmu"), "dmu")
    for t, n in ((actions, "actions"), (logp_old, "logp_old"), (mu_old, "mu_old"), (sigma_old, "sigma_old"), (adv, "adv")):
        _f32(t, n)
        assert t.is_contiguous(), n
    ws = scratch(lib.pm_ppo_actor_loss_ws_bytes(B, A), mu.device, "loss")
    check(lib.pm_ppo_actor_loss(_p(mu), ldmu, _p(log_std), _p(actions), _p(logp_old), _p(mu_old), _p(sigma_old), _p(adv),
                                _p(adv_stats), B, A, float(inv_batch), float(eps_clip), float(max_action),
                                int(bool(squash)), _p(stats), _p(dmu), lddmu, _p(dlog_std), _p(logp_out), _p(ws),
                                _stream()), "pm_ppo_actor_loss")


def ppo_actor_finalize(stats: Tensor, inv_batch: float, desired_kl: float, acc: Tensor, skip_flag: Tensor):
    assert skip_flag.dtype == torch.int32 and acc.numel() >= 4
    check(lib.pm_ppo_actor_finalize(_p(stats), float(inv_batch), float(desired_kl), _p(acc), _p(skip_flag), _stream()),
          "pm_ppo_actor_finalize")


def value_loss(v: Tensor, returns: Tensor, old_values: Optional[Tensor], clip_delta: Optional[Tensor], inv_batch: float,
               stats: Tensor, dv: Tensor):
    """ppo.py:368-374 forward + d/dv.  v, dv: (B,1) or (B,) possibly strided."""
    B = v.shape[0]
    ldv = v.stride(0) if v.dim() > 1 or B > 1 else 1
    lddv = dv.stride(0) if dv.dim() > 1 or B > 1 else 1
    ws = scratch(4 * (B // 256 + 2) + 4096, v.device, "loss")
    check(lib.pm_value_loss(_p(_f32(v, "v", False)), ldv, _p(returns), _p(old_values), _p(clip_delta), B,
                            float(inv_batch), _p(stats), _p(_f32(dv, "dv", False)), lddv, _p(ws), _stream()),
          "pm_value_loss")


def dagger_loss(mu: Tensor, tea_act: Tensor, max_action: float, squash: bool, inv_count: float, stats: Tensor,
                dmu: Tensor):
    """dagger.py:312-314 forward + d/dmu."""
    B, A, ldmu = _rows(_f32(mu, "mu"), "mu")
    _, _, lddmu = _rows(_f32(dmu, "dmu"), "dmu")
    assert tea_act.is_contiguous() and tea_act.shape == (B, A)
    ws = scratch(4096, mu.device, "loss")
    check(lib.pm_dagger_loss(_p(mu), ldmu, _p(_f32(tea_act, "tea_act")), B, A, float(max_action), int(bool(squash)),
                             float(inv_count), _p(stats), _p(dmu), lddmu, _p(ws), _stream()), "pm_dagger_loss")


def abs_sum(x: Tensor, scale: float, out: Tensor):
    assert x.is_contiguous()
    ws = scratch(4096, x.device, "loss")
    check(lib.pm_abs_sum(_p(_f32(x, "x")), x.numel(), float(scale), _p(out), _p(ws), _stream()), "pm_abs_sum")


def accumulate(stats: Tensor, scale: float, acc: Tensor, idx: int):
    check(lib.pm_accumulate(_p(stats), float(scale), _p(acc), int(idx), _stream()), "pm_accumulate")


# ------------------------------------------------------------------------------------------- K3 dense
def linear_forward(x: Tensor, W: Tensor, b: Optional[Tensor], act, out: Optional[Tensor] = None,
                   m_dev: Optional[Tensor] = None) -> Tensor:
    M, K, ldx = _rows(_f32(x, "x"), "x")
    N = W.shape[0]
    assert W.shape[1] == K and W.is_contiguous() and _f32(W, "W") is W
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    _, _, ldy = _rows(out, "out")
    check(lib.pm_linear_forward(_p(x), ldx, _p(W), _p(b), _p(out), ldy, M, N, K, PM_ACT[act], _p(m_dev), _stream()),
          "pm_linear_forward")
    return out


def linear_backward(x: Tensor, W: Tensor, dpre: Tensor, dW: Tensor, db: Optional[Tensor], dx: Optional[Tensor],
                    act_prev, m_dev: Optional[Tensor] = None):
    M, K, ldx = _rows(_f32(x, "x"), "x")
    N = W.shape[0]
    _, _, ldd = _rows(_f32(dpre, "dpre"), "dpre")
    lddx = 0
    if dx is not None:
        _, _, lddx = _rows(_f32(dx, "dx"), "dx")
    ws = scratch(lib.pm_linear_backward_ws_bytes(M, N, K), x.device, "linbwd")
    check(lib.pm_linear_backward(_p(x), ldx, _p(W), _p(dpre), ldd, _p(dW), _p(db), _p(dx), lddx, M, N, K,
                                 PM_ACT[act_prev], _p(m_dev), _p(ws), _stream()), "pm_linear_backward")


def linear_forward_tc(x: Tensor, W: Tensor, b: Optional[Tensor], act, precision: str, out: Optional[Tensor] = None,
                      m_dev: Optional[Tensor] = None) -> Tensor:
    """nn.Linear (+ activation) on tcgen05: precision "bf16" (1e-2 gate) or "fp32" (three-term bf16 split, 1e-4 gate)."""
    M, K, ldx = _rows(_f32(x, "x"), "x")
    N = W.shape[0]
    assert W.shape[1] == K and W.is_contiguous() and _f32(W, "W") is W
    if out is None:
        out = torch.empty(M, N, device=x.device, dtype=torch.float32)
    _, _, ldy = _rows(out, "out")
    check(lib.pm_linear_forward_tc(_p(x), ldx, _p(W), _p(b), _p(out), ldy, M, N, K, PM_ACT[act], PM_PREC[precision], _p(m_dev),
                                   _stream()), "pm_linear_forward_tc")
    return out


def linear_backward_tc(x: Tensor, W: Tensor, dpre: Tensor, dW: Tensor, db: Optional[Tensor], dx: Optional[Tensor], act_prev,
                       precision: str, m_dev: Optional[Tensor] = None):
    M, K, ldx = _rows(_f32(x, "x"), "x")
    N = W.shape[0]
    _, _, ldd = _rows(_f32(dpre, "dpre"), "dpre")
    lddx = 0
    if dx is not None:
        _, _, lddx = _rows(_f32(dx, "dx"), "dx")
    ws = scratch(lib.pm_linear_backward_tc_ws_bytes(M, N, K), x.device, "linbwd_tc")
    check(lib.pm_linear_backward_tc(_p(x), ldx, _p(W), _p(dpre), ldd, _p(dW), _p(db), _p(dx), lddx, M, N, K, PM_ACT[act_prev],
                                    PM_PREC[precision], _p(m_dev), _p(ws), _stream()), "pm_linear_backward_tc")


# ------------------------------------------------------------------------------------------- K3b fused head
def pointnet_head_forward(feat: Tensor, head_params: Sequence[Tensor], out_dim: int, act, h1: Tensor, h2: Tensor,
                          out: Tensor, precision: str = "fp32") -> Tensor:
    """network.py:152-159: out = L2(act(L1(act(L0(feat))))); h1 (B,128) / h2 (B,32) are saved for the backward."""
    B, F, ldf = _rows(_f32(feat, "feat"), "feat")
    _, _, ldo = _rows(_f32(out, "out"), "out")
    assert h1.shape == (B, 128) and h2.shape == (B, 32) and h1.is_contiguous() and h2.is_contiguous()
    ps = _enc_struct(head_params)
    check(lib.pm_pointnet_head_forward(_p(feat), ldf, B, F, ct.byref(ps), int(out_dim), PM_ACT[act], PM_PREC[precision], _p(h1),
                                       _p(h2), _p(out), ldo, _stream()), "pm_pointnet_head_forward")
    return out


def pointnet_head_backward(feat: Tensor, head_params: Sequence[Tensor], out_dim: int, act, h1: Tensor, h2: Tensor,
                           dout: Tensor, head_grads: Sequence[Tensor], dfeat: Optional[Tensor], dfeat_cols: int = 0,
                           precision: str = "fp32"):
    """autograd of network.py:152-159 w.r.t. the six head tensors (+ d/d feat[:, :dfeat_cols])."""
    B, F, ldf = _rows(_f32(feat, "feat"), "feat")
    _, _, lddo = _rows(_f32(dout, "dout"), "dout")
    lddf = 0
    if dfeat is not None:
        _, _, lddf = _rows(_f32(dfeat, "dfeat"), "dfeat")
    nbytes = lib.pm_pointnet_head_backward_ws_bytes(B, F)
    ws = scratch(nbytes, feat.device, "headbwd")
    ps, gs = _enc_struct(head_params), _enc_struct(head_grads)
    check(lib.pm_pointnet_head_backward(_p(feat), ldf, B, F, ct.byref(ps), int(out_dim), PM_ACT[act], PM_PREC[precision], _p(h1), _p(h2),
                                        _p(dout), lddo, ct.byref(gs), _p(dfeat), lddf, int(dfeat_cols), _p(ws), nbytes,
                                        _stream()), "pm_pointnet_head_backward")


# ------------------------------------------------------------------------------------------- K1/K2 encoder
def _enc_struct(ts: Sequence[Tensor]) -> EncoderParams:
    assert len(ts) == 6
    for t in ts:
        _f32(t, "encoder tensor")
        assert t.is_contiguous()
    return EncoderParams(*[t.data_ptr() for t in ts])


def pointnet_center_(x: Tensor, N: int, C: int):
    """network.py:172-173 — in place through the caller's tensor (Q3).  x: (B, >= N*C)."""
    B, _, ldx = _rows(_f32(x, "x"), "x")
    check(lib.pm_pointnet_center(_p(x), ldx, B, N, C, _stream()), "pm_pointnet_center")


def pointnet_encode_forward(x: Tensor, N: int, C: int, enc_params: Sequence[Tensor], act, precision: str, feat: Tensor,
                            feat_mean: Optional[Tensor] = None, argmax: Optional[Tensor] = None,
                            h2mean: Optional[Tensor] = None):
    """network.py:175-182.  x: (B, >= N*C) rows; feat/feat_mean: (B, 512) views with a common row stride."""
    B, width, ldx = _rows(_f32(x, "x"), "x")
    assert width >= N * C
    _, fw, ldf = _rows(_f32(feat, "feat"), "feat")
    assert fw == 512
    if feat_mean is not None:
        assert _rows(feat_mean, "feat_mean")[2] == ldf
    if argmax is not None:
        assert argmax.dtype == torch.int32 and argmax.is_contiguous() and argmax.shape == (B, 512)
    prec = PM_PREC[precision]
    nbytes = lib.pm_pointnet_encode_forward_ws_bytes(B, N, C, prec)
    ws = scratch(nbytes, x.device, "encfwd" if prec == PM_PREC["bf16"] else "encfwd_fp32") if nbytes else None
    ps = _enc_struct(enc_params)
    check(lib.pm_pointnet_encode_forward(_p(x), ldx, B, N, C, ct.byref(ps), PM_ACT[act], prec,
                                         _p(feat), _p(feat_mean), ldf, _p(argmax), _p(h2mean), _p(ws), nbytes, _stream()),
          "pm_pointnet_encode_forward")


def check_tc_errors():
    """Raise if any tcgen05 kernel on the current device reported a timed-out mbarrier wait since the last check (its outputs
    are garbage).  One 4-byte read; the algorithm classes call it right after their once-per-iteration read-back."""
    code = int(lib.pm_tc_sticky_error(1))
    if code:
        from ._lib import PMError
        raise PMError(f"a tcgen05 kernel reported protocol error {code} (bounded mbarrier wait timed out): the results of this "
                      "iteration are invalid")


def pointnet_tc_last_error(device) -> int:
    """Protocol error word of the last bf16 encoder launch on `device` (0 = clean).  Synchronises."""
    ws = _scratch.get((str(device), "encfwd"))
    if ws is None:
        return 0
    return int(lib.pm_pointnet_tc_last_error(_p(ws), _stream()))


def pointnet_tc3_last_error(device) -> int:
    """Protocol error word of the last fp32-mode (split-fp16 tcgen05) encoder launch on `device` (0 = clean).  Synchronises."""
    ws = _scratch.get((str(device), "encfwd_fp32"))
    if ws is None:
        return 0
    return int(lib.pm_pointnet_tc3_last_error(_p(ws), _stream()))


def pointnet_bwd_tc_last_error(device) -> int:
    """Protocol error word of the last bf16 encoder-backward launch on `device` (0 = clean).  Synchronises."""
    ws = _scratch.get((str(device), "encbwd_bf16"))
    if ws is None:
        return 0
    return int(lib.pm_pointnet_bwd_tc_last_error(_p(ws), _stream()))


def pointnet_encode_backward(x: Tensor, N: int, C: int, enc_params: Sequence[Tensor], act, dfeat: Tensor,
                             argmax: Tensor, enc_grads: Sequence[Tensor], dfeat_mean: Optional[Tensor] = None,
                             h2mean: Optional[Tensor] = None, precision: str = "fp32"):
    """autograd of network.py:175-182 w.r.t. the six encoder tensors (gradients overwritten)."""
    B, width, ldx = _rows(_f32(x, "x"), "x")
    _, fw, lddf = _rows(_f32(dfeat, "dfeat"), "dfeat")
    assert fw == 512 and argmax.dtype == torch.int32 and argmax.is_contiguous()
    prec = PM_PREC[precision]
    nbytes = lib.pm_pointnet_encode_backward_ws_bytes(B, N, C, int(dfeat_mean is not None), prec)
    ws = scratch(nbytes, x.device, "encbwd_bf16" if prec == PM_PREC["bf16"] else "encbwd")
    ps, gs = _enc_struct(enc_params), _enc_struct(enc_grads)
    check(lib.pm_pointnet_encode_backward(_p(x), ldx, B, N, C, ct.byref(ps), PM_ACT[act], prec, _p(dfeat), _p(dfeat_mean), lddf,
                                          _p(argmax), _p(h2mean), ct.byref(gs), _p(ws), nbytes, _stream()),
          "pm_pointnet_encode_backward")


# ------------------------------------------------------------------------------------------- K7 optimiser
def adam_step(params: Tensor, grads: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, n_clip: int, max_norm: float,
              opt_state: Tensor, skip_flag: Optional[Tensor], beta1: float = 0.9, beta2: float = 0.999,
              eps: float = 1e-8):
    n = params.numel()
    for t in (params, grads, exp_avg, exp_avg_sq):
        _f32(t, "adam buffer")
        assert t.is_contiguous() and t.numel() == n
    assert opt_state.numel() >= 8 and opt_state.dtype == torch.float32
    ws = scratch(lib.pm_adam_ws_bytes(n), params.device, "adam")
    check(lib.pm_adam_step(_p(params), _p(grads), _p(exp_avg), _p(exp_avg_sq), n, int(n_clip), float(max_norm),
                           float(beta1), float(beta2), float(eps), _p(opt_state), _p(skip_flag), _p(ws), _stream()),
          "pm_adam_step")


def fused_step_workspace(n: int, n_tail: int, device) -> Tensor:
    """Zeroed, optimiser-private workspace of pm_fused_step (it carries the launch sequence number and the grid-barrier counter)."""
    return torch.zeros(int(lib.pm_fused_step_ws_bytes(int(n), int(n_tail))), dtype=torch.uint8, device=device)


def fused_step(params: Tensor, grad_ext: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, n_clip: int, n_tail: int, max_norm: float,
               opt_state: Tensor, ws: Tensor, *, peers=None, finalize=None, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """One launch: [all-reduce over `peers`] + [KL-skip] + clip + Adam.  grad_ext: (n + n_tail,) this rank's gradients + tail.
    peers = (grad_ptrs_dev, flag_ptrs_dev, flags_local_tensor, rank, world) from parallel.SymmetricGrad, or None (one process).
    finalize = (inv_batch, desired_kl, acc, skip_flag) for the actor, None for the critic."""
    n = params.numel()
    for t in (params, exp_avg, exp_avg_sq):
        _f32(t, "adam buffer")
        assert t.is_contiguous() and t.numel() == n
    assert _f32(grad_ext, "grad_ext").is_contiguous() and grad_ext.numel() == n + n_tail
    gp, fp, fl, rank, world = peers if peers is not None else (None, None, None, 0, 1)
    inv_b, dkl, acc, skip = finalize if finalize is not None else (0.0, 0.0, None, None)
    check(lib.pm_fused_step(_p(params), _p(exp_avg), _p(exp_avg_sq), n, int(n_clip), int(n_tail), float(max_norm), float(beta1),
                            float(beta2), float(eps), _p(opt_state), _p(grad_ext), gp, fp, _p(fl), int(rank), int(world),
                            int(finalize is not None), float(inv_b), float(dkl), _p(acc), _p(skip), _p(ws), _stream()), "pm_fused_step")


# ------------------------------------------------------------------------------------------- K8 storage
def gather_rows(src: Tensor, idx: Tensor, out: Tensor) -> Tensor:
    assert idx.dtype == torch.int64 and idx.is_contiguous() and idx.is_cuda
    _, w, lds = _rows(_f32(src, "src"), "src")
    n, w2, ldo = _rows(_f32(out, "out"), "out")
    assert w == w2 and n == idx.numel()
    check(lib.pm_gather_rows(_p(src), lds, _p(idx), _p(out), ldo, n, w, _stream()), "pm_gather_rows")
    return out


def copy_rows(src: Tensor, dst: Tensor) -> Tensor:
    n, w, lds = _rows(_f32(src, "src"), "src")
    n2, w2, ldd = _rows(_f32(dst, "dst"), "dst")
    assert (n, w) == (n2, w2)
    check(lib.pm_copy_rows(_p(src), lds, _p(dst), ldd, n, w, _stream()), "pm_copy_rows")
    return dst


# ------------------------------------------------------------------------------------------- next row: Conv3D student
def conv3d_out_dim(Din: int, k: int, s: int) -> int:
    return int(lib.pm_conv3d_out_dim(int(Din), int(k), int(s)))


def conv3d_im2col(src: Tensor, ld_in: int, sample_stride: int, B: int, C: int, Din: int, k: int, s: int, cols: Tensor):
    """network.py:56-63 patches in TAP-major column order (kd, kh, kw, c): src holds voxel (d,h,w) of sample b at
    b*sample_stride + ((d*Din+h)*Din+w)*ld_in + c."""
    assert _f32(src, "src").is_cuda and _f32(cols, "cols").is_contiguous() and cols.dim() == 2
    check(lib.pm_conv3d_im2col(_p(src), int(ld_in), int(sample_stride), B, C, Din, k, s, _p(cols), cols.shape[1], _stream()),
          "pm_conv3d_im2col")
    return cols


def conv3d_col2im(dcols: Tensor, B: int, C: int, Din: int, k: int, s: int, y: Tensor, act, din: Tensor):
    assert _f32(dcols, "dcols").is_contiguous() and _f32(y, "y").is_contiguous() and _f32(din, "din").is_contiguous()
    check(lib.pm_conv3d_col2im(_p(dcols), dcols.shape[1], B, C, Din, k, s, _p(y), PM_ACT[act], _p(din), _stream()), "pm_conv3d_col2im")
    return din


def conv3d_weight_permute(src: Tensor, Cout: int, C: int, k: int, to_tap_major: bool, dst: Tensor):
    """nn.Conv3d weight (Cout, C, k^3) <-> the tap-major patch order (Cout, k^3, C) of conv3d_im2col."""
    assert _f32(src, "src").is_contiguous() and _f32(dst, "dst").is_contiguous() and src.numel() == dst.numel() == Cout * C * k ** 3
    check(lib.pm_conv3d_weight_permute(_p(src), Cout, C, k, int(to_tap_major), _p(dst), _stream()), "pm_conv3d_weight_permute")
    return dst


def conv3d_first_forward(x: Tensor, Din: int, w: Tensor, bias: Tensor, act, y: Tensor, stride: int = 3):
    """Conv3d(1,16,5,stride,padding 2) + act on volume rows x (B, >= Din^3) -> y ((b, voxel'), 16), exact fp32."""
    B, _, ldx = _rows(_f32(x, "x"), "x")
    assert _f32(w, "w").is_contiguous() and w.numel() == 16 * 125 and _f32(y, "y").is_contiguous()
    check(lib.pm_conv3d_first_forward(_p(x), ldx, B, int(Din), int(stride), _p(w), _p(bias), PM_ACT[act], _p(y), _stream()), "pm_conv3d_first_forward")
    return y


def conv3d_first_backward(x: Tensor, Din: int, dpre: Tensor, dW: Tensor, db: Tensor, stride: int = 3):
    B, _, ldx = _rows(_f32(x, "x"), "x")
    assert _f32(dpre, "dpre").is_contiguous() and dpre.shape[1] == 16 and dW.is_contiguous() and dW.numel() == 16 * 125
    ws = scratch(lib.pm_conv3d_first_backward_ws_bytes(), x.device, "conv1bwd")
    check(lib.pm_conv3d_first_backward(_p(x), ldx, B, int(Din), int(stride), _p(dpre), _p(dW), _p(ws), _stream()), "pm_conv3d_first_backward")
    rms_colsum(dpre, db)


def maxpool3d_forward(y: Tensor, B: int, C: int, Din: int, k: int, out: Tensor, argmax: Tensor):
    """nn.MaxPool3d(k) on channels-last rows ((b, voxel), C) -> out ((b, cell), C), argmax int32."""
    Dp = (Din - k) // k + 1
    assert _f32(y, "y").is_contiguous() and y.shape == (B * Din ** 3, C) and _f32(out, "out").is_contiguous() and out.shape == (B * Dp ** 3, C)
    assert argmax.dtype == torch.int32 and argmax.is_contiguous() and argmax.shape == out.shape and argmax.is_cuda
    check(lib.pm_maxpool3d_forward(_p(y), B, C, int(Din), int(k), _p(out), _p(argmax), _stream()), "pm_maxpool3d_forward")
    return out


def maxpool3d_backward(dout: Tensor, argmax: Tensor, y: Tensor, act, B: int, C: int, Din: int, k: int, dpre: Tensor):
    assert _f32(dout, "dout").is_contiguous() and dout.shape == argmax.shape and _f32(dpre, "dpre").is_contiguous() and dpre.shape == y.shape
    check(lib.pm_maxpool3d_backward(_p(dout), _p(argmax), _p(_f32(y, "y")), PM_ACT[act], B, C, int(Din), int(k), _p(dpre), _stream()),
          "pm_maxpool3d_backward")
    return dpre


def conv3d_flatten(src: Tensor, dst: Tensor, B: int, P: int, C: int, ld_row: int, to_rows: bool):
    check(lib.pm_conv3d_flatten(_p(_f32(src, "src")), _p(_f32(dst, "dst")), B, P, C, int(ld_row), int(bool(to_rows)), _stream()),
          "pm_conv3d_flatten")
    return dst


# ------------------------------------------------------------------------------------------- next row: mesh -> TSDF
def mesh2sdf_query(sdf_field: Tensor, sdf_res: Tensor, sdf_voxel: Tensor, sdf_bbox_min: Tensor, bbox_res_y: int, bbox_res_z: int,
                   pose_R: Tensor, pose_T: Tensor, init_tsdf: Tensor, resolution: int, vox_origin, size: float,
                   out: Optional[Tensor] = None) -> Tensor:
    """utils/mesh2sdf.py:119-139 + 239-272 in one launch -> (E, R, R, R)."""
    E, M = pose_R.shape[0], pose_R.shape[1]
    R = int(resolution)
    assert _f32(sdf_field, "sdf_field").is_contiguous() and sdf_field.shape[0] == M and sdf_res.dtype == torch.int32 and sdf_res.shape == (M, 3)
    assert _f32(pose_R, "pose_R").is_contiguous() and pose_R.shape == (E, M, 3, 3) and _f32(pose_T, "pose_T").is_contiguous() and pose_T.shape == (E, M, 3)
    assert _f32(init_tsdf, "init_tsdf").is_contiguous() and init_tsdf.numel() == E * R ** 3
    assert _f32(sdf_voxel, "sdf_voxel").numel() == M and _f32(sdf_bbox_min, "sdf_bbox_min").is_contiguous() and sdf_bbox_min.shape == (M, 3)
    if out is None:
        out = torch.empty(E, R, R, R, device=pose_R.device, dtype=torch.float32)
    org = (ct.c_float * 3)(*[float(v) for v in vox_origin])
    check(lib.pm_mesh2sdf_query(_p(sdf_field), sdf_field.shape[1], _p(sdf_res), _p(sdf_voxel), _p(sdf_bbox_min), M, int(bbox_res_y),
                                int(bbox_res_z), _p(pose_R), _p(pose_T), _p(init_tsdf), E, R, org, float(size), _p(out), _stream()),
          "pm_mesh2sdf_query")
    return out


# ------------------------------------------------------------------------------------------- next row: depth -> point cloud
def depth2pc_backproject(depth: Tensor, cam_intr, cam_pose: Tensor, vol_origin, size: float, out: Optional[Tensor] = None) -> Tensor:
    """utils/depth2tsdf.py:146-159.  depth (E,M,H,W) fp32 CUDA, cam_pose (M,4,4) fp32 CUDA -> masked world cloud (E, M*H*W, 3)."""
    E, M, H, W = depth.shape
    assert _f32(depth, "depth").is_contiguous() and _f32(cam_pose, "cam_pose").is_contiguous() and cam_pose.shape == (M, 4, 4)
    if out is None:
        out = torch.empty(E, M * H * W, 3, device=depth.device, dtype=torch.float32)
    intr = (ct.c_float * 9)(*[float(v) for row in cam_intr for v in row])
    org = (ct.c_float * 3)(*[float(v) for v in vol_origin])
    check(lib.pm_depth2pc_backproject(_p(depth), E, M, H, W, intr, _p(cam_pose), org, float(size), _p(out), _stream()),
          "pm_depth2pc_backproject")
    return out


def tsdf_voxel_tables(cam_pose: Tensor, cam_intr, im_h: int, im_w: int, size: float, resolution: int, vol_origin) -> Tuple[Tensor, Tensor]:
    """utils/depth2tsdf.py:14-62: per-view voxel -> pixel tables.  -> pix_off (M, R^3) int32 (row*W+col, -1 invalid), pix_z (M, R^3)."""
    M = cam_pose.shape[0]
    assert _f32(cam_pose, "cam_pose").is_contiguous() and cam_pose.shape == (M, 4, 4)
    R3 = int(resolution) ** 3
    pix_off = torch.empty(M, R3, device=cam_pose.device, dtype=torch.int32)
    pix_z = torch.empty(M, R3, device=cam_pose.device, dtype=torch.float32)
    intr = (ct.c_float * 9)(*[float(v) for row in cam_intr for v in row])
    org = (ct.c_float * 3)(*[float(v) for v in vol_origin])
    check(lib.pm_tsdf_voxel_tables(_p(cam_pose), M, intr, int(im_h), int(im_w), float(size), int(resolution), org, _p(pix_off), _p(pix_z),
                                   _stream()), "pm_tsdf_voxel_tables")
    return pix_off, pix_z


def tsdf_integrate(depth: Tensor, pix_off: Tensor, pix_z: Tensor, size: float, resolution: int, default_tsdf: float = 1.0,
                   out: Optional[Tensor] = None) -> Tensor:
    """utils/depth2tsdf.py:68-86: depth (E,M,H,W) fp32 -> fused TSDF volume (E, R, R, R)."""
    E, M, H, W = depth.shape
    R = int(resolution)
    assert _f32(depth, "depth").is_contiguous() and pix_off.dtype == torch.int32 and pix_off.shape == (M, R ** 3) and pix_off.is_contiguous()
    assert _f32(pix_z, "pix_z").is_contiguous() and pix_z.shape == (M, R ** 3)
    if out is None:
        out = torch.empty(E, R, R, R, device=depth.device, dtype=torch.float32)
    check(lib.pm_tsdf_integrate(_p(depth), E, M, H, W, _p(pix_off), _p(pix_z), float(size), R, float(default_tsdf), _p(out), _stream()),
          "pm_tsdf_integrate")
    return out


def tsdf_sparse_voxel(tsdf_vol: Tensor, K: int = 1024, lo: float = -0.2, hi: float = 0.2) -> Tensor:
    """utils/depth2tsdf.py:103-119: voxels with lo < tsdf < hi -> K farthest (on integer coordinates) -> (E, K, 4) = (x, y, z, tsdf)."""
    E, R = tsdf_vol.shape[0], tsdf_vol.shape[1]
    assert _f32(tsdf_vol, "tsdf_vol").is_contiguous() and tuple(tsdf_vol.shape) == (E, R, R, R)
    out = torch.empty(E, K, 4, device=tsdf_vol.device, dtype=torch.float32)
    nbytes = lib.pm_tsdf_sparse_voxel_ws_bytes(E, R, int(K))
    ws = scratch(nbytes, tsdf_vol.device, "fps")
    check(lib.pm_tsdf_sparse_voxel(_p(tsdf_vol), E, R, float(lo), float(hi), int(K), _p(out), _p(ws), nbytes, _stream()),
          "pm_tsdf_sparse_voxel")
    return out


def view_pointer_table(camera_tensor_list) -> Tuple[Tensor, bool]:
    """Device table of the E*M image pointers of `camera_tensor_list[env][view]` (each an (H,W) fp32 CUDA tensor, as Isaac Gym's
    camera tensors are) + whether all of them are 16-byte aligned.  The simulator reuses the buffers, so build it once."""
    ptrs = []
    for views in camera_tensor_list:
        for t in views:
            assert _f32(t, "camera tensor").is_contiguous()
            ptrs.append(t.data_ptr())
    dev = camera_tensor_list[0][0].device
    table = torch.tensor(ptrs, dtype=torch.int64).to(dev)
    return table, all(p % 16 == 0 for p in ptrs)


def depth2pc_backproject_views(table: Tensor, aligned16: bool, E: int, M: int, H: int, W: int, cam_intr, cam_pose: Tensor, vol_origin,
                               size: float, negate: bool = True, inf_value: float = 100.0, out: Optional[Tensor] = None) -> Tensor:
    """tasks/hand_base.py:317-324 (stack, negate, inf -> 100) + utils/depth2tsdf.py:146-159 in one pass over the camera images."""
    assert table.dtype == torch.int64 and table.numel() == E * M and table.is_cuda
    assert _f32(cam_pose, "cam_pose").is_contiguous() and cam_pose.shape == (M, 4, 4)
    if out is None:
        out = torch.empty(E, M * H * W, 3, device=table.device, dtype=torch.float32)
    intr = (ct.c_float * 9)(*[float(v) for row in cam_intr for v in row])
    org = (ct.c_float * 3)(*[float(v) for v in vol_origin])
    check(lib.pm_depth2pc_backproject_views(_p(table), E, M, H, W, int(aligned16), int(negate), float(inf_value), intr, _p(cam_pose), org,
                                            float(size), _p(out), _stream()), "pm_depth2pc_backproject_views")
    return out


def farthest_point_sample(points: Tensor, K: int, return_idx: bool = False, compact=None):
    """pytorch3d.ops.sample_farthest_points(points, K=K) semantics (start index 0, first index on ties): (E,P,3) -> (E,K,3).

    compact: None/True = auto (P > 48 Ki: one 8-CTA cluster per cloud with exact bounding-box pruning), False = no compaction,
    2 = streaming cluster kernel, 3 = one CTA per cloud, 4 = pruned cluster kernel."""
    E, P, three = points.shape
    assert three == 3 and _f32(points, "points").is_contiguous()
    out = torch.empty(E, K, 3, device=points.device, dtype=torch.float32)
    idx = torch.empty(E, K, device=points.device, dtype=torch.int64) if return_idx else None
    if compact is None:
        compact = P % 4 == 0          # identical picks, K passes over the valid points only
    nbytes = lib.pm_fps_ws_bytes(E, P)
    ws = scratch(nbytes, points.device, "fps")
    check(lib.pm_farthest_point_sample(_p(points), E, P, int(K), int(compact), _p(out), _p(idx), _p(ws), nbytes, _stream()),
          "pm_farthest_point_sample")
    return (out, idx) if return_idx else out


# ------------------------------------------------------------------------------------------- next row: env-side arithmetic (open_drawer)
def _i64(t: Tensor, name: str) -> Tensor:
    if not t.is_cuda or t.dtype != torch.int64 or not t.is_contiguous():
        raise TypeError(f"{name}: expected a contiguous CUDA int64 tensor")
    return t


class OpenDrawerPostPlan:
    """tasks/open_drawer.py:240-281 + 170-238 (+ load_robot.py:153-164, hand_base.py:388) in ONE launch.  The simulator tensors,
    index tables and result buffers are persistent, so the arguments are validated and converted once; a call costs one ctypes
    dispatch.  `out` holds the preallocated result tensors (see tasks/step_kernels.py:_pm_buffers)."""

    def __init__(self, dof_state_all: Tensor, rigid_body_all: Tensor, root_tensor: Tensor, obj_actor: int, dof_state_mask: Tensor,
                 rigid_body_mask: Tensor, ltip_rb_index: int, rtip_rb_index: int, dof_lower: Tensor, dof_upper: Tensor, part_bbox_init: Tensor,
                 part_axis_dir_init: Tensor, part_joint_lower: Tensor, part_joint_upper: Tensor, obj_lstid: Tensor, suc_prop: float,
                 progress_buf: Tensor, succ_objid: Tensor, out: dict):
        E, ndp = dof_state_mask.shape
        nd, nb = ndp - 1, rigid_body_mask.shape[1] - 2
        _i64(dof_state_mask, "dof_state_mask"), _i64(rigid_body_mask, "rigid_body_mask"), _i64(obj_lstid, "obj_lstid"), _i64(progress_buf, "progress_buf")
        for n, t in (("dof_state_all", dof_state_all), ("rigid_body_all", rigid_body_all), ("root_tensor", root_tensor), ("dof_lower", dof_lower),
                     ("dof_upper", dof_upper), ("part_bbox_init", part_bbox_init), ("part_axis_dir_init", part_axis_dir_init),
                     ("part_joint_lower", part_joint_lower), ("part_joint_upper", part_joint_upper)):
            assert _f32(t, n).is_contiguous(), n
        assert dof_state_all.shape[-1] == 2 and rigid_body_all.shape[-1] == 13 and root_tensor.shape[0] == E and root_tensor.shape[-1] == 13
        assert part_bbox_init.shape == (E, 8, 3) and part_axis_dir_init.numel() == 3 * E and part_joint_lower.numel() == E and part_joint_upper.numel() == E
        assert dof_lower.numel() == nd and dof_upper.numel() == nd and succ_objid.dtype in (torch.bool, torch.uint8) and succ_objid.is_cuda
        assert int(dof_state_mask.max()) < dof_state_all.shape[0] and int(rigid_body_mask.max()) < rigid_body_all.shape[0] and int(dof_state_mask.min()) >= 0
        assert int(obj_lstid.max()) < succ_objid.numel() and progress_buf.numel() == E
        shapes = dict(obs=(E, 29 + 2 * nd), part_bbox=(E, 8, 3), dof_state_tensor=(E, nd + 1, 2), rigid_body_tensor=(E, nb + 2, 13), tip_rb_tensor=(E, 13),
                      tip_rot_9d=(E, 3, 3), gripper_length=(E,), dof_qpos_normalized=(E, nd), rew_buf=(E,), success=(E,), extras_f=(6, E), extras_b=(3, E))
        for k, shp in shapes.items():
            t = out[k]
            assert t.is_cuda and t.is_contiguous() and tuple(t.shape) == shp, k
            assert t.dtype == (torch.bool if k in ("success", "extras_b") else torch.float32), k
        # the tensors whose addresses are baked into the argument list (kept alive; `bound_to` tells a caller whether to re-plan)
        self._keep = (dof_state_all, rigid_body_all, root_tensor, dof_state_mask, rigid_body_mask, dof_lower, dof_upper, part_bbox_init, part_axis_dir_init,
                      part_joint_lower, part_joint_upper, obj_lstid, progress_buf, succ_objid, out)
        self._bound = dict(dof_state_all=dof_state_all, rigid_body_all=rigid_body_all, root_tensor=root_tensor, progress_buf=progress_buf,
                           succ_objid=succ_objid)
        self._head = (_p(dof_state_all), _p(rigid_body_all), _p(root_tensor), root_tensor.shape[1], int(obj_actor), _p(dof_state_mask), _p(rigid_body_mask),
                      E, nd, nb, int(ltip_rb_index), int(rtip_rb_index), _p(dof_lower), _p(dof_upper), _p(part_bbox_init), _p(part_axis_dir_init),
                      _p(part_joint_lower), _p(part_joint_upper), _p(obj_lstid), float(suc_prop))
        self._tail = (_p(progress_buf), _p(out["obs"]), _p(out["part_bbox"]), _p(out["dof_state_tensor"]), _p(out["rigid_body_tensor"]),
                      _p(out["tip_rb_tensor"]), _p(out["tip_rot_9d"]), _p(out["gripper_length"]), _p(out["dof_qpos_normalized"]), _p(out["rew_buf"]),
                      _p(out["success"]), _p(succ_objid), _p(out["extras_f"]), _p(out["extras_b"]))

    def bound_to(self, **tensors) -> bool:
        """True if every named tensor is the very object this plan was built on."""
        return all(self._bound[k] is t for k, t in tensors.items())

    def __call__(self, do_obs: bool = True, do_reward: bool = True, advance_progress: bool = False) -> None:
        check(lib.pm_open_drawer_post_physics(*self._head, int(do_obs), int(do_reward), int(advance_progress), *self._tail, _stream()),
              "pm_open_drawer_post_physics")


class GraspCubePostPlan:
    """tasks/grasp_cube.py:118-138 + 66-115 (+ load_robot.py:153-164, hand_base.py:388) in ONE launch; arguments converted once."""

    def __init__(self, dof_state: Tensor, rigid_body: Tensor, root_tensor: Tensor, obj_actor: int, num_dofs: int, ltip_rb_index: int,
                 rtip_rb_index: int, dof_lower: Tensor, dof_upper: Tensor, pose_lower_limit, pose_upper_limit, success_pos, obj_default_pos,
                 goal_thresh: float, progress_buf: Tensor, out: dict):
        E = dof_state.shape[0]
        nd = int(num_dofs)
        for n, t in (("dof_state", dof_state), ("rigid_body", rigid_body), ("root_tensor", root_tensor), ("dof_lower", dof_lower), ("dof_upper", dof_upper)):
            assert _f32(t, n).is_contiguous(), n
        assert dof_state.dim() == 3 and dof_state.shape[2] == 2 and dof_state.shape[1] >= nd
        assert rigid_body.shape[0] == E and rigid_body.shape[2] == 13 and root_tensor.shape[0] == E and root_tensor.shape[2] == 13
        assert dof_lower.numel() == nd and dof_upper.numel() == nd and _i64(progress_buf, "progress_buf").numel() == E
        shapes = dict(obs=(E, 19 + 2 * nd), proprio=(E, 7 + 2 * nd), tip_rb_tensor=(E, 13), tip_rot_9d=(E, 3, 3), gripper_length=(E,),
                      dof_qpos_normalized=(E, nd), rew_buf=(E,), success=(E,), extras_f=(7, E), extras_b=(2, E))
        for k, shp in shapes.items():
            t = out[k]
            assert t.is_cuda and t.is_contiguous() and tuple(t.shape) == shp, k
            assert t.dtype == (torch.bool if k in ("success", "extras_b") else torch.float32), k
        arr = lambda v, n: (ct.c_float * n)(*[float(x) for x in v])
        self._keep = (dof_state, rigid_body, root_tensor, dof_lower, dof_upper, progress_buf, out)
        self._bound = dict(dof_state=dof_state, rigid_body=rigid_body, root_tensor=root_tensor, progress_buf=progress_buf)
        self._head = (_p(dof_state), dof_state.shape[1], _p(rigid_body), rigid_body.shape[1], _p(root_tensor), root_tensor.shape[1], int(obj_actor), E, nd,
                      int(ltip_rb_index), int(rtip_rb_index), _p(dof_lower), _p(dof_upper), arr(pose_lower_limit, 7), arr(pose_upper_limit, 7),
                      arr(success_pos, 3), arr(obj_default_pos, 3), float(goal_thresh))
        self._tail = (_p(progress_buf), _p(out["obs"]), _p(out["proprio"]), _p(out["tip_rb_tensor"]), _p(out["tip_rot_9d"]), _p(out["gripper_length"]),
                      _p(out["dof_qpos_normalized"]), _p(out["rew_buf"]), _p(out["success"]), _p(out["extras_f"]), _p(out["extras_b"]))

    def bound_to(self, **tensors) -> bool:
        return all(self._bound[k] is t for k, t in tensors.items())

    def __call__(self, do_obs: bool = True, do_reward: bool = True, advance_progress: bool = False) -> None:
        check(lib.pm_grasp_cube_post_physics(*self._head, int(do_obs), int(do_reward), int(advance_progress), *self._tail, _stream()),
              "pm_grasp_cube_post_physics")


def franka_control(raw_output: Tensor, drive_mode: str, mobile: bool, qpos: Tensor, num_dofs: int, dof_lower: Tensor, dof_upper: Tensor,
                   default_root_quat, dt: float, action_tensor: Tensor, dof_state_mask: Optional[Tensor] = None, jacobian: Optional[Tensor] = None,
                   ltip_rb_index: int = 0, rtip_rb_index: int = 0, jacobian_sum: Optional[Tensor] = None, damping: float = 0.05) -> Tensor:
    """tasks/load_robot.py:96-118 + 142-151 in one launch -> action_tensor (E, num_dofs), clamped to the joint limits.
    qpos: the simulator's dof_state_all (with dof_state_mask, the task's index table) or a strided (E, num_dofs) view of the
    current joint positions (franka.dof_qpos_raw)."""
    E = raw_output.shape[0]
    mode = {"pos": 0, "ik": 1}.get(drive_mode)
    if mode is None:
        raise NotImplementedError(drive_mode)
    lo = 3 if mobile else 0
    assert _f32(raw_output, "raw_output").is_contiguous() and raw_output.shape[1] == (7 if mode else 8) + lo
    assert _f32(action_tensor, "action_tensor").is_contiguous() and action_tensor.shape == (E, num_dofs)
    assert _f32(dof_lower, "dof_lower").numel() == num_dofs and _f32(dof_upper, "dof_upper").numel() == num_dofs
    _f32(qpos, "qpos", last_contig=False)
    if dof_state_mask is not None:
        assert qpos.is_contiguous() and qpos.shape[-1] == 2 and _i64(dof_state_mask, "dof_state_mask").shape[0] == E
        rs, es, mld = 0, 0, dof_state_mask.shape[1]
    else:
        assert qpos.shape == (E, num_dofs)
        rs, es, mld = qpos.stride(0), qpos.stride(1), 0
    n_links = 0
    if mode:
        assert _f32(jacobian, "jacobian").is_contiguous() and jacobian.shape[0] == E and jacobian.shape[2:] == (6, num_dofs)
        n_links = jacobian.shape[1]
    quat = (ct.c_float * 4)(*[float(v) for v in default_root_quat]) if mobile else None
    check(lib.pm_franka_control(_p(raw_output), E, int(num_dofs), int(mobile), mode, _p(qpos), rs, es, _p(dof_state_mask), mld, _p(jacobian), n_links,
                                int(ltip_rb_index), int(rtip_rb_index), _p(dof_lower), _p(dof_upper), quat, float(dt), float(damping),
                                _p(action_tensor), _p(jacobian_sum), _stream()), "pm_franka_control")
    return action_tensor


def episode_flags(train: bool, rew_buf: Optional[Tensor], progress_buf: Tensor, success: Optional[Tensor], epis_max_rew: Optional[Tensor],
                  epis_max_step: Optional[Tensor], explore_step: int, max_episode_length: int, reset_buf: Tensor, reset_succ: Optional[Tensor],
                  counts: Tensor, succ_rate: Optional[Tensor]) -> None:
    """tasks/hand_base.py:367-377 in one launch (reset_buf / reset_succ / success: bool or uint8 storage)."""
    E = progress_buf.shape[0]
    assert _i64(progress_buf, "progress_buf") is not None and counts.dtype == torch.int32 and counts.numel() >= 3 and counts.is_cuda
    assert reset_buf.dtype in (torch.bool, torch.uint8) and reset_buf.is_cuda and reset_buf.numel() == E
    if train:
        assert _f32(rew_buf, "rew_buf").numel() == E and _f32(epis_max_rew, "epis_max_rew").numel() == E and _i64(epis_max_step, "epis_max_step").numel() == E
        assert success.dtype in (torch.bool, torch.uint8) and reset_succ.dtype in (torch.bool, torch.uint8) and succ_rate.dtype == torch.float32
    check(lib.pm_episode_flags(E, int(train), _p(rew_buf), _p(progress_buf), _p(success), _p(epis_max_rew), _p(epis_max_step), int(explore_step),
                               int(max_episode_length), _p(reset_buf), _p(reset_succ), _p(counts), _p(succ_rate), _stream()), "pm_episode_flags")


def scatter_dof_targets(pos_act: Tensor, dof_state_mask: Tensor, num_dofs: int, pos_act_all: Tensor) -> None:
    """tasks/hand_base.py:382."""
    assert _f32(pos_act, "pos_act").is_contiguous() and pos_act.shape[1] == num_dofs and _f32(pos_act_all, "pos_act_all").is_contiguous()
    check(lib.pm_scatter_dof_targets(_p(pos_act), _p(_i64(dof_state_mask, "dof_state_mask")), dof_state_mask.shape[1], pos_act.shape[0], int(num_dofs),
                                     _p(pos_act_all), _stream()), "pm_scatter_dof_targets")
