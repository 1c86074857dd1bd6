"""GPU parity tests: every kernel is called through the C-ABI (partmanip_b200.ops -> ctypes -> libpartmanip_b200.so)
and compared with (a) the golden fixtures recorded from the unmodified reference and (b) the CPU oracle on seeded
inputs.  fp32 gate (north_star): |a-b| <= 1e-4 + 1e-4*|b|; integer/bit-exact where the kernel promises it."""
import math

import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import close, load_golden, max_err, sub

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from partmanip_b200 import ops as _ops
    return _ops


def cu(t):
    return t.to(DEV).contiguous()


# ------------------------------------------------------------------------------------------------ K6
def test_rms_golden_sequence(ops):
    g = load_golden("rms.npz")
    D = 11
    mean, S = torch.zeros(1, D, device=DEV), torch.ones(1, D, device=DEV) * 1e-4
    std = torch.sqrt(S)
    n = 0
    for k in range(4):
        upd = k != 3
        n += int(upd)
        x = cu(g["x"][k])
        y = ops.rms_forward(x, torch.empty_like(x), mean, S, std, n, upd)
        assert close(y.cpu(), g["y"][k], 1e-5, 1e-5), max_err(y.cpu(), g["y"][k])
        assert close(mean.cpu(), g["mean"][k], 1e-6, 1e-6) and close(S.cpu(), g["S"][k], 1e-5, 1e-6)
        assert close(std.cpu(), g["std"][k], 1e-5, 1e-6)


@pytest.mark.parametrize("E,D", [(512, 3072), (64, 37), (4096, 3072), (33, 3097)])
def test_rms_vs_oracle(ops, E, D):
    torch.manual_seed(E + D)
    rs = O.RunningStats(D)
    mean, S = torch.zeros(1, D, device=DEV), torch.ones(1, D, device=DEV) * 1e-4
    std = torch.sqrt(S)
    for k in range(3):
        x = torch.randn(E, D) * (1 + k) + 0.5 * k
        want = rs.normalize(x, True)
        got = ops.rms_forward(cu(x), torch.empty(E, D, device=DEV), mean, S, std, k + 1, True)
        assert close(got.cpu(), want, 1e-4, 1e-4), max_err(got.cpu(), want)
        assert close(mean.cpu(), rs.mean, 1e-5, 1e-6) and close(std.cpu(), rs.std, 1e-5, 1e-6)


# ------------------------------------------------------------------------------------------------ K5
def _gae_gpu(ops, rew, val, done, succ, last, gamma, lam, sv, norm):
    r, a = ops.gae(cu(rew), cu(val), cu(done), cu(succ), cu(last), gamma, lam, sv)
    if norm:
        ops.normalize_(a)
    return r.cpu(), a.cpu()


def test_gae_golden_bit_exact(ops):
    g = load_golden("gae.npz")
    r, a = _gae_gpu(ops, g["k2_rew"], g["k2_val"], g["k2_done"], g["k2_succ"], g["k2_last"], 0.99, 0.95, None, False)
    assert torch.equal(r, g["k2_ret"]) and torch.equal(a, g["k2_adv"])
    r, a = _gae_gpu(ops, g["k2_rew"], g["k2_val"], g["k2_done"], g["k2b_succ"], g["k2_last"], 0.99, 0.95, 500, True)
    assert torch.equal(r, g["k2b_ret"]) and close(a, g["k2b_adv"], 1e-5, 1e-6)
    r, a = _gae_gpu(ops, g["r_rew"], g["r_val"], g["r_done"], g["r_succ"], g["r_last"], 0.99, 0.95, 500, False)
    assert torch.equal(r, g["r3_ret"]) and torch.equal(a, g["r3_adv"])
    r, a = _gae_gpu(ops, g["r_rew"], g["r_val"], g["r_done"], g["r_succ"], g["r_last"], 0.97, 0.9, None, True)
    assert torch.equal(r, g["r4_ret"]) and close(a, g["r4_adv"], 1e-5, 1e-6)


@pytest.mark.parametrize("T,E", [(8, 4096), (1, 1), (8, 3), (16, 1000)])
def test_gae_vs_oracle_bit_exact(ops, T, E):
    g = torch.Generator().manual_seed(T * 1000 + E)
    rew, val = torch.randn(T, E, 1, generator=g), torch.randn(T, E, 1, generator=g) * 3
    done, succ = torch.rand(T, E, 1, generator=g) < 0.1, torch.rand(T, E, 1, generator=g) < 0.05
    last = torch.randn(E, 1, generator=g)
    for sv in (None, 500):
        want_r, want_a = O.gae(rew, val, done, succ, last, 0.99, 0.95, sv)
        r, a = ops.gae(cu(rew), cu(val), cu(done), cu(succ), cu(last).view(-1), 0.99, 0.95, sv)
        assert torch.equal(r.cpu(), want_r) and torch.equal(a.cpu(), want_a)


# ------------------------------------------------------------------------------------------------ K4
def test_randn_moments_and_determinism(ops):
    a = ops.randn(torch.empty(1 << 20, device=DEV), 7, 0)
    b = ops.randn(torch.empty(1 << 20, device=DEV), 7, 0)
    c = ops.randn(torch.empty(1 << 20, device=DEV), 8, 0)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 5e-3 and abs(float(a.std()) - 1) < 5e-3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05 and torch.isfinite(a).all()
    # odd sizes / offsets continue the stream
    d = ops.randn(torch.empty(5, device=DEV), 7, 1)
    assert torch.equal(d[:4], a[4:8])


def test_policy_sample_and_logprob_vs_oracle(ops):
    torch.manual_seed(3)
    E, A = 777, 10
    mu, ls, eps = torch.randn(E, A), torch.randn(A) * 0.3 - 0.5, torch.randn(E, A)
    want_a, want_lp = O.policy_sample(mu, ls, eps, 1.0)
    act, lp, sig = ops.policy_sample(cu(mu), cu(ls), cu(eps), 1.0, True)
    assert close(act.cpu(), want_a, 1e-5, 1e-6) and close(lp.cpu(), want_lp, 1e-5, 1e-5)
    assert torch.equal(sig.cpu(), ls.repeat(E, 1))
    lp2, ent = ops.policy_logprob(cu(mu), cu(ls), act, 1.0, True)
    want_lp2 = O.gaussian_logp(mu, ls, O.action_deactivation(want_a, 1.0))
    assert close(lp2.cpu(), want_lp2, 1e-4, 1e-4)
    assert close(ent.cpu(), O.gaussian_entropy(ls, E), 1e-5, 1e-5)
    act2, _, _ = ops.policy_sample(cu(mu), cu(ls), cu(eps), 2.5, False)
    assert close(act2.cpu(), mu + O.policy_std(ls) * eps, 1e-6, 1e-6)


@pytest.mark.parametrize("B,A,mini_norm", [(2048, 10, False), (300, 7, True), (1, 10, False)])
def test_actor_loss_forward_backward_vs_oracle(ops, B, A, mini_norm):
    torch.manual_seed(B + A)
    ls = (torch.randn(A) * 0.2 - 0.6).requires_grad_(True)
    mu = (torch.randn(B, A) * 0.5).requires_grad_(True)
    act = torch.tanh(torch.randn(B, A))
    act[0, 0] = 1.0                                            # saturated action -> clamp path (Q4)
    mu_old, sig_old = mu.detach() + 0.05 * torch.randn(B, A), (ls.detach() + 0.01).repeat(B, 1)
    logp_old = O.gaussian_logp(mu_old, ls.detach(), O.action_deactivation(act, 1.0)) + 0.3 * torch.randn(B)
    adv = torch.randn(B)
    adv[B // 2] = 0.0
    a_used = O.mini_adv_norm(adv[:, None]).squeeze(-1) if (mini_norm and B > 1) else adv
    logp = O.gaussian_logp(mu, ls, O.action_deactivation(act, 1.0))
    loss = O.surrogate_loss(logp, logp_old, a_used, 0.2)
    loss.backward()
    kl = O.kl_old_new(mu.detach(), ls.detach().repeat(B, 1), mu_old, sig_old)
    stats, dmu, dls = torch.zeros(2, device=DEV), torch.empty(B, A, device=DEV), torch.empty(A, device=DEV)
    adv_stats = None
    if mini_norm and B > 1:
        adv_stats = ops.normalize_stats(cu(adv), torch.zeros(2, device=DEV))
    lp_out = torch.empty(B, device=DEV)
    ops.ppo_actor_loss(cu(mu.detach()), cu(ls.detach()), cu(act), cu(logp_old), cu(mu_old), cu(sig_old), cu(adv), adv_stats,
                       1.0 / B, 0.2, 1.0, True, stats, dmu, dls, lp_out)
    assert close(lp_out.cpu(), logp.detach(), 1e-4, 1e-4)
    assert close(stats[0].cpu() / B, loss.detach(), 1e-4, 1e-5), (float(stats[0]) / B, float(loss))
    assert close(stats[1].cpu() / B, kl.mean(), 1e-4, 1e-5)
    assert close(dmu.cpu(), mu.grad, 1e-4, 1e-6), max_err(dmu.cpu(), mu.grad)
    assert close(dls.cpu(), ls.grad, 1e-4, 1e-5), max_err(dls.cpu(), ls.grad)


def test_actor_finalize_skip_logic(ops):
    acc = torch.zeros(8, device=DEV)
    skip = torch.zeros(1, device=DEV, dtype=torch.int32)
    ops.ppo_actor_finalize(torch.tensor([4.0, 0.2], device=DEV), 0.5, 0.2, acc, skip)       # kl_mean 0.1 <= 0.2
    assert int(skip) == 0 and acc[:4].tolist() == pytest.approx([2.0, 0.1, 1.0, 0.1])
    ops.ppo_actor_finalize(torch.tensor([8.0, 1.0], device=DEV), 0.5, 0.2, acc, skip)       # kl_mean 0.5 > 0.2 -> skip
    assert int(skip) == 1 and acc[:4].tolist() == pytest.approx([2.0, 0.1, 1.0, 0.5])


@pytest.mark.parametrize("clipped", [False, True])
def test_value_loss_vs_oracle(ops, clipped):
    torch.manual_seed(11)
    B = 1500
    v = torch.randn(B, 1, requires_grad=True)
    ret, old = torch.randn(B, 1) * 2, torch.randn(B, 1)
    loss = O.value_loss(v, ret, old, 0.2, clipped)
    loss.backward()
    stats, dv = torch.zeros(2, device=DEV), torch.empty(B, 1, device=DEV)
    delta = None
    if clipped:
        delta = torch.zeros(1, device=DEV)
        ops.abs_sum(cu(old).view(-1), 0.2 / B, delta)
        assert close(delta.cpu(), (0.2 * old).abs().mean(), 1e-5, 1e-7)
    ops.value_loss(cu(v.detach()), cu(ret).view(-1), cu(old).view(-1), delta, 1.0 / B, stats, dv)
    assert close(stats[0].cpu() / B, loss.detach(), 1e-4, 1e-6)
    assert close(dv.cpu(), v.grad, 1e-4, 1e-7)


# ------------------------------------------------------------------------------------------------ K7
def test_adam_clip_matches_torch_optimizer(ops):
    torch.manual_seed(0)
    n_clip, n = 1000, 1010
    p0 = torch.randn(n)
    ref = [p0[:n_clip].clone().requires_grad_(True), p0[n_clip:].clone().requires_grad_(True)]
    opt = torch.optim.Adam([{"params": [ref[0]]}, {"params": [ref[1]]}], lr=5e-5)
    p, m, v = cu(p0), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    st = torch.zeros(8, device=DEV)
    st[1] = 5e-5
    skip = torch.zeros(1, device=DEV, dtype=torch.int32)
    for step in range(6):
        gfull = torch.randn(n) * (3.0 if step % 2 else 0.01)
        if step == 3:                                       # a KL-skipped minibatch: no state change at all
            skip.fill_(1)
            before = p.clone()
            ops.adam_step(p, cu(gfull), m, v, n_clip, 0.5, st, skip)
            assert torch.equal(p, before) and float(st[0]) == 3.0
            skip.fill_(0)
            continue
        ref[0].grad, ref[1].grad = gfull[:n_clip].clone(), gfull[n_clip:].clone()
        tn = torch.nn.utils.clip_grad_norm_([ref[0]], 0.5)   # log_std tail excluded (Q8)
        opt.step()
        ops.adam_step(p, cu(gfull), m, v, n_clip, 0.5, st, skip)
        assert math.isclose(float(st[2]), float(tn), rel_tol=1e-5)
        want = torch.cat([ref[0].detach(), ref[1].detach()])
        assert (p.cpu() - want).abs().max() < 2e-7, float((p.cpu() - want).abs().max())
    assert float(st[0]) == 5.0


@pytest.mark.parametrize("n,n_clip", [(1010, 1000), (470197, 470187), (7, 7), (4096, 0)])
def test_fused_step_equals_finalize_plus_adam(ops, n, n_clip):
    """K7f at world == 1: one launch == pm_ppo_actor_finalize + pm_adam_step on the same inputs (weights, moments, step counter, the
    KL-skip flag and the loss / KL accumulators), including a skipped minibatch (ppo.py:337-338) and the critic form (no
    finalize).  The norm's fp64 partial sums are grouped differently, so weights agree to an ulp rather than bit for bit."""
    torch.manual_seed(n)
    p0 = torch.randn(n)
    mk = lambda: (cu(p0), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), torch.tensor([0, 5e-5, 0, 0, 0, 0, 0, 0], device=DEV))
    pa, ma, va, sa = mk()
    pb, mb, vb, sb = mk()
    acc_a, acc_b = torch.zeros(8, device=DEV), torch.zeros(8, device=DEV)
    skip_a, skip_b = torch.zeros(1, device=DEV, dtype=torch.int32), torch.zeros(1, device=DEV, dtype=torch.int32)
    ws = ops.fused_step_workspace(n, 4, DEV)
    B, desired_kl = 64, 0.02
    for step in range(5):
        gext = torch.cat([torch.randn(n) * (3.0 if step % 2 else 0.01), torch.tensor([-12.5 * (step + 1), 0.3 * B if step == 2 else 0.004 * B, 0.0, 0.0])])
        ga, gb = cu(gext), cu(gext)
        finalize = step != 4                                      # the last step in the critic form
        if finalize:
            ops.ppo_actor_finalize(ga[n:], 1.0 / B, desired_kl, acc_a, skip_a)
        ops.adam_step(pa, ga[:n], ma, va, n_clip, 0.5, sa, skip_a if finalize else None)
        ops.fused_step(pb, gb, mb, vb, n_clip, 4, 0.5, sb, ws, finalize=(1.0 / B, desired_kl, acc_b, skip_b) if finalize else None)
        assert int(skip_a) == int(skip_b) == (1 if step == 2 else 0) or not finalize
        assert float(sa[0]) == float(sb[0]) and torch.equal(acc_a, acc_b)
        assert math.isclose(float(sa[2]), float(sb[2]), rel_tol=1e-6, abs_tol=1e-30)
        assert float((pa - pb).abs().max()) <= 1e-7 and float((ma - mb).abs().max()) <= 1e-7 * float(ma.abs().max() + 1e-30)
        assert float((va - vb).abs().max()) <= 1e-6 * float(va.abs().max() + 1e-30)
    assert float(sb[0]) == 4.0 and float(acc_b[2]) == 3.0


# ------------------------------------------------------------------------------------------------ K3
@pytest.mark.parametrize("name,act", [("mlp_actor.npz", "tanh"), ("mlp_critic.npz", "elu")])
def test_mlp_golden_forward_backward(ops, name, act):
    from partmanip_b200.algorithms.algo_utils.network import MLP
    g = load_golden(name)
    w = sub(g, "w")
    D, out = g["x"].shape[1], g["y"].shape[1]
    net = MLP(D, out, dict(hid_dim=[96, 160, 64], activation=act), 0)
    net.load_state_dict(w)
    net.to(DEV)
    x = cu(g["x"])
    y = net(x)
    assert close(y.detach().cpu(), g["y"], 1e-4, 1e-5), max_err(y.detach().cpu(), g["y"])
    y.square().sum().backward()
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        assert close(got, v, 1e-4, 1e-5), (k, max_err(got, v))


@pytest.mark.parametrize("M,N,K,act", [(2048, 512, 53, "tanh"), (64, 7, 512, None), (300, 1, 32, "relu"),
                                       (1000, 130, 129, "selu"), (5, 512, 512, "lrelu"), (128, 128, 16, "sigmoid")])
def test_linear_vs_torch(ops, M, N, K, act):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, requires_grad=True)
    W = (torch.randn(N, K) / math.sqrt(K)).requires_grad_(True)
    b = torch.randn(N, requires_grad=True)
    pre = torch.nn.functional.linear(x, W, b)
    y = O.activation(act)(pre) if act else pre
    dy = torch.randn(M, N)
    y.backward(dy)
    yg = ops.linear_forward(cu(x.detach()), cu(W.detach()), cu(b.detach()), act)
    assert close(yg.cpu(), y.detach(), 1e-4, 1e-5), max_err(yg.cpu(), y.detach())
    # backward takes dL/d(pre); fold act' of THIS layer in on the host side of the test
    dpre = (dy * _act_prime(act, y.detach())) if act else dy
    dW, db, dx = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV), torch.empty(M, K, device=DEV)
    ops.linear_backward(cu(x.detach()), cu(W.detach()), cu(dpre), dW, db, dx, None)
    assert close(dW.cpu(), W.grad, 1e-4, 1e-4), max_err(dW.cpu(), W.grad)
    assert close(db.cpu(), b.grad, 1e-4, 1e-4) and close(dx.cpu(), x.grad, 1e-4, 1e-5)


def _act_prime(act, y):
    return {"tanh": 1 - y * y, "relu": (y > 0).float(), "selu": torch.where(y > 0, torch.full_like(y, 1.0507009873554805),
                                                                               y + 1.0507009873554805 * 1.6732632423543772),
            "lrelu": torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.01)), "sigmoid": y * (1 - y),
            "elu": torch.where(y > 0, torch.ones_like(y), y + 1)}[act]


def test_linear_device_row_limit(ops):
    torch.manual_seed(1)
    M, N, K = 700, 64, 48
    x, W, b = torch.randn(M, K), torch.randn(N, K), torch.randn(N)
    lim = torch.tensor([333], device=DEV, dtype=torch.int32)
    y = torch.full((M, N), 7.0, device=DEV)
    ops.linear_forward(cu(x), cu(W), cu(b), "tanh", out=y, m_dev=lim)
    want = torch.tanh(x @ W.T + b)
    assert close(y[:333].cpu(), want[:333], 1e-4, 1e-5) and bool((y[333:] == 7.0).all())
    dpre = torch.randn(M, N)
    dW, db = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV)
    ops.linear_backward(cu(x), cu(W), cu(dpre), dW, db, None, None, m_dev=lim)
    assert close(dW.cpu(), dpre[:333].T @ x[:333], 1e-4, 1e-4) and close(db.cpu(), dpre[:333].sum(0), 1e-4, 1e-4)


# ------------------------------------------------------------------------------------------------ K8
def test_gather_and_copy_rows(ops):
    torch.manual_seed(2)
    src = torch.randn(1000, 3072, device=DEV)
    idx = torch.randperm(1000, device=DEV)[:300]
    out = ops.gather_rows(src, idx, torch.empty(300, 3072, device=DEV))
    assert torch.equal(out, src[idx])
    narrow = torch.randn(50, 37, device=DEV)
    big = torch.zeros(50, 100, device=DEV)
    ops.copy_rows(narrow, big[:, 10:47])
    assert torch.equal(big[:, 10:47], narrow) and float(big[:, :10].abs().sum()) == 0.0


def test_gather_rows_beyond_2_31_elements(ops):
    """Maximum sizes: a DAgger ring past 2^31 floats (360 k rows of a 2048-point cloud = 8.8 GB; the shipped buf_size 1600 x 2048 envs
    would be 80 GB and fits a B200) — row offsets are 64-bit, rows at the far end gather and append correctly."""
    rows, D = 360_000, 6144
    assert rows * D > 2 ** 31
    ring = torch.empty(rows, D, device=DEV)
    far = torch.tensor([rows - 1, rows - 7, 350_001, 349_525, 5], device=DEV)          # 349_525 * 6144 is just below 2^31, 350_001 above
    marks = torch.arange(5, device=DEV, dtype=torch.float32)[:, None] + torch.linspace(0, 1, D, device=DEV)[None, :]
    ops.copy_rows(marks[:2], ring[rows - 2:rows])                                       # append at the very end (storage.py:84-91)
    assert torch.equal(ring[rows - 2:], marks[:2])
    ring[far] = marks
    out = ops.gather_rows(ring, far, torch.empty(5, D, device=DEV))
    assert torch.equal(out, marks)
    del ring
    torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ KAT-1
def test_kat1_shipped_checkpoint_through_actor_critic(ops):
    from partmanip_b200.algorithms.algo_utils import ActorCritic
    g = load_golden("kat1_ckpt_mlp.npz")
    cfg = dict(action_std=0.5, action_activate="tanh", clipAction=float(g["clipAction"]),
               network=dict(name="MLP", hid_dim=[512, 512, 512], activation="tanh"))
    ac = ActorCritic(53, 10, cfg)
    ac.load_state_dict(sub(g, "w"))
    ac.to(DEV).flatten_()
    xn = cu(g["xn"])
    act, val = ac.act_cri(xn)
    assert close(act.cpu(), g["actions"], 1e-4, 1e-5), max_err(act.cpu(), g["actions"])
    assert close(val.cpu(), g["values"], 1e-4, 1e-4), max_err(val.cpu(), g["values"])
    assert abs(float(act[0, 0]) - 0.6152026) < 1e-4 and abs(float(val[1, 0]) - 427.0086975) < 5e-2   # SURVEY literals
    logp, ent, val2, mu, sig = ac.update_act_cri(xn, cu(g["actions"]))
    assert close(logp.cpu(), g["logp"], 1e-4, 1e-4) and close(ent.cpu(), g["entropy"], 1e-4, 1e-4)
    assert close(mu.cpu(), g["mu"], 1e-4, 1e-4) and torch.equal(sig.cpu(), g["sigma"])
    # state_dict keys are the reference's
    assert set(ac.state_dict().keys()) == set(sub(g, "w").keys())


@pytest.mark.parametrize("B,F,out,act,cols", [(2048, 512, 10, "tanh", 512), (37, 537, 1, "elu", 512), (5, 1024, 7, "relu", 1024),
                                              (300, 512, 32, "tanh", 512), (1, 512, 10, "sigmoid", 512)])
def test_fused_head_forward_backward_vs_torch(ops, B, F, out, act, cols):
    """K3b: fused Linear(F,128)-act-Linear(128,32)-act-Linear(32,out) against torch autograd (fp32 gate 1e-4)."""
    torch.manual_seed(B + F + out)
    f = {"tanh": torch.tanh, "elu": torch.nn.functional.elu, "relu": torch.relu, "sigmoid": torch.sigmoid}[act]
    feat = torch.randn(B, F, requires_grad=True)
    Ws = [(torch.randn(128, F) / F ** 0.5), torch.randn(128) * 0.1, torch.randn(32, 128) / 128 ** 0.5, torch.randn(32) * 0.1,
          torch.randn(out, 32) / 32 ** 0.5, torch.randn(out) * 0.1]
    Ws = [w.requires_grad_(True) for w in Ws]
    h1 = f(feat @ Ws[0].T + Ws[1]); h2 = f(h1 @ Ws[2].T + Ws[3]); y = h2 @ Ws[4].T + Ws[5]
    dout = torch.randn(B, out)
    y.backward(dout)
    dW = [cu(w.detach()) for w in Ws]
    g = [torch.full_like(w, float("nan")) for w in dW]
    h1d, h2d, yd = torch.empty(B, 128, device=DEV), torch.empty(B, 32, device=DEV), torch.empty(B, out, device=DEV)
    featd = cu(feat.detach())
    ops.pointnet_head_forward(featd, dW, out, act, h1d, h2d, yd)
    assert close(yd.cpu(), y.detach(), 1e-4, 1e-5), max_err(yd.cpu(), y.detach())
    assert close(h1d.cpu(), h1.detach(), 1e-4, 1e-5) and close(h2d.cpu(), h2.detach(), 1e-4, 1e-5)
    dfeat = torch.full((B, cols), float("nan"), device=DEV)
    ops.pointnet_head_backward(featd, dW, out, act, h1d, h2d, cu(dout), g, dfeat, cols)
    for got, w in zip(g, Ws):
        tol = 1e-4 * float(w.grad.abs().max()) + 1e-6
        assert float((got.cpu() - w.grad).abs().max()) <= tol, (tuple(w.shape), float((got.cpu() - w.grad).abs().max()), tol)
    want = feat.grad[:, :cols]
    assert float((dfeat.cpu() - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-6
