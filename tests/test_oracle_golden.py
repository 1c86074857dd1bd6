"""Pins the CPU oracle (oracle/ppo_oracle.py) against fixtures recorded from the UNMODIFIED reference
(tests/golden/make_golden.py) and against the KATs of SURVEY.md §8(c).  CPU-only."""
import math

import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import ITER_CASES, close, load_golden, max_err, ppo_cfg, sub


def test_kat1_shipped_checkpoint_mlp():
    g = load_golden("kat1_ckpt_mlp.npz")
    w = sub(g, "w")
    xn = (g["x"] - g["rms_mean"]) / g["rms_std"]
    assert torch.equal(xn, g["xn"])
    mu = O.mlp_forward(sub(w, "actor"), xn)
    val = O.mlp_forward(sub(w, "critic"), xn)
    act = O.action_activation(mu, float(g["clipAction"]))
    assert close(act, g["actions"], 1e-5, 1e-6) and close(val, g["values"], 1e-5, 1e-5)
    # SURVEY KAT-1 literal values
    assert abs(float(act[0, 0]) - 0.6152026) < 1e-5 and abs(float(val[1, 0]) - 427.0086975) < 1e-2
    raw = O.action_deactivation(g["actions"], float(g["clipAction"]))
    logp = O.gaussian_logp(mu, w["log_std"], raw)
    assert close(logp, g["logp"], 1e-5, 1e-5)
    assert abs(float(logp[0]) + 5.7843256) < 1e-4
    assert close(O.gaussian_entropy(w["log_std"], 4), g["entropy"], 1e-5, 1e-5)
    assert torch.equal(w["log_std"].repeat(4, 1), g["sigma"])


def test_kat2_gae_bit_exact():
    g = load_golden("gae.npz")
    u = lambda t: t[..., None]
    r, a = O.gae(u(g["k2_rew"]), u(g["k2_val"]), u(g["k2_done"]), u(g["k2_succ"]), u(g["k2_last"]), 0.99, 0.95, None)
    assert torch.equal(r.squeeze(-1), g["k2_ret"]) and torch.equal(a.squeeze(-1), g["k2_adv"])
    assert abs(float(r[0, 0]) - 4.4708571) < 1e-6 and float(a[1, 1]) == 0.0       # SURVEY KAT-2 literals
    r, a = O.gae(u(g["k2_rew"]), u(g["k2_val"]), u(g["k2_done"]), u(g["k2b_succ"]), u(g["k2_last"]), 0.99, 0.95, 500, True)
    assert torch.equal(r.squeeze(-1), g["k2b_ret"]) and torch.equal(a.squeeze(-1), g["k2b_adv"])
    assert float(r[1, 1]) == 500.0 and abs(float(a[1, 1]) - 3.1752660) < 1e-5     # KAT-2b
    r, a = O.gae(u(g["r_rew"]), u(g["r_val"]), u(g["r_done"]), u(g["r_succ"]), u(g["r_last"]), 0.99, 0.95, 500)
    assert torch.equal(r.squeeze(-1), g["r3_ret"]) and torch.equal(a.squeeze(-1), g["r3_adv"])
    r, a = O.gae(u(g["r_rew"]), u(g["r_val"]), u(g["r_done"]), u(g["r_succ"]), u(g["r_last"]), 0.97, 0.9, None, True)
    assert torch.equal(r.squeeze(-1), g["r4_ret"]) and torch.equal(a.squeeze(-1), g["r4_adv"])


def test_kat3_sampler_geometry():
    g = load_golden("sampler_geometry.npz")["geo"]
    for E, T, nmb, count, size, first, last in g.tolist():
        s, c = O.minibatch_geometry(E * T, nmb)
        assert (s, c) == (size, count)
        idx = O.minibatch_indices(E * T, nmb)
        assert int(idx[0][0]) == first and int(idx[-1][-1]) == last
    assert O.minibatch_geometry(4096 * 8, 8) == (2048, 16)                          # KAT-3 literal


def test_rms_sequence():
    g = load_golden("rms.npz")
    rs = O.RunningStats(11)
    for k in range(4):
        y = rs.normalize(g["x"][k], update=(k != 3))
        assert torch.equal(y, g["y"][k]) and torch.equal(rs.mean, g["mean"][k])
        assert torch.equal(rs.S, g["S"][k]) and torch.equal(rs.std, g["std"][k])


PN_CASES = {
    "pointnet_base_a10.npz": dict(point_num=1024, proprio=0, max_mean=False, sub_mean=False, act="tanh"),
    "pointnet_maxmean_c4.npz": dict(point_num=1024, proprio=0, max_mean=True, sub_mean=False, act="tanh"),
    "pointnet_submean_proprio.npz": dict(point_num=1024, proprio=25, max_mean=False, sub_mean=True, act="tanh"),
    "pointnet_relu_submean.npz": dict(point_num=1024, proprio=0, max_mean=False, sub_mean=True, act="relu"),
    "pointnet_n2048_c3.npz": dict(point_num=2048, proprio=0, max_mean=False, sub_mean=False, act="tanh"),
}


@pytest.mark.parametrize("name", list(PN_CASES))
def test_kat4_pointnet_forward_backward(name):
    g = load_golden(name)
    kw = PN_CASES[name]
    w = {k: v.clone().requires_grad_(True) for k, v in sub(g, "w").items()}
    assert sum(v.numel() for v in w.values()) == int(g["n_params"])
    x = g["x"].clone()
    y = O.pointnet_forward(w, x, **kw)
    assert close(y, g["y"], 1e-5, 1e-6), max_err(y, g["y"])
    assert torch.equal(x, g["x_after"])                      # Q3: in-place centring reaches the caller's tensor
    if kw["sub_mean"]:
        assert not torch.equal(x, g["x"])
    y.square().sum().backward()
    for k, v in sub(g, "g").items():
        assert close(w[k].grad, v, 1e-4, 1e-6), (k, max_err(w[k].grad, v))
    if name == "pointnet_base_a10.npz":
        assert int(g["n_params"]) == 235242                  # KAT-4 literal


@pytest.mark.parametrize("name,act", [("mlp_actor.npz", "tanh"), ("mlp_critic.npz", "elu")])
def test_mlp_forward_backward(name, act):
    g = load_golden(name)
    w = {k: v.clone().requires_grad_(True) for k, v in sub(g, "w").items()}
    y = O.mlp_forward(w, g["x"], act)
    assert close(y, g["y"], 1e-5, 1e-6)
    y.square().sum().backward()
    for k, v in sub(g, "g").items():
        assert close(w[k].grad, v, 1e-4, 1e-6), k


def test_mlp_init_is_orthogonal_with_reference_gains():
    p = O.mlp_init(37, 7, [64, 48], torch.Generator().manual_seed(0))
    w0 = p["model.0.weight"]                                   # (64,37): columns orthogonal, gain sqrt2
    assert torch.allclose(w0.T @ w0, 2.0 * torch.eye(37), atol=1e-5)
    w2 = p["model.4.weight"]                                   # actor head gain 0.01
    assert torch.allclose(w2 @ w2.T, 1e-4 * torch.eye(7), atol=1e-8)
    pc = O.mlp_init(37, 1, [64, 48], torch.Generator().manual_seed(0))
    assert abs(float(pc["model.4.weight"].norm()) - 1.0) < 1e-5


def test_adam_and_clip_match_torch():
    torch.manual_seed(0)
    ps = [torch.randn(13, 7), torch.randn(7)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=5e-5)
    mine = {str(i): p.clone() for i, p in enumerate(ps)}
    st = O.AdamState(mine, 5e-5)
    for step in range(5):
        gs = [torch.randn_like(p) * (3.0 if step % 2 else 0.01) for p in ps]
        for r, g in zip(ref, gs):
            r.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        total, coef = O.clip_coef(gs, 0.5)
        assert torch.allclose(total, tn)
        opt.step()
        st.apply({str(i): g * coef for i, g in enumerate(gs)})
        for i, r in enumerate(ref):
            assert torch.allclose(mine[str(i)], r.detach(), rtol=1e-6, atol=1e-9)


def test_gaussian_matches_multivariate_normal():
    from torch.distributions import MultivariateNormal
    torch.manual_seed(1)
    mu, ls, a = torch.randn(6, 10), torch.randn(10) * 0.3 - 0.5, torch.randn(6, 10)
    d = MultivariateNormal(mu, scale_tril=torch.diag(ls.exp() * ls.exp()))
    assert torch.allclose(O.gaussian_logp(mu, ls, a), d.log_prob(a), atol=1e-5)
    assert torch.allclose(O.gaussian_entropy(ls, 6), d.entropy(), atol=1e-5)


def _replay_rollout(g, E, D, A, net, cfg):
    """Re-run the reference rollout (ppo.py:210-248) with the oracle from recorded env tensors + eps."""
    kind = net["name"]
    w = sub(g, "init")
    actor, critic, log_std = sub(w, "actor"), sub(w, "critic"), w["log_std"]
    rs = O.RunningStats(D) if cfg["tricks"]["use_state_norm"] else None
    norm = (lambda o: rs.normalize(o, True)) if rs else (lambda o: o)
    T = cfg["n_steps"]
    cur = norm(g["env.obs"][0].clone())
    out = {k: [] for k in ("obs", "act", "logp", "val", "mu")}
    with torch.no_grad():
        for t in range(T):
            mu = O.net_forward(kind, actor, cur, net)
            act, logp = O.policy_sample(mu, log_std, g["eps"][t], 1.0)
            val = O.net_forward(kind, critic, cur, net)
            for k, v in zip(out, (cur, act, logp, val, mu)):
                out[k].append(v)
            cur = norm(g["env.obs"][t + 1].clone())
        last = O.net_forward(kind, critic, cur, net)
    return {k: torch.stack(v) for k, v in out.items()}, last, rs


@pytest.mark.parametrize("name", list(ITER_CASES))
def test_full_iteration_replay(name):
    g = load_golden(name)
    E, D, A, net, over = ITER_CASES[name]
    cfg = ppo_cfg(E, net, **over)
    ro, last, rs = _replay_rollout(g, E, D, A, net, cfg)
    assert close(ro["obs"], g["buf.observations"], 1e-5, 1e-5)
    assert close(ro["mu"], g["buf.mu"], 1e-4, 1e-5)
    assert close(ro["act"], g["buf.actions"], 1e-4, 1e-5)
    assert close(ro["logp"], g["buf.actions_log_prob"].squeeze(-1), 1e-4, 1e-4)
    assert close(ro["val"], g["buf.values"], 1e-4, 1e-5)
    if rs is not None:
        assert close(rs.mean, g["rms.mean"], 1e-5, 1e-6) and close(rs.std, g["rms.std"], 1e-5, 1e-6)
        assert rs.n == int(g["rms.n"])
    # GAE on the recorded buffer
    ret, adv = O.gae(g["buf.rewards"], g["buf.values"], g["buf.dones"], g["buf.succs"], last, cfg["gamma"], cfg["lam"],
                     cfg["succ_value"], cfg["tricks"]["whole_adv_norm"])
    assert close(ret, g["buf.returns"], 1e-4, 1e-4) and close(adv, g["buf.advantages"], 1e-3, 1e-4)
    # update from the recorded buffer
    w = sub(g, "init")
    actor = {k: v.clone() for k, v in sub(w, "actor").items()}
    critic = {k: v.clone() for k, v in sub(w, "critic").items()}
    log_std = w["log_std"].clone()
    opt_a = O.AdamState({**actor, "log_std": log_std}, cfg["lr"])
    opt_c = O.AdamState(critic, cfg["lr"])
    flat = lambda t: t.reshape(-1, t.shape[-1])
    buf = dict(obs=flat(g["buf.observations"]), actions=flat(g["buf.actions"]), values=flat(g["buf.values"]),
               returns=flat(g["buf.returns"]), logp=flat(g["buf.actions_log_prob"]), adv=flat(g["buf.advantages"]),
               mu=flat(g["buf.mu"]), sigma=flat(g["buf.sigma"]))
    stats = O.ppo_update(actor, critic, log_std, opt_a, opt_c, buf, cfg, net["name"], net)
    assert stats["count"] == int(g["log.Train/kl_update_count"])
    assert opt_a.step == int(g["adam_actor.0.step"]) and opt_c.step == int(g["adam_critic.0.step"])
    assert close(stats["surrogate_loss"], g["log.Train/surrogate_loss"], 1e-3, 1e-5)
    assert close(stats["value_loss"], g["log.Train/value_function_loss"], 1e-3, 1e-5)
    assert close(stats["kl"], g["log.Train/kl"], 1e-3, 1e-6) and close(stats["kl_max"], g["log.Train/kl_max"], 1e-3, 1e-6)
    fin = sub(g, "final")
    lr = cfg["lr"]
    # Adam's first steps move each weight by ~lr regardless of gradient scale, so post-update weights are
    # compared with an absolute tolerance that is a small fraction of the total displacement (<= 40*lr).
    for k, v in sub(fin, "actor").items():
        assert (actor[k] - v).abs().max() <= 0.02 * 40 * lr, k
    for k, v in sub(fin, "critic").items():
        assert (critic[k] - v).abs().max() <= 0.02 * 40 * lr, k
    assert (log_std - fin["log_std"]).abs().max() <= 0.02 * 40 * lr
    assert math.isclose(float(log_std.exp().mean()), float(g["log.Train/mean_action_noise_std"]), rel_tol=1e-5)
