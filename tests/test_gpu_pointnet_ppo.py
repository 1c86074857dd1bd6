"""GPU parity tests for the PointNet encoder kernels (K1/K2) and for whole PPO iterations driven through the
reference-shaped `ppo` class, against fixtures recorded from the unmodified reference (tests/golden) and the
CPU oracle.  All kernels are reached through the C-ABI."""
import math

import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import ITER_CASES, close, load_golden, max_err, ppo_cfg, sub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(t):
    return t.to(DEV).contiguous()


PN_CASES = {
    "pointnet_base_a10.npz": dict(D=3072, out=10, proprio=0, cfg=dict(activation="tanh", max_mean=False, sub_mean=False)),
    "pointnet_submean_proprio.npz": dict(D=3097, out=7, proprio=25, cfg=dict(activation="tanh", max_mean=False, sub_mean=True)),
    "pointnet_relu_submean.npz": dict(D=3072, out=10, proprio=0, cfg=dict(activation="relu", max_mean=False, sub_mean=True)),
    "pointnet_n2048_c3.npz": dict(D=6144, out=10, proprio=0, cfg=dict(activation="tanh", max_mean=False, sub_mean=False, point_num=2048)),
    "pointnet_maxmean_c4.npz": dict(D=4096, out=10, proprio=0, cfg=dict(activation="tanh", max_mean=True, sub_mean=False)),
}


def _build(name, precision="fp32"):
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    g = load_golden(name)
    c = PN_CASES[name]
    net = PointNet(c["D"], c["out"], dict(name="PointNet", precision=precision, **c["cfg"]), c["proprio"])
    net.load_state_dict(sub(g, "w"))
    return g, net.to(DEV)


# "fp32" = the reference's precision on the tensor cores where the shape allows (split-fp16 tcgen05 forward, three-term bf16
# tcgen05 GEMMs in the backward), "fp32_ffma" = CUDA cores only: both must hold the 1e-4 gate on every recording
@pytest.mark.parametrize("precision", ["fp32", "fp32_ffma"])
@pytest.mark.parametrize("name", list(PN_CASES))
def test_pointnet_golden_forward(name, precision):
    from partmanip_b200 import ops
    g, net = _build(name, precision)
    assert sum(p.numel() for p in net.parameters()) == int(g["n_params"])
    x = cu(g["x"])
    with torch.no_grad():
        y = net(x)
    assert ops.pointnet_tc3_last_error(DEV) == 0
    assert close(y.cpu(), g["y"], 1e-4, 1e-5), max_err(y.cpu(), g["y"])
    # Q3: sub_mean centres the caller's tensor in place
    assert close(x.cpu(), g["x_after"], 1e-5, 1e-6)
    if PN_CASES[name]["cfg"]["sub_mean"]:
        assert not torch.equal(x.cpu(), g["x"])


@pytest.mark.parametrize("precision", ["fp32", "fp32_ffma"])
@pytest.mark.parametrize("name", list(PN_CASES))
def test_pointnet_golden_backward(name, precision):
    g, net = _build(name, precision)
    y = net(cu(g["x"]))
    y.square().sum().backward()
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        # gradient gate: 1e-4 relative to the tensor's scale (sums over 10^3..10^4 fp32 terms reorder freely)
        tol = 1e-4 * float(v.abs().max()) + 1e-6
        assert float((got - v).abs().max()) <= tol, (k, float((got - v).abs().max()), tol)


def test_pointnet_argmax_first_index_and_duplicates():
    """Ties between bit-identical points (the env's (0,0,0) padding) resolve to the first index, like torch.max."""
    from partmanip_b200 import ops
    torch.manual_seed(5)
    B, N, C = 3, 1024, 3
    x = torch.rand(B, N * C) * 2 - 1
    x.view(B, N, C)[:, 100:400] = 0.0                      # 300 duplicate points per cloud
    p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(1))
    h = O.pointnet_encode(p, x.view(B, N, C))
    want_v, want_i = h.max(dim=1)
    enc = [cu(p[k]) for k in ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")]
    feat = torch.empty(B, 512, device=DEV)
    am = torch.empty(B, 512, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(cu(x), N, C, enc, "tanh", "fp32", feat, None, am, None)
    assert close(feat.cpu(), want_v, 1e-4, 1e-5)
    am = am.cpu().long()
    # the chosen point attains the max (within fp32 reorder noise) and, among the duplicates, is the first one
    picked = h.gather(1, am[:, None, :]).squeeze(1)
    assert float((picked - want_v).abs().max()) < 1e-5
    dup = (am >= 100) & (am < 400)
    assert bool((am[dup] == 100).all())
    assert float((am == want_i).float().mean()) > 0.99


def test_encoder_backward_with_reference_argmax_is_tight():
    """Feeding the ORACLE's argmax removes the near-tie freedom: every encoder gradient then matches autograd to
    fp32 reorder noise, including the per-row mlp.4.weight entries."""
    from partmanip_b200 import ops
    torch.manual_seed(9)
    B, N, C = 8, 1024, 3
    x = torch.rand(B, N * C) * 2 - 1
    x.view(B, N, C)[:, ::10] = 0.0
    p = {k: v.requires_grad_(True) for k, v in O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(2)).items()}
    h = O.pointnet_encode(p, x.view(B, N, C))
    feat, am = h.max(dim=1)
    dfeat = torch.randn(B, 512)
    feat.backward(dfeat)
    names = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")
    enc = [cu(p[k].detach()) for k in names]
    grads = [torch.empty_like(t) for t in enc]
    ops.pointnet_encode_backward(cu(x), N, C, enc, "tanh", cu(dfeat), cu(am.int()), grads)
    for k, gt in zip(names, grads):
        v = p[k].grad
        assert float((gt.cpu() - v).abs().max()) <= 2e-5 * float(v.abs().max()) + 1e-7, (k, max_err(gt.cpu(), v))


@pytest.mark.parametrize("B", [1, 37])
def test_pointnet_vs_oracle_random_batch(B):
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    torch.manual_seed(B)
    net = PointNet(3072, 10, dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False), 0)
    w = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.to(DEV)
    x = torch.rand(B, 3072) * 2 - 1
    wl = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    y_ref = O.pointnet_forward(wl, x.clone())
    (y_ref * torch.arange(1, 11)).sum().backward()
    y = net(cu(x))
    (y * torch.arange(1, 11, device=DEV)).sum().backward()
    assert close(y.detach().cpu(), y_ref.detach(), 1e-4, 1e-5)
    for k, p in net.named_parameters():
        v = wl[k].grad
        tol = 1e-4 * float(v.abs().max()) + 1e-6
        assert float((p.grad.cpu() - v).abs().max()) <= tol, k


# ------------------------------------------------------------------------------------------------ full iterations
class _ReplayEnv:
    """Serves the env tensors recorded in a ppo_iter_* fixture (tests/golden/make_golden.py:FakeEnv)."""

    def __init__(self, g, E, D, A):
        self.g, self.t = g, 0
        self.num_envs, self.num_actions, self.max_episode_length = E, A, 200
        self.num_obs = {"obs": D, "proprio_state": 0}
        self.train_test_flag = "train"
        self.reset_succ = torch.zeros(E, dtype=torch.bool, device=DEV)
        self.rew_buf = torch.zeros(E, device=DEV)
        self.actions = []

    def reset(self):
        return {"obs": cu(self.g["env.obs"][0])}

    def step(self, actions, save_image_path=None):
        t = self.t
        self.t += 1
        self.actions.append(actions.clone())
        self.rew_buf = cu(self.g["env.rew"][t])
        self.reset_succ = cu(self.g["env.succ"][t])
        return {"obs": cu(self.g["env.obs"][t + 1])}, self.rew_buf, cu(self.g["env.done"][t]), {"succ_rate": torch.zeros(1, device=DEV)}


class _Logger:
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/pm_b200_test"

    def info(self, d, it):
        self.last = d


def _runner(name, precision="fp32"):
    from partmanip_b200.algorithms import ppo
    g = load_golden(name)
    E, D, A, net, over = ITER_CASES[name]
    if precision != "fp32":
        net = dict(net, precision=precision)
    cfg = ppo_cfg(E, net, device=DEV, **over)
    env = _ReplayEnv(g, E, D, A)
    r = ppo(env, cfg, _Logger())
    r.actor_critic.load_state_dict(sub(g, "init"))
    return g, cfg, env, r


@pytest.mark.parametrize("name", list(ITER_CASES))
def test_full_iteration_against_reference_recording(name):
    g, cfg, env, r = _runner(name)
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    last_obs, last_values = r.collect(curr, None, eps=cu(g["eps"]))
    st = r.storage
    # ---- rollout outputs (actions, values, log-probs, mu) vs the reference's buffer
    assert close(st.observations.cpu(), g["buf.observations"], 1e-4, 1e-4), max_err(st.observations.cpu(), g["buf.observations"])
    assert close(st.mu.cpu(), g["buf.mu"], 1e-4, 1e-4), max_err(st.mu.cpu(), g["buf.mu"])
    assert close(st.actions.cpu(), g["buf.actions"], 1e-4, 1e-4)
    assert close(st.actions_log_prob.cpu(), g["buf.actions_log_prob"], 1e-4, 1e-4)
    assert close(st.values.cpu(), g["buf.values"], 1e-4, 1e-4)
    assert torch.equal(st.sigma.cpu(), g["buf.sigma"]) and torch.equal(st.dones.cpu(), g["buf.dones"])
    assert torch.equal(st.rewards.cpu(), g["buf.rewards"]) and torch.equal(st.succs.cpu(), g["buf.succs"])
    if cfg["tricks"]["use_state_norm"]:
        ms = r.state_norm.running_ms
        assert ms.n == int(g["rms.n"]) and close(ms.mean.cpu(), g["rms.mean"], 1e-5, 1e-6) and close(ms.std.cpu(), g["rms.std"], 1e-5, 1e-6)
    # ---- GAE
    st.compute_returns(last_values, cfg["gamma"], cfg["lam"])
    assert close(st.returns.cpu(), g["buf.returns"], 1e-4, 1e-4)
    assert close(st.advantages.cpu(), g["buf.advantages"], 1e-3, 2e-4), max_err(st.advantages.cpu(), g["buf.advantages"])
    # ---- update: start from the reference's exact buffer so the comparison isolates the update kernels
    for k in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        getattr(st, k).copy_(cu(g["buf." + k]))
    r.update(1)
    log = r.log_dict
    assert log["Train/kl_update_count"] == int(g["log.Train/kl_update_count"])
    assert r.optimizer_actor.step_count == int(g["adam_actor.0.step"])
    assert r.optimizer_critic.step_count == int(g["adam_critic.0.step"])
    assert close(log["Train/surrogate_loss"], g["log.Train/surrogate_loss"], 1e-3, 1e-5)
    assert close(log["Train/value_function_loss"], g["log.Train/value_function_loss"], 1e-3, 1e-5)
    assert close(log["Train/kl"], g["log.Train/kl"], 1e-3, 1e-6) and close(log["Train/kl_max"], g["log.Train/kl_max"], 1e-3, 1e-6)
    fin = sub(g, "final")
    lr = cfg["lr"]
    sd = {k: v.cpu() for k, v in r.actor_critic.state_dict().items()}
    # Post-update weights.  Adam's early steps move every weight by ~lr per step whatever the gradient scale, so the
    # yardstick is the total displacement 40*lr.  The 40-step trajectory is CHAOTIC in the max-pool: a weight difference
    # of 1e-3*lr (fp32 summation order: MKL vs these kernels) flips the argmax between two near-tied points of some
    # cloud after a few steps, that row's gradient then differs by O(1e-2) and Adam amplifies it (measured in lockstep,
    # a lockstep debug run in round 1: two of OUR OWN kernel variants whose single-step gradients agree to 1e-7 drift apart by
    # 10*lr on the worst critic element after 40 steps, first flip at step 6).  Single-step gradient parity is gated
    # tightly elsewhere (test_pointnet_golden_backward, test_fused_head_*, test_encoder_backward_*: 1e-4); here the
    # gate is statistical: every element within half the displacement, >= 90 % of each tensor within 2 % of it, RMS
    # within 2 %.
    disp = 40 * lr
    for k, v in fin.items():
        d = (sd[k] - v).abs()
        assert float(d.max()) <= 0.5 * disp, (k, float(d.max()))
        assert float((d > 0.02 * disp).float().mean()) <= 0.10, (k, float((d > 0.02 * disp).float().mean()))
        assert float(d.pow(2).mean().sqrt()) <= 0.02 * disp, (k, float(d.pow(2).mean().sqrt()))
    assert math.isclose(float(sd["log_std"].exp().mean()), float(g["log.Train/mean_action_noise_std"]), rel_tol=1e-5)


@pytest.mark.parametrize("name", [n for n, c in ITER_CASES.items() if c[3]["name"] == "PointNet"])
def test_full_iteration_bf16_against_reference_recording(name):
    """The BENCHMARKED mode (tcgen05, bf16 operands) through a whole recorded `ppo.run()` iteration of the unmodified reference
    (algorithms/ppo.py:205-411): rollout outputs, GAE outputs, losses, KL and the skip count at north_star's bf16 gate
    |a-b| <= 1e-2 + 1e-2*|b|."""
    from partmanip_b200 import ops
    g, cfg, env, r = _runner(name, "bf16")
    gate = lambda a, b: close(a, b, 1e-2, 1e-2)
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    last_obs, last_values = r.collect(curr, None, eps=cu(g["eps"]))
    ops.check_tc_errors()
    st = r.storage
    assert close(st.observations.cpu(), g["buf.observations"], 1e-4, 1e-4)            # the normaliser stays fp32
    for k in ("mu", "actions", "actions_log_prob", "values"):
        got, want = getattr(st, k).cpu(), g["buf." + k]
        assert gate(got, want), (k, max_err(got, want))
    assert torch.equal(st.sigma.cpu(), g["buf.sigma"]) and torch.equal(st.dones.cpu(), g["buf.dones"])
    st.compute_returns(last_values, cfg["gamma"], cfg["lam"])
    assert gate(st.returns.cpu(), g["buf.returns"]), max_err(st.returns.cpu(), g["buf.returns"])
    assert gate(st.advantages.cpu(), g["buf.advantages"]), max_err(st.advantages.cpu(), g["buf.advantages"])
    # ---- update from the reference's exact buffer (isolates the update kernels, as in the fp32 test)
    for k in ("observations", "actions", "values", "returns", "advantages", "actions_log_prob", "mu", "sigma"):
        getattr(st, k).copy_(cu(g["buf." + k]))
    r.update(1)
    log = r.log_dict
    assert log["Train/kl_update_count"] == int(g["log.Train/kl_update_count"])
    assert r.optimizer_actor.step_count == int(g["adam_actor.0.step"])
    assert r.optimizer_critic.step_count == int(g["adam_critic.0.step"])
    for k in ("Train/surrogate_loss", "Train/value_function_loss", "Train/kl", "Train/kl_max"):
        assert gate(log[k], g["log." + k]), (k, float(log[k]), float(g["log." + k]))
    # post-update weights: every element moved by at most the 40-step Adam displacement, and the bulk follows the reference's
    # trajectory (bf16 gradients agree with fp32 ones to ~3e-3 relative L2; Adam's m/sqrt(v) keeps the per-step error at that
    # fraction of lr for all but the near-zero-gradient elements)
    disp = 40 * cfg["lr"]
    sd = {k: v.cpu() for k, v in r.actor_critic.state_dict().items()}
    worst = 0.0
    for k, v in sub(g, "final").items():
        d = (sd[k] - v).abs()
        assert float(d.max()) <= disp, (k, float(d.max()))
        worst = max(worst, float(d.pow(2).mean().sqrt()) / disp)
        assert float(d.pow(2).mean().sqrt()) <= 0.08 * disp, (k, float(d.pow(2).mean().sqrt()) / disp)   # measured 0.040-0.046
    print(f"bf16 iteration {name}: worst per-tensor RMS weight deviation = {worst:.4f} of the 40-step displacement")


def test_update_graph_is_dropped_when_a_workspace_moves():
    """ops.scratch() re-allocates a workspace when a larger one is requested; a captured update graph would then
    replay into freed memory.  The runner notices (scratch generation) and falls back to eager + re-capture; results stay
    bit-identical to a run that never used a graph."""
    from partmanip_b200 import ops
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    from tests.helpers import PN, ppo_cfg
    outs = []
    for disturb in (False, True):
        torch.manual_seed(4)
        env = FakeVecEnv(16, 3072, 10, DEV, cloud=True, seed=8)
        r = ppo(env, ppo_cfg(16, PN, device=DEV, cuda_graph=True), _Logger())
        curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
        gen = torch.Generator().manual_seed(0)
        for it in range(4):
            eps = torch.randn(8, 16, 10, generator=gen).to(DEV)
            last_obs, last_values = r.collect(curr, None, eps=eps)
            r.storage.compute_returns(last_values, r.gamma, r.lam)
            if disturb and it == 2:
                assert r._graph is not None
                key = next(k for k in ops._scratch if k[0] == DEV and k[1].endswith("loss"))       # ("cuda:0", "actor/loss") when the phases overlap
                with ops.scratch_ns(key[1][:-len("loss")]):
                    ops.scratch(ops._scratch[key].numel() + 4096, DEV, "loss")   # somebody else grows a workspace the graph uses
            r.update(it + 1)
            if disturb and it == 2:
                assert r._graph is None                                        # dropped, this update ran eagerly
            r.storage.clear()
            curr = r._ingest(last_obs.clone(), r.storage.obs_slot())
        outs.append({k: v.clone() for k, v in r.actor_critic.state_dict().items()})
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


def test_checkpoint_roundtrip_and_reference_format(tmp_path):
    g, cfg, env, r = _runner("ppo_iter_mlp_e64.npz")
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    _, last_values = r.collect(curr, None, eps=cu(g["eps"]))
    r.storage.compute_returns(last_values, cfg["gamma"], cfg["lam"])
    r.update(1)
    r.save_ckpt_dir = str(tmp_path)
    r.save(7)
    ck = torch.load(tmp_path / "model_7.pth", map_location="cpu", weights_only=False)
    assert set(ck) >= {"iteration", "model_state_dict", "optimizer_actor", "optimizer_critic", "total_steps", "tricks",
                       "obs_mode", "model_cfg", "state_running_ms"}                                   # ppo.py:86-98
    # torch.optim.Adam accepts the optimiser dicts (format compatibility with reference checkpoints)
    ref_params = [torch.nn.Parameter(v.clone()) for k, v in ck["model_state_dict"].items() if k.startswith("actor.")]
    ref_ls = torch.nn.Parameter(ck["model_state_dict"]["log_std"].clone())
    opt = torch.optim.Adam([{"params": ref_params}, {"params": [ref_ls]}], lr=1.0)
    opt.load_state_dict(ck["optimizer_actor"])
    assert float(opt.state_dict()["state"][0]["step"]) == r.optimizer_actor.step_count
    from partmanip_b200.algorithms import ppo
    cfg2 = dict(cfg, resume=str(tmp_path / "model_7.pth"))
    r2 = ppo(_ReplayEnv(g, 64, 37, 7), cfg2, _Logger())
    for k, v in r.actor_critic.state_dict().items():
        assert torch.equal(v, r2.actor_critic.state_dict()[k])
    assert torch.equal(r.optimizer_actor.exp_avg, r2.optimizer_actor.exp_avg)
    assert r2.optimizer_critic.step_count == r.optimizer_critic.step_count and r2.curr_iter == 7
    assert torch.equal(r.state_norm.running_ms.mean, r2.state_norm.running_ms.mean)
    assert r2.actor_critic.rng_state() == r.actor_critic.rng_state() and "b200_rng" in ck


def test_storage_overflow_and_sampler_geometry():
    from partmanip_b200.algorithms.algo_utils import RolloutStorage
    st = RolloutStorage(4, 2, 5, 3, DEV, None)
    z = lambda *s: torch.zeros(*s, device=DEV)
    for _ in range(2):
        st.add_transitions(z(4, 5), z(4, 3), z(4), z(4).bool(), z(4).bool(), z(4, 1), z(4), z(4, 3), z(4, 3))
    with pytest.raises(AssertionError, match="Rollout buffer overflow"):                           # storage.py:44-45
        st.add_transitions(z(4, 5), z(4, 3), z(4), z(4).bool(), z(4).bool(), z(4, 1), z(4), z(4, 3), z(4, 3))
    geo = load_golden("sampler_geometry.npz")["geo"]
    for E, T, nmb, count, size, first, last in geo.tolist():
        s2 = RolloutStorage(E, T, 1, 1, DEV, None)
        b = s2.mini_batch_generator(nmb)
        assert len(b) == count and len(b[0]) == size and b[0][0] == first and b[-1][-1] == last
    s3 = RolloutStorage(64, 8, 1, 1, DEV, None, sampler="random")
    idx = torch.cat(list(s3.mini_batch_generator(8)))
    assert idx.numel() == 512 and torch.equal(idx.sort()[0].cpu(), torch.arange(512))


def test_cuda_graph_update_is_bit_identical_to_eager():
    """Three iterations with the update replayed as one CUDA graph (iteration 1 eager, 2 captured, 3 replayed) leave exactly
    the weights three eager iterations leave: the kernels are deterministic and the update has no host round trip."""
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    from tests.helpers import MLP128, PN, ppo_cfg
    for net, D in ((PN, 3072), (MLP128, 53)):
        out = []
        for graph in (False, True):
            torch.manual_seed(4)
            env = FakeVecEnv(16, D, 10, DEV, cloud=net is PN, seed=8)
            r = ppo(env, ppo_cfg(16, net, device=DEV, cuda_graph=graph, desired_kl=0.02 if net is MLP128 else 0.1, lr=1e-3 if net is MLP128 else 5e-5), _Logger())
            curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
            g = torch.Generator().manual_seed(0)
            for it in range(3):
                eps = torch.randn(8, 16, 10, generator=g).to(DEV)
                last_obs, last_values = r.collect(curr, None, eps=eps)
                r.storage.compute_returns(last_values, r.gamma, r.lam)
                r.update(it + 1)
                r.storage.clear()
                curr = r._ingest(last_obs.clone(), r.storage.obs_slot())
            assert (r._graph is not None) == graph
            out.append(({k: v.clone() for k, v in r.actor_critic.state_dict().items()}, dict(r.log_dict)))
        for k in out[0][0]:
            assert torch.equal(out[0][0][k], out[1][0][k]), k
        for k in ("Train/surrogate_loss", "Train/value_function_loss", "Train/kl", "Train/kl_update_count"):
            assert float(out[0][1][k]) == float(out[1][1][k]) or (out[0][1][k] != out[0][1][k]), k


@pytest.mark.parametrize("schedule", ["linear_decay", "step_decay"])
def test_lr_schedule_reaches_the_replayed_graph(schedule):
    """ppo.py:390-400: the actor's learning rate changes after every update.  It is a device scalar, so the captured update graph must
    pick the new value up: four iterations replayed from the graph equal four eager iterations bit for bit, the logged rate follows the
    reference's formula, and the critic's rate never moves."""
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    from tests.helpers import MLP128, ppo_cfg
    out = []
    for graph in (False, True):
        torch.manual_seed(4)
        env = FakeVecEnv(16, 53, 10, DEV, cloud=False, seed=8)
        r = ppo(env, ppo_cfg(16, MLP128, device=DEV, cuda_graph=graph, lr=1e-3, lr_schedule=schedule, max_iterations=4), _Logger())
        curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
        g = torch.Generator().manual_seed(0)
        lrs = []
        for it in range(4):
            eps = torch.randn(8, 16, 10, generator=g).to(DEV)
            last_obs, last_values = r.collect(curr, None, eps=eps)
            r.storage.compute_returns(last_values, r.gamma, r.lam)
            r.update(it + 1)
            r.storage.clear()
            curr = r._ingest(last_obs.clone(), r.storage.obs_slot())
            lrs.append(r.log_dict["Train/learning_rate"])
            assert abs(float(r.optimizer_actor.opt_state[1]) - lrs[-1]) <= 1e-12 * max(1.0, lrs[-1]) + 1e-9       # the device scalar follows
            assert abs(float(r.optimizer_critic.opt_state[1]) - 1e-3) < 1e-9
        want = [max(1e-3 * (1 - (i + 1) / 4), 1e-5) for i in range(4)] if schedule == "linear_decay" else [1e-3, 1e-3, 1e-5, 1e-5]
        assert lrs == pytest.approx(want, rel=1e-12)
        out.append({k: v.clone() for k, v in r.actor_critic.state_dict().items()})
    for k in out[0]:
        assert torch.equal(out[0][k], out[1][k]), k


def test_update_with_the_random_sampler_vs_oracle():
    """`sampler: random` (storage.py:133-137): true gathers instead of slices, a fresh host permutation per epoch — actor epochs first,
    then the critic's.  One update from a collected buffer against the oracle's update fed by the same generator stream."""
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    from tests.helpers import MLP128, ppo_cfg
    E, D, A = 64, 53, 10
    torch.manual_seed(12)
    env = FakeVecEnv(E, D, A, DEV, cloud=False, seed=2)
    cfg = ppo_cfg(E, MLP128, device=DEV, sampler="random", lr=3e-4)
    r = ppo(env, cfg, _Logger())
    assert r._graph is None and not r.cuda_graph                           # host-side permutations: no captured update
    init = {k: v.detach().cpu().clone() for k, v in r.actor_critic.state_dict().items()}
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    last_obs, last_values = r.collect(curr, None)
    r.storage.compute_returns(last_values, r.gamma, r.lam)
    st = r.storage
    flat = lambda t: t.detach().cpu().reshape(-1, t.shape[-1]).clone()
    buf = dict(obs=flat(st.observations), actions=flat(st.actions), values=flat(st.values), returns=flat(st.returns),
               logp=flat(st.actions_log_prob), adv=flat(st.advantages), mu=flat(st.mu), sigma=flat(st.sigma))
    torch.manual_seed(77)                                                  # the permutations come from the default host generator
    r.update(1)
    actor = {k[len("actor."):]: v.clone() for k, v in init.items() if k.startswith("actor.")}
    critic = {k[len("critic."):]: v.clone() for k, v in init.items() if k.startswith("critic.")}
    log_std = init["log_std"].clone()
    opt_a, opt_c = O.AdamState({**actor, "log_std": log_std}, 3e-4), O.AdamState(critic, 3e-4)
    stats = O.ppo_update(actor, critic, log_std, opt_a, opt_c, buf, cfg, "MLP", MLP128, gen=torch.Generator().manual_seed(77))
    assert stats["count"] == int(r.log_dict["Train/kl_update_count"]) and opt_c.step == r.optimizer_critic.step_count == 40
    for k, want in (("Train/surrogate_loss", stats["surrogate_loss"]), ("Train/value_function_loss", stats["value_loss"]), ("Train/kl", stats["kl"])):
        assert abs(float(r.log_dict[k]) - float(want)) <= 1e-3 * max(1.0, abs(float(want))), (k, float(r.log_dict[k]), float(want))
    got = {k: v.detach().cpu() for k, v in r.actor_critic.state_dict().items()}
    disp = 40 * 3e-4
    for k, v in critic.items():
        assert float((got["critic." + k] - v).abs().max()) <= 0.05 * disp, k
    for k, v in actor.items():
        assert float((got["actor." + k] - v).abs().max()) <= 0.05 * disp, k
