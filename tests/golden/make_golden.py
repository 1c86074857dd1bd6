"""Generate golden fixtures from the UNMODIFIED reference (runs only in the build container).

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference has no tests of its own, so parity is pinned by executing its
hot-path modules on CPU with seeded inputs and recording inputs + outputs as .npz files next to
this script.  Two tiny stub modules stand in for `utils` (whose __init__ pulls isaacgym/skimage)
— see SURVEY.md §8(c).  The only hook into third-party code is capturing the standard-normal
draw inside torch.distributions so the sampled action can be reproduced from an explicit `eps`.
Nothing here is imported by the product or by the GPU tests; the .npz files are what travels.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    stub = types.ModuleType("utils")
    stub.path2video = lambda *a, **k: None
    sys.modules["utils"] = stub
    sys.modules["utils.torch_jit_utils"] = types.ModuleType("utils.torch_jit_utils")
    sys.path.insert(0, REF)
    import algorithms.algo_utils as au  # noqa
    from algorithms.ppo import ppo  # noqa
    from algorithms.algo_utils.network import PointNet, MLP  # noqa
    return au, ppo, PointNet, MLP


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **_np(arrs))
    print("wrote", name, {k: np.asarray(v).shape for k, v in _np(arrs).items()})


class _EpsTap:
    """Records every standard-normal draw MultivariateNormal makes (torch.distributions.utils._standard_normal)."""

    def __init__(self):
        import torch.distributions.multivariate_normal as mvn
        self.mvn = mvn
        self.orig = mvn._standard_normal
        self.draws = []

    def __enter__(self):
        def tapped(shape, dtype, device):
            e = self.orig(shape, dtype, device)
            self.draws.append(e.clone())
            return e
        self.mvn._standard_normal = tapped
        return self

    def __exit__(self, *a):
        self.mvn._standard_normal = self.orig


class FakeEnv:
    """Seeded zero-physics env with the attribute surface ppo.py consumes (hand_base.py:252-290)."""

    def __init__(self, E, D, A, obs_mode, seed, cloud=False, succ_p=0.0):
        self.g = torch.Generator().manual_seed(seed)
        self.num_envs, self.num_actions, self.max_episode_length = E, A, 200
        self.num_obs = {obs_mode: D, "proprio_state": 0}
        self.obs_mode, self.D, self.cloud, self.succ_p = obs_mode, D, cloud, succ_p
        self.train_test_flag = "train"
        self.reset_succ = torch.zeros(E, dtype=torch.bool)
        self.rew_buf = torch.zeros(E)
        self.log = {"obs": [], "rew": [], "done": [], "succ": []}

    def _obs(self):
        E, D = self.num_envs, self.D
        if self.cloud:
            n = D // 3
            pc = torch.rand(E, n, 3, generator=self.g)
            pc[..., :2] = pc[..., :2] * 2 - 1
            pc[..., 2] = pc[..., 2] * 2 + 0.05
            pad = torch.rand(E, n, generator=self.g) < 0.1
            pc[pad] = 0.0
            o = pc.reshape(E, D)
        else:
            o = torch.randn(E, D, generator=self.g)
        self.log["obs"].append(o.clone())
        return {self.obs_mode: o}

    def reset(self):
        return self._obs()

    def step(self, actions, save_image_path=None):
        E = self.num_envs
        self.rew_buf = torch.randn(E, generator=self.g)
        done = torch.rand(E, generator=self.g) < 0.05
        self.reset_succ = torch.rand(E, generator=self.g) < self.succ_p
        self.log["rew"].append(self.rew_buf.clone())
        self.log["done"].append(done.clone())
        self.log["succ"].append(self.reset_succ.clone())
        return self._obs(), self.rew_buf, done, {"succ_rate": torch.zeros(1)}


class _Logger:
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/pm_golden"

    def __init__(self):
        self.records = []

    def info(self, d, it):
        self.records.append({k: float(v) for k, v in d.items()})


def ppo_cfg(E, net, **over):
    cfg = dict(
        num_envs=E, obs_mode="obs", succ_value=None, max_iterations=1, n_steps=8, n_updates=5,
        n_minibatches=8, device="cpu", eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9,
        test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=5e-5, desired_kl=0.1,
        epsilon_clip=0.2, gamma=0.99, lam=0.95, sampler="sequential", resume=None,
        tricks=dict(mini_adv_norm=False, whole_adv_norm=False, use_state_norm=True,
                    use_clipped_value_loss=False, use_grad_clip=True, max_grad_norm=0.5),
        model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=net),
    )
    for k, v in over.items():
        if isinstance(v, dict) and k in cfg:
            cfg[k] = {**cfg[k], **v}
        else:
            cfg[k] = v
    return cfg


def golden_ppo_iteration(ppo, name, E, D, A, net, seed, cloud, **over):
    """One full ppo.run() iteration of the unmodified reference class; records the initial weights,
    every env tensor, the captured eps, the filled rollout buffer, the logged scalars and the
    post-update weights + Adam moments."""
    torch.manual_seed(seed)
    cfg = ppo_cfg(E, net, **over)
    env = FakeEnv(E, D, A, "obs", seed + 1, cloud=cloud, succ_p=0.02 if cfg["succ_value"] is not None else 0.0)
    log = _Logger()
    runner = ppo(env, cfg, log)
    init = {"init." + k: v.clone() for k, v in runner.actor_critic.state_dict().items()}
    # keep a handle on the storage tensors that compute_returns() rebinds / update() reads
    seen = {}
    orig_update = runner.update

    def tapped_update(it):
        st = runner.storage
        for k in ("observations", "actions", "rewards", "dones", "succs", "values", "returns",
                  "advantages", "actions_log_prob", "mu", "sigma"):
            seen["buf." + k] = getattr(st, k).clone()
        return orig_update(it)

    runner.update = tapped_update
    with _EpsTap() as tap:
        runner.run()
    out = {}
    out.update(init)
    out.update(seen)
    out.update({"final." + k: v for k, v in runner.actor_critic.state_dict().items()})
    for tag, opt in (("adam_actor", runner.optimizer_actor), ("adam_critic", runner.optimizer_critic)):
        sd = opt.state_dict()["state"]
        for i, s in sd.items():
            out[f"{tag}.{i}.step"] = torch.as_tensor(float(s["step"]))
            if s["exp_avg"].numel() <= 40000:   # keep fixtures small: moments of the big matrices are implied by the weights
                out[f"{tag}.{i}.exp_avg"] = s["exp_avg"]
                out[f"{tag}.{i}.exp_avg_sq"] = s["exp_avg_sq"]
    out["eps"] = torch.stack(tap.draws)                       # (T, E, A)
    out["env.obs"] = torch.stack(env.log["obs"])              # (T+1, E, D) raw (pre-normalisation; pre-centring)
    out["env.rew"] = torch.stack(env.log["rew"])
    out["env.done"] = torch.stack(env.log["done"])
    out["env.succ"] = torch.stack(env.log["succ"])
    if cfg["tricks"]["use_state_norm"]:
        ms = runner.state_norm.running_ms
        out["rms.mean"], out["rms.S"], out["rms.std"], out["rms.n"] = ms.mean, ms.S, ms.std, torch.as_tensor(ms.n)
    rec = log.records[-1]
    for k in ("Train/value_function_loss", "Train/surrogate_loss", "Train/kl", "Train/kl_max",
              "Train/kl_update_count", "Train/value_gt_return_mean", "Train/value_gt_return_max",
              "Train/mean_action_noise_std"):
        out["log." + k] = torch.as_tensor(rec[k])
    save(name, **out)


def main():
    au, ppo, PointNet, MLP = _import_reference()
    torch.set_num_threads(8)

    # ---- KAT-1: shipped checkpoint, RNG-free --------------------------------
    ck = torch.load(os.path.join(REF, "assets/ckpts/model_200000.pth"), map_location="cpu", weights_only=False)
    ac = au.ActorCritic(53, 10, ck["model_cfg"])
    ac.load_state_dict(ck["model_state_dict"])
    i = torch.arange(4, dtype=torch.float32)[:, None]
    j = torch.arange(53, dtype=torch.float32)[None, :]
    x = torch.sin(0.7 * i + 0.13 * j)
    ms = ck["state_running_ms"]
    xn = (x - ms["mean"]) / ms["std"]
    with torch.no_grad():
        act, val = ac.act_cri(xn)
        logp, ent, val2, mu, sig = ac.update_act_cri(xn, act)
    sd = {"w." + k: v for k, v in ck["model_state_dict"].items()}
    save("kat1_ckpt_mlp.npz", x=x, rms_mean=ms["mean"], rms_std=ms["std"], xn=xn, actions=act, values=val,
         logp=logp, entropy=ent, mu=mu, sigma=sig, clipAction=ck["model_cfg"]["clipAction"], **sd)
    # Adam states of the checkpoint are summarised only (step counters), for the resume test
    sa = ck["optimizer_actor"]["state"]
    sc = ck["optimizer_critic"]["state"]
    print("ckpt adam steps", float(sa[0]["step"]), float(sc[0]["step"]))

    # ---- GAE (KAT-2 / 2b + random) ------------------------------------------------------
    def run_gae(rew, val, done, succ, last, gamma, lam, succ_value, norm):
        T, E = rew.shape
        st = au.RolloutStorage(E, T, 4, 2, "cpu", succ_value, norm)
        st.rewards.copy_(rew[..., None]); st.values.copy_(val[..., None])
        st.dones.copy_(done[..., None]); st.succs.copy_(succ[..., None])
        st.compute_returns(last[:, None], gamma, lam)
        return st.returns.squeeze(-1), st.advantages.squeeze(-1)

    rew = torch.tensor([[1, 0, 2], [0, 1, 0], [3, 0, 1], [0, 2, 0]], dtype=torch.float32)
    val = torch.tensor([[.5, .1, .2], [.4, .3, .1], [.2, .6, .7], [.9, .8, .3]])
    done = torch.tensor([[0, 0, 0], [0, 1, 0], [0, 0, 0], [1, 0, 0]], dtype=torch.bool)
    succ = torch.zeros(4, 3, dtype=torch.bool)
    last = torch.tensor([1.0, 0.5, 0.25])
    r2, a2 = run_gae(rew, val, done, succ, last, 0.99, 0.95, None, False)
    succ_b = succ.clone(); succ_b[1, 1] = True
    r2b, a2b = run_gae(rew, val, done, succ_b, last, 0.99, 0.95, 500, True)
    g = torch.Generator().manual_seed(7)
    T, E = 8, 257
    rr = torch.randn(T, E, generator=g); vv = torch.randn(T, E, generator=g) * 3
    dd = torch.rand(T, E, generator=g) < 0.1; ss = torch.rand(T, E, generator=g) < 0.05
    ll = torch.randn(E, generator=g)
    r3, a3 = run_gae(rr, vv, dd, ss, ll, 0.99, 0.95, 500, False)
    r4, a4 = run_gae(rr, vv, dd, ss, ll, 0.97, 0.9, None, True)
    save("gae.npz", k2_rew=rew, k2_val=val, k2_done=done, k2_succ=succ, k2_last=last, k2_ret=r2, k2_adv=a2,
         k2b_succ=succ_b, k2b_ret=r2b, k2b_adv=a2b,
         r_rew=rr, r_val=vv, r_done=dd, r_succ=ss, r_last=ll, r3_ret=r3, r3_adv=a3, r4_ret=r4, r4_adv=a4)

    # ---- sampler geometry (KAT-3) ---------------------------------------------------------
    geo = []
    for (E_, T_, nmb) in [(64, 8, 8), (2048, 8, 8), (4096, 8, 8), (100, 8, 8), (3, 8, 8)]:
        st = au.RolloutStorage(E_, T_, 1, 1, "cpu")
        b = list(st.mini_batch_generator(nmb))
        geo.append([E_, T_, nmb, len(b), len(b[0]), b[0][0], b[-1][-1]])
    save("sampler_geometry.npz", geo=np.array(geo))

    # ---- RMS (Q9) -------------------------------------------------------------------------
    nrm = au.Normalization(11, "cpu")
    g = torch.Generator().manual_seed(3)
    xs, ys, means, Ss, stds = [], [], [], [], []
    for k in range(4):
        x = torch.randn(37, 11, generator=g) * (1 + k) + k
        y = nrm(x, update=(k != 3))
        xs.append(x); ys.append(y); means.append(nrm.running_ms.mean.clone())
        Ss.append(nrm.running_ms.S.clone()); stds.append(nrm.running_ms.std.clone())
    save("rms.npz", x=torch.stack(xs), y=torch.stack(ys), mean=torch.stack(means), S=torch.stack(Ss), std=torch.stack(stds))

    # ---- PointNet forward variants (KAT-4 shapes) + input mutation ------------------------
    def pn_case(tag, D, out, net_cfg, proprio, B, point_num=1024, seed=11):
        torch.manual_seed(seed)
        net = PointNet(D, out, net_cfg, proprio)
        if point_num != 1024:   # H6: the oracle for 2048-pt clouds is the class with point_num patched pre-construction
            torch.manual_seed(seed)
            net = PointNet.__new__(PointNet)
            torch.nn.Module.__init__(net)
            import torch.nn as nn
            from algorithms.algo_utils.network import get_activation
            net.activation = get_activation(net_cfg["activation"])
            net.max_mean_concat = net_cfg["max_mean"]
            net.point_num = point_num
            c = (D - proprio) // point_num
            net.mlp = nn.Sequential(nn.Linear(c, 128), net.activation, nn.Linear(128, 256), net.activation, nn.Linear(256, 512))
            net.final_mlp = nn.Sequential(nn.Linear(512 * (1 + net.max_mean_concat) + proprio, 128), net.activation,
                                          nn.Linear(128, 32), net.activation, nn.Linear(32, out))
            net.proprio_shape = proprio
            net.substract_mean = net_cfg["sub_mean"]
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.rand(B, D, generator=g) * 2 - 1
        x.view(-1)[::17] = 0.0
        x_in = x.clone()
        y = net(x_in)
        y.square().sum().backward()
        grads = {"g." + k: p.grad for k, p in net.named_parameters()}
        save(f"pointnet_{tag}.npz", x=x, x_after=x_in, y=y, n_params=sum(p.numel() for p in net.parameters()),
             **{"w." + k: v for k, v in net.state_dict().items()}, **grads)

    base = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False)
    pn_case("base_a10", 3072, 10, base, 0, 6)
    pn_case("maxmean_c4", 4096, 10, {**base, "max_mean": True}, 0, 4)
    pn_case("submean_proprio", 3072 + 25, 7, {**base, "sub_mean": True}, 25, 4)
    pn_case("relu_submean", 3072, 10, {**base, "activation": "relu", "sub_mean": True}, 0, 3)
    pn_case("n2048_c3", 6144, 10, base, 0, 3, point_num=2048)
    # (H6: the *unpatched* class given a 6144-d input builds Linear(6,128) over 1024 "points"; same code path as maxmean_c4's C=4.)

    # ---- MLP forward/backward (orthogonal init; state configs) -----------------------------
    torch.manual_seed(5)
    for tag, D, out, actn in (("mlp_actor", 37, 7, "tanh"), ("mlp_critic", 53, 1, "elu")):
        net = MLP(D, out, dict(hid_dim=[96, 160, 64], activation=actn), 0)
        x = torch.randn(9, D)
        y = net(x)
        y.square().sum().backward()
        save(f"{tag}.npz", x=x, y=y, **{"w." + k: v for k, v in net.state_dict().items()},
             **{"g." + k: p.grad for k, p in net.named_parameters()})

    # ---- full PPO iterations through the reference `ppo` class ------------------------------
    pn = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False)
    golden_ppo_iteration(ppo, "ppo_iter_pointnet_e16.npz", 16, 3072, 10, pn, seed=100, cloud=True)
    # NB sub_mean=True cannot be TRAINED in the reference: update_act_cri runs actor then critic on the same
    # obs_batch and the critic's in-place centring invalidates the tensor the actor's first Linear saved
    # (RuntimeError "modified by an inplace operation").  sub_mean is therefore pinned forward-only (pn_case).
    golden_ppo_iteration(ppo, "ppo_iter_pointnet_e8_nonorm.npz", 8, 3072, 10, pn,
                         seed=101, cloud=True, tricks=dict(use_state_norm=False, mini_adv_norm=True,
                                                           use_clipped_value_loss=True))
    mlp = dict(name="MLP", hid_dim=[128, 128, 128], activation="tanh")
    golden_ppo_iteration(ppo, "ppo_iter_mlp_e64.npz", 64, 37, 7, mlp, seed=102, cloud=False, succ_value=500,
                         tricks=dict(whole_adv_norm=True))
    golden_ppo_iteration(ppo, "ppo_iter_mlp_e64_klskip.npz", 64, 53, 10, mlp, seed=103, cloud=False, lr=3e-3,
                         desired_kl=0.02)


if __name__ == "__main__":
    main()
