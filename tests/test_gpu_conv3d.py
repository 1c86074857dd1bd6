"""GPU parity of the next row §8f(3), second half: the `Conv3DNet` TSDF student (algorithms/algo_utils/network.py:56-135) on the
kernels (pm_conv3d_im2col / pm_conv3d_col2im + the tcgen05 dense layers) against the recording of the UNMODIFIED reference module
(tests/golden/conv3d_student.npz: outputs and every parameter gradient) and the numpy oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import conv3d_oracle as C

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv3d_student.npz"))


def _cfg(tag):
    params = {k[len(tag) + 7:]: G[k] for k in G.files if k.startswith(tag + "_param_")}
    grads = {k[len(tag) + 6:]: G[k] for k in G.files if k.startswith(tag + "_grad_")}
    return params, grads, G[tag + "_x"], G[tag + "_y"], G[tag + "_gy"]


def _net(params, D, out, act, proprio, precision):
    from partmanip_b200.algorithms.algo_utils.network import Conv3DNet
    net = Conv3DNet(D - proprio, out, dict(name="Conv3DNet", activation=act, precision=precision), proprio)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    return net.to(DEV)


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("tag,act,proprio", [("tanh_p0", "tanh", 0), ("relu_p7", "relu", 7)])
def test_conv3dnet_matches_the_reference_recording(tag, act, proprio, precision, rtol):
    params, want_grads, x, want_y, gy = _cfg(tag)
    net = _net(params, x.shape[1], want_y.shape[1], act, proprio, precision)
    assert [k for k, _ in net.named_parameters()] == list(params)            # same names and order as the reference's state_dict
    y = net(torch.from_numpy(x).to(DEV))
    got = y.detach().cpu().numpy()
    assert float(np.abs(got - want_y).max()) <= rtol + rtol * float(np.abs(want_y).max()), float(np.abs(got - want_y).max())
    (y * torch.from_numpy(gy).to(DEV)).sum().backward()
    for k, p in net.named_parameters():
        g, w = p.grad.cpu().numpy(), want_grads[k]
        assert g.shape == w.shape and np.isfinite(g).all(), k
        if precision == "fp32":
            tol = rtol * max(1e-3, float(np.abs(w).max()))
            assert float(np.abs(g - w).max()) <= tol, (k, float(np.abs(g - w).max()), tol)
        else:
            # bf16 operands: relative L2 per tensor.  relu's derivative is a step — a pre-activation within bf16 rounding of 0
            # flips a whole patch's contribution, and the recording has only 3 samples to average over
            rel = float(np.linalg.norm(g - w) / (np.linalg.norm(w) + 1e-12))
            assert rel <= (3e-2 if act == "tanh" else 2e-1), (k, rel)


def test_conv3dnet_batch_of_volumes_vs_oracle():
    """A larger batch of smooth synthetic volumes (values in [-1, 1] like a fused TSDF) against the numpy oracle, fp32 gate."""
    rng = np.random.default_rng(3)
    B, R, p = 9, 50, 0
    params, _, _, _, _ = _cfg("tanh_p0")
    zz, yy, xx = np.meshgrid(np.arange(R), np.arange(R), np.arange(R), indexing="ij")
    x = np.stack([np.clip(np.sin(0.11 * xx + e) * np.cos(0.07 * yy) + 0.02 * (zz - 25) + 0.1 * rng.standard_normal((R, R, R)), -1, 1)
                  for e in range(B)]).reshape(B, -1).astype(np.float32)
    gy = rng.standard_normal((B, 10)).astype(np.float32)
    want_y, want_g = C.conv3dnet_backward(x, params, "tanh", p, gy)
    net = _net(params, x.shape[1], 10, "tanh", p, "fp32")
    y = net(torch.from_numpy(x).to(DEV))
    assert float(np.abs(y.detach().cpu().numpy() - want_y).max()) <= 1e-4 + 1e-4 * float(np.abs(want_y).max())
    (y * torch.from_numpy(gy).to(DEV)).sum().backward()
    for k, pp in net.named_parameters():
        w = want_g[k]
        assert float(np.abs(pp.grad.cpu().numpy() - w).max()) <= 1e-4 * max(1e-3, float(np.abs(w).max())), k


def test_conv3d_patch_gather_and_its_adjoint():
    """im2col / col2im against the oracle's stride-trick windows for the three layer geometries; col2im is the exact adjoint:
    <im2col(x), g> == <x, col2im(g)>."""
    from partmanip_b200 import ops
    rng = np.random.default_rng(1)
    for (Cin, k, s, Din) in ((1, 5, 3, 50), (16, 3, 3, 17), (32, 3, 2, 6), (3, 3, 1, 5)):
        B = 2
        x = rng.standard_normal((B, Cin, Din, Din, Din)).astype(np.float32)
        want, _ = C._im2col(x, k, s)                                         # (B, Do, Ho, Wo, Cin*k^3)
        Do = want.shape[1]
        assert ops.conv3d_out_dim(Din, k, s) == Do
        xcl = torch.from_numpy(np.moveaxis(x, 1, -1).copy()).to(DEV)         # channels-last (B, D, H, W, C)
        K = Cin * k ** 3
        kpad = (K + 3) // 4 * 4
        cols = torch.full((B * Do ** 3, kpad), float("nan"), device=DEV)
        ops.conv3d_im2col(xcl, Cin, Din ** 3 * Cin, B, Cin, Din, k, s, cols)
        want_tm = want.reshape(-1, Cin, k ** 3).transpose(0, 2, 1).reshape(-1, K)     # the kernels' tap-major column order (kd, kh, kw, c)
        assert np.array_equal(cols[:, :K].cpu().numpy(), want_tm) and bool((cols[:, K:] == 0).all())
        w = torch.from_numpy(rng.standard_normal((5, Cin, k ** 3)).astype(np.float32)).to(DEV)
        wp = ops.conv3d_weight_permute(w, 5, Cin, k, True, torch.empty(5, k ** 3 * Cin, device=DEV))
        assert torch.equal(wp.view(5, k ** 3, Cin), w.permute(0, 2, 1)) and torch.equal(ops.conv3d_weight_permute(wp, 5, Cin, k, False, torch.empty_like(w)), w)
        g = torch.from_numpy(rng.standard_normal((B * Do ** 3, kpad)).astype(np.float32)).to(DEV)
        din = torch.empty(B * Din ** 3, Cin, device=DEV)
        ops.conv3d_col2im(g, B, Cin, Din, k, s, torch.zeros_like(din), None, din)        # act = none: derivative 1
        lhs = float((cols[:, :K].double() * g[:, :K].double()).sum())
        rhs = float((xcl.reshape(-1, Cin).double() * din.double()).sum())
        scale = float(cols[:, :K].double().norm() * g[:, :K].double().norm())      # din is summed in fp32: error ~1e-7 of this scale
        assert abs(lhs - rhs) <= 1e-5 * scale, (Cin, k, s, lhs, rhs, scale)


def test_dagger_update_with_the_conv3d_student_vs_oracle(tmp_path):
    """What dagger_tsdf.yaml ships: a Conv3DNet student on TSDF volumes (+ proprio tail) imitating a frozen state teacher.  One DAgger
    update step (dagger.py:299-337) through the `dagger` plugin against the numpy Conv3D oracle + the restated Adam step: loss and
    every updated student tensor."""
    from oracle import ppo_oracle as O
    from partmanip_b200.algorithms import dagger
    from partmanip_b200.envs import FakeVecEnv
    torch.manual_seed(2)
    E, A, Dt, p, R = 4, 10, 53, 7, 50
    D = R ** 3 + p
    g = torch.Generator().manual_seed(21)
    tea_cfg = dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="MLP", hid_dim=[64, 64], activation="elu"))
    tea_p = O.mlp_init(Dt, A, [64, 64], gen=g)
    tea_sd = {f"actor.{k}": v for k, v in tea_p.items()}
    tea_sd.update({f"critic.{k}": v for k, v in O.mlp_init(Dt, 1, [64, 64], gen=g).items()})
    tea_sd["log_std"] = torch.full((A,), -0.69)
    path = str(tmp_path / "teacher.pth")
    torch.save(dict(obs_mode="state", model_cfg=tea_cfg, model_state_dict=tea_sd, tricks=dict(use_state_norm=False)), path)

    class _Logger:
        save_ckpt_dir = save_video_dir = save_pose_dir = str(tmp_path)

        def info(self, d, it):
            pass
    cfg = dict(num_envs=E, obs_mode="obs", max_iterations=10, n_steps=4, n_updates=1, n_minibatches=1, device=DEV, buf_size=4,
               reward_reset=False, add_proprio_obs=True, offline_data_pth=None, eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9,
               test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=1e-4, teacher=path, resume=None, pretrain=None,
               sampler="sequential", model=dict(action_std=0.1, action_activate="tanh", clipAction=1.0,
                                                network=dict(name="Conv3DNet", activation="tanh")))
    env = FakeVecEnv(E, D, A, DEV, cloud=False, seed=5, extra_obs={"state": Dt, "proprio_state": p})
    r = dagger(env, cfg, _Logger())
    assert type(r.student.actor).__name__ == "Conv3DNet" and r.student.actor.proprio_shape == p and r.student.actor.res == R
    w0 = {k[len("actor."):]: v.detach().cpu().numpy().copy() for k, v in r.student.state_dict().items() if k.startswith("actor.")}
    stu_obs, tea_obs = r._reset_env()
    stu_obs, tea_obs, _ = r._collect(stu_obs, tea_obs)
    S, Tt = r.storage.observations[:16].cpu(), r.storage.tea_obs[:16].cpu()
    r.update(1)
    # ---- oracle: loss = mean((tea_act - tanh(mu))^2), d/dmu, Conv3D backward, Adam step 1
    tea_act = O.action_activation(O.mlp_forward(tea_p, Tt, "elu"), 1.0).numpy()
    mu, _ = C.conv3dnet_forward(S.numpy(), w0, "tanh", p, keep=True)
    t = np.tanh(mu)
    loss = float(((tea_act - t) ** 2).mean())
    gy = (2.0 * (t - tea_act) * (1 - t * t) / t.size).astype(np.float32)
    _, grads = C.conv3dnet_backward(S.numpy(), w0, "tanh", p, gy)
    assert abs(r.log_dict["Train/dagger_loss"] - loss) <= 1e-4 * max(1.0, abs(loss)), (r.log_dict["Train/dagger_loss"], loss)
    opt = O.AdamState({k: torch.from_numpy(v.copy()) for k, v in w0.items()}, 1e-4)
    opt.apply({k: torch.from_numpy(v) for k, v in grads.items()})
    got = {k[len("actor."):]: v.detach().cpu() for k, v in r.student.state_dict().items() if k.startswith("actor.")}
    for k, v in opt.params.items():
        d = (got[k] - v).abs()
        # Adam's first step moves every weight by ~lr * sign(g): elements whose gradient is within rounding of 0 may differ by up to 2 lr
        assert float((d > 0.05 * 1e-4).float().mean()) <= 0.02 and float(d.max()) <= 2.1e-4, (k, float((d > 5e-6).float().mean()), float(d.max()))


# ---------------------------------------------------------------------------------------------------------------- PoolConv3DNet
GP = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "poolconv3d_student.npz"))


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("tag", ["tanh", "relu"])
def test_poolconv3dnet_matches_the_reference_recording(tag, precision, rtol):
    """network.py:100-117 (stride-2 encoder, MaxPool3d(4) that keeps the [0,4)^3 corner of the 7^3 map, 64 -> 32 -> out head) against the
    recording of the unmodified module: outputs and every parameter gradient."""
    from partmanip_b200.algorithms.algo_utils.network import PoolConv3DNet
    params = {k[len(tag) + 7:]: GP[k] for k in GP.files if k.startswith(tag + "_param_")}
    want_grads = {k[len(tag) + 6:]: GP[k] for k in GP.files if k.startswith(tag + "_grad_")}
    x, want_y, gy = GP[tag + "_x"], GP[tag + "_y"], GP[tag + "_gy"]
    net = PoolConv3DNet(x.shape[1], 10, dict(name="PoolConv3DNet", activation=tag, precision=precision), 0)
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    net.to(DEV)
    assert [k for k, _ in net.named_parameters()] == list(params)
    y = net(torch.from_numpy(x).to(DEV))
    got = y.detach().cpu().numpy()
    assert float(np.abs(got - want_y).max()) <= rtol + rtol * float(np.abs(want_y).max()), float(np.abs(got - want_y).max())
    (y * torch.from_numpy(gy).to(DEV)).sum().backward()
    for k, p in net.named_parameters():
        g, w = p.grad.cpu().numpy(), want_grads[k]
        assert g.shape == w.shape and np.isfinite(g).all(), k
        if precision == "fp32":
            tol = rtol * max(1e-3, float(np.abs(w).max()))
            assert float(np.abs(g - w).max()) <= tol, (k, float(np.abs(g - w).max()), tol)
        else:       # bf16 operands: a pooled maximum may move to a neighbouring voxel; relative L2 per tensor
            rel = float(np.linalg.norm(g - w) / (np.linalg.norm(w) + 1e-12))
            assert rel <= 2e-1, (k, rel)


def test_maxpool3d_vs_torch():
    """pm_maxpool3d_forward / _backward against nn.MaxPool3d + autograd: ties (first maximum wins), a resolution that is not a multiple of
    the kernel (the tail voxels belong to no cell and get zero gradient), several cells per axis."""
    from partmanip_b200 import ops
    torch.manual_seed(0)
    B, C, Din, k = 3, 5, 9, 4                                      # Dp = 2: voxels 8 of each axis are dropped
    y = torch.randn(B, C, Din, Din, Din)
    y[0, 0, :4, :4, :4] = 1.5                                      # a whole cell tied
    y[1, 2, 4:8, 0:4, 4:8] = torch.round(y[1, 2, 4:8, 0:4, 4:8])   # many ties
    yr = torch.tanh(y).requires_grad_(True)                        # the pooled tensor is an activation output
    pooled = torch.nn.functional.max_pool3d(yr, k)
    dout = torch.randn_like(pooled)
    (gy,) = torch.autograd.grad(pooled, yr, dout)
    want_dpre = gy * (1 - yr.detach() ** 2)                        # times tanh' expressed through the output
    rows = yr.detach().permute(0, 2, 3, 4, 1).reshape(-1, C).contiguous().to(DEV)
    Dp = (Din - k) // k + 1
    out = torch.empty(B * Dp ** 3, C, device=DEV)
    arg = torch.empty(B * Dp ** 3, C, device=DEV, dtype=torch.int32)
    ops.maxpool3d_forward(rows, B, C, Din, k, out, arg)
    assert torch.equal(out.cpu(), pooled.detach().permute(0, 2, 3, 4, 1).reshape(-1, C))
    dpre = torch.full_like(rows, float("nan"))
    ops.maxpool3d_backward(dout.permute(0, 2, 3, 4, 1).reshape(-1, C).contiguous().to(DEV), arg, rows, "tanh", B, C, Din, k, dpre)
    want = want_dpre.permute(0, 2, 3, 4, 1).reshape(-1, C)
    assert float((dpre.cpu() - want).abs().max()) <= 1e-6


def test_actor_critic_resolves_poolconv3dnet():
    from partmanip_b200.algorithms.algo_utils import ActorCritic
    ac = ActorCritic(50 ** 3, 7, dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="PoolConv3DNet", activation="tanh")), 0).to(DEV)
    x = torch.rand(2, 50 ** 3, device=DEV) * 2 - 1
    a = ac.act(x)
    v = ac.cri(x)
    assert a.shape == (2, 7) and v.shape == (2, 1) and bool(torch.isfinite(a).all()) and float(a.abs().max()) <= 1.0
