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
        assert np.array_equal(cols[:, :K].cpu().numpy(), want.reshape(-1, K)) and bool((cols[:, K:] == 0).all())
        g = torch.from_numpy(rng.standard_normal((B * Do ** 3, kpad)).astype(np.float32)).to(DEV)
        din = torch.empty(B * Din ** 3, Cin, device=DEV)
        ops.conv3d_col2im(g, B, Cin, Din, k, s, torch.zeros_like(din), None, din)        # act = none: derivative 1
        lhs = float((cols[:, :K].double() * g[:, :K].double()).sum())
        rhs = float((xcl.reshape(-1, Cin).double() * din.double()).sum())
        scale = float(cols[:, :K].double().norm() * g[:, :K].double().norm())      # din is summed in fp32: error ~1e-7 of this scale
        assert abs(lhs - rhs) <= 1e-5 * scale, (Cin, k, s, lhs, rhs, scale)
