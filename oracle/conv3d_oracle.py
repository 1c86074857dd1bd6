"""TEST INFRASTRUCTURE ONLY — CPU oracle of the "next" row §8f(3), second half: the Conv3D student that consumes the fused TSDF
volume.  Restates /root/reference/algorithms/algo_utils/network.py:56-135 (`conv_stride`, `Encoder`, `Conv3DNet`) with numpy,
float32 data, forward and backward (im2col; torch's `nn.Conv3d` cross-correlation, zero padding k//2):
    tsdf (B, R^3 [+p]) -> Conv3d(1,16,k5,s3) -> act -> Conv3d(16,32,k3,s3) -> act -> Conv3d(32,32,k3,s2) -> act
                       -> flatten (B, 32*27) [cat proprio] -> Linear(.,256) -> act -> Linear(256, out)
PINNED: tests/golden/conv3d_student.npz holds inputs, parameters, outputs and parameter gradients of the UNMODIFIED reference
module for two seeded configurations (tests/golden/make_golden_conv3d.py).  No CUDA kernels consume this oracle yet: it is the
first step (oracle + boundary) of that row.  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np

FILTERS, KERNELS, STRIDES = (16, 32, 32), (5, 3, 3), (3, 3, 2)       # network.py:70


def act_fwd(name: str, x):
    if name == "tanh":
        return np.tanh(x)
    if name in ("relu", "crelu"):
        return np.maximum(x, 0)
    raise NotImplementedError(name)                                   # the other activations of get_activation are not restated here


def act_bwd(name: str, y, gy):
    """gradient through the activation given its OUTPUT y"""
    if name == "tanh":
        return gy * (1 - y * y)
    return gy * (y > 0)


def _im2col(x, k, s):
    """x (B, C, D, H, W) -> columns (B, Do, Ho, Wo, C*k^3) of the zero-padded input (padding k//2), plus the padded shape."""
    p = k // 2
    xp = np.pad(x, ((0, 0), (0, 0), (p, p), (p, p), (p, p)))
    B, C, D, H, W = xp.shape
    Do, Ho, Wo = (D - k) // s + 1, (H - k) // s + 1, (W - k) // s + 1
    sb, sc, sd, sh, sw = xp.strides
    win = np.lib.stride_tricks.as_strided(xp, (B, Do, Ho, Wo, C, k, k, k), (sb, sd * s, sh * s, sw * s, sc, sd, sh, sw), writeable=False)
    return win.reshape(B, Do, Ho, Wo, C * k ** 3), xp.shape


def conv3d_fwd(x, w, b, s):
    """nn.Conv3d(stride=s, padding=k//2): x (B,Cin,D,H,W), w (Cout,Cin,k,k,k) -> (B,Cout,Do,Ho,Wo), and the columns for backward."""
    k = w.shape[-1]
    cols, _ = _im2col(x.astype(np.float32), k, s)
    y = cols @ w.reshape(w.shape[0], -1).T.astype(np.float32) + b.astype(np.float32)
    return np.moveaxis(y, -1, 1).astype(np.float32), cols


def conv3d_bwd(x_shape, cols, w, gy, s, need_gx=True):
    """-> (gx or None, gw, gb) for y = conv3d_fwd(x, w, b, s)."""
    k = w.shape[-1]
    Cout = w.shape[0]
    g = np.moveaxis(gy, 1, -1).astype(np.float32)                     # (B, Do, Ho, Wo, Cout)
    gw = (g.reshape(-1, Cout).T @ cols.reshape(-1, cols.shape[-1])).reshape(w.shape).astype(np.float32)
    gb = g.reshape(-1, Cout).sum(0).astype(np.float32)
    if not need_gx:
        return None, gw, gb
    B, C, D, H, W = x_shape
    p = k // 2
    gcols = (g @ w.reshape(Cout, -1).astype(np.float32)).reshape(g.shape[:4] + (C, k, k, k))
    gxp = np.zeros((B, C, D + 2 * p, H + 2 * p, W + 2 * p), np.float32)
    Do, Ho, Wo = g.shape[1:4]
    for a in range(k):                                                # scatter-add the k^3 taps back (col2im)
        for bb in range(k):
            for c in range(k):
                gxp[:, :, a:a + s * Do:s, bb:bb + s * Ho:s, c:c + s * Wo:s] += np.moveaxis(gcols[..., a, bb, c], -1, 1)
    return gxp[:, :, p:p + D, p:p + H, p:p + W], gw, gb


def conv3dnet_forward(x_in, params: dict, activation: str, proprio: int, keep=False):
    """Conv3DNet.forward (network.py:82-97).  params uses the reference's state_dict names."""
    x_in = np.asarray(x_in, np.float32)
    B = x_in.shape[0]
    res = round((x_in.shape[1] - proprio) ** (1 / 3))
    tsdf = x_in[:, :x_in.shape[1] - proprio].reshape(B, 1, res, res, res)
    h, saved = tsdf, []
    for i, s in enumerate(STRIDES, start=1):
        pre, cols = conv3d_fwd(h, params[f"encoder.conv{i}.weight"], params[f"encoder.conv{i}.bias"], s)
        y = act_fwd(activation, pre).astype(np.float32)
        saved.append((h.shape, cols, y))
        h = y
    flat = h.reshape(B, -1)
    if proprio:
        flat = np.concatenate([flat, x_in[:, -proprio:]], axis=-1)
    h1 = act_fwd(activation, flat @ params["final_mlp.0.weight"].T + params["final_mlp.0.bias"]).astype(np.float32)
    out = (h1 @ params["final_mlp.2.weight"].T + params["final_mlp.2.bias"]).astype(np.float32)
    return (out, (saved, flat, h1)) if keep else out


def conv3dnet_backward(x_in, params: dict, activation: str, proprio: int, gy):
    """Parameter gradients of sum(y * gy) (what autograd gives the reference)."""
    out, (saved, flat, h1) = conv3dnet_forward(x_in, params, activation, proprio, keep=True)
    gy = np.asarray(gy, np.float32)
    grads = {"final_mlp.2.weight": gy.T @ h1, "final_mlp.2.bias": gy.sum(0)}
    gh1 = act_bwd(activation, h1, gy @ params["final_mlp.2.weight"])
    grads["final_mlp.0.weight"] = gh1.T @ flat
    grads["final_mlp.0.bias"] = gh1.sum(0)
    gflat = gh1 @ params["final_mlp.0.weight"]
    n_enc = flat.shape[1] - proprio
    g = gflat[:, :n_enc].reshape(saved[-1][2].shape)
    for i in (3, 2, 1):
        x_shape, cols, y = saved[i - 1]
        g = act_bwd(activation, y, g)
        g, gw, gb = conv3d_bwd(x_shape, cols, params[f"encoder.conv{i}.weight"], g, STRIDES[i - 1], need_gx=i > 1)
        grads[f"encoder.conv{i}.weight"], grads[f"encoder.conv{i}.bias"] = gw, gb
    return out, {k: np.asarray(v, np.float32) for k, v in grads.items()}
