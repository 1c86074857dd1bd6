"""Network plugins `MLP` and `PointNet` — same constructor signature, parameter names and init order as
the reference (algorithms/algo_utils/network.py:27-54, 141-198) so checkpoints and seeds carry over, but
forward/backward run in the CUDA kernels of libpartmanip_b200.so (K1/K2/K3) instead of nn.Linear chains.

`module(x)` is autograd-compatible (a torch.autograd.Function around the kernels, used by DAgger-style
callers); the PPO engine drives `module.runner` directly on flat parameter/gradient buffers.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ... import ops

import os
_FUSED_HEAD = os.environ.get("PM_GENERIC_HEAD", "0") != "1"   # dev A/B switch: generic per-layer GEMM kernels instead
_ACT_NAMES = ("elu", "selu", "relu", "crelu", "lrelu", "tanh", "sigmoid")


class _Act(nn.Module):
    """Stateless placeholder keeping nn.Sequential indices identical to the reference (0,2,4: Linear; 1,3: act)."""

    def __init__(self, name: str):
        super().__init__()
        self.name = name

    def forward(self, x):  # pragma: no cover - the kernels apply the activation
        raise RuntimeError("activation modules are placeholders; use the owning network's forward")


def get_activation(act_name: str):
    """network.py:7-24 — unknown names print and return None there; here they raise (a None module would
    only fail later inside nn.Sequential)."""
    if act_name not in _ACT_NAMES:
        raise NotImplementedError(f"invalid activation function {act_name!r}")
    return _Act(act_name)


class _Runner:
    """Per-network kernel sequencing with buffers cached per batch size."""

    def __init__(self, net: "nn.Module"):
        self.net = net
        self._bufs: Dict[int, dict] = {}

    def params(self) -> List[torch.Tensor]:
        return [p for p in self.net.parameters()]


class _MLPRunner(_Runner):
    def _get(self, B: int, dev):
        b = self._bufs.get(B)
        if b is None or b["dev"] != dev:
            dims = self.net.dims
            b = {"dev": dev,
                 "h": [torch.empty(B, d, device=dev) for d in dims[1:]],
                 "d": [torch.empty(B, d, device=dev) for d in dims[1:-1]]}
            self._bufs[B] = b
        return b

    def forward(self, x: torch.Tensor, need_backward: bool = True) -> torch.Tensor:
        net = self.net
        B = x.shape[0]
        buf = self._get(B, x.device)
        h = x
        L = len(net.linears)
        for i, lin in enumerate(net.linears):
            act = net.act_name if i != L - 1 else None
            if net.precision == "fp32_ffma":
                h = ops.linear_forward(h, lin.weight, lin.bias, act, out=buf["h"][i])
            else:
                h = ops.linear_forward_tc(h, lin.weight, lin.bias, act, net.precision, out=buf["h"][i])
        return h

    def backward(self, x: torch.Tensor, dout: torch.Tensor, grads: List[torch.Tensor]):
        """grads: [W0,b0,W1,b1,...] tensors to overwrite; uses activations saved by the last forward(x)."""
        net = self.net
        B = x.shape[0]
        buf = self._get(B, x.device)
        L = len(net.linears)
        dpre = dout
        for i in reversed(range(L)):
            lin = net.linears[i]
            inp = x if i == 0 else buf["h"][i - 1]
            dx = None if i == 0 else buf["d"][i - 1]
            if net.precision == "fp32_ffma":
                ops.linear_backward(inp, lin.weight, dpre, grads[2 * i], grads[2 * i + 1], dx, net.act_name if i > 0 else None)
            else:
                ops.linear_backward_tc(inp, lin.weight, dpre, grads[2 * i], grads[2 * i + 1], dx,
                                       net.act_name if i > 0 else None, net.precision)
            dpre = dx


class MLP(nn.Module):
    """network.py:27-54.  Linear(in,h0)-act-...-Linear(h_last,out); orthogonal init, gains sqrt2.. then 1 / 0.01.

    Build extension: `net_cfg['precision']` selects the arithmetic of the dense layers — "fp32" (default): tcgen05 with every
    fp32 value split into three bf16 terms, 1e-4 parity gate; "bf16": tcgen05 with bf16 operands, 1e-2 gate; "fp32_ffma": the
    CUDA-core kernels of dense.cu."""

    def __init__(self, input_dim, output_dim, net_cfg, proprio_shape=0):
        super().__init__()
        hidden_dim = net_cfg['hid_dim']
        activation = get_activation(net_cfg['activation'])
        layers = [nn.Linear(input_dim, hidden_dim[0]), activation]
        for l in range(len(hidden_dim)):
            if l == len(hidden_dim) - 1:
                layers.append(nn.Linear(hidden_dim[l], output_dim))
            else:
                layers.append(nn.Linear(hidden_dim[l], hidden_dim[l + 1]))
                layers.append(activation)
        self.model = nn.Sequential(*layers)
        init_weights = [math.sqrt(2)] * len(hidden_dim)
        self.output_dim = output_dim
        init_weights.append(1 if output_dim == 1 else 0.01)
        for idx, module in enumerate(m for m in self.model if isinstance(m, nn.Linear)):
            torch.nn.init.orthogonal_(module.weight, gain=init_weights[idx])
        self.act_name = net_cfg['activation']
        self.dims = [input_dim, *hidden_dim, output_dim]
        self.precision = net_cfg.get('precision', 'fp32')
        if self.precision not in ("fp32", "bf16", "fp32_ffma"):
            raise NotImplementedError(f"precision {self.precision!r}")
        self.runner = _MLPRunner(self)

    @property
    def linears(self):
        return [m for m in self.model if isinstance(m, nn.Linear)]

    def forward(self, x):
        return _NetFunction.apply(self, x, *self.parameters())


class _PointNetRunner(_Runner):
    def _get(self, B: int, dev):
        b = self._bufs.get(B)
        if b is None or b["dev"] != dev:
            net = self.net
            F = net.feat_dim
            b = {"dev": dev,
                 "feat": torch.empty(B, F, device=dev),
                 "argmax": torch.empty(B, 512, device=dev, dtype=torch.int32),
                 "h2mean": torch.empty(B, 256, device=dev) if net.max_mean_concat else None,
                 "h1": torch.empty(B, 128, device=dev), "h2": torch.empty(B, 32, device=dev),
                 "out": torch.empty(B, net.output_dim, device=dev),
                 "dh2": torch.empty(B, 32, device=dev), "dh1": torch.empty(B, 128, device=dev),
                 "dfeat": torch.empty(B, F, device=dev)}
            self._bufs[B] = b
        return b

    def enc_params(self):
        m = self.net.mlp
        return [m[0].weight, m[0].bias, m[2].weight, m[2].bias, m[4].weight, m[4].bias]

    def head_params(self):
        f = self.net.final_mlp
        return [f[0].weight, f[0].bias, f[2].weight, f[2].bias, f[4].weight, f[4].bias]

    def forward(self, x: torch.Tensor, need_backward: bool = True) -> torch.Tensor:
        """need_backward=False (rollout / evaluation): the encoder skips the argmax bookkeeping of the max-pool."""
        net = self.net
        B = x.shape[0]
        N, C, p = net.point_num, net.in_channels, net.proprio_shape
        buf = self._get(B, x.device)
        if net.substract_mean:
            ops.pointnet_center_(x, N, C)                      # in place through the caller's tensor (Q3)
        feat = buf["feat"]
        fmean = feat[:, 512:1024] if net.max_mean_concat else None
        prec = net.precision if not net.max_mean_concat else "fp32"
        ops.pointnet_encode_forward(x, N, C, self.enc_params(), net.act_name, prec, feat[:, :512], fmean,
                                    buf["argmax"] if (need_backward or net.max_mean_concat) else None, buf["h2mean"])
        if p:
            ops.copy_rows(x[:, N * C:N * C + p], feat[:, net.feat_dim - p:])
        f = net.final_mlp
        if net.output_dim <= 32 and _FUSED_HEAD:      # fused head: one launch
            return ops.pointnet_head_forward(feat, self.head_params(), net.output_dim, net.act_name, buf["h1"], buf["h2"],
                                             buf["out"])
        ops.linear_forward(feat, f[0].weight, f[0].bias, net.act_name, out=buf["h1"])
        ops.linear_forward(buf["h1"], f[2].weight, f[2].bias, net.act_name, out=buf["h2"])
        return ops.linear_forward(buf["h2"], f[4].weight, f[4].bias, None, out=buf["out"])

    def backward(self, x: torch.Tensor, dout: torch.Tensor, grads: List[torch.Tensor]):
        """grads in parameter order: mlp.{0,2,4}.{weight,bias}, final_mlp.{0,2,4}.{weight,bias}."""
        net = self.net
        B = x.shape[0]
        N, C = net.point_num, net.in_channels
        buf = self._get(B, x.device)
        f = net.final_mlp
        if net.output_dim <= 32 and _FUSED_HEAD:      # fused head backward: three launches
            ops.pointnet_head_backward(buf["feat"], self.head_params(), net.output_dim, net.act_name, buf["h1"], buf["h2"],
                                       dout, grads[6:12], buf["dfeat"], 512 * (1 + net.max_mean_concat))
        else:
            ops.linear_backward(buf["h2"], f[4].weight, dout, grads[10], grads[11], buf["dh2"], net.act_name)
            ops.linear_backward(buf["h1"], f[2].weight, buf["dh2"], grads[8], grads[9], buf["dh1"], net.act_name)
            ops.linear_backward(buf["feat"], f[0].weight, buf["dh1"], grads[6], grads[7], buf["dfeat"], None)
        dfm = buf["dfeat"][:, 512:1024] if net.max_mean_concat else None
        prec = net.precision if not net.max_mean_concat else "fp32"
        ops.pointnet_encode_backward(x, N, C, self.enc_params(), net.act_name, buf["dfeat"][:, :512], buf["argmax"],
                                     grads[0:6], dfm, buf["h2mean"], precision=prec)


class PointNet(nn.Module):
    """network.py:141-198.  Per-point Linear(C,128)-act-Linear(128,256)-act-Linear(256,512), max (| mean) pool,
    [cat proprio], head Linear(F,128)-act-Linear(128,32)-act-Linear(32,out).  PyTorch default Linear init.

    Build extensions over the reference (which hard-codes point_num=1024, network.py:146): `net_cfg['point_num']`
    (default 1024) and `net_cfg['precision']` (encoder): "fp32" (default; split-operand tcgen05 kernels at the reference's
    precision), "bf16" (tcgen05, 1e-2 gate) or "fp32_ffma" (CUDA cores only).  The head is fp32 in every mode."""

    def __init__(self, input_dim, output_dim, net_cfg, proprio_shape=0):
        super().__init__()
        self.activation = get_activation(net_cfg['activation'])
        self.max_mean_concat = bool(net_cfg['max_mean'])
        self.point_num = int(net_cfg.get('point_num', 1024))
        # network.py:148 — channels per point = input_dim // point_num (the floor absorbs a proprio tail < point_num)
        self.in_channels = input_dim // self.point_num
        self.mlp = nn.Sequential(
            nn.Linear(input_dim // self.point_num, 128), self.activation,
            nn.Linear(128, 256), self.activation,
            nn.Linear(256, 512),
        )
        self.final_mlp = nn.Sequential(
            nn.Linear(512 * (1 + self.max_mean_concat) + proprio_shape, 128), self.activation,
            nn.Linear(128, 32), self.activation,
            nn.Linear(32, output_dim),
        )
        self.proprio_shape = proprio_shape
        self.substract_mean = bool(net_cfg['sub_mean'])
        self.act_name = net_cfg['activation']
        self.output_dim = output_dim
        self.precision = net_cfg.get('precision', 'fp32')
        if self.precision not in ("fp32", "bf16", "fp32_ffma"):
            raise NotImplementedError(f"precision {self.precision!r}")
        self.feat_dim = 512 * (1 + self.max_mean_concat) + proprio_shape
        self.runner = _PointNetRunner(self)

    def forward(self, x):
        return _NetFunction.apply(self, x, *self.parameters())


class _Conv3DRunner(_Runner):
    """Conv3DNet on the kernels: per convolution pm_conv3d_im2col + a dense layer on tcgen05 (channels-last activations), the head as
    two dense layers; backward = the dense backward + pm_conv3d_col2im."""
    SPECS = ((1, 16, 5, 3), (16, 32, 3, 3), (32, 32, 3, 2))          # (Cin, Cout, k, stride): network.py:70
    POOL = 0                                                         # PoolConv3DNet: nn.MaxPool3d(POOL) after the encoder
    HID = 256                                                        # width of final_mlp's hidden layer

    def _get(self, B: int, dev):
        b = self._bufs.get(B)
        if b is None or b["dev"] != dev:
            net = self.net
            dims = [net.res]
            for _, _, k, s in self.SPECS:
                dims.append(ops.conv3d_out_dim(dims[-1], k, s))
            b = {"dev": dev, "dims": dims, "cols": [], "y": [], "dcols": [], "dpre": []}
            for i, (ci, co, k, s) in enumerate(self.SPECS):
                rows, kpad = B * dims[i + 1] ** 3, (ci * k ** 3 + 3) // 4 * 4
                b["cols"].append(torch.empty(rows, kpad, device=dev) if i > 0 else None)    # layer 1 runs directly on the volume
                b["y"].append(torch.empty(rows, co, device=dev))
                b["dcols"].append(torch.empty(rows, kpad, device=dev) if i > 0 else None)
                b["dpre"].append(torch.empty(rows, co, device=dev))
            # conv2 / conv3 weights (and their gradients) in the tap-major order of the patch columns
            b["wp"] = [None] + [torch.empty(co, ci * k ** 3, device=dev) for ci, co, k, _ in self.SPECS[1:]]
            b["dwp"] = [None] + [torch.empty(co, ci * k ** 3, device=dev) for ci, co, k, _ in self.SPECS[1:]]
            F, H, c3 = net.feat_dim, self.HID, self.SPECS[2][1]
            b.update(flat=torch.empty(B, F, device=dev), h=torch.empty(B, H, device=dev), out=torch.empty(B, net.output_dim, device=dev),
                     dh=torch.empty(B, H, device=dev), dflat=torch.empty(B, F, device=dev))
            if self.POOL:
                dp = (dims[3] - self.POOL) // self.POOL + 1
                b.update(pdim=dp, pooled=torch.empty(B * dp ** 3, c3, device=dev), parg=torch.empty(B * dp ** 3, c3, device=dev, dtype=torch.int32),
                         dpooled=torch.empty(B * dp ** 3, c3, device=dev))
            self._bufs[B] = b
        return b

    def _convs(self):
        e = self.net.encoder
        return [e.conv1, e.conv2, e.conv3]

    def forward(self, x: torch.Tensor, need_backward: bool = True) -> torch.Tensor:
        net, prec = self.net, self.net.precision
        B = x.shape[0]
        buf = self._get(B, x.device)
        dims, act = buf["dims"], net.act_name
        src, ld_in, sstride = x, 1, x.stride(0)                      # the TSDF volume: one channel, voxels contiguous per sample
        for i, ((ci, co, k, s), conv) in enumerate(zip(self.SPECS, self._convs())):
            if i == 0:       # one input channel: direct fp32 kernel on the raw volume, no 5 GB patch matrix
                ops.conv3d_first_forward(x, dims[0], conv.weight.view(co, -1), conv.bias, act, buf["y"][0], stride=s)
                src, ld_in, sstride = buf["y"][0], co, dims[1] ** 3 * co
                continue
            ops.conv3d_im2col(src, ld_in, sstride, B, ci, dims[i], k, s, buf["cols"][i])
            ops.conv3d_weight_permute(conv.weight, co, ci, k, True, buf["wp"][i])
            ops.linear_forward_tc(buf["cols"][i][:, :ci * k ** 3], buf["wp"][i], conv.bias, act, prec, out=buf["y"][i])
            src, ld_in, sstride = buf["y"][i], co, dims[i + 1] ** 3 * co
        c3 = self.SPECS[2][1]
        feat_src, P = buf["y"][2], dims[3] ** 3
        if self.POOL:                                                # network.py:113: max-pool the encoder output
            ops.maxpool3d_forward(buf["y"][2], B, c3, dims[3], self.POOL, buf["pooled"], buf["parg"])
            feat_src, P = buf["pooled"], buf["pdim"] ** 3
        ops.conv3d_flatten(feat_src, buf["flat"], B, P, c3, net.feat_dim, True)
        if net.proprio_shape:
            ops.copy_rows(x[:, x.shape[1] - net.proprio_shape:], buf["flat"][:, c3 * P:])
        f = net.final_mlp
        ops.linear_forward_tc(buf["flat"], f[0].weight, f[0].bias, act, prec, out=buf["h"])
        return ops.linear_forward_tc(buf["h"], f[2].weight, f[2].bias, None, prec, out=buf["out"])

    def backward(self, x: torch.Tensor, dout: torch.Tensor, grads: List[torch.Tensor]):
        """grads in parameter order: encoder.conv{1,2,3}.{weight,bias}, final_mlp.{0,2}.{weight,bias}; uses the last forward(x)."""
        net, prec = self.net, self.net.precision
        B = x.shape[0]
        buf = self._get(B, x.device)
        dims, act = buf["dims"], net.act_name
        f = net.final_mlp
        c3 = self.SPECS[2][1]
        ops.linear_backward_tc(buf["h"], f[2].weight, dout, grads[8], grads[9], buf["dh"], act, prec)
        if self.POOL:
            # the pooled features are not an activation output of their own: plain d flat, routed to the winning voxels, times act'(y3)
            ops.linear_backward_tc(buf["flat"], f[0].weight, buf["dh"], grads[6], grads[7], buf["dflat"], None, prec)
            ops.conv3d_flatten(buf["dflat"], buf["dpooled"], B, buf["pdim"] ** 3, c3, net.feat_dim, False)
            ops.maxpool3d_backward(buf["dpooled"], buf["parg"], buf["y"][2], act, B, c3, dims[3], self.POOL, buf["dpre"][2])
        else:
            # d flat = (dh W0) * act'(flat): on the encoder columns flat IS act(conv3 pre-activation), so this is dPre3 (flatten order)
            ops.linear_backward_tc(buf["flat"], f[0].weight, buf["dh"], grads[6], grads[7], buf["dflat"], act, prec)
            ops.conv3d_flatten(buf["dflat"], buf["dpre"][2], B, dims[3] ** 3, c3, net.feat_dim, False)
        for i in (2, 1, 0):
            ci, co, k, s = self.SPECS[i]
            conv = self._convs()[i]
            if i == 0:
                ops.conv3d_first_backward(x, dims[0], buf["dpre"][0], grads[0].view(co, -1), grads[1], stride=s)
                continue
            ops.linear_backward_tc(buf["cols"][i][:, :ci * k ** 3], buf["wp"][i], buf["dpre"][i], buf["dwp"][i],
                                   grads[2 * i + 1], buf["dcols"][i][:, :ci * k ** 3] if i > 0 else None, None, prec)
            ops.conv3d_weight_permute(buf["dwp"][i], co, ci, k, False, grads[2 * i])
            if i > 0:
                ops.conv3d_col2im(buf["dcols"][i], B, ci, dims[i], k, s, buf["y"][i - 1], act, buf["dpre"][i - 1])


class _Conv3DEncoder(nn.Module):
    """network.py:119-135 `Encoder` — parameter container (conv1..conv3 as nn.Conv3d so names, shapes and default init match)."""

    def __init__(self, in_channels, filters, kernels, stride):
        super().__init__()
        self.conv1 = nn.Conv3d(in_channels, filters[0], kernels[0], stride=stride[0], padding=kernels[0] // 2)
        self.conv2 = nn.Conv3d(filters[0], filters[1], kernels[1], stride=stride[1], padding=kernels[1] // 2)
        self.conv3 = nn.Conv3d(filters[1], filters[2], kernels[2], stride=stride[2], padding=kernels[2] // 2)


class Conv3DNet(nn.Module):
    """network.py:67-97 — the TSDF student of dagger_tsdf.yaml / bc.yaml: Encoder(1, [16,32,32], [5,3,3], [3,3,2]) on the
    (res, res, res) volume, flatten (32*27) [cat proprio], Linear(.,256)-act-Linear(256,out).  Same constructor, parameter names
    (`encoder.conv{1,2,3}`, `final_mlp.{0,2}`) and init order as the reference.  `net_cfg['precision']`: "fp32" (default,
    three-term bf16 split on tcgen05, 1e-4 gate) or "bf16"."""

    def __init__(self, input_dim, output_dim, net_cfg, proprio_shape):
        super().__init__()
        self.res = round(input_dim ** (1 / 3))
        self.encoder = _Conv3DEncoder(1, [16, 32, 32], [5, 3, 3], [3, 3, 2])
        self.activation = get_activation(net_cfg['activation'])
        self.final_mlp = nn.Sequential(nn.Linear(32 * 27 + proprio_shape, 256), self.activation, nn.Linear(256, output_dim))
        self.proprio_shape = proprio_shape
        self.act_name = net_cfg['activation']
        self.output_dim = output_dim
        self.feat_dim = 32 * 27 + proprio_shape
        self.precision = net_cfg.get('precision', 'fp32')
        if self.precision not in ("fp32", "bf16"):
            raise NotImplementedError(f"precision {self.precision!r}")
        self.runner = _Conv3DRunner(self)

    def forward(self, x):
        return _NetFunction.apply(self, x, *self.parameters())


class _PoolConv3DRunner(_Conv3DRunner):
    SPECS = ((1, 16, 5, 2), (16, 32, 3, 2), (32, 64, 3, 2))          # network.py:104
    POOL = 4                                                         # network.py:106
    HID = 32


class PoolConv3DNet(nn.Module):
    """network.py:100-117: Encoder(1, [16,32,64], [5,3,3], [2,2,2]) -> MaxPool3d(4) -> Linear(64,32)-act-Linear(32,out).  The reference
    hard-codes 64 pooled features, i.e. one pooling cell (res 50: 25 -> 13 -> 7 voxels, of which the pool keeps the [0,4)^3 corner); the
    proprioceptive columns are not used by this network (its forward reshapes the whole row into the volume)."""

    def __init__(self, input_dim, output_dim, net_cfg, proprio_shape):
        super().__init__()
        self.res = round(input_dim ** (1 / 3))
        self.encoder = _Conv3DEncoder(1, [16, 32, 64], [5, 3, 3], [2, 2, 2])
        self.activation = get_activation(net_cfg['activation'])
        self.maxpool = nn.MaxPool3d(kernel_size=4)
        self.final_mlp = nn.Sequential(nn.Linear(64, 32), self.activation, nn.Linear(32, output_dim))
        self.proprio_shape = 0
        self.act_name = net_cfg['activation']
        self.output_dim = output_dim
        self.feat_dim = 64
        self.precision = net_cfg.get('precision', 'fp32')
        if self.precision not in ("fp32", "bf16"):
            raise NotImplementedError(f"precision {self.precision!r}")
        d = self.res
        for _, _, k, s in _PoolConv3DRunner.SPECS:
            d = ops.conv3d_out_dim(d, k, s)
        if d < 4 or (d - 4) // 4 + 1 != 1:
            raise ValueError(f"PoolConv3DNet: resolution {self.res} does not pool to the single cell final_mlp expects")
        self.runner = _PoolConv3DRunner(self)

    def forward(self, x):
        return _NetFunction.apply(self, x, *self.parameters())


class _NetFunction(torch.autograd.Function):
    """Autograd bridge: forward/backward through the kernels; gradients w.r.t. the parameters only
    (the observation is an input — the reference never differentiates through it either)."""

    @staticmethod
    def forward(ctx, net, x, *params):
        if x.shape[0] == 0:
            raise ValueError("empty batch")
        if x.dim() != 2:
            raise ValueError(f"expected (batch, obs_dim), got {tuple(x.shape)}")
        out = net.runner.forward(x)
        ctx.net = net
        ctx.save_for_backward(x)
        return out.clone()

    @staticmethod
    def backward(ctx, dout):
        net = ctx.net
        (x,) = ctx.saved_tensors
        grads = [torch.empty_like(p) for p in net.parameters()]
        net.runner.backward(x, dout.contiguous(), grads)
        return (None, None, *grads)
