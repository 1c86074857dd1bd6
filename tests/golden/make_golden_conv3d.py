"""Golden fixture for the Conv3D-student oracle (runs only in the build container): the UNMODIFIED reference `Conv3DNet`
(/root/reference/algorithms/algo_utils/network.py:56-135) forward + backward on a seeded TSDF-like input, with and without the
proprioceptive tail.  The module is loaded from its file; it needs torch + torchvision only.

    python tests/golden/make_golden_conv3d.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_network", "/root/reference/algorithms/algo_utils/network.py")
net = importlib.util.module_from_spec(spec)
spec.loader.exec_module(net)

out = {}
for tag, act, proprio, res, odim in (("tanh_p0", "tanh", 0, 50, 10), ("relu_p7", "relu", 7, 50, 10)):
    torch.manual_seed(5 if proprio == 0 else 6)
    m = net.Conv3DNet(res ** 3, odim, {"activation": act}, proprio)
    B = 3
    x = torch.clamp(torch.randn(B, res ** 3) * 0.6 + 0.5, -1, 1)
    if proprio:
        x = torch.cat([x, torch.randn(B, proprio)], dim=-1)
    x.requires_grad_(False)
    y = m(x)
    gy = torch.randn_like(y)
    grads = torch.autograd.grad((y * gy).sum(), list(m.parameters()))
    out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_gy"] = x.numpy(), y.detach().numpy(), gy.numpy()
    for (name, p), g in zip(m.named_parameters(), grads):
        out[f"{tag}_param_{name}"] = p.detach().numpy()
        out[f"{tag}_grad_{name}"] = g.numpy()
    print(tag, "y", tuple(y.shape), "params", sum(p.numel() for p in m.parameters()),
          "encoder out", tuple(m.encoder(x[:, :res ** 3].reshape(B, 1, res, res, res)).shape))
np.savez_compressed(os.path.join(HERE, "conv3d_student.npz"), **out)
print("wrote conv3d_student.npz", os.path.getsize(os.path.join(HERE, "conv3d_student.npz")) // 1024, "KB")
