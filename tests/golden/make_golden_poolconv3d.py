"""Golden fixture for the pooled Conv3D student (build container only): the UNMODIFIED reference `PoolConv3DNet`
(/root/reference/algorithms/algo_utils/network.py:100-117) forward + backward on a seeded TSDF-like batch.

    python tests/golden/make_golden_poolconv3d.py
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
for tag, act in (("tanh", "tanh"), ("relu", "relu")):
    torch.manual_seed(15 if act == "tanh" else 16)
    m = net.PoolConv3DNet(50 ** 3, 10, {"activation": act}, 0)
    B = 3
    x = torch.clamp(torch.randn(B, 50 ** 3) * 0.6 + 0.5, -1, 1)
    y = m(x)
    gy = torch.randn_like(y)
    grads = torch.autograd.grad((y * gy).sum(), list(m.parameters()))
    out[f"{tag}_x"], out[f"{tag}_y"], out[f"{tag}_gy"] = x.numpy(), y.detach().numpy(), gy.numpy()
    for (name, p), g in zip(m.named_parameters(), grads):
        out[f"{tag}_param_{name}"] = p.detach().numpy()
        out[f"{tag}_grad_{name}"] = g.numpy()
    enc = m.encoder(x.reshape(B, 1, 50, 50, 50))
    print(tag, "y", tuple(y.shape), "params", sum(p.numel() for p in m.parameters()), "encoder out", tuple(enc.shape), "pooled", tuple(m.maxpool(enc).shape))
np.savez_compressed(os.path.join(HERE, "poolconv3d_student.npz"), **out)
print("wrote poolconv3d_student.npz", os.path.getsize(os.path.join(HERE, "poolconv3d_student.npz")) // 1024, "KB")
