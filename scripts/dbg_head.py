import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from partmanip_b200 import ops
DEV = "cuda:0"
for (B, F, out, act) in [(16, 512, 1, "tanh"), (16, 512, 10, "tanh"), (2048, 512, 1, "tanh"), (37, 537, 1, "tanh")]:
    torch.manual_seed(1)
    f = torch.tanh
    feat = torch.randn(B, F, dtype=torch.float64, requires_grad=True)
    Ws = [(torch.randn(128, F) / F ** 0.5), torch.randn(128) * 0.1, torch.randn(32, 128) / 128 ** 0.5, torch.randn(32) * 0.1,
          torch.randn(out, 32) / 32 ** 0.5, torch.randn(out) * 0.1]
    Ws = [w.double().requires_grad_(True) for w in Ws]
    h1 = f(feat @ Ws[0].T + Ws[1]); h2 = f(h1 @ Ws[2].T + Ws[3]); y = h2 @ Ws[4].T + Ws[5]
    dout = torch.randn(B, out, dtype=torch.float64)
    y.backward(dout)
    dW = [w.detach().float().to(DEV).contiguous() for w in Ws]
    g = [torch.full_like(w, float("nan")) for w in dW]
    h1d, h2d, yd = torch.empty(B, 128, device=DEV), torch.empty(B, 32, device=DEV), torch.empty(B, out, device=DEV)
    featd = feat.detach().float().to(DEV)
    ops.pointnet_head_forward(featd, dW, out, act, h1d, h2d, yd)
    dfeat = torch.full((B, 512), float("nan"), device=DEV)
    ops.pointnet_head_backward(featd, dW, out, act, h1d, h2d, dout.float().to(DEV), g, dfeat, 512)
    rel = lambda a, b: float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))
    print(B, F, out, "y", rel(yd, y.detach()), "grads", [f"{rel(a, w.grad):.1e}" for a, w in zip(g, Ws)], "dfeat", rel(dfeat, feat.grad[:, :512]))
