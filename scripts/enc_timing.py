"""Dev tool: CUDA-event timing of the encoder forward / backward kernels on one 2048-cloud minibatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ppo_oracle as O
from partmanip_b200 import ops

dev = "cuda:0"
B, N, C = 2048, 1024, 3
NAMES = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")
xs = [torch.rand(B, N * C, device=dev) * 2 - 1 for _ in range(8)]     # 8 x 25 MB > L2 in rotation
p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(3))
enc = [p[k].to(dev) for k in NAMES]
grads = [torch.empty_like(t) for t in enc]
feat = torch.empty(B, 512, device=dev)
am = torch.empty(B, 512, device=dev, dtype=torch.int32)
dfeat = torch.randn(B, 512, device=dev) * 0.01


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


only = sys.argv[1].split(",") if len(sys.argv) > 1 else ("bf16", "fp32", "fp32_ffma")
for prec in ("bf16", "fp32", "fp32_ffma"):
    if prec not in only:
        continue
    f = timeit(lambda i: ops.pointnet_encode_forward(xs[i % 8], N, C, enc, "tanh", prec, feat, None, am, None), 3 if prec == "fp32_ffma" else 20)
    b = timeit(lambda i: ops.pointnet_encode_backward(xs[i % 8], N, C, enc, "tanh", dfeat, am, grads, precision=prec), 3 if prec == "fp32_ffma" else 20)
    fl = B * 2 * N * (C * 128 + 128 * 256 + 256 * 512)
    print(f"{prec}: forward {f:.3f} ms ({fl / f / 1e9:.0f} TFLOP/s)  backward {b:.3f} ms  uniq-crit/cloud "
          f"{float(torch.tensor([am[i].unique().numel() for i in range(16)]).float().mean()):.0f}")
print("errs", ops.pointnet_tc_last_error(dev), ops.pointnet_bwd_tc_last_error(dev), ops.pointnet_tc3_last_error(dev))
