"""Dev tool: per-phase clock64 stamps of tile #3 of CTA 0 (thread 0) of the tcgen05 encoder backward (build with EXTRA=-DPM_TC_TIMING)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ppo_oracle as O
from partmanip_b200 import ops
dev = "cuda:0"
B, N, C = 2048, 1024, 3
NAMES = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")
x = torch.rand(B, N * C, device=dev) * 2 - 1
p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(3))
enc = [p[k].to(dev) for k in NAMES]
grads = [torch.empty_like(t) for t in enc]
feat = torch.empty(B, 512, device=dev); am = torch.empty(B, 512, device=dev, dtype=torch.int32)
ops.pointnet_encode_forward(x, N, C, enc, "tanh", "bf16", feat, None, am, None)
dfeat = torch.randn(B, 512, device=dev) * 0.01
for _ in range(3):
    ops.pointnet_encode_backward(x, N, C, enc, "tanh", dfeat, am, grads, precision="bf16")
torch.cuda.synchronize()
ws = ops._scratch[(dev, "encbwd_bf16")]
off = 65536 + 8 * 32768 + 64
d = ws[off:off + 64 * 8].view(torch.int64).cpu()
lab = ["tile start", "S1 done", "sync1 passed", "M1 issued", "ACC2_FULL ok", "S2 done", "sync2 passed", "M3+M5+M2 issued", "-",
       "ACC1_FULL ok", "S3 done", "sync3 passed", "M4 done + read back", "sync4 passed (tile end)"]
base = int(d[0])
for i, l in enumerate(lab):
    if l != "-": print(f"{int(d[i]) - base:8d}  {l}")
