"""Dev tool: dump the per-role clock64 stamps of tile #3 of the tcgen05 encoder (build with `make EXTRA=-DPM_TC_TIMING`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ppo_oracle as O
from partmanip_b200 import ops
dev = "cuda:0"
B, N, C = 2048, 1024, 3
x = torch.rand(B, N * C, device=dev) * 2 - 1
p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(3))
names = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")
enc = [p[k].to(dev) for k in names]
feat = torch.empty(B, 512, device=dev)
am = torch.empty(B, 512, device=dev, dtype=torch.int32)
for _ in range(3):
    ops.pointnet_encode_forward(x, N, C, enc, "tanh", "bf16", feat, None, am, None)
torch.cuda.synchronize()
print("err", ops.pointnet_tc_last_error(dev))
ws = ops._scratch[(dev, "encfwd")]
off = 2 * 163840 + 64
d = ws[off:off + 2 * 64 * 8].view(torch.int64).cpu().view(2, 64)
for cta in range(2):
    t = d[cta]
    base = int(t[0])
    lab = {0: "A tile start", 1: "A L1 done", 2: "A L3_DONE(prev) ok", 3: "A H1 stored+arrive", 4: "A ACC2_FULL ok", 5: "A E2 done",
           20: "B wait ACC2_FULL", 21: "B E2 half done", 24: "M wait H1", 25: "M H1_FULL ok", 26: "M L2 issued",
           27: "M H2_KB0 ok", 28: "M H2_KB2 ok", 29: "M H2_KB1 ok", 30: "M H2_KB3 ok", }
    for s in range(2):
        lab[8 + 3 * s] = f"B g{s} wait"; lab[9 + 3 * s] = f"B g{s} FULL ok"; lab[10 + 3 * s] = f"B g{s} done"
    lab[32] = "M g1 issued"
    print(f"--- CTA {cta}")
    for k in sorted(lab, key=lambda k: int(t[k])):
        if int(t[k]):
            print(f"{int(t[k]) - base:9d}  {lab[k]}")
