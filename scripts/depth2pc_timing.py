"""Dev tool: CUDA-event timing of the next-row kernels at the reference's shapes (3 views of 288x512 depth per env)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from partmanip_b200 import ops

dev = "cuda:0"
M, H, W = 3, 288, 512
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}


def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


intr = np.array([[366.0, 0, W // 2], [0, 366.0, H // 2], [0, 0, 1]])
pose = torch.eye(4, device=dev).repeat(M, 1, 1).contiguous()
pose[:, 2, 3] = -0.3
E = 256
depth = (0.3 + 0.6 * torch.rand(E, M, H, W, device=dev))
out = torch.empty(E, M * H * W, 3, device=dev)
ms = timeit(lambda: ops.depth2pc_backproject(depth, intr, pose, [-0.25, -0.25, -0.0503], 0.5, out), 10)
byt = E * M * H * W * 16
print(f"backproject: E={E} x {M}x{H}x{W}: {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s = {byt / ms / 1e6 / peaks['hbm_gbs']:.2f} of measured HBM copy peak "
      f"(algorithmic 16 B/pixel)")
valid = float((out.abs().sum(-1) > 0).float().mean())
for Efps, P in ((148, M * H * W), (148, 32768)):
    pts = out[:Efps, :P].contiguous()
    ms = timeit(lambda: ops.farthest_point_sample(pts, 1024), 1)
    print(f"fps: {Efps} clouds x {P} points -> 1024: {ms:.1f} ms per wave of {Efps} clouds  ({ms / Efps * 1e3:.0f} us/cloud amortised; "
          f"{Efps * P * 1023 / ms / 1e6:.1f} G point-updates/s); valid fraction of the full cloud {valid:.3f}")
