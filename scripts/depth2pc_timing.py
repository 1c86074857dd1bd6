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
for Efps, P, mode in ((148, M * H * W, 3), (148, M * H * W, 2), (16, M * H * W, 2), (148, 32768, 3), (148, 32768, 2)):
    pts = out[:Efps, :P].contiguous()
    ms1 = timeit(lambda: ops.farthest_point_sample(pts, 1, compact=mode), 1)
    ms = timeit(lambda: ops.farthest_point_sample(pts, 1024, compact=mode), 1)
    print(f"[compact={mode}: {'8-CTA cluster per cloud' if mode == 2 else 'one CTA per cloud'}; compaction alone {ms1:.2f} ms] ", end="")
    print(f"fps: {Efps} clouds x {P} points -> 1024: {ms:.1f} ms per wave of {Efps} clouds  ({ms / Efps * 1e3:.0f} us/cloud amortised; "
          f"{Efps * P * 1023 / ms / 1e6:.1f} G point-updates/s); valid fraction of the full cloud {valid:.3f}")

from partmanip_b200._lib import lib
print("cluster sampler: max co-resident clusters (clouds) =", lib.pm_fps_cluster_max_active())


def smooth_scene_depth(E, seed=5):
    g = torch.Generator().manual_seed(seed)
    v, u = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    ph = torch.rand(E, M, 4, generator=g) * 6.28
    d = 0.55 + 0.1 * u + 0.05 * v + 0.03 * torch.sin(5 * u + ph[..., 0, None, None]) * torch.cos(4 * v + ph[..., 1, None, None])
    box = ((u - 0.2 * torch.cos(ph[..., 2, None, None])).abs() < 0.25) & ((v - 0.2 * torch.sin(ph[..., 3, None, None])).abs() < 0.2)
    return torch.where(box, d - 0.12, d).float().contiguous()


Es = 150
pose2 = pose.clone(); pose2[1, 0, 3] = 0.05; pose2[2, 1, 3] = -0.05
smooth = ops.depth2pc_backproject(smooth_scene_depth(Es).to(dev), intr, pose2, [-0.25, -0.25, -0.0503], 0.5)
print(f"smooth-scene clouds (tilted plane + bumps + box, 3 views): valid fraction {float((smooth.abs().sum(-1) > 0).float().mean()):.3f}")
for name, src in (("random-depth clouds (no spatial coherence: pruning worst case)", out), ("smooth-scene clouds", smooth)):
    for mode, label in ((3, "one CTA per cloud"), (2, "8-CTA cluster, streaming"), (4, "8-CTA cluster, exact bbox pruning")):
        for Efps in (15, 150):
            pts = src[:Efps]
            ms = min(timeit(lambda: ops.farthest_point_sample(pts, 1024, compact=mode), 1) for _ in range(2))
            print(f"  {name}: {label}: {Efps} clouds x {M * H * W} -> 1024: {ms:.2f} ms  ({ms / Efps * 1e3:.0f} us/cloud)")
