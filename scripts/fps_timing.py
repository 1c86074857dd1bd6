"""Dev tool: phase stamps of the cluster samplers (build with `make EXTRA=-DPM_FPS_TIMING`): clock64 at
0 pick start | 1 scan / block updates done | 2 warp candidate stored | 3 after __syncthreads | 4 candidate pushed (warp 0) |
5 eight candidates landed | 6 winner known — for pick 300 of cloud 0, thread 0 (warp 0) and thread 480 (warp 15) of every CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from partmanip_b200 import ops
from partmanip_b200._lib import lib

dev = "cuda:0"
M, H, W = 3, 288, 512


def align(x, a=256):
    return (x + a - 1) // a * a


def smooth_scene_depth(E, seed=5):
    g = torch.Generator().manual_seed(seed)
    v, u = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    ph = torch.rand(E, M, 4, generator=g) * 6.28
    d = 0.55 + 0.1 * u + 0.05 * v + 0.03 * torch.sin(5 * u + ph[..., 0, None, None]) * torch.cos(4 * v + ph[..., 1, None, None])
    box = ((u - 0.2 * torch.cos(ph[..., 2, None, None])).abs() < 0.25) & ((v - 0.2 * torch.sin(ph[..., 3, None, None])).abs() < 0.2)
    return torch.where(box, d - 0.12, d).float().contiguous()


intr = np.array([[366.0, 0, W // 2], [0, 366.0, H // 2], [0, 0, 1]])
pose = torch.eye(4, device=dev).repeat(M, 1, 1).contiguous()
pose[:, 2, 3] = -0.3
pose[1, 0, 3] = 0.05
pose[2, 1, 3] = -0.05
E = 4
smooth = ops.depth2pc_backproject(smooth_scene_depth(E).to(dev), intr, pose, [-0.25, -0.25, -0.0503], 0.5)
rand = torch.rand(E, M * H * W, 3, device=dev) * 2 - 1
rand[torch.rand(E, M * H * W, device=dev) < 0.47] = 0.0
for name, pts in (("random 53% valid", rand), ("smooth scene", smooth), ("smooth scene, 0.45 m box", None)):
    if pts is None:
        pts = ops.depth2pc_backproject(smooth_scene_depth(E).to(dev), intr, pose, [-0.225, -0.225, -0.0503], 0.45)
    P = pts.shape[1]
    for mode in (2, 4):
        ops.farthest_point_sample(pts, 512, compact=mode)
        torch.cuda.synchronize()
        ws = ops.scratch(lib.pm_fps_ws_bytes(E, P), pts.device, "fps")
        off = align(E * (P + 4) * 12) + align(E * (P + 4) * 4) + align(E * 4)
        st = ws[off:off + 8 * 2 * 8 * 8].view(torch.int64).view(8, 2, 8).cpu()
        print(f"{name} (valid {float((pts.abs().sum(-1) > 0).float().mean()):.2f}) mode {mode}: cycles per phase "
              f"(scan, warp candidate, syncthreads, reduce+push, wait, final)")
        for r in (0, 3, 7):
            for w in range(2):
                d = (st[r, w, 1:7] - st[r, w, 0:6]).tolist()
                print(f"  rank {r} {'warp0 ' if w == 0 else 'warp15'}: {d}  total {int(st[r, w, 6] - st[r, w, 0])}")
