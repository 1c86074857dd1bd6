"""Dev tool: one launch of the pruned cluster sampler on 15 smooth-scene clouds (for `ncu -k regex:fps_cluster_pruned`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from partmanip_b200 import ops

dev = "cuda:0"
M, H, W = 3, 288, 512
g = torch.Generator().manual_seed(5)
v, u = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
ph = torch.rand(15, M, 4, generator=g) * 6.28
d = 0.55 + 0.1 * u + 0.05 * v + 0.03 * torch.sin(5 * u + ph[..., 0, None, None]) * torch.cos(4 * v + ph[..., 1, None, None])
box = ((u - 0.2 * torch.cos(ph[..., 2, None, None])).abs() < 0.25) & ((v - 0.2 * torch.sin(ph[..., 3, None, None])).abs() < 0.2)
depth = torch.where(box, d - 0.12, d).float().contiguous().to(dev)
intr = np.array([[366.0, 0, W // 2], [0, 366.0, H // 2], [0, 0, 1]])
pose = torch.eye(4, device=dev).repeat(M, 1, 1).contiguous()
pose[:, 2, 3] = -0.3
pose[1, 0, 3] = 0.05
pose[2, 1, 3] = -0.05
cloud = ops.depth2pc_backproject(depth, intr, pose, [-0.25, -0.25, -0.0503], 0.5)
for _ in range(2):
    ops.farthest_point_sample(cloud, 1024)
torch.cuda.synchronize()
