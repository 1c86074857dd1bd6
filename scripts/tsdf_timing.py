"""Dev tool: CUDA-event timing of the TSDF observation kernels at the reference's shapes (3 views of 288x512 depth, 50^3 voxels)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from partmanip_b200 import ops

dev = "cuda:0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hbm = float(json.load(open(os.path.join(root, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6550.0)) if os.path.exists(os.path.join(root, "MEASURED_PEAKS.json")) else 6550.0
M, H, W, R = 3, 288, 512, 50


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def look_at(eye):
    eye = np.asarray(eye, np.float64)
    z = -eye / np.linalg.norm(eye)
    x = np.cross([0.0, 0.0, 1.0], z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4); T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T


fx = W / 2.0 / np.tan(np.deg2rad(69.75) / 2.0)
intr = np.array([[fx, 0, W // 2], [0, fx, H // 2], [0, 0, 1]])
poses = np.stack([look_at([0.6, 0.1, 0.5]), look_at([-0.2, 0.55, 0.45]), look_at([0.05, -0.5, 0.6])])
pose_d = torch.from_numpy(poses).float().to(dev).contiguous()
pix_off, pix_z = ops.tsdf_voxel_tables(pose_d, intr, H, W, 0.5, R, [-0.25, -0.25, -0.0503])
print(f"tables: {float((pix_off >= 0).float().mean()):.2f} of the (view, voxel) pairs project into the image")
for E in (256, 1024):
    g = torch.Generator(device=dev).manual_seed(E)
    v, u = torch.meshgrid(torch.linspace(-1, 1, H, device=dev), torch.linspace(-1, 1, W, device=dev), indexing="ij")
    depth = (0.62 + 0.12 * torch.sin(3 * u + torch.rand(E, M, 1, 1, device=dev, generator=g) * 6) * torch.cos(2 * v)
             + 0.02 * torch.rand(E, M, H, W, device=dev, generator=g)).contiguous()
    out = torch.empty(E, R, R, R, device=dev)
    ms = timeit(lambda: ops.tsdf_integrate(depth, pix_off, pix_z, 0.5, R, out=out))
    wr, rd = E * R ** 3 * 4 / 1e6, E * M * H * W * 4 / 1e6
    print(f"integrate: E={E}: {ms:.3f} ms ({ms / E * 1e3:.2f} us/env); writes {wr:.0f} MB, gathers from {rd:.0f} MB of depth: "
          f"{wr / ms:.0f} GB/s written, <= {(wr + rd) / ms:.0f} GB/s total = {wr / ms / hbm:.2f} .. {(wr + rd) / ms / hbm:.2f} of the HBM copy peak")
    band = float(((out < 0.2) & (out > -0.2)).float().mean())
    ms = timeit(lambda: ops.tsdf_sparse_voxel(out, 1024), 3)
    print(f"sparse_voxel (band compaction + 1024 farthest voxels + gather): E={E}: {ms:.2f} ms ({ms / E * 1e3:.1f} us/env); band = {band * R ** 3:.0f} voxels per env")
