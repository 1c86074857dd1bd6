"""Golden fixture for the TSDF-integration oracle (runs only in the build container): the UNMODIFIED reference
TSDFVolume.register_camera + integrate (/root/reference/utils/depth2tsdf.py:31-86) on a seeded input, `skimage` stubbed
(import-time only).  Records the per-view voxel->pixel tables register_camera builds and the fused TSDF volume.

    python tests/golden/make_golden_tsdf.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sk = types.ModuleType("skimage"); sk.measure = types.ModuleType("skimage.measure")
sys.modules["skimage"], sys.modules["skimage.measure"] = sk, sk.measure
spec = importlib.util.spec_from_file_location("ref_depth2tsdf", "/root/reference/utils/depth2tsdf.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

torch.manual_seed(11)
E, M, H, W, R = 3, 3, 36, 64, 12
fx = W / 2.0 / np.tan(np.deg2rad(69.75) / 2.0)
intr = np.array([[fx, 0, W // 2], [0, fx, H // 2], [0, 0, 1]], dtype=np.float64)


def look_at(eye):
    eye = np.asarray(eye, np.float64)
    z = -eye / np.linalg.norm(eye)
    x = np.cross([0.0, 0.0, 1.0], z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4); T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T


poses = np.stack([look_at([0.6, 0.1, 0.5]), look_at([-0.2, 0.55, 0.45]), look_at([0.05, -0.5, 0.6])])
vol = mod.TSDFVolume("cpu", size=0.5, resolution=R)
vol.register_camera(poses, intr, H, W, E)
# a smooth surface near the workspace so that many voxels fall inside the truncation band, plus holes (0) and background (100)
v, u = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
depth = 0.62 + 0.12 * torch.sin(3 * u + torch.rand(E, M, 1, 1) * 6) * torch.cos(2 * v) + 0.02 * torch.rand(E, M, H, W)
depth[torch.rand(E, M, H, W) < 0.05] = 0.0
depth[torch.rand(E, M, H, W) < 0.05] = 100.0
tsdf = vol.integrate(depth.float())
np.savez_compressed(os.path.join(HERE, "tsdf_small.npz"), depth=depth.numpy().astype(np.float32), cam_intr=intr, cam_pose=poses,
                    vol_origin=np.asarray([-0.25, -0.25, -0.0503]), size=np.float64(0.5), resolution=np.int64(R),
                    pix_x=vol.valid_pix_x.numpy().astype(np.int32), pix_y=vol.valid_pix_y.numpy().astype(np.int32),
                    pix_z=vol.pix_z.numpy().astype(np.float32), valid_pix=vol.valid_pix.numpy(), tsdf=tsdf.numpy().astype(np.float32))
t = tsdf.numpy()
print("tsdf", t.shape, "in band", int(((t < 1) & (t > -1)).sum()), "default", int((t == 1).sum()), "min", float(t.min()))
