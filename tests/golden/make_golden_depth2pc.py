"""Golden fixture for the next-row oracle (runs only in the build container): the UNMODIFIED reference TSDFVolume.depth2pc
(/root/reference/utils/depth2tsdf.py:136-173) on a seeded input, with `skimage` stubbed (import-time only) and
`pytorch3d.ops.sample_farthest_points` replaced by a stub that CAPTURES the masked world cloud it is handed.

    python tests/golden/make_golden_depth2pc.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
captured = {}


def _stub_fps(points, K=50, **kw):
    captured["cloud"] = points.detach().clone()
    return points[:, :K], torch.arange(K)[None].repeat(points.shape[0], 1)


sk = types.ModuleType("skimage"); sk.measure = types.ModuleType("skimage.measure")
sys.modules["skimage"], sys.modules["skimage.measure"] = sk, sk.measure
p3d, ops = types.ModuleType("pytorch3d"), types.ModuleType("pytorch3d.ops")
ops.sample_farthest_points = _stub_fps
p3d.ops = ops
sys.modules["pytorch3d"], sys.modules["pytorch3d.ops"] = p3d, ops
spec = importlib.util.spec_from_file_location("ref_depth2tsdf", "/root/reference/utils/depth2tsdf.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

torch.manual_seed(7)
E, M, H, W = 2, 2, 24, 32
fx = W / 2.0 / np.tan(np.deg2rad(69.75) / 2.0)
intr = np.array([[fx, 0, W // 2], [0, fx, H // 2], [0, 0, 1]], dtype=np.float64)


def look_at(eye):
    eye = np.asarray(eye, np.float64)
    z = -eye / np.linalg.norm(eye)
    x = np.cross([0.0, 0.0, 1.0], z); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4); T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, eye
    return T


poses = np.stack([look_at([0.6, 0.1, 0.5]), look_at([-0.2, 0.55, 0.45])])
vol = mod.TSDFVolume("cpu", size=0.5, resolution=8)
vol.register_camera(poses, intr, H, W, E)
depth = 0.45 + 0.5 * torch.rand(E, M, H, W)
out = vol.depth2pc(depth)
np.savez_compressed(os.path.join(HERE, "depth2pc_small.npz"), depth=depth.numpy(), cam_intr=intr, cam_pose=poses,
                    vol_origin=np.asarray([-0.25, -0.25, -0.0503]), size=np.float64(0.5), cloud=captured["cloud"].numpy(),
                    n_valid=np.int64((captured["cloud"].abs().sum(-1) > 0).sum()))
print("masked cloud", tuple(captured["cloud"].shape), "valid points", int((captured["cloud"].abs().sum(-1) > 0).sum()))
