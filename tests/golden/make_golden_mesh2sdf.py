"""Golden fixture for the mesh -> TSDF query (runs only in the build container): the UNMODIFIED reference methods
TSDFfromMesh.merge_sdf_field / query_tsdf_parallel / triplet_interpolation_query_parallel
(/root/reference/utils/mesh2sdf.py:119-139, 169-198, 239-272) on seeded inputs.  `trimesh` / `skimage` are stubbed (import-time
only); the object is built with __new__ and the attributes __init__ (mesh2sdf.py:16-37) sets, because __init__ itself loads mesh
files that are not needed here: the per-part SDF grids are analytic (spheres / boxes of different resolutions).

    python tests/golden/make_golden_mesh2sdf.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ("trimesh", "skimage", "skimage.measure"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["skimage"].measure = sys.modules["skimage.measure"]
spec = importlib.util.spec_from_file_location("ref_mesh2sdf", "/root/reference/utils/mesh2sdf.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)


def sdf_grid(shape, voxel, kind, rng):
    """an analytic signed-distance grid in the part frame, clamped like the reference's pre-stored fields (mesh2sdf.py:231)"""
    nx, ny, nz = shape
    bbox_min = -0.5 * voxel * np.array(shape) + rng.uniform(-0.004, 0.004, 3)
    gx, gy, gz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    p = np.stack([gx, gy, gz], -1) * voxel + bbox_min
    if kind == "sphere":
        d = np.linalg.norm(p, axis=-1) - 0.3 * voxel * min(shape)
    else:
        q = np.abs(p) - 0.25 * voxel * np.array(shape)
        d = np.linalg.norm(np.maximum(q, 0), axis=-1) + np.minimum(q.max(-1), 0)
    return dict(sdf=np.clip(d, -0.04, 0.04).astype(np.float32), bbox_min=bbox_min.astype(np.float32), voxel_size=np.float32(voxel))


def main():
    rng = np.random.default_rng(5)
    torch.manual_seed(5)
    E, R, size = 3, 16, 0.5
    obj = mod.TSDFfromMesh.__new__(mod.TSDFfromMesh)
    obj.num_envs, obj.parallel, obj.device, obj.debug = E, True, "cpu", False
    obj.resolution, obj.size = R, size
    obj.vox_size = size / R                                                     # mesh2sdf.py:24-26
    obj.sdf_trunc = 4 * obj.vox_size
    obj.vox_origin = torch.tensor([-0.25, -0.25, -0.0503])
    tmp = torch.arange(0, R)
    xv, yv, zv = torch.meshgrid(tmp, tmp, tmp, indexing="ij")
    vox = torch.stack([xv.flatten(), yv.flatten(), zv.flatten()], dim=1).long()
    obj.vox_coords = vox * obj.vox_size + obj.vox_origin                        # mesh2sdf.py:29-33
    obj.point_num = obj.vox_coords.shape[0]
    obj.init_tsdf = obj.vox_coords[:, -1].clone().unsqueeze(0).repeat(E, 1)     # ground plane (mesh2sdf.py:36)
    parts = [sdf_grid((40, 36, 50), 0.004, "box", rng), sdf_grid((30, 30, 30), 0.005, "sphere", rng),
             sdf_grid((24, 52, 28), 0.003, "box", rng), sdf_grid((20, 20, 44), 0.006, "sphere", rng)]
    obj.sdf_dict_list = parts
    obj.merge_sdf_field()
    M = len(parts)
    # part poses: rotations (world <- part) and translations inside the workspace
    A = torch.randn(E, M, 3, 3)
    Q, _ = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.linalg.det(Q))[..., None, None]
    T = torch.stack([torch.rand(E, M) * 0.3 - 0.15, torch.rand(E, M) * 0.3 - 0.15, torch.rand(E, M) * 0.25 + 0.05], dim=-1)
    tsdf = obj.query_tsdf_parallel(Q, T)
    out = dict(tsdf=tsdf.numpy().astype(np.float32), pose_R=Q.numpy().astype(np.float32), pose_T=T.numpy().astype(np.float32),
               init_tsdf=obj.init_tsdf.numpy().astype(np.float32), vox_origin=obj.vox_origin.numpy().astype(np.float32),
               size=np.float64(size), resolution=np.int64(R), sdf_field=obj.sdf_field.numpy().astype(np.float32),
               sdf_field_res=obj.sdf_field_res.reshape(M, 3).numpy().astype(np.int64),
               sdf_voxel_size=obj.sdf_voxel_size.reshape(M).numpy().astype(np.float32),
               sdf_bbox_min=obj.sdf_bbox_min.reshape(M, 3).numpy().astype(np.float32),
               bbox_res=np.array([int(obj.sdf_field_res.reshape(M, 3).max(0)[0][0]), int(obj.bboxResy), int(obj.bboxResz)], np.int64))
    for i, p in enumerate(parts):
        out[f"part{i}_sdf"], out[f"part{i}_bbox_min"], out[f"part{i}_voxel_size"] = p["sdf"], p["bbox_min"], p["voxel_size"]
    np.savez_compressed(os.path.join(HERE, "mesh2sdf_small.npz"), **out)
    t = out["tsdf"]
    print("wrote mesh2sdf_small.npz", t.shape, "frac inside band", float((np.abs(t) < 1).mean()), "min", float(t.min()))


if __name__ == "__main__":
    main()
