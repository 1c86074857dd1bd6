"""`TSDFfromMesh` — the query half of the reference class (utils/mesh2sdf.py:15-139, 169-198, 239-272): the scene's TSDF volume from
pre-computed signed-distance grids of the robot links / objects and their current poses (`mesh_tsdf` observations of
dagger_tsdf.yaml).  `query_tsdf(pose_R, pose_T)` / `query_tsdf_parallel` run as ONE kernel (pm_mesh2sdf_query) instead of the
reference's chain of (b, m, n, .) temporaries; `merge_sdf_field` is the same one-time padding / stacking of the part grids.

What is NOT mirrored: turning meshes into grids (`mesh2sdf`, `load_franka`, `preprocess_mesh`: trimesh / kaolin / ManifoldPlus,
offline asset preparation) and the marching-cubes debug dumps — pass the grids in with `add_sdf(sdf_dict)` (the dicts `load_sdf`
reads from `sdf.npy`: {'sdf': (X,Y,Z) array, 'bbox_min': (3,), 'voxel_size': float})."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class TSDFfromMesh:
    def __init__(self, num_envs, size, resolution, device, parallel=True, debug=False, vox_origin=None):
        if not str(device).startswith("cuda"):
            raise RuntimeError("partmanip_b200 runs on CUDA devices only (no CPU fallback); got device=%r" % (device,))
        if debug:
            raise NotImplementedError("debug dumps (marching cubes via skimage) are not mirrored")
        self.num_envs, self.parallel, self.device, self.debug = num_envs, parallel, device, debug
        self.resolution, self.size = resolution, size
        self.vox_size = self.size / self.resolution                                       # mesh2sdf.py:24-25
        self.sdf_trunc = 4 * self.vox_size
        self.vox_origin = [-0.25, -0.25, -0.0503] if vox_origin is None else [float(v) for v in vox_origin]
        self.point_num = resolution ** 3
        # the ground plane: every voxel's height (mesh2sdf.py:36), float32 index * vox + origin like the reference's vox_coords
        z = (torch.arange(resolution, dtype=torch.float32) * torch.tensor(self.vox_size, dtype=torch.float32)
             + torch.tensor(self.vox_origin[2], dtype=torch.float32))
        self.init_tsdf = z.repeat(resolution * resolution).unsqueeze(0).repeat(num_envs, 1).to(device).contiguous()
        self.ground_tsdf = self.init_tsdf.clone()
        self.sdf_dict_list = []

    def initialize_sdf(self, nerf_pred_tsdf):
        """mesh2sdf.py:57-61: start from a predicted (normalised) volume instead of the ground plane."""
        self.init_tsdf = (torch.as_tensor(nerf_pred_tsdf, device=self.device, dtype=torch.float32) * self.sdf_trunc).reshape(self.num_envs, -1).contiguous()

    def add_sdf(self, sdf_dict):
        """what load_sdf appends after reading / computing a part's grid (mesh2sdf.py:64-82)"""
        self.sdf_dict_list.append(sdf_dict)

    def merge_sdf_field(self):
        """mesh2sdf.py:169-198: pad every grid with +1 to the common resolution and stack."""
        parts = self.sdf_dict_list
        self.part_num = len(parts)
        res = np.array([p['sdf'].shape for p in parts], np.int64)
        tgt = res.max(0)
        field = np.ones((self.part_num, *tgt), np.float32)
        for i, (p, r) in enumerate(zip(parts, res)):
            field[i, :r[0], :r[1], :r[2]] = p['sdf']
        dev = self.device
        self.sdf_field = torch.from_numpy(field.reshape(self.part_num, -1)).to(dev).contiguous()
        self.sdf_field_res = torch.from_numpy(res.astype(np.int32)).to(dev).contiguous()
        self.sdf_voxel_size = torch.tensor([float(p['voxel_size']) for p in parts], dtype=torch.float32, device=dev)
        self.sdf_bbox_min = torch.from_numpy(np.stack([np.asarray(p['bbox_min'], np.float32) for p in parts])).to(dev).contiguous()
        self.bboxResy, self.bboxResz = int(tgt[1]), int(tgt[2])

    def query_tsdf_parallel(self, pose_R, pose_T):
        """mesh2sdf.py:119-139.  pose_R (b, m, 3, 3), pose_T (b, m, 3) -> (b, R, R, R) in [-1, 1]."""
        assert tuple(pose_R.shape) == (self.num_envs, self.part_num, 3, 3) and tuple(pose_T.shape) == (self.num_envs, self.part_num, 3)
        return ops.mesh2sdf_query(self.sdf_field, self.sdf_field_res, self.sdf_voxel_size, self.sdf_bbox_min, self.bboxResy, self.bboxResz,
                                  pose_R.float().contiguous(), pose_T.float().contiguous(), self.init_tsdf, self.resolution,
                                  self.vox_origin, self.size)

    def query_tsdf(self, pose_R, pose_T):
        """mesh2sdf.py:84-88 (the naive per-part loop computes the same volume; both run the one kernel here)."""
        return self.query_tsdf_parallel(pose_R, pose_T)
