"""`TSDFVolume` — the point-cloud half of the reference class (utils/depth2tsdf.py:6-66, 136-173): same constructor,
`register_camera(cam_pose, cam_intr, im_h, im_w, num_env)` and `depth2pc(depth_im) -> (num_env, 1024, 3)`, with the
back-projection / workspace mask and the farthest-point sampling (pytorch3d there) running in libpartmanip_b200.so.
`depth2pc_from_views(camera_tensor_list)` additionally folds in the caller's stacking step (tasks/hand_base.py:317-324 +
:333): the simulator's per-env, per-view camera tensors are read in place.
`integrate(depth_im) -> (num_env, R, R, R)` (depth2tsdf.py:68-86) runs on the voxel -> pixel tables `register_camera` builds with
`pm_tsdf_voxel_tables` (the reference's `valid_pix*` / `pix_z` tensors, packed); `sparse_voxel(depth_im) -> (num_env, 1024, 4)`
(depth2tsdf.py:88-120, the `depth_sparse` observation) fuses, selects the |tsdf| < 0.2 band and farthest-point-samples it.
`extract_point_cloud` (CPU marching cubes via skimage, debugging only) is not mirrored.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


def env_chunks(num_env: int, chunk: int):
    """[lo, hi) env ranges of at most `chunk` envs covering 0..num_env in order."""
    chunk = max(1, int(chunk))
    return [(lo, min(lo + chunk, num_env)) for lo in range(0, num_env, chunk)]


class TSDFVolume(object):
    def __init__(self, device, size=0.5, resolution=50, _vol_origin=(-0.25, -0.25, -0.0503)):
        if not str(device).startswith("cuda"):
            raise RuntimeError("partmanip_b200 runs on CUDA devices only (no CPU fallback); got device=%r" % (device,))
        self._size = size
        self._resolution = resolution
        self._voxel_size = self._size / self._resolution
        self.device = device
        self._vol_origin = [float(v) for v in _vol_origin]
        self.num_points = 1024                                  # depth2tsdf.py:160 hard-codes K=1024
        # depth2pc materialises the full-resolution cloud (m*h*w points per env: 5.3 MB at 3 x 288 x 512) and the samplers'
        # workspace (20 B per point) before reducing it to 1024 points; envs are processed this many at a time so that 4096
        # envs need 3.7 GB of transient memory instead of 58 GB.  The result does not depend on it (clouds are independent).
        self.env_chunk = 256

    def register_camera(self, cam_pose, cam_intr, im_h, im_w, num_env):
        """depth2tsdf.py:31-66 (the camera bookkeeping depth2pc needs)."""
        cam_pose = np.asarray(cam_pose, dtype=np.float32)
        self.registered_shape = (num_env, cam_pose.shape[0], im_h, im_w)
        self.cam_pose = torch.tensor(cam_pose, device=self.device).float().contiguous()     # (m, 4, 4); identical for every env
        self.cam_intr = np.asarray(cam_intr, dtype=np.float64)
        self.pix_off, self.pix_z = ops.tsdf_voxel_tables(self.cam_pose, self.cam_intr, im_h, im_w, self._size, self._resolution,
                                                         self._vol_origin)
        self.default_tsdf = 1

    def integrate(self, depth_im):
        """depth2tsdf.py:68-86: depth_im (b, m, h, w) -> fused TSDF volume (b, R, R, R); also kept as `_tsdf_vol` like the reference."""
        assert tuple(depth_im.shape) == tuple(self.registered_shape)
        self._tsdf_vol = ops.tsdf_integrate(depth_im.float().contiguous(), self.pix_off, self.pix_z, self._size, self._resolution,
                                            float(self.default_tsdf))
        return self._tsdf_vol

    def sparse_voxel(self, depth_im):
        """depth2tsdf.py:88-120: (b, m, h, w) -> (b, 1024, 4) = 1024 farthest band voxels (x, y, z, tsdf).  Unlike `integrate`, the
        reference does not keep the volume in `_tsdf_vol` here."""
        assert tuple(depth_im.shape) == tuple(self.registered_shape)
        vol = ops.tsdf_integrate(depth_im.float().contiguous(), self.pix_off, self.pix_z, self._size, self._resolution,
                                 float(self.default_tsdf))
        return ops.tsdf_sparse_voxel(vol, self.num_points, -0.2, 0.2)

    def depth2pc(self, depth_im):
        """depth2tsdf.py:136-173: depth_im (b, m, h, w) -> (b, 1024, 3)."""
        assert tuple(depth_im.shape) == tuple(self.registered_shape)
        depth_im = depth_im.float().contiguous()
        E = depth_im.shape[0]
        out = None
        for lo, hi in env_chunks(E, self.env_chunk):
            cloud = ops.depth2pc_backproject(depth_im[lo:hi], self.cam_intr, self.cam_pose, self._vol_origin, self._size)
            pc = ops.farthest_point_sample(cloud, self.num_points)
            if lo == 0 and hi == E:
                return pc
            if out is None:
                out = torch.empty(E, self.num_points, 3, device=pc.device, dtype=pc.dtype)
            out[lo:hi] = pc
        return out

    def depth2pc_from_views(self, camera_tensor_list):
        """tasks/hand_base.py:317-324,333 + depth2tsdf.py:136-173: `camera_tensor_list[env][view]` are the simulator's (h, w) fp32
        depth images (negative z, -inf background).  Equivalent to
        `depth2pc(where(isinf(-stack), 100, -stack))` without materialising the stack or the temporaries."""
        E, M, H, W = self.registered_shape
        assert len(camera_tensor_list) == E and all(len(v) == M for v in camera_tensor_list)
        key = tuple(t.data_ptr() for v in camera_tensor_list for t in v)
        if getattr(self, "_view_key", None) != key:             # Isaac Gym reuses its camera buffers: the table is built once
            assert all(tuple(t.shape) == (H, W) for v in camera_tensor_list for t in v)
            self._view_table, self._view_aligned = ops.view_pointer_table(camera_tensor_list)
            self._view_key = key
        out = None
        for lo, hi in env_chunks(E, self.env_chunk):
            cloud = ops.depth2pc_backproject_views(self._view_table[lo * M:hi * M], self._view_aligned, hi - lo, M, H, W, self.cam_intr,
                                                   self.cam_pose, self._vol_origin, self._size, negate=True, inf_value=100.0)
            pc = ops.farthest_point_sample(cloud, self.num_points)
            if lo == 0 and hi == E:
                return pc
            if out is None:
                out = torch.empty(E, self.num_points, 3, device=pc.device, dtype=pc.dtype)
            out[lo:hi] = pc
        return out
