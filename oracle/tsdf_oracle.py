"""TEST INFRASTRUCTURE ONLY — CPU oracle of the "next" row §8f(3), first half: depth images -> fused TSDF volume.

Restates /root/reference/utils/depth2tsdf.py with numpy, float32 throughout:
  * `voxel_pixel_tables` = the per-view voxel -> pixel tables of `TSDFVolume.__init__` + `register_camera` (:14-62);
  * `integrate` = `TSDFVolume.integrate` (:68-86): gather one depth pixel per (view, voxel), truncated signed distance,
    equal-weight average over the views that see the voxel in front of / inside the band, 1 where none does.
  * `sparse_voxel` = the rest of `TSDFVolume.sparse_voxel` (:103-119): band select in torch.where order, pytorch3d farthest-point
    sampling on the integer voxel coordinates (restated in oracle/depth2pc_oracle.py — pytorch3d is absent: that step is
    PARITY UNPINNED), gather of the TSDF values.
PINNED (tables and fusion): tests/golden/tsdf_small.npz holds the tables and the volume the UNMODIFIED reference computes for a seeded input
(tests/golden/make_golden_tsdf.py).  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np


def voxel_pixel_tables(cam_pose, cam_intr, im_h: int, im_w: int, size: float, resolution: int, vol_origin):
    """-> pix_x, pix_y (M, R^3) int32 (0 where invalid), pix_z (M, R^3) fp32, valid (M, R^3) bool.  Voxel v = x*R*R + y*R + z
    (torch.meshgrid 'ij' order, depth2tsdf.py:22-23)."""
    R = int(resolution)
    voxel_size = size / R                                              # Python float, as in the reference (:16)
    h = np.arange(R)
    xv, yv, zv = np.meshgrid(h, h, h, indexing="ij")
    vox = np.stack([xv.ravel(), yv.ravel(), zv.ravel()], axis=1)
    world = np.asarray(vol_origin, np.float32) + (np.float32(voxel_size) * vox.astype(np.float32)).astype(np.float32)   # :27
    pose = np.asarray(cam_pose, np.float32)
    d = (world[None] - pose[:, None, :3, 3]).astype(np.float32)                                  # (M, R^3, 3)
    Rm = pose[:, :3, :3]
    cam = ((d[..., 0:1] * Rm[:, None, 0, :] + d[..., 1:2] * Rm[:, None, 1, :]).astype(np.float32)
           + d[..., 2:3] * Rm[:, None, 2, :]).astype(np.float32)                                 # bmm(world - t, R)  (:47)
    fx, fy = np.float32(cam_intr[0][0]), np.float32(cam_intr[1][1])
    cx, cy = np.float32(cam_intr[0][2]), np.float32(cam_intr[1][2])
    pz = cam[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        px = np.rint((cam[..., 0] * fx / pz).astype(np.float32) + cx)                            # torch.round = half to even
        py = np.rint((cam[..., 1] * fy / pz).astype(np.float32) + cy)
    px = np.nan_to_num(px, nan=-1.0, posinf=2.0 ** 40, neginf=-2.0 ** 40).astype(np.int64)
    py = np.nan_to_num(py, nan=-1.0, posinf=2.0 ** 40, neginf=-2.0 ** 40).astype(np.int64)
    valid = (px >= 0) & (px < im_w) & (py >= 0) & (py < im_h) & (pz > 0)
    return np.where(valid, px, 0).astype(np.int32), np.where(valid, py, 0).astype(np.int32), pz.astype(np.float32), valid


def integrate(depth, pix_x, pix_y, pix_z, valid, size: float, resolution: int, default_tsdf: float = 1.0):
    """depth (E, M, H, W) fp32 -> (E, R, R, R) fp32 (depth2tsdf.py:68-86)."""
    depth = np.asarray(depth, np.float32)
    E, M = depth.shape[:2]
    R = int(resolution)
    trunc = np.float32(4 * (size / R))                                 # :17 (Python float product, used as an fp32 scalar)
    m_idx = np.arange(M)[:, None]
    dv = depth[:, m_idx, pix_y, pix_x]                                 # (E, M, R^3)
    diff = (dv - pix_z[None]).astype(np.float32)
    tsdf = np.minimum((diff / trunc).astype(np.float32), np.float32(1))
    vp = valid[None] & (dv > 0) & (diff >= -trunc)
    cnt = vp.sum(1).astype(np.float32)                                 # (E, R^3)
    with np.errstate(divide="ignore"):
        w = np.where(vp, (np.float32(1) / cnt)[:, None, :], np.float32(0)).astype(np.float32)
    prod = (tsdf * w).astype(np.float32)
    acc = prod[:, 0]
    for m in range(1, M):
        acc = (acc + prod[:, m]).astype(np.float32)
    vol = (acc + np.float32(default_tsdf) * (cnt == 0)).astype(np.float32)
    return vol.reshape(E, R, R, R)


def sparse_voxel(tsdf_vol, K: int = 1024, lo: float = -0.2, hi: float = 0.2):
    """tsdf_vol (E, R, R, R) fp32 -> (E, K, 4) fp32 = (x, y, z, tsdf) of K farthest-point-sampled band voxels (:103-119)."""
    from .depth2pc_oracle import farthest_point_sample
    vol = np.asarray(tsdf_vol, np.float32)
    out = np.zeros((vol.shape[0], K, 4), np.float32)
    for e in range(vol.shape[0]):
        ind = np.argwhere((vol[e] < np.float32(hi)) & (vol[e] > np.float32(lo)))                 # row-major, as torch.where
        if len(ind) == 0:                                              # the reference fails on an empty band; the kernel returns voxel 0
            ind = np.zeros((1, 3), np.int64)
        _, idx = farthest_point_sample(ind[None].astype(np.float32), K)
        sel = ind[idx[0]]
        if len(sel) < K:                                               # fewer band voxels than K: the samplers repeat the first one
            sel = np.concatenate([sel, np.repeat(sel[:1], K - len(sel), axis=0)])
        out[e, :, :3] = sel
        out[e, :, 3] = vol[e][sel[:, 0], sel[:, 1], sel[:, 2]]
    return out
