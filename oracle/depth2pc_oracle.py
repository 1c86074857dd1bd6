"""TEST INFRASTRUCTURE ONLY — CPU oracle of the "next" row §8f(1): depth images -> world point cloud -> 1024 farthest points.

Restates /root/reference/utils/depth2tsdf.py:136-173 (TSDFVolume.depth2pc) with numpy, float32 throughout:
  * back-projection + camera->world transform + workspace mask (:146-159) — PINNED: tests/golden/depth2pc_small.npz holds the
    masked cloud the UNMODIFIED reference computes for a seeded input (tests/golden/make_golden_depth2pc.py runs the
    reference's own depth2pc with `pytorch3d.ops.sample_farthest_points` replaced by a capturing stub);
  * farthest-point sampling (:160) — the reference calls pytorch3d (`from pytorch3d.ops import sample_farthest_points`,
    not vendored and not installed; the reference pins no version, README mentions pytorch3d without one).  Restated from
    pytorch3d's published algorithm (pytorch3d/ops/sample_farthest_points.py, `sample_farthest_points_naive`, v0.7.x):
    start at index 0 (random_start_point=False), keep the running minimum squared distance of every point to the selected
    set, pick the arg-max next (first index on ties).  PARITY UNPINNED against pytorch3d itself; the squared distance is
    evaluated as (dx*dx + dy*dy) + dz*dz in float32 without FMA contraction, which the CUDA kernel reproduces bit for bit.
Only tests/ may import this module."""
from __future__ import annotations

import numpy as np


def pixel_maps(im_h: int, im_w: int):
    """depth2tsdf.py:64-65 — NB the reference's `xmap` holds the ROW index j and `ymap` the COLUMN index i."""
    xmap = np.repeat(np.arange(im_h), im_w).astype(np.float32)
    ymap = np.tile(np.arange(im_w), im_h).astype(np.float32)
    return xmap, ymap


def stack_views(camera_tensor_list) -> np.ndarray:
    """tasks/hand_base.py:317-324: per-env lists of per-view (H, W) depth images -> (E, M, H, W), negated, +-inf -> 100."""
    d = np.stack([np.stack([np.asarray(v, np.float32) for v in views], axis=0) for views in camera_tensor_list], axis=0)
    d = -d
    return np.where(np.isinf(d), np.float32(100), d).astype(np.float32)


def backproject(depth: np.ndarray, cam_intr: np.ndarray, cam_pose: np.ndarray, vol_origin, size: float) -> np.ndarray:
    """depth (E, M, H, W) fp32, cam_intr (3,3), cam_pose (M,4,4) -> masked world cloud (E, M*H*W, 3): points outside the
    open box (origin, origin+size) are zeroed (depth2tsdf.py:146-159)."""
    depth = np.asarray(depth, np.float32)
    E, M, H, W = depth.shape
    xmap, ymap = pixel_maps(H, W)
    cx, cy = np.float32(cam_intr[0, 2]), np.float32(cam_intr[1, 2])
    fx, fy = np.float32(cam_intr[0, 0]), np.float32(cam_intr[1, 1])
    pt2 = depth.reshape(E, M, H * W)
    pt0 = (ymap - cx) * pt2 / fx
    pt1 = (xmap - cy) * pt2 / fy
    cld = np.stack((pt0, pt1, pt2), axis=-1).astype(np.float32)                       # (E, M, HW, 3)
    R = np.asarray(cam_pose, np.float32)[:, :3, :3]
    t = np.asarray(cam_pose, np.float32)[:, :3, 3]
    world = np.einsum("emkj,mij->emki", cld, R).astype(np.float32) + t[None, :, None, :]
    world = world.reshape(E, M * H * W, 3).astype(np.float32)
    o = np.asarray(vol_origin, np.float32)
    valid = ((world < np.float32(size) + o) & (world > o)).sum(-1, keepdims=True) == 3
    return (world * valid).astype(np.float32)


def farthest_point_sample(points: np.ndarray, K: int):
    """points (E, P, 3) fp32 -> (selected points (E, K, 3), indices (E, K) int64); pytorch3d semantics, see module docstring."""
    pts = np.asarray(points, np.float32)
    E, P, _ = pts.shape
    K = min(K, P)
    idx = np.zeros((E, K), np.int64)
    for e in range(E):
        p = pts[e]
        mind = np.full(P, np.finfo(np.float32).max, np.float32)
        last = 0
        for i in range(1, K):
            d = p - p[last]
            dist = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + (d[:, 2] * d[:, 2]).astype(np.float32)
            mind = np.minimum(mind, dist.astype(np.float32))
            last = int(np.argmax(mind))                      # first index among ties
            idx[e, i] = last
    return np.take_along_axis(pts, idx[:, :, None].repeat(3, axis=2), axis=1), idx


def depth2pc(depth, cam_intr, cam_pose, vol_origin, size, K: int = 1024):
    """TSDFVolume.depth2pc (depth2tsdf.py:136-173): returns (E, K, 3)."""
    return farthest_point_sample(backproject(depth, cam_intr, cam_pose, vol_origin, size), K)[0]
