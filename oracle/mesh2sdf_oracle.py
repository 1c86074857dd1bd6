"""TEST INFRASTRUCTURE ONLY — CPU oracle (numpy, float32) of the "next" row §8f(3), third part: TSDF of the scene from per-part
signed-distance grids (`mesh_tsdf` observations of dagger_tsdf.yaml).  Restates /root/reference/utils/mesh2sdf.py:
  merge_sdf_field                       169-198   (pad every part's grid with +1 to the common resolution, flatten)
  query_tsdf_parallel                   119-139   (voxel centres into each part's frame, interpolate, min over parts and the initial
                                                   volume, divide by the truncation distance, clamp to [-1, 1])
  triplet_interpolation_query_parallel  239-272   (trilinear lookup; queries outside [1, res - 2] of a part's own grid give +1)
PINNED: tests/golden/mesh2sdf_small.npz records the UNMODIFIED reference methods on seeded inputs (tests/golden/make_golden_mesh2sdf.py);
tests/test_oracle_mesh2sdf.py replays it.  Only tests/ may import this module."""
from __future__ import annotations

import numpy as np


def voxel_centres(size: float, resolution: int, vox_origin) -> np.ndarray:
    """mesh2sdf.py:24-33: (R^3, 3) float32 centres, x slowest (torch.meshgrid 'ij')."""
    vox = np.float32(size / resolution)
    g = np.arange(resolution)
    xv, yv, zv = np.meshgrid(g, g, g, indexing="ij")
    c = np.stack([xv.ravel(), yv.ravel(), zv.ravel()], 1).astype(np.float32)
    return (c * vox + np.asarray(vox_origin, np.float32)).astype(np.float32)


def merge_sdf_field(parts):
    """mesh2sdf.py:169-198.  parts: list of dicts {sdf (X,Y,Z), voxel_size, bbox_min}.  -> field (M, Xm*Ym*Zm), res (M,3), voxel (M,),
    bbox_min (M,3), (Xm, Ym, Zm)."""
    res = np.array([p["sdf"].shape for p in parts], np.int64)
    tgt = res.max(0)
    field = []
    for p, r in zip(parts, res):
        f = np.ones(tuple(tgt), np.float32)                                    # F.pad(..., "constant", 1)
        f[:r[0], :r[1], :r[2]] = p["sdf"]
        field.append(f.reshape(-1))
    return (np.stack(field), res, np.array([p["voxel_size"] for p in parts], np.float32),
            np.stack([np.asarray(p["bbox_min"], np.float32) for p in parts]), tuple(int(v) for v in tgt))


def trilinear(field, res, voxel, bbox_min, bbox_res, q):
    """mesh2sdf.py:239-272.  q (b, m, n, 3) part-frame points -> (b, m, n)."""
    _, ry, rz = bbox_res
    qi = ((q - bbox_min[None, :, None, :]) / voxel[None, :, None, None]).astype(np.float32)
    valid3 = (qi >= 1) & (qi - res[None, :, None, :].astype(np.float32) <= -2)
    qi = qi * valid3
    valid = valid3.sum(-1) == 3
    li = qi.astype(np.int64)                                                   # .long(): truncation (values are >= 0)
    d = (qi - li).astype(np.float32)
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    i000 = (li[..., 0] * ry + li[..., 1]) * rz + li[..., 2]
    m = np.arange(field.shape[0])[None, :, None]
    f = lambda off: field[m, i000 + off]
    one = np.float32(1)
    v = ((f(0) * (one - z) + f(1) * z) * (one - y) + (f(rz) * (one - z) + f(rz + 1) * z) * y) * (one - x) \
        + ((f(rz * ry) * (one - z) + f(rz * ry + 1) * z) * (one - y) + (f(rz * ry + rz) * (one - z) + f(rz * ry + rz + 1) * z) * y) * x
    return (v * valid + one * (~valid)).astype(np.float32)


def query_tsdf(field, res, voxel, bbox_min, bbox_res, centres, init_tsdf, sdf_trunc, pose_R, pose_T):
    """mesh2sdf.py:119-139.  pose_R (b, m, 3, 3), pose_T (b, m, 3), init_tsdf (b, n) -> (b, R, R, R) in [-1, 1]."""
    b, m = pose_R.shape[:2]
    q = np.einsum("bmnk,bmkj->bmnj", (centres[None, None] - pose_T[:, :, None, :]).astype(np.float32), pose_R.astype(np.float32)).astype(np.float32)
    upd = trilinear(field, res, voxel, bbox_min, bbox_res, q)
    t = np.minimum(upd.min(1), init_tsdf.astype(np.float32))
    t = np.clip(t / np.float32(sdf_trunc), -1, 1).astype(np.float32)
    R = round(centres.shape[0] ** (1 / 3))
    return t.reshape(b, R, R, R)
