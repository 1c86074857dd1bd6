"""GPU parity of the next row §8f(3), third part: mesh -> TSDF query (utils/mesh2sdf.py:119-139, 239-272) — pm_mesh2sdf_query through
the host mirror against the recording of the UNMODIFIED reference methods and the numpy oracle at the reference's 50^3 resolution."""
import os

import numpy as np
import pytest
import torch

from oracle import mesh2sdf_oracle as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh2sdf_small.npz"))


def _parts():
    n = sum(1 for k in G.files if k.endswith("_sdf") and k.startswith("part"))
    return [dict(sdf=G[f"part{i}_sdf"], bbox_min=G[f"part{i}_bbox_min"], voxel_size=G[f"part{i}_voxel_size"]) for i in range(n)]


def _volume(E, R, parts, init=None):
    from partmanip_b200.utils.mesh2sdf import TSDFfromMesh
    vol = TSDFfromMesh(E, 0.5, R, DEV)
    for p in parts:
        vol.add_sdf(p)
    vol.merge_sdf_field()
    if init is not None:
        vol.init_tsdf = torch.from_numpy(init).to(DEV).contiguous()
    return vol


def test_query_matches_the_reference_recording():
    E, R = G["tsdf"].shape[0], int(G["resolution"])
    vol = _volume(E, R, _parts())
    assert torch.equal(vol.sdf_field.cpu(), torch.from_numpy(G["sdf_field"])) and (vol.bboxResy, vol.bboxResz) == tuple(G["bbox_res"][1:])
    assert float((vol.init_tsdf.cpu() - torch.from_numpy(G["init_tsdf"])).abs().max()) == 0.0      # the ground plane, bit for bit
    out = vol.query_tsdf(torch.from_numpy(G["pose_R"]).to(DEV), torch.from_numpy(G["pose_T"]).to(DEV)).cpu().numpy()
    d = np.abs(out - G["tsdf"])
    assert out.shape == G["tsdf"].shape and float((d > 1e-5).mean()) <= 2e-4 and float(np.median(d)) <= 1e-6, (float((d > 1e-5).mean()), float(d.max()))


@pytest.mark.parametrize("E,R", [(5, 50), (1, 7), (33, 20)])
def test_query_matches_oracle(E, R):
    rng = np.random.default_rng(E + R)
    parts = _parts()
    Mn = len(parts)
    A = rng.standard_normal((E, Mn, 3, 3))
    Q, _ = np.linalg.qr(A)
    Q = (Q * np.sign(np.linalg.det(Q))[..., None, None]).astype(np.float32)
    T = np.stack([rng.uniform(-0.15, 0.15, (E, Mn)), rng.uniform(-0.15, 0.15, (E, Mn)), rng.uniform(0.05, 0.3, (E, Mn))], -1).astype(np.float32)
    init = rng.uniform(-0.05, 0.4, (E, R ** 3)).astype(np.float32)
    vol = _volume(E, R, parts, init)
    out = vol.query_tsdf_parallel(torch.from_numpy(Q).to(DEV), torch.from_numpy(T).to(DEV)).cpu().numpy()
    field, res, voxel, bmin, bres = M.merge_sdf_field(parts)
    want = M.query_tsdf(field, res, voxel, bmin, bres, M.voxel_centres(0.5, R, [-0.25, -0.25, -0.0503]), init, 4 * 0.5 / R, Q, T)
    d = np.abs(out - want)
    assert float((d > 1e-5).mean()) <= 2e-4 and float(np.median(d)) <= 1e-6, (float((d > 1e-5).mean()), float(d.max()))
    assert float(out.min()) >= -1.0 and float(out.max()) <= 1.0


def test_parts_outside_the_workspace_leave_the_initial_volume():
    E, R = 2, 10
    vol = _volume(E, R, _parts())
    Rm = torch.eye(3, device=DEV).repeat(E, vol.part_num, 1, 1)
    T = torch.full((E, vol.part_num, 3), 50.0, device=DEV)                # every part far away: all queries invalid (+1)
    out = vol.query_tsdf(Rm, T)
    want = torch.clamp(vol.init_tsdf / vol.sdf_trunc, -1, 1).reshape(E, R, R, R)
    assert torch.equal(out, want)
