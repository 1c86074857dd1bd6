"""GPU parity of the next row §8f(3), first half (depth -> fused TSDF volume): CUDA kernels through the C-ABI against the
recording of the unmodified reference (tests/golden/tsdf_small.npz) and the CPU oracle at the reference's resolution."""
import os

import numpy as np
import pytest
import torch

from oracle import tsdf_oracle as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tsdf_small.npz"))


def _unpack(pix_off, W):
    off = pix_off.cpu().numpy()
    valid = off >= 0
    return np.where(valid, off % W, 0), np.where(valid, off // W, 0), valid


def test_voxel_tables_match_reference_recording():
    from partmanip_b200 import ops
    E, M, H, W = G["depth"].shape
    pose = torch.from_numpy(G["cam_pose"]).float().to(DEV).contiguous()
    pix_off, pix_z = ops.tsdf_voxel_tables(pose, G["cam_intr"], H, W, float(G["size"]), int(G["resolution"]), G["vol_origin"])
    px, py, valid = _unpack(pix_off, W)
    assert (valid == G["valid_pix"]).all()
    assert (px == G["pix_x"]).all() and (py == G["pix_y"]).all()
    assert float(np.abs(pix_z.cpu().numpy() - G["pix_z"]).max()) <= 1e-6


def test_integrate_matches_reference_recording():
    from partmanip_b200.utils.depth2tsdf import TSDFVolume
    E, M, H, W = G["depth"].shape
    vol = TSDFVolume(DEV, size=float(G["size"]), resolution=int(G["resolution"]), _vol_origin=G["vol_origin"].tolist())
    vol.register_camera(G["cam_pose"], G["cam_intr"], H, W, E)
    out = vol.integrate(torch.from_numpy(G["depth"]).to(DEV)).cpu().numpy()
    assert out.shape == G["tsdf"].shape
    assert float(np.abs(out - G["tsdf"]).max()) <= 1e-6
    assert ((out == 1) == (G["tsdf"] == 1)).all()


# M <= 4 runs the blocked one-pass kernel (E = 19 spans three 8-env groups with a ragged tail), M = 5 the generic two-pass kernel
@pytest.mark.parametrize("E,M,H,W,R", [(4, 3, 72, 128, 50), (1, 1, 9, 7, 5), (3, 2, 40, 40, 17), (19, 3, 36, 64, 50), (2, 5, 40, 40, 13),
                                       (9, 4, 30, 50, 6)])
def test_integrate_matches_oracle(E, M, H, W, R):
    from partmanip_b200 import ops
    rng = np.random.default_rng(E * 100 + R)
    fx = W / 2.0 / np.tan(np.deg2rad(69.75) / 2.0)
    intr = np.array([[fx, 0, W // 2], [0, fx, H // 2], [0, 0, 1]])
    poses = np.concatenate([G["cam_pose"]] * 2)[:M]
    org = [-0.25, -0.25, -0.0503]
    depth = (0.62 + 0.15 * rng.standard_normal((E, M, H, W))).astype(np.float32)
    depth[rng.uniform(size=depth.shape) < 0.05] = 0.0
    depth[rng.uniform(size=depth.shape) < 0.05] = 100.0
    px, py, pz, valid = T.voxel_pixel_tables(poses, intr, H, W, 0.5, R, org)
    want = T.integrate(depth, px, py, pz, valid, 0.5, R)
    pose_d = torch.from_numpy(poses).float().to(DEV).contiguous()
    pix_off, pix_z = ops.tsdf_voxel_tables(pose_d, intr, H, W, 0.5, R, org)
    gx, gy, gv = _unpack(pix_off, W)
    assert (gv == valid).all() and (gx == px).all() and (gy == py).all()
    assert float(np.abs(pix_z.cpu().numpy() - pz).max()) <= 1e-6
    got = ops.tsdf_integrate(torch.from_numpy(depth).to(DEV), pix_off, pix_z, 0.5, R).cpu().numpy()
    assert float(np.abs(got - want).max()) <= 1e-5 and ((got == 1) == (want == 1)).mean() > 0.9999
    # on the kernel's own tables the fusion itself follows the oracle's fp32 operation order: bit-exact
    want_same = T.integrate(depth, gx, gy, pix_z.cpu().numpy(), gv, 0.5, R)
    assert np.array_equal(got, want_same)


def test_integrate_empty_scene_keeps_default():
    from partmanip_b200 import ops
    E, M, H, W = G["depth"].shape
    R = int(G["resolution"])
    pose = torch.from_numpy(G["cam_pose"]).float().to(DEV).contiguous()
    pix_off, pix_z = ops.tsdf_voxel_tables(pose, G["cam_intr"], H, W, float(G["size"]), R, G["vol_origin"])
    for fill in (0.0, 100.0, 1e-3):
        out = ops.tsdf_integrate(torch.full((2, M, H, W), fill, device=DEV), pix_off, pix_z, float(G["size"]), R)
        assert bool((out == 1).all())


def _smooth_volume(E, R, seed):
    """A synthetic fused volume: signed distance to a wavy sheet, clamped like the reference's (values in (-1, 1], 1 = free)."""
    rng = np.random.default_rng(seed)
    x, y, z = np.meshgrid(np.arange(R), np.arange(R), np.arange(R), indexing="ij")
    vols = []
    for e in range(E):
        a, b, c = rng.uniform(0.1, 0.4, 3)
        h = R * (0.5 + 0.2 * np.sin(a * x + e) * np.cos(b * y)) + c
        vols.append(np.clip((h - z) / 4.0, -1, 1))
    v = np.stack(vols).astype(np.float32)
    v[v <= -0.99] = 1.0
    return v


@pytest.mark.parametrize("E,R,K", [(3, 12, 48), (2, 50, 1024), (2, 5, 16), (1, 31, 200)])
def test_sparse_voxel_matches_oracle(E, R, K):
    from partmanip_b200 import ops
    vol = G["tsdf"] if R == 12 else _smooth_volume(E, R, R)
    want = T.sparse_voxel(vol, K)
    got = ops.tsdf_sparse_voxel(torch.from_numpy(vol).to(DEV), K).cpu().numpy()
    assert np.array_equal(got, want)                                   # integer coordinates: index-exact picks, gathered values


def test_sparse_voxel_empty_and_small_bands():
    from partmanip_b200 import ops
    v = np.ones((2, 6, 6, 6), np.float32)
    v[1, 1, 2, 3] = 0.05
    v[1, 5, 0, 1] = -0.1
    v[1, 2, 2, 2] = 0.2                                                  # on the threshold: excluded (strict comparisons)
    got = ops.tsdf_sparse_voxel(torch.from_numpy(v).to(DEV), 7).cpu().numpy()
    assert np.array_equal(got, T.sparse_voxel(v, 7))
    assert (got[0] == np.array([0, 0, 0, 1], np.float32)).all()
    assert got[1, :2].tolist() == [[1, 2, 3, np.float32(0.05)], [5, 0, 1, np.float32(-0.1)]]


def test_tsdfvolume_sparse_voxel_end_to_end():
    from partmanip_b200.utils.depth2tsdf import TSDFVolume
    E, M, H, W = G["depth"].shape
    vol = TSDFVolume(DEV, size=float(G["size"]), resolution=int(G["resolution"]), _vol_origin=G["vol_origin"].tolist())
    vol.register_camera(G["cam_pose"], G["cam_intr"], H, W, E)
    vol.num_points = 40
    out = vol.sparse_voxel(torch.from_numpy(G["depth"]).to(DEV)).cpu().numpy()
    fused = vol.integrate(torch.from_numpy(G["depth"]).to(DEV)).cpu().numpy()
    assert np.array_equal(out, T.sparse_voxel(fused, 40))


def test_config5_standin_sparse_voxels_into_pointnet_bf16():
    """BASELINE config 5's stand-in (the reference has no Sparse-UNet; its `depth_sparse` observation feeds PointNet as
    1024 x 4): depth images -> TSDF fusion -> 1024 farthest band voxels (x, y, z, tsdf) (utils/depth2tsdf.py:88-120) -> PointNet
    with C = 4 on the tcgen05 path (network.py:141-198), end to end against the oracle chain at the bf16 gate; the observation
    itself (integer coordinates + gathered values) must be exact."""
    from oracle import ppo_oracle as O
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    from partmanip_b200.utils.depth2tsdf import TSDFVolume
    from tests.helpers import close, max_err
    E, M, H, W, R = 6, 3, 72, 128, 50
    rng = np.random.default_rng(77)
    fx = W / 2.0 / np.tan(np.deg2rad(69.75) / 2.0)
    intr = np.array([[fx, 0, W // 2], [0, fx, H // 2], [0, 0, 1]])
    poses, org = G["cam_pose"][:M], [-0.25, -0.25, -0.0503]
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = np.stack([np.stack([0.62 + 0.08 * np.sin(0.05 * xx + e) * np.cos(0.07 * yy + m) for m in range(M)]) for e in range(E)])
    depth = (depth + 0.002 * rng.standard_normal(depth.shape)).astype(np.float32)
    vol = TSDFVolume(DEV, size=0.5, resolution=R, _vol_origin=org)
    vol.register_camera(poses, intr, H, W, E)
    obs = vol.sparse_voxel(torch.from_numpy(depth).to(DEV))
    assert tuple(obs.shape) == (E, 1024, 4)
    fused = vol.integrate(torch.from_numpy(depth).to(DEV)).cpu().numpy()
    want_obs = T.sparse_voxel(fused, 1024)
    assert np.array_equal(obs.cpu().numpy(), want_obs)
    torch.manual_seed(5)
    net = PointNet(4096, 10, dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, precision="bf16"), 0)
    assert net.in_channels == 4
    w = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net.to(DEV)
    # the voxel coordinates are 0..49: scale into the unit box like a `depth_sparse` consumer would before the first Linear
    scale = torch.tensor([1 / R, 1 / R, 1 / R, 1.0], device=DEV)
    x = (obs * scale).reshape(E, 4096).contiguous()
    with torch.no_grad():
        y = net(x)
    want = O.pointnet_forward(w, (torch.from_numpy(want_obs) * scale.cpu()).reshape(E, 4096))
    assert close(y.cpu(), want, 1e-2, 1e-2), max_err(y.cpu(), want)
