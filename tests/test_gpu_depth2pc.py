"""GPU parity of the next row §8f(1) (depth -> world point cloud -> farthest-point subsample): CUDA kernels through the
C-ABI against the CPU oracle and the reference recording.  Index work (the FPS picks) is bit-exact; the floating-point
back-projection agrees to 1e-6 with an identical workspace mask."""
import os

import numpy as np
import pytest
import torch

from oracle import depth2pc_oracle as D

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth2pc_small.npz"))


def test_backproject_matches_reference_recording():
    from partmanip_b200 import ops
    depth = torch.from_numpy(G["depth"]).to(DEV)
    pose = torch.from_numpy(G["cam_pose"]).float().to(DEV)
    c = ops.depth2pc_backproject(depth, G["cam_intr"], pose, G["vol_origin"], float(G["size"])).cpu().numpy()
    want = G["cloud"]
    assert ((np.abs(c).sum(-1) > 0) == (np.abs(want).sum(-1) > 0)).all()
    assert float(np.abs(c - want).max()) <= 1e-6


@pytest.mark.parametrize("E,P,K", [(2, 1536, 1024), (3, 5000, 64), (1, 1, 1), (2, 60000, 32), (4, 1024, 1024), (2, 4999, 40), (1, 50001, 20)])
def test_fps_picks_bit_exact_vs_oracle(E, P, K):
    from partmanip_b200 import ops
    rng = np.random.default_rng(E * 1000 + P)
    pts = rng.uniform(-1, 1, (E, P, 3)).astype(np.float32)
    if P > 200:
        pts[:, 17:17 + P // 5] = 0.0                               # zeroed (masked) points: exact duplicates
    want_pts, want_idx = D.farthest_point_sample(pts, K)
    got, idx = ops.farthest_point_sample(torch.from_numpy(pts).to(DEV), K, return_idx=True)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(got.cpu().numpy(), want_pts)


def test_tsdfvolume_depth2pc_end_to_end():
    """The host mirror of TSDFVolume (same ctor / register_camera / depth2pc) on the recorded input: 1024 farthest points of
    the reference's own masked cloud, in the oracle's order."""
    from partmanip_b200.utils.depth2tsdf import TSDFVolume
    E, M, H, W = G["depth"].shape
    vol = TSDFVolume(DEV, size=float(G["size"]), resolution=8, _vol_origin=G["vol_origin"].tolist())
    vol.register_camera(G["cam_pose"], G["cam_intr"], H, W, E)
    out = vol.depth2pc(torch.from_numpy(G["depth"]).to(DEV)).cpu().numpy()
    want = D.farthest_point_sample(G["cloud"], 1024)[0]
    assert out.shape == (E, 1024, 3)
    # picks are decided on the kernel's own back-projection (1e-6 from the recording): compare as point SETS per env
    for e in range(E):
        a = np.unique(np.round(out[e], 5), axis=0)
        b = np.unique(np.round(want[e], 5), axis=0)
        assert a.shape == b.shape and float(np.abs(a - b).max()) <= 2e-5


@pytest.mark.parametrize("zero_frac", [0.0, 0.5, 0.97, 1.0])
def test_fps_compacted_equals_plain(zero_frac):
    """Compacting to the non-zero points + the first zero point leaves the picks unchanged (same coordinates AND the same
    original indices), also when fewer distinct points than K remain (repeats of point 0)."""
    from partmanip_b200 import ops
    rng = np.random.default_rng(int(zero_frac * 100))
    E, P, K = 3, 4096, 256
    pts = rng.uniform(-1, 1, (E, P, 3)).astype(np.float32)
    pts[rng.uniform(size=(E, P)) < zero_frac] = 0.0
    d = torch.from_numpy(pts).to(DEV)
    a, ia = ops.farthest_point_sample(d, K, return_idx=True, compact=False)
    b, ib = ops.farthest_point_sample(d, K, return_idx=True, compact=True)
    assert torch.equal(a, b) and torch.equal(ia, ib)
    want_pts, want_idx = D.farthest_point_sample(pts, K)
    assert np.array_equal(ib.cpu().numpy(), want_idx)


@pytest.mark.parametrize("E,P,K,zero_frac", [(3, 4096, 256, 0.5),       # small slices (64 groups per CTA)
                                             (2, 40, 8, 0.0),            # fewer groups than CTAs x threads: idle CTAs / threads
                                             (1, 4, 4, 1.0),             # one distinct point: repeats of point 0
                                             (2, 200000, 48, 0.0),       # slices cross the shared-memory / L2 boundary (18 432 points)
                                             (1, 300000, 24, 0.0),       # more than 32 Ki points per CTA (min distances in the global scratch)
                                             (1, 600000, 8, 0.0),        # more than 64 Ki points per CTA (pruned kernel: unsummarised, streamed blocks)
                                             (2, 65536, 1024, 0.9)])
@pytest.mark.parametrize("mode", [2, 4])
def test_fps_cluster_kernel_bit_exact(E, P, K, zero_frac, mode):
    """The 8-CTA cluster kernels (registers + DSMEM candidate exchange; mode 4 = with exact bounding-box pruning) make the
    oracle's picks, index for index."""
    from partmanip_b200 import ops
    rng = np.random.default_rng(P + K)
    pts = rng.uniform(-1, 1, (E, P, 3)).astype(np.float32)
    pts[rng.uniform(size=(E, P)) < zero_frac] = 0.0
    d = torch.from_numpy(pts).to(DEV)
    got, idx = ops.farthest_point_sample(d, K, return_idx=True, compact=mode)
    want_pts, want_idx = D.farthest_point_sample(pts, K)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(got.cpu().numpy(), want_pts)


def test_fps_cluster_equals_one_cta_per_cloud_with_ties():
    """Quantised coordinates produce many exactly tied distances: the first index must win in every reduction stage
    (thread, warp, CTA, cluster) exactly as in the one-CTA kernel."""
    from partmanip_b200 import ops
    rng = np.random.default_rng(7)
    pts = (rng.integers(-8, 9, (20, 60000, 3)) / 8.0).astype(np.float32)
    d = torch.from_numpy(pts).to(DEV)
    a, ia = ops.farthest_point_sample(d, 200, return_idx=True, compact=3)
    b, ib = ops.farthest_point_sample(d, 200, return_idx=True, compact=2)
    c, ic = ops.farthest_point_sample(d, 200, return_idx=True, compact=False)
    p, ip = ops.farthest_point_sample(d, 200, return_idx=True, compact=4)
    assert torch.equal(ia, ib) and torch.equal(a, b) and torch.equal(ia, ic)
    assert torch.equal(ia, ip) and torch.equal(a, p)


def _smooth_scene_depth(E, M, H, W, seed):
    """Depth images of a smooth scene (tilted plane + bumps + a box): neighbouring pixels are neighbouring points."""
    g = torch.Generator().manual_seed(seed)
    v, u = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    ph = torch.rand(E, M, 4, generator=g) * 6.28
    d = 0.55 + 0.1 * u + 0.05 * v + 0.03 * torch.sin(5 * u + ph[..., 0, None, None]) * torch.cos(4 * v + ph[..., 1, None, None])
    box = ((u - 0.2 * torch.cos(ph[..., 2, None, None])).abs() < 0.25) & ((v - 0.2 * torch.sin(ph[..., 3, None, None])).abs() < 0.2)
    return torch.where(box, d - 0.12, d).float().contiguous()


def test_fps_pruned_on_depth_image_clouds():
    """Spatially coherent clouds (what depth2pc produces) are where the pruning skips almost everything: the picks must still
    be the one-CTA kernel's and the oracle's."""
    from partmanip_b200 import ops
    E, M, H, W = 3, 3, 288, 512
    depth = _smooth_scene_depth(E, M, H, W, 5).to(DEV)
    intr = np.array([[366.0, 0, W // 2], [0, 366.0, H // 2], [0, 0, 1]])
    pose = torch.eye(4, device=DEV).repeat(M, 1, 1).contiguous()
    pose[:, 2, 3] = -0.3
    pose[1, 0, 3] = 0.05
    pose[2, 1, 3] = -0.05
    cloud = ops.depth2pc_backproject(depth, intr, pose, [-0.25, -0.25, -0.0503], 0.5)
    valid = float((cloud.abs().sum(-1) > 0).float().mean())
    assert 0.05 < valid < 0.95
    a, ia = ops.farthest_point_sample(cloud, 1024, return_idx=True, compact=3)
    b, ib = ops.farthest_point_sample(cloud, 1024, return_idx=True, compact=4)
    c, ic = ops.farthest_point_sample(cloud, 1024, return_idx=True)          # auto -> pruned cluster at this size
    assert torch.equal(ia, ib) and torch.equal(a, b) and torch.equal(ia, ic)
    want_pts, want_idx = D.farthest_point_sample(cloud[:1].cpu().numpy(), 1024)
    assert np.array_equal(ib[:1].cpu().numpy(), want_idx) and np.array_equal(b[:1].cpu().numpy(), want_pts)


@pytest.mark.parametrize("H,W,offset", [(36, 64, 0), (9, 7, 0), (36, 64, 1)])
def test_backproject_from_camera_tensors_in_place(H, W, offset):
    """tasks/hand_base.py:317-324 folded into the back-projection: E*M separate simulator images (negative depth, -inf
    background, some of them not 16-byte aligned) give exactly the cloud of the stacked / negated / inf-replaced tensor."""
    from partmanip_b200 import ops
    E, M = 5, 3
    g = torch.Generator().manual_seed(H * W + offset)
    views, host = [], []
    for e in range(E):
        row, hrow = [], []
        for m in range(M):
            raw = -(0.3 + 0.6 * torch.rand(H * W + offset, generator=g))
            raw[torch.rand(H * W + offset, generator=g) < 0.1] = float("-inf")
            t = raw.to(DEV)[offset:].view(H, W)                          # offset=1: pointer only 4-byte aligned
            row.append(t); hrow.append(raw[offset:].view(H, W).numpy())
        views.append(row); host.append(hrow)
    intr = np.array([[45.0, 0, W // 2], [0, 45.0, H // 2], [0, 0, 1]])
    pose = torch.eye(4, device=DEV).repeat(M, 1, 1).contiguous()
    pose[:, 2, 3] = -0.3
    table, aligned = ops.view_pointer_table(views)
    assert aligned == (offset == 0)
    got = ops.depth2pc_backproject_views(table, aligned, E, M, H, W, intr, pose, [-0.25, -0.25, -0.0503], 0.5)
    stacked = D.stack_views(host)
    assert float(stacked.max()) == 100.0
    same = ops.depth2pc_backproject(torch.from_numpy(stacked).to(DEV), intr, pose, [-0.25, -0.25, -0.0503], 0.5)
    assert torch.equal(got, same)
    want = D.backproject(stacked, intr, pose.cpu().numpy(), [-0.25, -0.25, -0.0503], 0.5)
    g_, w_ = got.cpu().numpy(), want
    agree = (np.abs(g_).sum(-1) > 0) == (np.abs(w_).sum(-1) > 0)            # a point within 1 ulp of a box face may flip
    assert agree.mean() > 0.999 and float(np.abs(g_ - w_)[agree].max()) <= 1e-5


def test_tsdfvolume_depth2pc_from_views_equals_stacked_path():
    from partmanip_b200.utils.depth2tsdf import TSDFVolume
    E, M, H, W = 2, 3, 72, 128
    g = torch.Generator().manual_seed(3)
    views = [[(-(0.35 + 0.4 * torch.rand(H, W, generator=g))).to(DEV) for _ in range(M)] for _ in range(E)]
    views[1][2][:5] = float("-inf")
    vol = TSDFVolume(DEV, size=0.5, resolution=8)
    intr = np.array([[90.0, 0, W // 2], [0, 90.0, H // 2], [0, 0, 1]])
    pose = np.tile(np.eye(4), (M, 1, 1)); pose[:, 2, 3] = -0.3
    vol.register_camera(pose, intr, H, W, E)
    a = vol.depth2pc_from_views(views)
    stacked = -torch.stack([torch.stack(v, 0) for v in views], 0)
    stacked = torch.where(torch.isinf(stacked), torch.full_like(stacked, 100), stacked)      # hand_base.py:317-324 verbatim semantics
    b = vol.depth2pc(stacked)
    assert a.shape == (E, 1024, 3) and torch.equal(a, b)
