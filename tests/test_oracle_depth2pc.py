"""CPU: the next-row oracle (oracle/depth2pc_oracle.py) against the masked cloud recorded from the UNMODIFIED reference
(tests/golden/depth2pc_small.npz, made by tests/golden/make_golden_depth2pc.py) and against the defining properties of
farthest-point sampling (the pytorch3d step has no recording: pytorch3d is not installed — parity unpinned there)."""
import os

import numpy as np

from oracle import depth2pc_oracle as D

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "depth2pc_small.npz"))


def test_backprojection_and_mask_match_the_reference_recording():
    c = D.backproject(G["depth"], G["cam_intr"], G["cam_pose"], G["vol_origin"], float(G["size"]))
    assert c.shape == G["cloud"].shape and c.dtype == np.float32
    valid, want_valid = np.abs(c).sum(-1) > 0, np.abs(G["cloud"]).sum(-1) > 0
    assert (valid == want_valid).all() and int(valid.sum()) == int(G["n_valid"])
    assert float(np.abs(c - G["cloud"]).max()) <= 1e-6


def test_fps_greedy_max_min_property_and_tie_rule():
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1, 1, (2, 400, 3)).astype(np.float32)
    pts[:, 50:120] = 0.0                                            # the env's invalid-point padding: exact duplicates
    sel, idx = D.farthest_point_sample(pts, 64)
    assert (idx[:, 0] == 0).all() and sel.shape == (2, 64, 3)
    for e in range(2):
        chosen = [0]
        for k in range(1, 64):
            d = ((pts[e][:, None, :] - pts[e][chosen][None]) ** 2).sum(-1).min(1)
            assert abs(d[idx[e, k]] - d.max()) <= 1e-6 * max(1.0, d.max())      # every pick maximises the min distance
            chosen.append(int(idx[e, k]))
        dup = [i for i in idx[e] if 50 <= i < 120]
        assert dup in ([], [50])                                    # at most one representative of the duplicates: the first one
        assert len(set(idx[e].tolist())) == 64
    # with K beyond the number of distinct points the duplicates' representative is picked exactly once, then only repeats remain
    sel, idx = D.farthest_point_sample(pts[:1], 340)
    assert [i for i in idx[0, :331] if 50 <= i < 120] == [50] and len(set(idx[0, :331].tolist())) == 331


def test_fps_on_the_reference_cloud_is_deterministic():
    a = D.farthest_point_sample(G["cloud"], 128)[1]
    b = D.farthest_point_sample(G["cloud"].copy(), 128)[1]
    assert (a == b).all() and (a[:, 0] == 0).all()


def test_stack_views_matches_torch_semantics():
    """hand_base.py:317-324: stack per env, stack over envs, negate, isinf -> 100."""
    import torch
    rng = np.random.default_rng(0)
    views = [[-(rng.uniform(0.2, 1.0, (4, 6)).astype(np.float32)) for _ in range(2)] for _ in range(3)]
    views[1][0][2, 3] = -np.inf
    views[2][1][0, 0] = np.inf
    got = D.stack_views(views)
    t = torch.stack([torch.stack([torch.from_numpy(v) for v in vs], dim=0) for vs in views], dim=0)
    t = -t
    t = torch.where(torch.isinf(t), torch.full_like(t, 100), t)
    assert got.shape == (3, 2, 4, 6) and np.array_equal(got, t.numpy())
