"""CPU: host logic of the TSDFVolume mirror that needs no GPU — the env chunking of depth2pc / depth2pc_from_views composes the
per-chunk results in order (kernels replaced by CPU fakes), and the class refuses non-CUDA devices (no CPU fallback)."""
import numpy as np
import pytest
import torch

from partmanip_b200.utils import depth2tsdf as DT


def test_env_chunks_cover_in_order():
    assert DT.env_chunks(5, 256) == [(0, 5)]
    assert DT.env_chunks(600, 256) == [(0, 256), (256, 512), (512, 600)]
    assert DT.env_chunks(512, 256) == [(0, 256), (256, 512)]
    assert DT.env_chunks(3, 1) == [(0, 1), (1, 2), (2, 3)]


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        DT.TSDFVolume("cpu")


def _fake_volume(E, M, H, W, chunk):
    vol = object.__new__(DT.TSDFVolume)
    vol.registered_shape = (E, M, H, W)
    vol.cam_intr, vol.cam_pose, vol._vol_origin, vol._size = np.eye(3), torch.eye(4).repeat(M, 1, 1), [0.0, 0.0, 0.0], 0.5
    vol.num_points, vol.env_chunk = 4, chunk
    return vol


@pytest.mark.parametrize("chunk", [1, 2, 3, 256])
def test_depth2pc_chunking_composes(monkeypatch, chunk):
    E, M, H, W = 5, 2, 3, 4
    calls = []

    def fake_backproject(depth, intr, pose, org, size):
        calls.append(depth.shape[0])
        return depth.reshape(depth.shape[0], -1, 1).repeat(1, 1, 3)           # "cloud" = depth values

    def fake_fps(cloud, K):
        return cloud[:, :K].clone()

    monkeypatch.setattr(DT.ops, "depth2pc_backproject", fake_backproject)
    monkeypatch.setattr(DT.ops, "farthest_point_sample", fake_fps)
    depth = torch.arange(E * M * H * W, dtype=torch.float32).reshape(E, M, H, W)
    out = _fake_volume(E, M, H, W, chunk).depth2pc(depth)
    want = depth.reshape(E, -1, 1).repeat(1, 1, 3)[:, :4]
    assert out.shape == (E, 4, 3) and torch.equal(out, want)
    assert calls == [hi - lo for lo, hi in DT.env_chunks(E, chunk)]


@pytest.mark.parametrize("chunk", [2, 256])
def test_depth2pc_from_views_chunking_slices_the_pointer_table(monkeypatch, chunk):
    E, M, H, W = 5, 3, 2, 2
    views = [[torch.full((H, W), float(e * M + m)) for m in range(M)] for e in range(E)]
    seen = []

    def fake_table(lst):
        return torch.arange(E * M, dtype=torch.int64), True                  # stands in for the device pointer table

    def fake_backproject_views(table, aligned, e, m, h, w, intr, pose, org, size, negate=True, inf_value=100.0):
        assert table.numel() == e * m and aligned and negate and inf_value == 100.0
        seen.append(table.tolist())
        return table.float().reshape(e, m, 1).repeat(1, 2, 3)                # (e, 2m, 3)

    monkeypatch.setattr(DT.ops, "view_pointer_table", fake_table)
    monkeypatch.setattr(DT.ops, "depth2pc_backproject_views", fake_backproject_views)
    monkeypatch.setattr(DT.ops, "farthest_point_sample", lambda cloud, K: cloud[:, :K].clone())
    out = _fake_volume(E, M, H, W, chunk).depth2pc_from_views(views)
    assert out.shape == (E, 4, 3)
    assert [i for part in seen for i in part] == list(range(E * M))          # every env's views, once, in order
    assert torch.equal(out[:, 0, 0], torch.arange(E, dtype=torch.float32) * M)
