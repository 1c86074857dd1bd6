"""CPU: the TSDF-integration oracle (oracle/tsdf_oracle.py) against the tables and the fused volume recorded from the UNMODIFIED
reference (tests/golden/tsdf_small.npz, made by tests/golden/make_golden_tsdf.py)."""
import os

import numpy as np

from oracle import tsdf_oracle as T

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tsdf_small.npz"))


def _tables():
    E, M, H, W = G["depth"].shape
    return T.voxel_pixel_tables(G["cam_pose"], G["cam_intr"], H, W, float(G["size"]), int(G["resolution"]), G["vol_origin"])


def test_voxel_pixel_tables_match_the_reference_recording():
    px, py, pz, valid = _tables()
    assert (valid == G["valid_pix"]).all()
    assert (px == G["pix_x"]).all() and (py == G["pix_y"]).all()
    assert float(np.abs(pz - G["pix_z"]).max()) <= 1e-6


def test_integrate_matches_the_reference_recording():
    px, py, pz, valid = _tables()
    vol = T.integrate(G["depth"], px, py, pz, valid, float(G["size"]), int(G["resolution"]))
    assert vol.shape == G["tsdf"].shape and vol.dtype == np.float32
    assert float(np.abs(vol - G["tsdf"]).max()) <= 1e-6
    assert ((vol == 1) == (G["tsdf"] == 1)).all()


def test_integrate_edge_cases():
    px, py, pz, valid = _tables()
    E, M, H, W = G["depth"].shape
    R = int(G["resolution"])
    # no surface anywhere (depth 0 = hole in every pixel): every voxel keeps the default value 1
    vol = T.integrate(np.zeros((1, M, H, W), np.float32), px, py, pz, valid, float(G["size"]), R)
    assert (vol == 1).all()
    # background far behind the workspace: every visible voxel is in free space -> clamp at +1
    vol = T.integrate(np.full((1, M, H, W), 100, np.float32), px, py, pz, valid, float(G["size"]), R)
    assert (vol == 1).all()
    # a surface in front of every voxel (occluded beyond the band): no view is valid -> default 1
    vol = T.integrate(np.full((1, M, H, W), 1e-3, np.float32), px, py, pz, valid, float(G["size"]), R)
    assert (vol == 1).all()


def test_sparse_voxel_properties():
    """depth2tsdf.py:103-119: the picks are band voxels, start at the first one in row-major order, are distinct while the band
    lasts, and carry their own TSDF value."""
    vol = G["tsdf"]
    band = (vol < 0.2) & (vol > -0.2)
    K = 48
    out = T.sparse_voxel(vol, K)
    assert out.shape == (vol.shape[0], K, 4) and out.dtype == np.float32
    for e in range(vol.shape[0]):
        xyz = out[e, :, :3].astype(np.int64)
        assert band[e][xyz[:, 0], xyz[:, 1], xyz[:, 2]].all()
        assert (xyz[0] == np.argwhere(band[e])[0]).all()
        n = min(K, int(band[e].sum()))
        assert len({tuple(v) for v in xyz[:n]}) == n
        assert np.array_equal(out[e, :, 3], vol[e][xyz[:, 0], xyz[:, 1], xyz[:, 2]])
    # empty band / band smaller than K: defined behaviour of the kernel (voxel 0 / repeats of the first band voxel)
    v2 = np.ones((1, 4, 4, 4), np.float32)
    assert (T.sparse_voxel(v2, 8) == np.array([0, 0, 0, 1], np.float32)).all()
    v2[0, 1, 2, 3] = 0.05
    v2[0, 3, 0, 1] = -0.1
    o = T.sparse_voxel(v2, 5)[0]
    assert o[:2].tolist() == [[1, 2, 3, np.float32(0.05)], [3, 0, 1, np.float32(-0.1)]] and (o[2:] == o[0]).all()
