"""CPU: the mesh -> TSDF oracle (oracle/mesh2sdf_oracle.py) against the recording of the UNMODIFIED reference methods
(tests/golden/mesh2sdf_small.npz, made by tests/golden/make_golden_mesh2sdf.py)."""
import os

import numpy as np

from oracle import mesh2sdf_oracle as M

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh2sdf_small.npz"))


def _parts():
    n = sum(1 for k in G.files if k.endswith("_sdf") and k.startswith("part"))
    return [dict(sdf=G[f"part{i}_sdf"], bbox_min=G[f"part{i}_bbox_min"], voxel_size=G[f"part{i}_voxel_size"]) for i in range(n)]


def test_merge_matches_the_reference_recording():
    field, res, voxel, bmin, bres = M.merge_sdf_field(_parts())
    assert np.array_equal(field, G["sdf_field"]) and np.array_equal(res, G["sdf_field_res"])
    assert np.array_equal(voxel, G["sdf_voxel_size"]) and np.array_equal(bmin, G["sdf_bbox_min"]) and list(bres) == G["bbox_res"].tolist()


def test_query_matches_the_reference_recording():
    field, res, voxel, bmin, bres = M.merge_sdf_field(_parts())
    R, size = int(G["resolution"]), float(G["size"])
    centres = M.voxel_centres(size, R, G["vox_origin"])
    got = M.query_tsdf(field, res, voxel, bmin, bres, centres, G["init_tsdf"], 4 * size / R, G["pose_R"], G["pose_T"])
    want = G["tsdf"]
    assert got.shape == want.shape
    d = np.abs(got - want)
    # trilinear interpolation is continuous; only the validity test at a part's grid border is a step, so a query within rounding of
    # the border may differ: allow a handful of voxels, everything else to 1e-5
    assert float((d > 1e-5).mean()) <= 2e-4, float((d > 1e-5).mean())
    assert float(np.median(d)) <= 1e-6
    assert float((np.abs(want) < 1).mean()) > 0.2                         # the scene really intersects the band
