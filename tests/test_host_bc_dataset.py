"""Host-side pieces of the behaviour-cloning mirror (no GPU): the offline dataset indexes files exactly like algorithms/bc.py:12-31."""
import os

import numpy as np

from tests.helpers_bc import write_dataset


def test_tsdf_dataset_index_mapping(tmp_path):
    import importlib
    import sys
    import types
    n = write_dataset(str(tmp_path), seed=1, scenes=3, steps=4, R=6, A=10, P=5)
    # the dataset class lives next to code that needs the CUDA library at import; load just the class definition's module lazily
    try:
        from partmanip_b200.algorithms.bc import Tsdf_Dataset
    except Exception:                                   # library not built: the class itself has no CUDA dependency
        src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "partmanip_b200", "algorithms", "bc.py")).read()
        mod = types.ModuleType("bc_dataset_only")
        exec(src[src.index("class Tsdf_Dataset"):src.index("class bc:")], {"torch": importlib.import_module("torch"), "np": np, "os": os,
                                                                         "pjoin": os.path.join}, mod.__dict__)
        Tsdf_Dataset = mod.__dict__["Tsdf_Dataset"]
        sys.modules.pop("bc_dataset_only", None)
    ds = Tsdf_Dataset(str(tmp_path))
    assert len(ds) == n == 12 and ds.step_num == 4 and ds.env_num == 3
    scenes = os.listdir(str(tmp_path))
    for idx in (0, 3, 4, 11):
        tsdf, action, state = ds[idx]
        want = np.load(os.path.join(str(tmp_path), scenes[idx // 4], f"step_{idx % 4:05d}.npy"), allow_pickle=True).item()
        assert np.array_equal(tsdf, want["tsdf"]) and np.array_equal(action, want["action"]) and np.array_equal(state, want["proprio_state"])
        assert tsdf.shape == (6, 6, 6) and action.shape == (10,) and state.shape == (1, 5)
