"""CPU: the Conv3D-student oracle (oracle/conv3d_oracle.py) against outputs and parameter gradients recorded from the UNMODIFIED
reference `Conv3DNet` (tests/golden/conv3d_student.npz, made by tests/golden/make_golden_conv3d.py)."""
import os

import numpy as np
import pytest

from oracle import conv3d_oracle as C

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "conv3d_student.npz"))


def _cfg(tag):
    params = {k[len(tag) + 7:]: G[k] for k in G.files if k.startswith(tag + "_param_")}
    grads = {k[len(tag) + 6:]: G[k] for k in G.files if k.startswith(tag + "_grad_")}
    return params, grads, G[tag + "_x"], G[tag + "_y"], G[tag + "_gy"]


@pytest.mark.parametrize("tag,act,proprio", [("tanh_p0", "tanh", 0), ("relu_p7", "relu", 7)])
def test_conv3dnet_forward_and_backward_match_the_reference_recording(tag, act, proprio):
    params, want_grads, x, want_y, gy = _cfg(tag)
    assert params["encoder.conv1.weight"].shape == (16, 1, 5, 5, 5) and params["final_mlp.0.weight"].shape == (256, 32 * 27 + proprio)
    y, grads = C.conv3dnet_backward(x, params, act, proprio, gy)
    assert y.shape == want_y.shape
    assert float(np.abs(y - want_y).max()) <= 1e-4 + 1e-4 * float(np.abs(want_y).max())       # north_star fp32 gate
    assert set(grads) == set(want_grads)
    for k, g in grads.items():
        w = want_grads[k]
        assert g.shape == w.shape, k
        assert float(np.abs(g - w).max()) <= 1e-4 * max(1.0, float(np.abs(w).max())), k


def test_conv3d_shapes_follow_torch_padding_rule():
    x = np.zeros((1, 1, 50, 50, 50), np.float32)
    w1 = np.zeros((16, 1, 5, 5, 5), np.float32)
    y, _ = C.conv3d_fwd(x, w1, np.zeros(16, np.float32), 3)
    assert y.shape == (1, 16, 17, 17, 17)                              # network.py:84 comment: 50 -> 17 -> 6 -> 3
    y2, _ = C.conv3d_fwd(y, np.zeros((32, 16, 3, 3, 3), np.float32), np.zeros(32, np.float32), 3)
    y3, _ = C.conv3d_fwd(y2, np.zeros((32, 32, 3, 3, 3), np.float32), np.zeros(32, np.float32), 2)
    assert y2.shape == (1, 32, 6, 6, 6) and y3.shape == (1, 32, 3, 3, 3)
