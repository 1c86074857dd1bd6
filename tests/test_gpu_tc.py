"""GPU parity tests of the tcgen05 (bf16 operands, fp32 TMEM accumulators) encoder — the throughput mode.
north_star bf16 gate: |a-b| <= 1e-2 + 1e-2*|b| on outputs (features, actions/means, values)."""
import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import close, load_golden, max_err, sub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")


def cu(t):
    return t.to(DEV).contiguous()


def _has_tc():
    from partmanip_b200._lib import lib
    return bool(lib.pm_has_tcgen05())


@pytest.mark.parametrize("B,N,C,act", [(1, 1024, 3, "tanh"), (5, 1024, 3, "tanh"), (200, 1024, 3, "tanh"),
                                       (3, 2048, 3, "tanh"), (4, 1024, 4, "relu"), (2, 256, 1, "tanh")])
def test_tc_encoder_features_and_argmax(B, N, C, act):
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200 import ops
    torch.manual_seed(B * 7 + N + C)
    x = torch.rand(B, N * C) * 2 - 1
    x.view(B, N, C)[:, ::9] = 0.0
    p = O.pointnet_init(N * C, 10, point_num=N, gen=torch.Generator().manual_seed(3))
    h = O.pointnet_encode(p, x.view(B, N, C), act)
    want = h.max(dim=1)[0]
    enc = [cu(p[k]) for k in NAMES]
    feat = torch.full((B, 512), float("nan"), device=DEV)
    am = torch.full((B, 512), -1, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "bf16", feat, None, am, None)
    assert ops.pointnet_tc_last_error(DEV) == 0
    assert close(feat.cpu(), want, 1e-2, 1e-2), max_err(feat.cpu(), want)
    am = am.cpu().long()
    assert int(am.min()) >= 0 and int(am.max()) < N
    picked = h.gather(1, am[:, None, :]).squeeze(1)          # fp32 value at the point the bf16 kernel chose
    assert float((picked - want).abs().max()) <= 2e-2
    # without argmax (rollout variant) the features are identical
    feat2 = torch.empty(B, 512, device=DEV)
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "bf16", feat2, None, None, None)
    assert ops.pointnet_tc_last_error(DEV) == 0 and torch.equal(feat, feat2)


def test_tc_pointnet_golden_outputs_and_training_step():
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    g = load_golden("pointnet_base_a10.npz")
    net = PointNet(3072, 10, dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, precision="bf16"), 0)
    net.load_state_dict(sub(g, "w"))
    net.to(DEV)
    y = net(cu(g["x"]))
    assert close(y.detach().cpu(), g["y"], 1e-2, 1e-2), max_err(y.detach().cpu(), g["y"])
    y.square().sum().backward()
    # gradients flow through the fp32 critical-point backward; with a bf16 forward they agree with the reference's
    # to the forward's accuracy except on rows whose argmax flipped between near-tied points
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        rel = float((got - v).norm() / (v.norm() + 1e-12))
        assert rel < 0.1, (k, rel)
