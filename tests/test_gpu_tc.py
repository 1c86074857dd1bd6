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


def _oracle_grads_given_argmax(w, x, am, act, point_num=1024, proprio=0):
    """Reference gradients of sum(y^2) with the max-pool routed to the KERNEL's winning points (`am`, (B,512)).  Two points
    whose layer-3 values differ by less than the bf16 forward's error may swap places in the argmax; with the routing fixed
    the remaining difference is operand rounding only, so the gate can be tight.  x: already-centred input rows."""
    import torch.nn.functional as F
    wl = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    b = x.shape[0]
    pc = (x[:, :-proprio] if proprio else x).reshape(b, point_num, -1)
    h = O.pointnet_encode(wl, pc, act)
    feat = h.gather(1, am.long()[:, None, :]).squeeze(1)
    if proprio:
        feat = torch.cat((feat, x[:, -proprio:]), dim=-1)
    f = O.activation(act)
    o = f(F.linear(feat, wl["final_mlp.0.weight"], wl["final_mlp.0.bias"]))
    o = f(F.linear(o, wl["final_mlp.2.weight"], wl["final_mlp.2.bias"]))
    y = F.linear(o, wl["final_mlp.4.weight"], wl["final_mlp.4.bias"])
    y.square().sum().backward()
    return y.detach(), {k: v.grad for k, v in wl.items()}


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
    # without argmax (rollout variant) the features agree up to the 5 mantissa bits the argmax variant borrows for the
    # column index inside its max-pool key (2^-18 relative)
    feat2 = torch.empty(B, 512, device=DEV)
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "bf16", feat2, None, None, None)
    assert ops.pointnet_tc_last_error(DEV) == 0
    assert float((feat - feat2).abs().max()) <= 1e-5 * float(feat2.abs().max()) + 1e-7


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
    # Gradients through the tcgen05 backward.  (a) against the reference recording itself: every tensor within 5 % relative L2
    # (the residue is rows whose argmax flipped between near-tied points — measured in round 1: <= 3 %); (b) with the max-pool
    # routed to the kernel's own winners the flip freedom is gone and what remains is bf16 operand rounding: 2 % gate.
    am = net.runner._bufs[g["x"].shape[0]]["argmax"].cpu()
    _, forced = _oracle_grads_given_argmax(sub(g, "w"), g["x"], am, "tanh")
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        rel = float((got - v).norm() / (v.norm() + 1e-12))
        rel_f = float((got - forced[k]).norm() / (forced[k].norm() + 1e-12))
        assert rel < 5e-2 and rel_f < 2e-2, (k, rel, rel_f)


@pytest.mark.parametrize("B,N,C,act", [(1, 1024, 3, "tanh"), (8, 1024, 3, "tanh"), (75, 1024, 3, "tanh"), (301, 256, 3, "tanh"),
                                       (6, 2048, 4, "tanh"), (6, 2048, 4, "relu"), (5, 1024, 3, "elu")])
def test_tc_encoder_backward_vs_autograd(B, N, C, act):
    """The fused tcgen05 backward (bf16 operands, every (cloud, channel) pair a row) against the oracle's autograd
    with the SAME argmax: relative L2 error per gradient tensor within the bf16 operand rounding (gate 1e-2·sqrt-ish;
    measured ~3e-3), and the protocol error word clean."""
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200 import ops
    torch.manual_seed(B + N + C)
    x = torch.rand(B, N * C) * 2 - 1
    x.view(B, N, C)[:, ::10] = 0.0
    p = {k: v.requires_grad_(True) for k, v in O.pointnet_init(N * C, 10, point_num=N, gen=torch.Generator().manual_seed(2)).items()}
    h = O.pointnet_encode(p, x.view(B, N, C), act)
    feat, am = h.max(dim=1)
    dfeat = torch.randn(B, 512) * 0.01
    feat.backward(dfeat)
    enc = [cu(p[k].detach()) for k in NAMES]
    grads = [torch.full_like(t, float("nan")) for t in enc]
    ops.pointnet_encode_backward(cu(x), N, C, enc, act, cu(dfeat), cu(am.int()), grads, precision="bf16")
    assert ops.pointnet_bwd_tc_last_error(DEV) == 0
    for k, gt in zip(NAMES, grads):
        v = p[k].grad
        got = gt.cpu()
        assert bool(torch.isfinite(got).all()), k
        rel = float((got - v).norm() / (v.norm() + 1e-20))
        # relu's derivative is a step: rows whose bf16-recomputed pre-activation lands on the other side of 0 flip
        assert rel < (1e-2 if act != "relu" else 8e-2), (k, rel)
    # mlp.4.bias is a plain sum of dfeat: exact up to fp32 summation order
    assert float((grads[5].cpu() - p["mlp.4.bias"].grad).abs().max()) <= 1e-5 * float(dfeat.abs().sum(0).max()) + 1e-7


def test_tc_pointnet_submean_proprio_golden_bf16():
    """bf16 mode through the full network with in-place centring (Q3) and a proprio tail (F = 537): outputs within the
    1e-2 gate of the reference recording, the caller's tensor centred exactly like the reference centres it."""
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    g = load_golden("pointnet_submean_proprio.npz")
    p = int(sub(g, "w")["final_mlp.0.weight"].shape[1]) - 512
    net = PointNet(int(g["x"].shape[1]), int(g["y"].shape[1]), dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=True,
                                                                     precision="bf16"), p)
    net.load_state_dict(sub(g, "w"))
    net.to(DEV)
    x = cu(g["x"])
    y = net(x)
    assert close(y.detach().cpu(), g["y"], 1e-2, 1e-2), max_err(y.detach().cpu(), g["y"])
    assert close(x.cpu(), g["x_after"], 1e-5, 1e-6)
    y.square().sum().backward()
    am = net.runner._bufs[g["x"].shape[0]]["argmax"].cpu()
    _, forced = _oracle_grads_given_argmax(sub(g, "w"), g["x_after"], am, "tanh", proprio=p)
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        rel = float((got - v).norm() / (v.norm() + 1e-12))
        rel_f = float((got - forced[k]).norm() / (forced[k].norm() + 1e-12))
        assert bool(torch.isfinite(got).all()) and rel < 5e-2 and rel_f < 2e-2, (k, rel, rel_f)


def test_tc_pointnet_critic_head_golden_bf16():
    """The critic (one output, network.py:152-159 with output_dim=1) in bf16 mode against a recording of the unmodified reference
    class (tests/golden/pointnet_critic_out1.npz): value within the 1e-2 gate, gradients as above."""
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    g = load_golden("pointnet_critic_out1.npz")
    net = PointNet(3072, 1, dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, precision="bf16"), 0)
    net.load_state_dict(sub(g, "w"))
    net.to(DEV)
    y = net(cu(g["x"]))
    assert tuple(y.shape) == (g["x"].shape[0], 1)
    assert close(y.detach().cpu(), g["y"], 1e-2, 1e-2), max_err(y.detach().cpu(), g["y"])
    y.square().sum().backward()
    am = net.runner._bufs[g["x"].shape[0]]["argmax"].cpu()
    _, forced = _oracle_grads_given_argmax(sub(g, "w"), g["x"], am, "tanh")
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        rel = float((got - v).norm() / (v.norm() + 1e-12))
        rel_f = float((got - forced[k]).norm() / (forced[k].norm() + 1e-12))
        assert rel < 5e-2 and rel_f < 2e-2, (k, rel, rel_f)


def test_tc_unsupported_shapes_fail_loudly():
    """The tcgen05 path has shape limits (N multiple of 256, C <= 4); outside them the C-ABI returns an error code and a
    message (no silent fallback) and the fp32 path serves the shape."""
    from partmanip_b200 import ops
    from partmanip_b200._lib import PMError
    p = O.pointnet_init(1000 * 3, 10, point_num=1000, gen=torch.Generator().manual_seed(1))
    enc = [cu(p[k]) for k in NAMES]
    x = cu(torch.rand(2, 3000))
    feat = torch.empty(2, 512, device=DEV)
    am = torch.empty(2, 512, device=DEV, dtype=torch.int32)
    with pytest.raises(PMError, match="multiple of 256"):
        ops.pointnet_encode_forward(x, 1000, 3, enc, "tanh", "bf16", feat, None, am, None)
    ops.pointnet_encode_forward(x, 1000, 3, enc, "tanh", "fp32", feat, None, am, None)
    want = O.pointnet_encode(p, x.cpu().view(2, 1000, 3)).max(dim=1)[0]
    assert close(feat.cpu(), want, 1e-4, 1e-5)
    p6 = O.pointnet_init(256 * 6, 10, point_num=256, gen=torch.Generator().manual_seed(1))
    with pytest.raises(PMError, match="channels per point"):
        ops.pointnet_encode_forward(cu(torch.rand(2, 1536)), 256, 6, [cu(p6[k]) for k in NAMES], "tanh", "bf16", feat, None, am, None)
    with pytest.raises(ValueError):
        ops.pointnet_encode_forward(torch.rand(2, 3000), 1000, 3, enc, "tanh", "fp32", feat, None, am, None)   # CPU tensor


def test_tc_encoder_full_size_properties():
    """BASELINE config-2 minibatch (2048 clouds x 1024 pts) through size-independent properties: (1) a random subset of
    clouds against the oracle, (2) the symmetric pool is exactly invariant to the order of the points of a cloud (argmax
    follows the permutation), (3) a cloud's features do not depend on which other clouds share the launch."""
    if not _has_tc():
        pytest.skip("library built without the tcgen05 encoder")
    from partmanip_b200 import ops
    torch.manual_seed(11)
    B, N, C = 2048, 1024, 3
    x = torch.rand(B, N, C) * 2 - 1
    x[:, ::11] = 0.0
    p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(4))
    enc = [cu(p[k]) for k in NAMES]
    xd = cu(x.reshape(B, N * C))
    feat = torch.empty(B, 512, device=DEV)
    am = torch.empty(B, 512, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(xd, N, C, enc, "tanh", "bf16", feat, None, am, None)
    assert ops.pointnet_tc_last_error(DEV) == 0
    # (1) subset parity
    idx = torch.randint(0, B, (8,))
    want = O.pointnet_encode(p, x[idx]).max(dim=1)[0]
    assert close(feat[idx.to(DEV)].cpu(), want, 1e-2, 1e-2), max_err(feat[idx.to(DEV)].cpu(), want)
    # (2) permutation invariance (exact: every point goes through identical arithmetic wherever it sits)
    perm = torch.randperm(N)
    xp = cu(x[:, perm].reshape(B, N * C))
    feat_p = torch.empty_like(feat)
    am_p = torch.empty_like(am)
    ops.pointnet_encode_forward(xp, N, C, enc, "tanh", "bf16", feat_p, None, am_p, None)
    assert float((feat_p - feat).abs().max()) <= 1e-5 * float(feat.abs().max())      # equal up to the 4 index bits in the key
    pts = x.reshape(B, N, C)
    a0 = pts.gather(1, am.cpu().long()[:, :, None].expand(-1, -1, C))
    a1 = pts[:, perm].gather(1, am_p.cpu().long()[:, :, None].expand(-1, -1, C))
    assert float(((a0 - a1).abs().amax(dim=-1) > 0).float().mean()) < 0.02                # same winning POINT (near-ties aside)
    # (3) batch independence: the first 37 clouds alone
    feat_s = torch.empty(37, 512, device=DEV)
    am_s = torch.empty(37, 512, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(xd[:37], N, C, enc, "tanh", "bf16", feat_s, None, am_s, None)
    assert torch.equal(feat_s, feat[:37]) and torch.equal(am_s, am[:37])


def test_empty_batch_is_rejected():
    from partmanip_b200.algorithms.algo_utils.network import PointNet
    net = PointNet(3072, 10, dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False), 0).to(DEV)
    with pytest.raises(ValueError, match="empty batch"):
        net(torch.empty(0, 3072, device=DEV))
