"""GPU parity of the fp32 parity mode on the tensor cores (pointnet_tc3.cu: split-fp16 operands, three tcgen05 MMAs per product,
weights streamed by the TMA engine) — network.py:148-150,182 at the reference's precision.  Gate: north_star's fp32
|a-b| <= 1e-4 + 1e-4*|b| on the pooled features; the winning points must attain the fp32 maximum to 1e-5."""
import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import close, max_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias", "mlp.4.weight", "mlp.4.bias")


def cu(t):
    return t.to(DEV).contiguous()


@pytest.mark.parametrize("B,N,C,act", [(1, 1024, 3, "tanh"), (5, 1024, 3, "tanh"), (200, 1024, 3, "tanh"), (3, 2048, 3, "tanh"),
                                       (4, 1024, 4, "relu"), (2, 256, 1, "tanh"), (7, 512, 2, "elu"), (3, 1024, 3, "sigmoid"),
                                       (3, 1024, 3, "lrelu"), (2, 768, 4, "selu")])
def test_tc3_features_and_argmax(B, N, C, act):
    from partmanip_b200 import ops
    torch.manual_seed(B * 7 + N + C)
    x = torch.rand(B, N * C) * 2 - 1
    x.view(B, N, C)[:, ::9] = 0.0
    p = O.pointnet_init(N * C, 10, point_num=N, gen=torch.Generator().manual_seed(3))
    h = O.pointnet_encode(p, x.view(B, N, C), act)
    want, want_i = h.max(dim=1)
    enc = [cu(p[k]) for k in NAMES]
    feat = torch.full((B, 512), float("nan"), device=DEV)
    am = torch.full((B, 512), -1, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "fp32", feat, None, am, None)
    assert ops.pointnet_tc3_last_error(DEV) == 0
    assert close(feat.cpu(), want, 1e-4, 1e-4), max_err(feat.cpu(), want)
    am = am.cpu().long()
    assert int(am.min()) >= 0 and int(am.max()) < N
    picked = h.gather(1, am[:, None, :]).squeeze(1)
    assert float((picked - want).abs().max()) <= 1e-5 * (1 + float(want.abs().max()))
    assert float((am == want_i).float().mean()) > 0.98
    feat2 = torch.empty(B, 512, device=DEV)                     # rollout variant (no argmax): same values up to the index bits
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "fp32", feat2, None, None, None)
    assert ops.pointnet_tc3_last_error(DEV) == 0
    assert float((feat - feat2).abs().max()) <= 1e-5 * float(feat2.abs().max()) + 1e-7
    # and the CUDA-core path agrees with both
    feat3 = torch.empty(B, 512, device=DEV)
    ops.pointnet_encode_forward(cu(x), N, C, enc, act, "fp32_ffma", feat3, None, None, None)
    assert close(feat3.cpu(), want, 1e-4, 1e-4)


def test_tc3_row_stride_with_proprio_tail_uses_the_unaligned_path():
    """ldx = N*C + 25 (a proprio tail): tile starts are not 16-byte aligned, so the points are fetched with plain loads instead
    of bulk copies — same results."""
    from partmanip_b200 import ops
    torch.manual_seed(2)
    B, N, C = 6, 1024, 3
    x = torch.rand(B, N * C + 25) * 2 - 1
    p = O.pointnet_init(N * C, 10, gen=torch.Generator().manual_seed(3))
    want = O.pointnet_encode(p, x[:, :N * C].reshape(B, N, C)).max(dim=1)[0]
    feat = torch.empty(B, 512, device=DEV)
    am = torch.empty(B, 512, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(cu(x), N, C, [cu(p[k]) for k in NAMES], "tanh", "fp32", feat, None, am, None)
    assert ops.pointnet_tc3_last_error(DEV) == 0
    assert close(feat.cpu(), want, 1e-4, 1e-4), max_err(feat.cpu(), want)


def test_tc3_full_size_properties():
    """BASELINE config-2 minibatch (2048 clouds x 1024 pts): subset parity against the oracle at the fp32 gate, permutation
    invariance of the pool, batch independence (bit-identical features whatever else shares the launch)."""
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
    ops.pointnet_encode_forward(xd, N, C, enc, "tanh", "fp32", feat, None, am, None)
    assert ops.pointnet_tc3_last_error(DEV) == 0
    idx = torch.randint(0, B, (8,))
    want = O.pointnet_encode(p, x[idx]).max(dim=1)[0]
    assert close(feat[idx.to(DEV)].cpu(), want, 1e-4, 1e-4), max_err(feat[idx.to(DEV)].cpu(), want)
    perm = torch.randperm(N)
    feat_p = torch.empty_like(feat)
    am_p = torch.empty_like(am)
    ops.pointnet_encode_forward(cu(x[:, perm].reshape(B, N * C)), N, C, enc, "tanh", "fp32", feat_p, None, am_p, None)
    assert float((feat_p - feat).abs().max()) <= 1e-5 * float(feat.abs().max())
    feat_s = torch.empty(37, 512, device=DEV)
    am_s = torch.empty(37, 512, device=DEV, dtype=torch.int32)
    ops.pointnet_encode_forward(xd[:37], N, C, enc, "tanh", "fp32", feat_s, None, am_s, None)
    assert torch.equal(feat_s, feat[:37]) and torch.equal(am_s, am[:37])
