"""GPU parity of K3t (dense_tc.cu): nn.Linear forward / backward on tcgen05 for fp32 tensors — the state-policy MLP of
cfg/algos/ppo.yaml (53 -> 512^3 -> 10, network.py:27-54) and the dense layers of the fp32 encoder backward — against torch
autograd in fp32 and the recordings of the unmodified reference class.  Gates: "fp32" (three-term bf16 split) at north_star's
1e-4, "bf16" at 1e-2."""
import pytest
import torch

from tests.helpers import close, load_golden, max_err, sub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ACTS = {"tanh": torch.tanh, "elu": torch.nn.functional.elu, "relu": torch.relu, None: lambda t: t}


def cu(t):
    return t.to(DEV).contiguous()


SHAPES = [(9, 96, 37, "tanh"), (2048, 512, 53, "tanh"), (2048, 512, 512, "tanh"), (300, 10, 512, None), (129, 1, 512, "elu"),
          (1000, 256, 128, "relu"), (1, 7, 5, "tanh"), (257, 130, 200, "elu"),
          (100001, 256, 128, "tanh"), (77777, 128, 256, "relu")]   # ~5 and ~2 M tiles per CTA: the round-robin tile walk


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
@pytest.mark.parametrize("M,N,K,act", SHAPES)
def test_linear_forward_backward_tc_vs_torch(M, N, K, act, precision, rtol):
    from partmanip_b200 import ops
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, dtype=torch.float64)
    W = (torch.randn(N, K, dtype=torch.float64) / K ** 0.5).requires_grad_(True)
    b = (torch.randn(N, dtype=torch.float64) * 0.1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    f = ACTS[act]
    hprev = f(xr)                                   # the layer input is itself an activation output (DX multiplies by act'(hprev))
    y = f(torch.nn.functional.linear(hprev, W, b))
    dy = torch.randn(M, N, dtype=torch.float64)
    # dpre = dL/d(pre-activation of THIS layer) is what the kernels receive
    pre = torch.nn.functional.linear(hprev, W, b)
    dpre = torch.autograd.grad(f(pre), pre, dy, retain_graph=True)[0]
    gW, gb, gin = torch.autograd.grad(y, (W, b, xr), dy)
    h32 = cu(hprev.detach().float())
    W32, b32 = cu(W.detach().float()), cu(b.detach().float())
    got = ops.linear_forward_tc(h32, W32, b32, act, precision)
    want = y.detach().float()
    assert close(got.cpu(), want, rtol, rtol), max_err(got.cpu(), want)
    dW, db, dx = torch.full((N, K), float("nan"), device=DEV), torch.full((N,), float("nan"), device=DEV), torch.full((M, K), float("nan"), device=DEV)
    ops.linear_backward_tc(h32, W32, cu(dpre.float()), dW, db, dx, act, precision)
    for name, g, w in (("dW", dW, gW), ("db", db, gb), ("dx", dx, gin)):
        w = w.float()
        tol = rtol * float(w.abs().max()) + 1e-7
        assert bool(torch.isfinite(g).all()), name
        assert float((g.cpu() - w).abs().max()) <= tol, (name, float((g.cpu() - w).abs().max()), tol)


def test_tc_split_keeps_tiny_gradients():
    """The three-term split is bf16-based: it has fp32's exponent range, so gradients of 1e-12 keep their relative accuracy
    (an fp16-based split would flush them to zero)."""
    from partmanip_b200 import ops
    torch.manual_seed(0)
    M, N, K = 512, 256, 128
    x, W = torch.randn(M, K), torch.randn(N, K) / K ** 0.5
    dpre = torch.randn(M, N) * 1e-12
    dW, db, dx = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV), torch.empty(M, K, device=DEV)
    ops.linear_backward_tc(cu(x), cu(W), cu(dpre), dW, db, dx, None, "fp32")
    want_dW, want_dx = dpre.double().T @ x.double(), dpre.double() @ W.double()
    assert float((dW.cpu().double() - want_dW).abs().max()) <= 1e-5 * float(want_dW.abs().max())
    assert float((dx.cpu().double() - want_dx).abs().max()) <= 1e-5 * float(want_dx.abs().max())
    assert float((db.cpu().double() - dpre.double().sum(0)).abs().max()) <= 1e-5 * float(dpre.double().sum(0).abs().max())


def test_tc_device_row_limit():
    """m_dev (device-side live row count, used by the critical-point backward): rows beyond it are neither read nor written."""
    from partmanip_b200 import ops
    torch.manual_seed(1)
    M, N, K, live = 700, 256, 128, 333
    x, W, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    x[live:] = float("nan")
    lim = torch.tensor([live], dtype=torch.int32, device=DEV)
    out = torch.full((M, N), -7.0, device=DEV)
    ops.linear_forward_tc(cu(x), cu(W), cu(b), "tanh", "fp32", out=out, m_dev=lim)
    want = torch.tanh(x[:live] @ W.T + b)
    assert close(out[:live].cpu(), want, 1e-4, 1e-4) and bool((out[live:] == -7.0).all())
    dpre = torch.randn(M, N)
    dpre[live:] = float("nan")
    dW, db, dx = torch.empty(N, K, device=DEV), torch.empty(N, device=DEV), torch.full((M, K), -7.0, device=DEV)
    ops.linear_backward_tc(cu(x), cu(W), cu(dpre), dW, db, dx, None, "fp32", m_dev=lim)
    wdW = dpre[:live].double().T @ x[:live].double()
    assert float((dW.cpu().double() - wdW).abs().max()) <= 1e-4 * float(wdW.abs().max())
    assert float((db.cpu() - dpre[:live].sum(0)).abs().max()) <= 1e-4 * float(dpre[:live].sum(0).abs().max())
    assert close(dx[:live].cpu(), dpre[:live] @ W, 1e-4, 1e-4) and bool((dx[live:] == -7.0).all())


@pytest.mark.parametrize("name,D,out,hid,act", [("mlp_actor.npz", 37, 7, [96, 160, 64], "tanh"), ("mlp_critic.npz", 53, 1, [96, 160, 64], "elu"),
                                                ("mlp_state_512.npz", 53, 10, [512, 512, 512], "tanh"),
                                                ("mlp_state_512_critic.npz", 53, 1, [512, 512, 512], "tanh")])
@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 1e-2), ("fp32_ffma", 1e-4)])
def test_mlp_reference_recordings(name, D, out, hid, act, precision, rtol):
    """The `MLP` plugin (network.py:27-54) on every arithmetic path against recordings of the unmodified reference class."""
    from partmanip_b200.algorithms.algo_utils.network import MLP
    g = load_golden(name)
    net = MLP(D, out, dict(hid_dim=hid, activation=act, precision=precision), 0)
    net.load_state_dict(sub(g, "w"))
    net.to(DEV)
    y = net(cu(g["x"]))
    assert close(y.detach().cpu(), g["y"], rtol, rtol), max_err(y.detach().cpu(), g["y"])
    y.square().sum().backward()
    for k, v in sub(g, "g").items():
        got = dict(net.named_parameters())[k].grad.cpu()
        tol = (rtol if precision != "bf16" else 3e-2) * float(v.abs().max()) + 1e-6
        assert float((got - v).abs().max()) <= tol, (k, float((got - v).abs().max()), tol)


def test_tc_device_row_limit_many_tiles():
    """Device-side row limit with more M tiles than resident CTAs: every CTA walks several tiles, the walk stops at the limit."""
    from partmanip_b200 import ops
    torch.manual_seed(2)
    M, N, K, live = 90000, 128, 64, 61111
    x, W, b = torch.randn(M, K), torch.randn(N, K) / K ** 0.5, torch.randn(N)
    x[live:] = float("nan")
    lim = torch.tensor([live], dtype=torch.int32, device=DEV)
    out = torch.full((M, N), -7.0, device=DEV)
    ops.linear_forward_tc(cu(x), cu(W), cu(b), "tanh", "fp32", out=out, m_dev=lim)
    want = torch.tanh(x[:live].double() @ W.double().T + b.double()).float()
    assert float((out[:live].cpu() - want).abs().max()) <= 1e-4
    assert bool((out[live:] == -7.0).all())
