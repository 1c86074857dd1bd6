"""world_size-2 `gloo` tests (CPU) of the env-sharding contract in partmanip_b200/parallel.py:
two ranks that each hold half of the envs must take exactly the optimiser step a single process takes on the
concatenated batch.  The per-rank arithmetic is the CPU oracle (the CUDA kernels cannot run here); what is under test
is the collective plumbing the product uses on the GPU: summed gradients scaled by 1/(B*world), the summed
[surrogate, KL] pair that makes the KL-skip rank-consistent, and the two-stage global observation statistics."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    g = torch.Generator().manual_seed(7)
    B, D, A = 64, 37, 7
    p = {k: v for k, v in O.mlp_init(D, A, [32, 32], gen=g).items()}
    log_std = torch.full((A,), -0.7)
    obs = torch.randn(B, D, generator=g)
    adv = torch.randn(B, generator=g)
    mu_old = O.mlp_forward(p, obs) + torch.randn(B, A, generator=g) * 0.02       # the policy that collected the data
    act = torch.tanh(mu_old + O.policy_std(log_std) * torch.randn(B, A, generator=g))
    logp_old = O.gaussian_logp(mu_old, log_std, O.action_deactivation(act, 1.0))
    return p, log_std, obs, act, adv, mu_old, logp_old


def _actor_grads(p, log_std, obs, act, adv, mu_old, logp_old, inv_batch):
    """sum-reduced actor loss scaled by inv_batch (what pm_ppo_actor_loss does) -> grads, [sum surrogate, sum kl]."""
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ls = log_std.clone().requires_grad_(True)
    mu = O.mlp_forward(p, obs)
    logp = O.gaussian_logp(mu, ls, O.action_deactivation(act, 1.0))
    ratio = torch.exp(logp - logp_old)
    sur = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2))
    kl = O.kl_old_new(mu, ls.expand_as(mu), mu_old, log_std.expand_as(mu))
    (sur.sum() * inv_batch).backward()
    flat = torch.cat([v.grad.reshape(-1) for v in p.values()] + [ls.grad.reshape(-1)])
    return flat, torch.stack([sur.sum().detach(), kl.sum().detach()])


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from partmanip_b200 import parallel            # host-only module: importable without the CUDA library
    try:
        p, log_std, obs, act, adv, mu_old, logp_old = _problem()
        B = obs.shape[0] // world
        sl = slice(rank * B, (rank + 1) * B)
        # ---- gradients + KL-skip statistics
        flat, stats = _actor_grads(p, log_std, obs[sl], act[sl], adv[sl], mu_old[sl], logp_old[sl], parallel.inv_global_batch(B))
        parallel.all_reduce_sum_(flat)
        parallel.all_reduce_sum_(stats)
        # ---- global observation statistics, two stages (RMS.py:14-16 over the concatenated env batch)
        x = obs[sl]
        colsum = x.sum(0)
        parallel.all_reduce_sum_(colsum)
        count = parallel.global_count(B)
        sqdev = (x - colsum / count).pow(2).sum(0)
        parallel.all_reduce_sum_(sqdev)
        w = torch.zeros(4)
        parallel.broadcast_(w.add_(rank + 1.0))
        q.put((rank, flat, stats, colsum / count, sqdev / count, w, parallel.world(), parallel.rank()))
    finally:
        dist.destroy_process_group()


def _run_two_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    try:
        outs = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    finally:
        for pr in procs:
            pr.join(60)
            if pr.exitcode is None:          # results are in; a worker stuck in gloo's teardown is not a parity failure
                pr.terminate()
                pr.join(10)
    assert all(pr.exitcode in (0, None, -15) for pr in procs), [pr.exitcode for pr in procs]
    return outs


@pytest.mark.timeout(400)
def test_two_rank_update_equals_single_process_on_concatenated_batch():
    try:
        outs = _run_two_ranks()
    except Exception:                        # the probed port can be taken between the probe and the rendezvous: one retry on a new port
        outs = _run_two_ranks()
    p, log_std, obs, act, adv, mu_old, logp_old = _problem()
    want_flat, want_stats = _actor_grads(p, log_std, obs, act, adv, mu_old, logp_old, 1.0 / obs.shape[0])
    new_mean = obs.mean(0)
    want_var = (obs - new_mean).pow(2).mean(0)
    for rank, flat, stats, mean, var, w, ws, rk in outs:
        assert ws == 2 and rk == rank
        assert float((flat - want_flat).abs().max()) <= 1e-5 * float(want_flat.abs().max()), float((flat - want_flat).abs().max())
        assert torch.allclose(stats, want_stats, rtol=1e-5, atol=1e-6)
        assert torch.allclose(mean, new_mean, rtol=1e-5, atol=1e-6) and torch.allclose(var, want_var, rtol=1e-5, atol=1e-6)
        assert torch.equal(w, torch.ones(4))                     # broadcast from rank 0
    # both ranks hold bit-identical reduced buffers => identical Adam steps and an identical KL-skip decision
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])


def test_single_process_helpers_are_noops():
    from partmanip_b200 import parallel
    t = torch.arange(4.0)
    assert parallel.world() == 1 and parallel.rank() == 0
    assert torch.equal(parallel.all_reduce_sum_(t.clone()), t) and torch.equal(parallel.broadcast_(t.clone()), t)
    assert parallel.inv_global_batch(2048) == 1.0 / 2048 and parallel.global_count(4096) == 4096.0
