"""Multi-GPU parity (run under torchrun, one rank per GPU): R ranks x E envs must reproduce ONE process with R*E envs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mgpu_parity.py

With n_minibatches == n_steps the global minibatch k is "time step k, all envs" in both layouts, so the optimiser steps
must agree up to fp32 summation order.  Each process first runs the single-process reference (before the process group
exists), then the sharded run; rank 0 prints the per-tensor deviations and exits non-zero on a violation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from partmanip_b200.algorithms import ppo
from partmanip_b200.envs import FakeVecEnv
from tests.helpers import ppo_cfg

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"


class _Logger:
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/pm_b200_mgpu"
    def info(self, d, it): pass


def run(E, shard, net, D, A, eps_all, init=None):
    torch.manual_seed(11)
    env = FakeVecEnv(E, D, A, dev, cloud=net["name"] == "PointNet", seed=99, shard=shard)
    cfg = ppo_cfg(E, net, device=dev)
    r = ppo(env, cfg, _Logger())
    if init is not None:
        r.actor_critic.load_state_dict(init)
    init_sd = {k: v.clone() for k, v in r.actor_critic.state_dict().items()}
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    sr, sw = shard if shard else (0, 1)
    eps = eps_all[:, sr * E:(sr + 1) * E].contiguous().to(dev)
    last_obs, last_values = r.collect(curr, None, eps=eps)
    r.storage.compute_returns(last_values, r.gamma, r.lam)
    r.update(1)
    torch.cuda.synchronize()
    return init_sd, {k: v.clone() for k, v in r.actor_critic.state_dict().items()}, dict(r.log_dict), r


cases = [("MLP", dict(name="MLP", hid_dim=[64, 64], activation="tanh"), 53, 10, 32),
         ("PointNet-fp32", dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False), 3072, 10, 8),
         ("PointNet-bf16", dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, precision="bf16"), 3072, 10, 8)]
refs = []
for name, net, D, A, E in cases:        # single-process references with E*world envs (no process group yet)
    g = torch.Generator().manual_seed(5)
    eps_all = torch.randn(8, E * world, A, generator=g)
    refs.append((eps_all,) + run(E * world, None, net, D, A, eps_all)[:3])

dist.init_process_group("nccl", device_id=torch.device(dev))
bad = 0
for (name, net, D, A, E), (eps_all, init_sd, ref_sd, ref_log) in zip(cases, refs):
    _, sd, log, r = run(E, (rank, world), net, D, A, eps_all, init=init_sd)
    lr, steps = 5e-5, 40
    worst = max(float((sd[k] - ref_sd[k]).abs().max()) for k in sd) / (lr * steps)
    # replicas must be bit-identical across ranks
    flat = torch.cat([v.reshape(-1) for v in sd.values()])
    other = flat.clone()
    dist.broadcast(other, 0)
    same = bool(torch.equal(flat, other))
    lo = {k: (float(log[k]), float(ref_log[k])) for k in ("Train/surrogate_loss", "Train/value_function_loss", "Train/kl", "Train/kl_update_count")}
    # fraction of the 40*lr Adam displacement; the 40-step trajectory is chaotic in the max-pool argmax (see
    # tests/test_gpu_pointnet_ppo.py), so the weights get a loose gate and the logged losses/KL the tight one
    # (fp32 = the split-operand tensor-core kernels since round 2: single-step gradients agree to 1e-5 instead of 1e-7, so the
    # chaotic 40-step trajectories separate a little more: measured 0.135 at world 2)
    tol = 0.25 if "bf16" not in name else 1.0
    ok = same and worst <= tol and all(abs(a - b) <= 1e-3 * max(1.0, abs(b)) + (1e-2 if "bf16" in name else 0) for a, b in lo.values())
    bad += 0 if ok else 1
    if rank == 0:
        print(f"\n[{name}] world={world} E/rank={E}: worst |dW|/(40 lr) = {worst:.4f}, replicas identical = {same}, logs (sharded, single) = {lo}  -> {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(1 if bad else 0)
