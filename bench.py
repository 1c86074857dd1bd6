#!/usr/bin/env python
"""bench.py — env·steps/sec of (encoder + PPO update) on synthetic open_drawer-shaped point clouds.

One "step" = one full PPO iteration of the hot path with a zero-cost synthetic env (BASELINE.md §3 workload):
T=8 x [state-norm -> random_act_cri -> add_transitions] + cri + compute_returns + update (5 epochs, actor phase then
critic phase, 16 minibatches of 2048 at E=4096).  value = E*T*world / time per iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--envs E] [--precision bf16|fp32]

Under torchrun (N>1) every rank owns E envs (weak scaling); gradients / KL sums / obs statistics are all-reduced.
`--impl reference` times the reference's algorithm on the host cores (the CPU oracle port — the reference is
Python/PyTorch and cannot be vendored) on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS, N_PTS, CH, ACT = 8, 1024, 3, 10
F_POINTNET = 2 * N_PTS * (CH * 128 + 128 * 256 + 256 * 512) + 2 * (512 * 128 + 128 * 32 + 32 * ACT)   # 336.5 MFLOP
ENC_FLOPS_PER_CLOUD = 2 * N_PTS * (CH * 128 + 128 * 256 + 256 * 512)


def ppo_cfg(E, device, precision):
    """Shipped cfg/algos/ppo.yaml hyper-parameters + the PointNet keys the reference's YAML lacks."""
    return dict(
        num_envs=E, obs_mode="obs", succ_value=None, max_iterations=10 ** 9, n_steps=T_STEPS, n_updates=5,
        n_minibatches=8, device=device, eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9,
        test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=5e-5, desired_kl=0.1,
        epsilon_clip=0.2, gamma=0.99, lam=0.95, sampler="sequential", resume=None,
        cuda_graph=os.environ.get("PM_CUDA_GRAPH", "1") == "1", cuda_graph_multi_rank=os.environ.get("PM_CUDA_GRAPH_MULTI", "1") == "1",
        tricks=dict(mini_adv_norm=False, whole_adv_norm=False, use_state_norm=True, use_clipped_value_loss=False,
                    use_grad_clip=True, max_grad_norm=0.5),
        model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0,
                   network=dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False,
                                point_num=N_PTS, precision=precision)),
    )


class _Logger:
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/pm_b200_bench"

    def info(self, d, it):
        pass


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = max([float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()] or [0.0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 9 and r[5 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def cpu_port_iteration(E, threads, iters=1, warm=0, device="cpu"):
    """The reference's algorithm (CPU oracle port, unmodified math) for one PPO iteration at E envs; returns
    env·steps/s on the host cores.  Only bench.py's cpu_baseline / --impl reference legs call this.
    device="cuda:0" runs the same eager-PyTorch restatement on the GPU (`--impl torch_gpu`: the reference's PyTorch-GPU
    path — fp32, TF32 off, eager — as the denominator of north_star's ">= 10x" target; not a driver arm)."""
    import torch
    from oracle import ppo_oracle as O
    torch.set_num_threads(threads)
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(1234)
    D = N_PTS * CH
    cfg = ppo_cfg(E, "cpu", "fp32")
    net = cfg["model"]["network"]
    dv = lambda t: t.to(device)
    actor = {k: dv(v) for k, v in O.pointnet_init(D, ACT, gen=g).items()}
    critic = {k: dv(v) for k, v in O.pointnet_init(D, 1, gen=g).items()}
    log_std = dv(torch.full((ACT,), float(torch.log(torch.tensor(0.5)))))
    opt_a = O.AdamState({**actor, "log_std": log_std}, cfg["lr"])
    opt_c = O.AdamState(critic, cfg["lr"])
    rs = O.RunningStats(D)
    rs.mean, rs.S, rs.std = dv(rs.mean), dv(rs.S), dv(rs.std)

    def obs():
        pc = torch.rand(E, N_PTS, CH, generator=g)
        pc[..., :2] = pc[..., :2] * 2 - 1
        pc[..., 2] = pc[..., 2] * 2 + 0.05
        pc[torch.rand(E, N_PTS, generator=g) < 0.1] = 0.0
        return dv(pc.reshape(E, D))

    pool = [obs() for _ in range(T_STEPS + 1)]
    rnd = lambda *sh: dv(torch.randn(*sh, generator=g))
    times = []
    for it in range(warm + iters):
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        buf = {k: [] for k in ("obs", "actions", "values", "logp", "mu", "sigma", "rew", "done")}
        with torch.no_grad():
            cur = rs.normalize(pool[0].clone(), True)
            for t in range(T_STEPS):
                mu = O.pointnet_forward(actor, cur)
                a, lp = O.policy_sample(mu, log_std, rnd(E, ACT), 1.0)
                v = O.pointnet_forward(critic, cur)
                for k, x in zip(("obs", "actions", "values", "logp", "mu", "sigma"), (cur, a, v, lp[:, None], mu, log_std.repeat(E, 1))):
                    buf[k].append(x)
                buf["rew"].append(rnd(E, 1))
                buf["done"].append(dv(torch.rand(E, 1, generator=g) < 0.05))
                cur = rs.normalize(pool[t + 1].clone(), True)
            last = O.pointnet_forward(critic, cur)
            st = {k: torch.stack(v) for k, v in buf.items()}
            ret, adv = O.gae(st["rew"], st["values"], st["done"], torch.zeros_like(st["done"]), last, 0.99, 0.95, None)
        flat = lambda x: x.reshape(-1, x.shape[-1])
        b = dict(obs=flat(st["obs"]), actions=flat(st["actions"]), values=flat(st["values"]), returns=flat(ret),
                 logp=flat(st["logp"]), adv=flat(adv), mu=flat(st["mu"]), sigma=flat(st["sigma"]))
        O.ppo_update(actor, critic, log_std, opt_a, opt_c, b, cfg, "PointNet", net)
        if on_gpu:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = sum(times[warm:]) / max(iters, 1)
    return E * T_STEPS / t, t


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    E = args.ref_envs
    v, t = cpu_port_iteration(E, threads, iters=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec (encoder+PPO update)", "value": v, "unit": "env*steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ppo + open_drawer shapes, PointNet, {N_PTS} pts x {CH} ch, A={ACT}, T={T_STEPS}, 5 epochs; "
                               f"CPU sample E={E} envs per step (minibatch {min(E * T_STEPS // 8, 2048)})"},
        "cpu_baseline": {"value": v, "unit": "env*steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} full PPO iterations at E={E} (intensive metric; reference is Python/PyTorch, "
                                   f"timed via the CPU oracle port with torch {torch.__version__} on {threads} threads)"},
        "e2e": {"value": v, "unit": "env*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--precision", default=None, choices=[None, "bf16", "fp32"])
    ap.add_argument("--ref-envs", type=int, default=8)
    ap.add_argument("--cpu-baseline-envs", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "torch_gpu":          # context only: eager PyTorch fp32 (TF32 off) on cuda:0, the reference's GPU path
        if rank == 0:
            v, t = cpu_port_iteration(args.envs, os.cpu_count() or 1, iters=args.steps, warm=args.warmup, device="cuda:0")
            print(json.dumps({"impl": "torch_gpu", "metric": "env_steps_per_sec (encoder+PPO update)", "value": v,
                              "unit": "env*steps/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": t * 1e3, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": f"eager PyTorch restatement of the reference path on cuda:0, E={args.envs} envs x "
                                                     f"{N_PTS} pts, fp32, TF32 off"}}), flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from partmanip_b200 import _lib, ops
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv

    precision = args.precision or ("bf16" if _lib.lib.pm_has_tcgen05() else "fp32")
    E, D = args.envs, N_PTS * CH
    torch.manual_seed(1234)

    def make(host):
        env = FakeVecEnv(E, D, ACT, dev, cloud=True, channels=CH, pool=16, seed=1234 + rank, host=host)
        runner = ppo(env, ppo_cfg(E, dev, precision), _Logger())
        return env, runner

    def iteration(runner, curr):
        last_obs, last_values = runner.collect(curr, None)
        runner.storage.compute_returns(last_values, runner.gamma, runner.lam)
        runner.update(1)
        runner.storage.clear()
        return ops.copy_rows(last_obs, runner.storage.obs_slot())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(runner, env, steps, warmup, sample_clocks):
        curr = runner._ingest(env.reset()["obs"], runner.storage.obs_slot())
        for _ in range(warmup):
            curr = iteration(runner, curr)
        barrier()
        calls0 = ops.launch_count()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            curr = iteration(runner, curr)
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, ops.launch_count() - calls0, clocks

    # ---- device-resident arm (value)
    env, runner = make(host=False)
    ms_step, launches, clocks = timed(runner, env, args.steps, args.warmup, rank == 0)
    value = E * T_STEPS * world / (ms_step * 1e-3)

    # ---- dominant kernel: PointNet encoder forward on one 2048-cloud minibatch, timed live with CUDA events
    B = min(2048, E * T_STEPS)
    obs_flat = runner.storage.observations.view(-1, D)
    nslices = obs_flat.shape[0] // B
    actor = runner.actor_critic.actor
    enc = actor.runner.enc_params()
    feat = torch.empty(B, 512, device=dev)
    am = torch.empty(B, 512, device=dev, dtype=torch.int32)
    for i in range(3):
        ops.pointnet_encode_forward(obs_flat[(i % nslices) * B:(i % nslices + 1) * B], N_PTS, CH, enc, "tanh", precision, feat, None, am, None)
    torch.cuda.synchronize()
    reps = 16
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for i in range(reps):   # successive launches read different 25 MB slices of the 403 MB buffer (> L2)
        ops.pointnet_encode_forward(obs_flat[(i % nslices) * B:(i % nslices + 1) * B], N_PTS, CH, enc, "tanh", precision, feat, None, am, None)
    k1.record()
    torch.cuda.synchronize()
    enc_ms = k0.elapsed_time(k1) / reps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    achieved_tf = B * ENC_FLOPS_PER_CLOUD / (enc_ms * 1e-3) / 1e12
    # encoder-forward launches per iteration: T rollout steps x (actor + critic) + the last-value critic pass, each over
    # E clouds (E/B launch-equivalents of B clouds), + 2 networks x 5 epochs x (E*T/B) minibatches
    enc_equiv_per_iter = (T_STEPS * 2 + 1) * (E / B) + 2 * 5 * (E * T_STEPS // B)
    # companion kernel: encoder backward on the same minibatch
    grads = [torch.empty_like(t) for t in enc]
    dfeat = torch.randn(B, 512, device=dev) * 0.01
    for i in range(2):
        ops.pointnet_encode_backward(obs_flat[:B], N_PTS, CH, enc, "tanh", dfeat, am, grads, precision=precision)
    k0.record()
    for i in range(reps):
        ops.pointnet_encode_backward(obs_flat[(i % nslices) * B:(i % nslices + 1) * B], N_PTS, CH, enc, "tanh", dfeat, am, grads,
                                     precision=precision)
    k1.record()
    torch.cuda.synchronize()
    bwd_ms = k0.elapsed_time(k1) / reps
    # DRAM traffic of the forward kernel per launch from the committed `ncu --set full` capture
    # (profiles/r01_fwd_r01c_metrics.txt: dram__bytes_read.sum 25.57 MB, dram__bytes_write.sum ~0: the 8 MB of outputs
    # stay in the 126 MB L2) — the algorithmic input is B*N*C*4 = 25.17 MB
    traffic = 25571072 if (precision == "bf16" and B == 2048) else None
    roofline = {"bound": "tensor", "kernel": "pointnet encoder forward (%s), %d clouds x %d pts per launch" % (precision, B, N_PTS),
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst), of measured" if peaks else "fallback 1.59 PF, of fallback",
                "traffic": traffic, "algorithmic_bytes": B * N_PTS * CH * 4, "algorithmic_flops": B * ENC_FLOPS_PER_CLOUD,
                "ms_per_launch": enc_ms, "share_of_step": enc_ms * enc_equiv_per_iter / ms_step,
                "hbm_frac_for_transparency": (B * N_PTS * CH * 4 / (enc_ms * 1e-3) / 1e9) / peaks.get("hbm_gbs", 6650.0),
                "companion": {"kernel": "pointnet encoder backward (%s), same minibatch" % precision, "ms_per_launch": bwd_ms,
                              "share_of_step": bwd_ms * 2 * 5 * (E * T_STEPS // B) / ms_step}}

    # ---- end-to-end arm: host-resident observations, H2D copy of every step's obs + D2H of the actions inside the timed region
    e2e = None
    if not args.no_e2e:
        runner.release_graph()
        del runner, env
        torch.cuda.empty_cache()
        env_h, runner_h = make(host=True)
        runner = runner_h
        e2e_warm = 2                                   # eager + graph-capture iterations stay outside the timed region
        ms_e2e, _, _ = timed(runner_h, env_h, max(2, args.steps // 2), e2e_warm, False)
        iters_run = max(2, args.steps // 2) + e2e_warm
        e2e = {"value": E * T_STEPS * world / (ms_e2e * 1e-3), "unit": "env*steps/s",
               "h2d_bytes_per_step": int(env_h.h2d_bytes / iters_run), "d2h_bytes_per_step": int(env_h.d2h_bytes / iters_run),
               "ms_per_step": ms_e2e}

    # ---- CPU baseline (oracle port) on rank 0, N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        Ec = args.cpu_baseline_envs
        v, t = cpu_port_iteration(Ec, threads, iters=1, warm=0)
        cpu_baseline = {"value": v, "unit": "env*steps/s", "cores": threads, "kind": "port",
                        "sample": f"1 full PPO iteration at E={Ec} envs x {N_PTS} pts ({t:.1f} s) with the CPU oracle port"}

    if rank == 0:
        line = {
            "metric": "env_steps_per_sec (encoder+PPO update)", "value": value, "unit": "env*steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"ppo + open_drawer shapes: {E} envs/GPU x {N_PTS}-pt cloud x {CH} ch, PointNet encoder, A={ACT}, "
                                   f"T={T_STEPS}, 5 epochs x {E * T_STEPS // B} minibatches of {B}; one step = one PPO iteration",
                       "envs_per_gpu": E, "precision": precision, "cache": "inputs 403 MB rollout buffer > 126 MB L2",
                       "parallelism": f"env-sharded x{world}" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # leave without tearing NCCL down: captured graphs hold NCCL work objects and destroy_process_group() can block on them
        runner.release_graph()
        barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
