#!/usr/bin/env python
"""bench.py — env·steps/sec of (encoder + PPO update) on synthetic open_drawer-shaped point clouds.

One "step" = one full PPO iteration of the hot path with a zero-cost synthetic env (BASELINE.md §3 workload):
T=8 x [state-norm -> random_act_cri -> add_transitions] + cri + compute_returns + update (5 epochs, actor phase then
critic phase, 16 minibatches of 2048 at E=4096).  value = E*T*world / time per iteration.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|torch_gpu] [--envs E] [--precision bf16|fp32]

The headline line is the bf16 (tcgen05) mode.  At N=1 the same line also carries
  modes.fp32      the same workload at the REFERENCE's precision (1e-4 gate; split-operand tcgen05 kernels),
  gpu_baseline    the reference's path as eager PyTorch on the same GPU (fp32, TF32 off) — north_star's ">= 10x" denominator,
  configs         BASELINE config 4 (DAgger, 2048 envs x 2048 pts, random sampler) and the shipped state-MLP PPO config,
  cpu_baseline    the CPU oracle port on the host cores.
Under torchrun (N>1) every rank owns E envs (weak scaling); gradients / KL sums / obs statistics are all-reduced, and the
line carries `multi_gpu_parity` (replicas bit-identical across ranks, losses of N ranks x E envs vs ONE process with N*E envs).
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
_OUT = sys.stdout

T_STEPS, N_PTS, CH, ACT = 8, 1024, 3, 10
F_POINTNET = 2 * N_PTS * (CH * 128 + 128 * 256 + 256 * 512) + 2 * (512 * 128 + 128 * 32 + 32 * ACT)   # 336.5 MFLOP
ENC_FLOPS_PER_CLOUD = 2 * N_PTS * (CH * 128 + 128 * 256 + 256 * 512)
# encoder backward, per (cloud, channel) ROW the kernel executes: recompute layer 2 (128x256), dH1 = dPre2.W2 (256x128),
# dW2 += dPre2^T.H1 (256x128) as MMAs + the per-row dW3 / dPre2 epilogue (2 x 256 FMAs)
BWD_FLOPS_PER_ROW = 2 * (3 * 128 * 256 + 2 * 256)


def ppo_cfg(E, device, precision, net=None, **over):
    """Shipped cfg/algos/ppo.yaml hyper-parameters + the PointNet keys the reference's YAML lacks."""
    cfg = dict(
        num_envs=E, obs_mode="obs", succ_value=None, max_iterations=10 ** 9, n_steps=T_STEPS, n_updates=5,
        n_minibatches=8, device=device, eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9,
        test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=5e-5, desired_kl=0.1,
        epsilon_clip=0.2, gamma=0.99, lam=0.95, sampler="sequential", resume=None,
        cuda_graph=os.environ.get("PM_CUDA_GRAPH", "1") == "1", cuda_graph_multi_rank=os.environ.get("PM_CUDA_GRAPH_MULTI", "1") == "1",
        fused_step=os.environ.get("PM_FUSED_STEP", "1") == "1", overlap_phases=os.environ.get("PM_OVERLAP", "1") == "1",
        tricks=dict(mini_adv_norm=False, whole_adv_norm=False, use_state_norm=True, use_clipped_value_loss=False,
                    use_grad_clip=True, max_grad_norm=0.5),
        model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0,
                   network=net or dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False,
                                       point_num=N_PTS, precision=precision)),
    )
    cfg.update(over)
    return cfg


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


def cpu_port_iteration(E, threads, iters=1, warm=0, device="cpu", budget_s=None, mlp=None):
    """The reference's algorithm (CPU oracle port, unmodified math) for one PPO iteration at E envs; returns
    (env·steps/s, seconds per iteration, timed iterations).  Only bench.py's cpu_baseline / --impl reference / gpu_baseline legs
    call this.  device="cuda:0" runs the same eager-PyTorch restatement on the GPU: the reference's PyTorch-GPU path — fp32,
    TF32 off, eager — the denominator of north_star's ">= 10x" target.  budget_s bounds the timed region (at least one
    timed iteration always runs).  mlp = (obs_dim, hid_dims): the state policy of cfg/algos/ppo.yaml instead of the PointNet encoder."""
    import torch
    from oracle import ppo_oracle as O
    torch.set_num_threads(threads)
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(1234)
    dv = lambda t: t.to(device)
    if mlp is None:
        D, kind = N_PTS * CH, "PointNet"
        cfg = ppo_cfg(E, "cpu", "fp32")
        net = cfg["model"]["network"]
        actor = {k: dv(v) for k, v in O.pointnet_init(D, ACT, gen=g).items()}
        critic = {k: dv(v) for k, v in O.pointnet_init(D, 1, gen=g).items()}
        fwd = lambda p, x: O.pointnet_forward(p, x)
    else:
        D, kind = mlp[0], "MLP"
        net = dict(name="MLP", hid_dim=list(mlp[1]), activation="tanh")
        cfg = ppo_cfg(E, "cpu", "fp32", net=net)
        actor = {k: dv(v) for k, v in O.mlp_init(D, ACT, list(mlp[1]), gen=g).items()}
        critic = {k: dv(v) for k, v in O.mlp_init(D, 1, list(mlp[1]), gen=g).items()}
        fwd = lambda p, x: O.mlp_forward(p, x, "tanh")
    log_std = dv(torch.full((ACT,), float(torch.log(torch.tensor(0.5)))))
    opt_a = O.AdamState({**actor, "log_std": log_std}, cfg["lr"])
    opt_c = O.AdamState(critic, cfg["lr"])
    rs = O.RunningStats(D)
    rs.mean, rs.S, rs.std = dv(rs.mean), dv(rs.S), dv(rs.std)

    def obs():
        if mlp is not None:
            return dv(torch.randn(E, D, generator=g))
        pc = torch.rand(E, N_PTS, CH, generator=g)
        pc[..., :2] = pc[..., :2] * 2 - 1
        pc[..., 2] = pc[..., 2] * 2 + 0.05
        pc[torch.rand(E, N_PTS, generator=g) < 0.1] = 0.0
        return dv(pc.reshape(E, D))

    pool = [obs() for _ in range(T_STEPS + 1)]
    rnd = lambda *sh: dv(torch.randn(*sh, generator=g))
    times = []
    t_start = None
    for it in range(warm + iters):
        if it == warm:
            t_start = time.perf_counter()
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        buf = {k: [] for k in ("obs", "actions", "values", "logp", "mu", "sigma", "rew", "done")}
        with torch.no_grad():
            cur = rs.normalize(pool[0].clone(), True)
            for t in range(T_STEPS):
                mu = fwd(actor, cur)
                a, lp = O.policy_sample(mu, log_std, rnd(E, ACT), 1.0)
                v = fwd(critic, cur)
                for k, x in zip(("obs", "actions", "values", "logp", "mu", "sigma"), (cur, a, v, lp[:, None], mu, log_std.repeat(E, 1))):
                    buf[k].append(x)
                buf["rew"].append(rnd(E, 1))
                buf["done"].append(dv(torch.rand(E, 1, generator=g) < 0.05))
                cur = rs.normalize(pool[t + 1].clone(), True)
            last = fwd(critic, cur)
            st = {k: torch.stack(v) for k, v in buf.items()}
            ret, adv = O.gae(st["rew"], st["values"], st["done"], torch.zeros_like(st["done"]), last, 0.99, 0.95, None)
        flat = lambda x: x.reshape(-1, x.shape[-1])
        b = dict(obs=flat(st["obs"]), actions=flat(st["actions"]), values=flat(st["values"]), returns=flat(ret),
                 logp=flat(st["logp"]), adv=flat(adv), mu=flat(st["mu"]), sigma=flat(st["sigma"]))
        O.ppo_update(actor, critic, log_std, opt_a, opt_c, b, cfg, kind, net)
        if on_gpu:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and it >= warm and time.perf_counter() - t_start > budget_s:
            break
    timed = times[warm:]
    t = sum(timed) / max(len(timed), 1)
    return E * T_STEPS / t, t, len(timed)


def run_reference(args, rank):
    """The driver's reference arm: the oracle port on all host cores; every step = one full PPO iteration at E_ref envs (an
    intensive-metric sample of the E=4096 workload: same clouds, same hyper-parameters, minibatch min(E*T/8, 2048)).
    E_ref is the largest of 64..512 that keeps the whole --steps/--warmup run within the budget (SURVEY §8d asks for 512)."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    E = args.ref_envs
    if E <= 0:
        _, t16, _ = cpu_port_iteration(16, threads, iters=1, warm=1)        # calibrate: seconds per env of a warm iteration
        per_env = t16 / 16
        E = 64
        for cand in (128, 256, 512):
            if (args.steps + args.warmup) * per_env * cand <= args.ref_budget_s:
                E = cand
    v, t, n = cpu_port_iteration(E, threads, iters=args.steps, warm=args.warmup)
    line = {
        "impl": "reference", "metric": "env_steps_per_sec (encoder+PPO update)", "value": v, "unit": "env*steps/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ppo + open_drawer shapes, PointNet, {N_PTS} pts x {CH} ch, A={ACT}, T={T_STEPS}, 5 epochs; "
                               f"CPU sample E={E} envs per step (minibatch {min(E * T_STEPS // 8, 2048)})"},
        "cpu_baseline": {"value": v, "unit": "env*steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n} full PPO iterations at E={E} after {args.warmup} warm-up (intensive metric; reference is "
                                   f"Python/PyTorch, timed via the CPU oracle port with torch {torch.__version__} on {threads} threads)"},
        "e2e": {"value": v, "unit": "env*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ multi-GPU parity evidence
PARITY_E = 8


def _parity_run(E, shard, dev, eps_all, init=None):
    """One PPO iteration (bf16 PointNet) on a shard of a fixed synthetic global batch; returns (init weights, final weights, log)."""
    import torch
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    torch.manual_seed(11)
    env = FakeVecEnv(E, N_PTS * CH, ACT, dev, cloud=True, seed=99, shard=shard)
    r = ppo(env, ppo_cfg(E, dev, "bf16", max_iterations=1), _Logger())
    if init is not None:
        r.actor_critic.load_state_dict(init)
    init_sd = {k: v.clone() for k, v in r.actor_critic.state_dict().items()}
    curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())
    sr, _ = shard if shard else (0, 1)
    last_obs, last_values = r.collect(curr, None, eps=eps_all[:, sr * E:(sr + 1) * E].contiguous().to(dev))
    r.storage.compute_returns(last_values, r.gamma, r.lam)
    r.update(1)
    torch.cuda.synchronize()
    out = {k: v.clone() for k, v in r.actor_critic.state_dict().items()}
    log = {k: float(r.log_dict[k]) for k in ("Train/surrogate_loss", "Train/value_function_loss", "Train/kl", "Train/kl_update_count")}
    r.release_graph()
    return init_sd, out, log


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch_gpu"])
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--precision", default=None, choices=[None, "bf16", "fp32", "fp32_ffma"])
    ap.add_argument("--ref-envs", type=int, default=0, help="reference arm: envs per step (0 = calibrate, 64..512)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0)
    ap.add_argument("--cpu-baseline-envs", type=int, default=64)
    ap.add_argument("--gpu-baseline-iters", type=int, default=10)
    ap.add_argument("--gpu-baseline-budget-s", type=float, default=75.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip modes.fp32 / configs (DAgger, state-MLP)")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: whatever the algorithm classes print ("load teacher ckpt ...") goes to stderr
    global _OUT
    _OUT, sys.stdout = sys.stdout, sys.stderr
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.impl == "torch_gpu":          # context only: eager PyTorch fp32 (TF32 off) on cuda:0, the reference's GPU path
        if rank == 0:
            v, t, n = cpu_port_iteration(args.envs, os.cpu_count() or 1, iters=args.steps, warm=args.warmup, device="cuda:0")
            print(json.dumps({"impl": "torch_gpu", "metric": "env_steps_per_sec (encoder+PPO update)", "value": v,
                              "unit": "env*steps/s", "n_gpus": 1, "steps": n, "warmup": args.warmup,
                              "ms_per_step": t * 1e3, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": f"eager PyTorch restatement of the reference path on cuda:0, E={args.envs} envs x "
                                                     f"{N_PTS} pts, fp32, TF32 off"}}), file=_OUT, flush=True)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    from partmanip_b200 import _lib, ops
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv

    # ---- multi-GPU parity, part 1: ONE process with world*E envs, before the process group exists (every rank, own GPU)
    parity_ref = None
    if world > 1:
        g = torch.Generator().manual_seed(5)
        eps_all = torch.randn(T_STEPS, PARITY_E * world, ACT, generator=g)
        parity_ref = (eps_all,) + _parity_run(PARITY_E * world, None, dev, eps_all)
        dist.init_process_group("nccl", device_id=torch.device(dev))

    precision = args.precision or ("bf16" if _lib.lib.pm_has_tcgen05() else "fp32")
    E, D = args.envs, N_PTS * CH
    torch.manual_seed(1234)

    def make(host, prec=precision):
        env = FakeVecEnv(E, D, ACT, dev, cloud=True, channels=CH, pool=16, seed=1234 + rank, host=host)
        runner = ppo(env, ppo_cfg(E, dev, prec), _Logger())
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

    def kernel_ms(fn, reps=16, warm=3):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(reps):
            fn(i)
        k1.record()
        torch.cuda.synchronize()
        return k0.elapsed_time(k1) / reps

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst), of measured" if peaks else "fallback 1.59 PF, of fallback"

    def encoder_roofline(runner, prec, ms_step):
        """Dominant kernel (encoder forward) and its companion (encoder backward) on one 2048-cloud minibatch, timed live with CUDA
        events; successive launches read different 25 MB slices of the 403 MB rollout buffer (> L2)."""
        B = min(2048, E * T_STEPS)
        obs_flat = runner.storage.observations.view(-1, D)
        nsl = obs_flat.shape[0] // B
        enc = runner.actor_critic.actor.runner.enc_params()
        feat = torch.empty(B, 512, device=dev)
        am = torch.empty(B, 512, device=dev, dtype=torch.int32)
        sl = lambda i: obs_flat[(i % nsl) * B:(i % nsl + 1) * B]
        reps = 16 if prec != "fp32_ffma" else 3
        enc_ms = kernel_ms(lambda i: ops.pointnet_encode_forward(sl(i), N_PTS, CH, enc, "tanh", prec, feat, None, am, None), reps)
        grads = [torch.empty_like(t) for t in enc]
        dfeat = torch.randn(B, 512, device=dev) * 0.01
        bwd_ms = kernel_ms(lambda i: ops.pointnet_encode_backward(sl(i), N_PTS, CH, enc, "tanh", dfeat, am, grads, precision=prec), reps)
        srt = am.sort(dim=1)[0]
        uniq = int((srt[:, 1:] != srt[:, :-1]).sum()) + B                 # unique critical points of the minibatch
        achieved = B * ENC_FLOPS_PER_CLOUD / (enc_ms * 1e-3) / 1e12
        # what the tensor pipe executes per algorithmic flop: 1 MMA (bf16) or 3 (split fp16) per product of layers 2-3
        mma_mult = {"bf16": 1, "fp32": 3}.get(prec)
        # encoder-forward launch-equivalents per iteration: T rollout steps x (actor + critic) + the last-value critic pass, each
        # over E clouds (E/B launches of B clouds), + 2 networks x 5 epochs x (E*T/B) minibatches
        fwd_per_iter = (T_STEPS * 2 + 1) * (E / B) + 2 * 5 * (E * T_STEPS // B)
        bwd_per_iter = 2 * 5 * (E * T_STEPS // B)
        rows = B * 512
        roof = {"bound": "tensor", "kernel": "pointnet encoder forward (%s), %d clouds x %d pts per launch" % (prec, B, N_PTS),
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "peak_source": peak_src,
                "traffic": 25563136 if (prec == "bf16" and B == 2048) else None,
                "traffic_source": "profiles/r02z_encoder_fwd_tc_ncu_metrics.txt: ncu --set full capture (dram__bytes_read.sum + write.sum of one launch); "
                                  "not measured live by bench.py" if (prec == "bf16" and B == 2048) else None,
                "algorithmic_bytes": B * N_PTS * CH * 4, "algorithmic_flops": B * ENC_FLOPS_PER_CLOUD,
                "ms_per_launch": enc_ms, "share_of_step": enc_ms * fwd_per_iter / ms_step,
                "hbm_frac_for_transparency": (B * N_PTS * CH * 4 / (enc_ms * 1e-3) / 1e9) / peaks.get("hbm_gbs", 6650.0)}
        if mma_mult and mma_mult > 1:
            roof["executed_mma_tflops"] = achieved * mma_mult
            roof["executed_frac_of_bf16_peak"] = achieved * mma_mult / peak_tf
            roof["note"] = "frac counts ALGORITHMIC flops; every product runs as %d fp16 MMAs (hi.hi + lo.hi + hi.lo)" % mma_mult
        comp = {"kernel": "pointnet encoder backward (%s), same minibatch" % prec, "ms_per_launch": bwd_ms,
                "share_of_step": bwd_ms * bwd_per_iter / ms_step, "unit": "TFLOP/s", "peak": peak_tf,
                "unique_critical_points": uniq, "rows_executed": rows if prec == "bf16" else uniq,
                "algorithmic_flops_unique_points": uniq * BWD_FLOPS_PER_ROW}
        uniq_tf = uniq * BWD_FLOPS_PER_ROW / (bwd_ms * 1e-3) / 1e12
        comp["achieved_unique"] = uniq_tf
        comp["frac_unique"] = uniq_tf / peak_tf
        if prec == "bf16":       # the fused kernel treats every (cloud, channel) pair as a row: rows/uniq x duplicate work
            ex_tf = rows * BWD_FLOPS_PER_ROW / (bwd_ms * 1e-3) / 1e12
            comp["achieved_executed"] = ex_tf
            comp["frac_executed"] = ex_tf / peak_tf
        roof["companion"] = comp
        return roof

    # ---- device-resident arm (value)
    env, runner = make(host=False)
    ms_step, launches, clocks = timed(runner, env, args.steps, args.warmup, rank == 0)
    value = E * T_STEPS * world / (ms_step * 1e-3)
    roofline = encoder_roofline(runner, precision, ms_step)

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
               "ms_per_step": ms_e2e,
               "h2d": "every env step's observation is copied from pinned host memory inside the timed region, double-buffered on a copy "
                      "stream so that the copy of step t+1 overlaps the learner's compute on step t (the synthetic pool does not depend "
                      "on the actions); actions are copied back to the host every step"}
        del env_h, runner_h

    # ---- multi-GPU parity, part 2: world ranks x PARITY_E envs against the single-process run
    multi_gpu_parity = None
    if world > 1:
        runner.release_graph()
        eps_all, init_sd, ref_sd, ref_log = parity_ref
        _, sd, log = _parity_run(PARITY_E, (rank, world), dev, eps_all, init=init_sd)
        flat = torch.cat([v.reshape(-1) for v in sd.values()])
        ref0 = flat.clone()
        dist.broadcast(ref0, 0)
        same = torch.tensor([1 if torch.equal(flat, ref0) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        lr, steps40 = 5e-5, 40
        worst = max(float((sd[k] - ref_sd[k]).abs().max()) for k in sd) / (lr * steps40)
        multi_gpu_parity = {
            "workload": f"one PPO iteration, PointNet bf16, {world} ranks x {PARITY_E} envs vs ONE process with {world * PARITY_E} envs "
                        "(same synthetic global batch, same initial weights, same exploration noise)",
            "replicas_bit_identical_across_ranks": bool(int(same)),
            "loss_vs_1proc": {k: {"ranks": log[k], "one_process": ref_log[k]} for k in log},
            "max_rel_diff_losses": max(abs(log[k] - ref_log[k]) / max(1.0, abs(ref_log[k])) for k in log),
            "worst_weight_diff_over_40lr": worst,
        }

    modes, configs, gpu_baseline, cpu_baseline = None, None, None, None
    extras = rank == 0 and world == 1 and not args.no_extras
    if extras:
        runner.release_graph()
        del runner
        torch.cuda.empty_cache()
        # ---- the reference's precision: same workload, precision fp32 (split-operand tcgen05 kernels, 1e-4 parity gate)
        if precision == "bf16":
            env32, r32 = make(host=False, prec="fp32")
            ms32, _, _ = timed(r32, env32, max(2, min(args.steps, 3)), 2, False)
            roof32 = encoder_roofline(r32, "fp32", ms32)
            modes = {"fp32": {"value": E * T_STEPS / (ms32 * 1e-3), "unit": "env*steps/s", "ms_per_step": ms32, "dtype": "f32",
                              "parity_gate": "1e-4 (north_star fp32)", "roofline": roof32}}
            r32.release_graph()
            del env32, r32
            torch.cuda.empty_cache()
        configs = {}
        # ---- BASELINE config 4: DAgger, state expert -> vision student, 2048 envs x 2048-pt clouds, random sampler
        try:
            configs["dagger_2048x2048"] = bench_dagger(dev, precision)
        except Exception as e:  # pragma: no cover - keep the headline line alive
            configs["dagger_2048x2048"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        # ---- BASELINE config 5's stand-in on one GPU: the `depth_sparse` observation (1024 band voxels x (x, y, z, tsdf)) into
        #      PointNet with C = 4, bf16 (the reference has no Sparse-UNet: SURVEY H4)
        try:
            env4 = FakeVecEnv(E, N_PTS * 4, ACT, dev, cloud=True, channels=4, pool=8, seed=99)
            net4 = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=N_PTS, precision="bf16")
            r4 = ppo(env4, ppo_cfg(E, dev, "bf16", net=net4), _Logger())
            ms4, _, _ = timed(r4, env4, 3, 3, False)
            configs["config5_standin_pointnet_c4"] = {
                "value": E * T_STEPS / (ms4 * 1e-3), "unit": "env*steps/s", "ms_per_step": ms4, "steps": 3, "warmup": 3, "dtype": "bf16",
                "workload": f"ppo, {E} envs x {N_PTS} sparse voxels x 4 ch (x, y, z, tsdf) into PointNet (C = 4), one GPU of config 5's 8"}
            r4.release_graph()
            del env4, r4
        except Exception as e:  # pragma: no cover
            configs["config5_standin_pointnet_c4"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
        # ---- the shipped state policy: ppo + MLP 53 -> 512^3 -> 10 at E = 2048 (cfg/algos/ppo.yaml)
        try:
            configs["state_mlp_2048"] = bench_state_mlp(dev, peak_tf, peak_src, kernel_ms)
        except Exception as e:  # pragma: no cover
            configs["state_mlp_2048"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- the reference's PyTorch-GPU path on this GPU (eager fp32, TF32 off): the ">= 10x" denominator
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        try:
            v, t, n = cpu_port_iteration(E, os.cpu_count() or 1, iters=args.gpu_baseline_iters, warm=3, device=dev,
                                         budget_s=args.gpu_baseline_budget_s)
            gpu_baseline = {"value": v, "unit": "env*steps/s", "ms_per_step": t * 1e3, "steps": n, "warmup": 3, "dtype": "f32",
                            "kind": "port", "what": "eager PyTorch restatement of the reference's encoder+PPO path (oracle port) on the "
                                                    "same GPU: fp32, TF32 off, default stream, same E / clouds / hyper-parameters",
                            "speedup_bf16_mode": value / v,
                            "speedup_fp32_mode": (modes["fp32"]["value"] / v) if modes else (value / v if precision != "bf16" else None)}
        except Exception as e:  # pragma: no cover
            gpu_baseline = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()

    # ---- CPU baseline (oracle port) on rank 0, N=1 only: one warm-up + timed iterations at E=64
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        Ec = args.cpu_baseline_envs
        v, t, n = cpu_port_iteration(Ec, threads, iters=2, warm=1, budget_s=5.0)
        cpu_baseline = {"value": v, "unit": "env*steps/s", "cores": threads, "kind": "port",
                        "sample": f"{n} full PPO iteration(s) at E={Ec} envs x {N_PTS} pts after 1 warm-up ({t:.1f} s each) with the CPU oracle port"}

    if rank == 0:
        B = min(2048, E * T_STEPS)
        line = {
            "metric": "env_steps_per_sec (encoder+PPO update)", "value": value, "unit": "env*steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"ppo + open_drawer shapes: {E} envs/GPU x {N_PTS}-pt cloud x {CH} ch, PointNet encoder, A={ACT}, "
                                   f"T={T_STEPS}, 5 epochs x {E * T_STEPS // B} minibatches of {B}; one step = one PPO iteration",
                       "envs_per_gpu": E, "precision": precision, "cache": "inputs 403 MB rollout buffer > 126 MB L2",
                       "parallelism": f"env-sharded x{world}" if world > 1 else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "modes": modes, "gpu_baseline": gpu_baseline, "configs": configs, "multi_gpu_parity": multi_gpu_parity,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        # leave without tearing NCCL down: captured graphs hold NCCL work objects and destroy_process_group() can block on them
        barrier()
        _OUT.flush()
        sys.stderr.flush()
        os._exit(0)


def bench_dagger(dev, precision):
    """BASELINE config 4 at full size on one GPU: E = 2048 envs x 2048-pt clouds (D = 6144), vision student (PointNet), frozen
    state teacher (MLP 53 -> 512^3 -> 10), ring buffer of 16 steps (32768 rows, 805 MB), random sampler, n_steps = 1,
    2 epochs x 16 minibatches of 2048 per iteration (cfg/algos/dagger_tsdf.yaml hyper-parameters; PointNet instead of the
    Conv3D student).  One step = rollout step (student acts, both observations stored) + update; env·steps/s = E / time."""
    import torch
    from oracle import ppo_oracle as O
    from partmanip_b200.algorithms import dagger
    from partmanip_b200.envs import FakeVecEnv
    E, NP, A, Dt = 2048, 2048, 10, 53
    D = NP * CH
    g = torch.Generator().manual_seed(21)
    tea_cfg = dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="MLP", hid_dim=[512, 512, 512], activation="tanh"))
    sd = {f"actor.{k}": v for k, v in O.mlp_init(Dt, A, [512, 512, 512], gen=g).items()}
    sd.update({f"critic.{k}": v for k, v in O.mlp_init(Dt, 1, [512, 512, 512], gen=g).items()})
    sd["log_std"] = torch.full((A,), -0.69)
    os.makedirs("/tmp/pm_b200_bench", exist_ok=True)
    path = "/tmp/pm_b200_bench/teacher.pth"
    torch.save(dict(obs_mode="state", model_cfg=tea_cfg, model_state_dict=sd, tricks=dict(use_state_norm=False)), path)
    net = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=NP, precision=precision)
    cfg = dict(num_envs=E, obs_mode="obs", max_iterations=10 ** 9, n_steps=1, n_updates=2, n_minibatches=16, device=dev, buf_size=16,
               reward_reset=False, add_proprio_obs=False, offline_data_pth=None, eval_round=1, eval_frequence=10 ** 9,
               save_frequence=10 ** 9, test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=5e-5,
               teacher=path, resume=None, pretrain=None, sampler="random",
               model=dict(action_std=0.1, action_activate="tanh", clipAction=1.0, network=net))
    env = FakeVecEnv(E, D, A, dev, cloud=True, channels=CH, pool=8, seed=77, extra_obs={"state": Dt})
    r = dagger(env, cfg, _Logger())
    stu, tea = r._reset_env()
    for _ in range(16):                                   # fill the ring
        stu, tea, _ = r._collect(stu, tea)
    for _ in range(2):
        stu, tea, _ = r._collect(stu, tea)
        r.update(1)
    torch.cuda.synchronize()
    iters = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        stu, tea, _ = r._collect(stu, tea)
        r.update(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out = {"value": E / (ms * 1e-3), "unit": "env*steps/s", "ms_per_step": ms, "steps": iters, "warmup": 2, "dtype": precision,
           "workload": f"dagger: {E} envs x {NP}-pt clouds, buffer 16 x {E} rows, random sampler, 2 epochs x 16 minibatches of 2048, "
                       "teacher MLP 53->512^3->10; one step = 1 env step for all envs + update",
           "dagger_loss": float(r.log_dict["Train/dagger_loss"])}
    del r, env, stu, tea
    torch.cuda.empty_cache()
    try:                                                  # the same step as eager PyTorch on this GPU (oracle port; one warm + one timed step)
        v, t, n = torch_dagger_step(dev, E=E, NP=NP, iters=1, warm=1)
        out["gpu_baseline"] = {"value": v, "unit": "env*steps/s", "ms_per_step": t * 1e3, "steps": n, "warmup": 1, "dtype": "f32", "kind": "port",
                               "speedup": out["value"] / v}
    except Exception as e:  # pragma: no cover
        out["gpu_baseline"] = {"error": repr(e)[:200]}
    torch.cuda.empty_cache()
    return out


def torch_dagger_step(device, E=2048, NP=2048, iters=1, warm=1, mb=2048, n_mb=16, epochs=2, buf_steps=16, hid=(512, 512, 512)):
    """BASELINE config 4 as eager PyTorch (the oracle port of dagger.py:205-337, fp32, TF32 off) on `device`: one step = the student's
    exploration forward on E clouds, the ring append, and n_updates x n_minibatches update steps, each with its own teacher forward
    (dagger.py:311) and a gathered minibatch.  Returns (env·steps/s, seconds per step, timed steps).  Only the DAgger config's
    gpu_baseline calls this."""
    import torch
    from oracle import ppo_oracle as O
    on_gpu = str(device).startswith("cuda")
    if on_gpu:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
    A, Dt, D = 10, 53, NP * CH
    g = torch.Generator().manual_seed(21)
    dv = lambda t: t.to(device)
    student = {k: dv(v) for k, v in O.pointnet_init(D, A, point_num=NP, gen=g).items()}
    teacher = {k: dv(v) for k, v in O.mlp_init(Dt, A, list(hid), gen=g).items()}
    log_std = dv(torch.full((A,), -2.3))
    opt = O.AdamState(student, 5e-5)
    net = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=NP)
    rows = buf_steps * E
    ring_s, ring_t = dv(torch.rand(rows, D, generator=g) * 2 - 1), dv(torch.randn(rows, Dt, generator=g))
    obs, tea_obs = dv(torch.rand(E, D, generator=g) * 2 - 1), dv(torch.randn(E, Dt, generator=g))
    times, slot = [], 0
    for it in range(warm + iters):
        if on_gpu:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            mu = O.pointnet_forward(student, obs.clone(), point_num=NP)
            O.policy_sample(mu, log_std, dv(torch.randn(E, A, generator=g)), 1.0)
            ring_s[slot:slot + E].copy_(obs)
            ring_t[slot:slot + E].copy_(tea_obs)
            slot = (slot + E) % rows
        for _ in range(epochs):
            perm = torch.randperm(rows, generator=g)
            for k in range(n_mb):
                idx = dv(perm[k * mb:(k + 1) * mb])
                with torch.no_grad():
                    tea_act = O.action_activation(O.mlp_forward(teacher, ring_t[idx], "tanh"), 1.0)
                O.dagger_update_step(opt.params, opt, ring_s[idx], tea_act, "PointNet", net, 1.0)
        if on_gpu:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = sum(times[warm:]) / max(len(times[warm:]), 1)
    return E / t, t, len(times[warm:])


def bench_state_mlp(dev, peak_tf, peak_src, kernel_ms):
    """The configuration the shipped `--algocfg ppo --taskcfg open_drawer` runs (cfg/algos/ppo.yaml: E = 2048, MLP 53 -> 512^3 ->
    10, tanh): full PPO iterations in the fp32 parity mode (three-term bf16 split on tcgen05) and in bf16, and the roofline of
    its dominant GEMM (2048 x 512 x 512 fc layer) against the measured bf16 tensor peak."""
    import torch
    from partmanip_b200 import ops
    from partmanip_b200.algorithms import ppo
    from partmanip_b200.envs import FakeVecEnv
    E, D, A = 2048, 53, 10
    out = {"workload": f"ppo + state obs: {E} envs, MLP {D}->512^3->{A} (cfg/algos/ppo.yaml), T={T_STEPS}, 5 epochs x 8 minibatches of 2048"}
    for prec in ("fp32", "bf16", "fp32_ffma"):
        net = dict(name="MLP", hid_dim=[512, 512, 512], activation="tanh", precision=prec)
        env = FakeVecEnv(E, D, A, dev, cloud=False, pool=16, seed=5)
        r = ppo(env, ppo_cfg(E, dev, prec, net=net), _Logger())
        curr = r._ingest(env.reset()["obs"], r.storage.obs_slot())

        def it(curr):
            last_obs, last_values = r.collect(curr, None)
            r.storage.compute_returns(last_values, r.gamma, r.lam)
            r.update(1)
            r.storage.clear()
            return ops.copy_rows(last_obs, r.storage.obs_slot())
        for _ in range(3):
            curr = it(curr)
        torch.cuda.synchronize()
        iters = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            curr = it(curr)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[prec] = {"value": E * T_STEPS / (ms * 1e-3), "unit": "env*steps/s", "ms_per_step": ms, "steps": iters, "warmup": 3}
        r.release_graph()
    # the same configuration as eager PyTorch on this GPU (the oracle port: fp32, TF32 off) — what the shipped command line runs today
    try:
        v, t, n = cpu_port_iteration(E, os.cpu_count() or 1, iters=10, warm=3, device=dev, budget_s=20, mlp=(D, (512, 512, 512)))
        out["gpu_baseline"] = {"value": v, "unit": "env*steps/s", "ms_per_step": t * 1e3, "steps": n, "warmup": 3, "dtype": "f32", "kind": "port",
                               "speedup_fp32_mode": out["fp32"]["value"] / v, "speedup_bf16_mode": out["bf16"]["value"] / v}
    except Exception as e:  # pragma: no cover
        out["gpu_baseline"] = {"error": repr(e)[:200]}
    # dominant GEMM: one 512 x 512 fc layer over a 2048-row minibatch (forward), rotating over 8 input buffers
    M, N, K = 2048, 512, 512
    xs = [torch.randn(M, K, device=dev) for _ in range(8)]
    W, b = torch.randn(N, K, device=dev) / K ** 0.5, torch.zeros(N, device=dev)
    y = torch.empty(M, N, device=dev)
    roof = {}
    for prec, mult in (("bf16", 1), ("fp32", 6)):
        ms = kernel_ms(lambda i: ops.linear_forward_tc(xs[i % 8], W, b, "tanh", prec, out=y), 50, 5)
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        roof[prec] = {"bound": "tensor", "kernel": f"fc {K}->{N} forward over {M} rows on tcgen05 ({prec})", "ms_per_launch": ms,
                      "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "peak_source": peak_src,
                      "executed_mma_tflops": tf * mult,
                      "note": "1 GFLOP per launch on 64 CTAs: latency-bound (fp32 operands are converted to bf16 images in the kernel)"}
    out["roofline"] = roof
    return out


if __name__ == "__main__":
    main()
