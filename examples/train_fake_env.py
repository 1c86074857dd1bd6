#!/usr/bin/env python
"""Minimal end-to-end use of the plugin surface: what the reference's train.py does after parsing its YAML
(train.py:68-72: env = eval(task)(...); runner = eval(algo)(env, cfg['algo'], logger); runner.run()), with the synthetic
zero-physics env standing in for the closed Isaac Gym stepper.

    python examples/train_fake_env.py [--envs 256] [--iters 5] [--precision bf16|fp32] [--net PointNet|MLP]
    torchrun --nproc-per-node 2 examples/train_fake_env.py        # env-sharded data parallel, one rank per GPU
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partmanip_b200.algorithms import ppo            # noqa: E402  (same name / ctor / run() as the reference's algorithms.ppo)
from partmanip_b200.envs import FakeVecEnv           # noqa: E402


class ScreenLogger:
    """The reference's Logger interface as the algo uses it (utils/logger.py:20-22,57-71)."""
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/partmanip_b200_example"

    def info(self, d, it):
        keys = ("Progress/FPS", "Train/surrogate_loss", "Train/value_function_loss", "Train/kl", "Train/kl_update_count")
        print(f"iter {it}: " + "  ".join(f"{k.split('/')[1]}={float(d[k]):.5g}" for k in keys if k in d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--net", default="PointNet", choices=["PointNet", "MLP"])
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    pointnet = args.net == "PointNet"
    D, A = (3072, 10) if pointnet else (53, 10)
    net = (dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=1024, precision=args.precision)
           if pointnet else dict(name="MLP", hid_dim=[512, 512, 512], activation="tanh"))
    cfg = dict(   # cfg/algos/ppo.yaml of the reference + the PointNet keys its YAML lacks
        num_envs=args.envs, obs_mode="obs", succ_value=None, max_iterations=args.iters, n_steps=8, n_updates=5, n_minibatches=8,
        device=dev, eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9, test_only=False, save_pose=False,
        save_video=False, lr_schedule="fixed", lr=5e-5, desired_kl=0.1, epsilon_clip=0.2, gamma=0.99, lam=0.95,
        sampler="sequential", resume=None,
        tricks=dict(mini_adv_norm=False, whole_adv_norm=False, use_state_norm=True, use_clipped_value_loss=False,
                    use_grad_clip=True, max_grad_norm=0.5),
        model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=net))
    env = FakeVecEnv(args.envs, D, A, dev, cloud=pointnet, seed=1 + int(os.environ.get("RANK", "0")))
    runner = ppo(env, cfg, ScreenLogger())
    runner.run()
    runner.release_graph()
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
