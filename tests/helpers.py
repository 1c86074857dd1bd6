"""Shared test helpers: golden fixture loading and the cfg the goldens were generated with."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(np.asarray(d[k])) for k in d.files}


def sub(d, prefix):
    """{'prefix.a.b': t} -> {'a.b': t}"""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in d.items() if k.startswith(prefix + ".")}


def ppo_cfg(E, net, **over):
    """Mirror of tests/golden/make_golden.py:ppo_cfg (shipped cfg/algos/ppo.yaml hyper-parameters)."""
    cfg = dict(
        num_envs=E, obs_mode="obs", succ_value=None, max_iterations=1, n_steps=8, n_updates=5,
        n_minibatches=8, device="cpu", eval_round=1, eval_frequence=10 ** 9, save_frequence=10 ** 9,
        test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=5e-5, desired_kl=0.1,
        epsilon_clip=0.2, gamma=0.99, lam=0.95, sampler="sequential", resume=None,
        tricks=dict(mini_adv_norm=False, whole_adv_norm=False, use_state_norm=True,
                    use_clipped_value_loss=False, use_grad_clip=True, max_grad_norm=0.5),
        model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=net),
    )
    for k, v in over.items():
        if isinstance(v, dict) and k in cfg:
            cfg[k] = {**cfg[k], **v}
        else:
            cfg[k] = v
    return cfg


PN = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False)
MLP128 = dict(name="MLP", hid_dim=[128, 128, 128], activation="tanh")

# name -> (E, D, A, net_cfg, cfg overrides) exactly as generated
ITER_CASES = {
    "ppo_iter_pointnet_e16.npz": (16, 3072, 10, PN, {}),
    "ppo_iter_pointnet_e8_nonorm.npz": (8, 3072, 10, PN, dict(tricks=dict(use_state_norm=False, mini_adv_norm=True,
                                                                          use_clipped_value_loss=True))),
    "ppo_iter_mlp_e64.npz": (64, 37, 7, MLP128, dict(succ_value=500, tricks=dict(whole_adv_norm=True))),
    "ppo_iter_mlp_e64_klskip.npz": (64, 53, 10, MLP128, dict(lr=3e-3, desired_kl=0.02)),
}


def close(a, b, rtol=1e-4, atol=1e-4):
    """north_star fp32 gate: |a-b| <= atol + rtol*|b|."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return bool(((a - b).abs() <= atol + rtol * b.abs()).all())


def max_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float(((a - b).abs() / (1.0 + b.abs())).max())
