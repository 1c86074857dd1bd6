"""Synthetic offline TSDF dataset in the layout algorithms/bc.py:12-31 reads: <root>/scene_xxxxx/step_yyyyy.npy, each a pickled
dict(tsdf (R, R, R) float32, action (A,), proprio_state (1, P), tea_obs (T,)).  Deterministic in `seed`; shared by the golden generator and the
GPU test (a 50^3 volume is 500 KB: the files are generated, not committed)."""
import os

import numpy as np


def write_dataset(root, seed=0, scenes=3, steps=4, R=50, A=10, P=31, T=53):
    rng = np.random.default_rng(seed)
    rng_tea = np.random.default_rng(seed + 1000003)        # separate stream: the volumes / actions / states above stay what the goldens saw
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, R, dtype=np.float32)] * 3, indexing="ij")
    for s in range(scenes):
        d = os.path.join(root, f"scene_{s:05d}")
        os.makedirs(d, exist_ok=True)
        for t in range(steps):
            c = rng.uniform(-0.4, 0.4, size=3).astype(np.float32)
            r = np.float32(rng.uniform(0.2, 0.5))
            tsdf = np.clip((np.sqrt((xx - c[0]) ** 2 + (yy - c[1]) ** 2 + (zz - c[2]) ** 2) - r) * 4, -1, 1).astype(np.float32)
            tsdf += rng.normal(0, 0.02, size=tsdf.shape).astype(np.float32)
            action = np.tanh(rng.normal(0, 1, size=A)).astype(np.float32)
            state = rng.normal(0, 1, size=(1, P)).astype(np.float32)
            tea = rng_tea.normal(0, 1, size=T).astype(np.float32)          # storage.py:75 `tea_obs` (DAgger's offline prefill; bc ignores it)
            np.save(os.path.join(d, f"step_{t:05d}.npy"), dict(tsdf=tsdf, action=action, proprio_state=state, tea_obs=tea), allow_pickle=True)
    return scenes * steps


def bc_cfg(data_path, device, **over):
    """cfg/algos/bc.yaml with a short schedule."""
    cfg = dict(num_envs=16, obs_mode="mesh_tsdf", add_proprio_obs=True, max_iterations=3, n_minibatches=3, data_path=data_path, device=device,
               eval_round=3, eval_frequence=200, save_frequence=10 ** 9, test_only=False, save_pose=False, save_video=False,
               lr_schedule="step_decay", lr=5e-4, resume=None,
               model=dict(action_std=0.0, action_activate="tanh", clipAction=1.0, network=dict(name="Conv3DNet", activation="tanh")))
    cfg.update(over)
    return cfg


class FakeBCEnv:
    """What algorithms/bc.py:34-41 reads from the vectorised env."""
    num_actions = 10
    max_episode_length = 200

    def __init__(self, R=50, P=31):
        self.num_obs = {"mesh_tsdf": R ** 3, "proprio_state": P}


class Logger:
    def __init__(self, d):
        self.save_ckpt_dir = self.save_video_dir = self.save_pose_dir = d
        self.rows = []

    def info(self, d, it):
        self.rows.append((it, {k: float(v) for k, v in d.items()}))
