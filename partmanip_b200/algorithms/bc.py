"""`bc` — mirror of the reference behaviour-cloning runner (algorithms/bc.py:12-179): a student ActorCritic regresses recorded
expert actions from offline TSDF volumes (+ proprioception) with an MSE loss and one Adam over the student.

Same constructor `(vec_env, cfg, logger)`, cfg keys, dataset layout (`Tsdf_Dataset`), checkpoint keys and log keys.  The student
forward / backward (Conv3DNet on the conv + tcgen05 dense kernels by default, or any other plugin), the loss and the Adam step run in
libpartmanip_b200.so; torch's DataLoader stays what it is in the reference — file IO and shuffling on the host (the shuffle draws
from torch's global generator exactly like the reference's, so a seeded run visits the same minibatches).  `cfg['num_workers']`
(optional, default 10 = bc.py:117) sets the loader's worker count.
"""
from __future__ import annotations

import os
from os.path import join as pjoin

import numpy as np
import torch

from .. import ops
from .algo_utils import ActorCritic
from .ppo import FlatAdam


class Tsdf_Dataset(torch.utils.data.Dataset):
    """bc.py:12-31: <data_path>/<scene>/step_00000.npy ... each a pickled dict(tsdf, action, proprio_state)."""

    def __init__(self, data_path):
        super().__init__()
        self.data_path = data_path
        self.env_lst = os.listdir(data_path)
        self.env_num = len(self.env_lst)
        self.step_num = len(os.listdir(pjoin(data_path, 'scene_00000')))

    def __getitem__(self, index):
        env_ind, step_ind = self.env_lst[index // self.step_num], index % self.step_num
        data = np.load(pjoin(self.data_path, f'{env_ind}/step_{str(step_ind).zfill(5)}.npy'), allow_pickle=True).item()
        return data['tsdf'], data['action'], data['proprio_state']

    def __len__(self):
        return self.env_num * self.step_num


class bc:
    def __init__(self, vec_env, cfg, logger):
        self.vec_env, self.logger = vec_env, logger
        self.num_envs = cfg['num_envs']
        self.stu_obs_mode = cfg['obs_mode']
        self.stu_num_obs = vec_env.num_obs[self.stu_obs_mode]
        self.num_actions = vec_env.num_actions
        self.max_episode_length = vec_env.max_episode_length
        self.model_cfg, self.max_iter, self.device = cfg['model'], cfg['max_iterations'], cfg['device']
        self.data_path, self.n_minibatches, self.add_proprio_obs = cfg['data_path'], cfg['n_minibatches'], cfg['add_proprio_obs']
        self.eval_round, self.eval_freq, self.save_freq = cfg['eval_round'], cfg['eval_frequence'], cfg['save_frequence']
        self.test_only, self.save_pose, self.save_video = cfg['test_only'], cfg['save_pose'], cfg['save_video']
        self.save_ckpt_dir = logger.save_ckpt_dir
        self.lr_schedule, self.lr = cfg['lr_schedule'], cfg['lr']
        self.num_workers = cfg.get('num_workers', 10)
        # bc.py:65-68: one Adam over student.parameters() = [log_std, actor.*, critic.*]; only the actor ever receives a gradient and
        # torch's Adam skips grad-less tensors, so the flat optimiser steps the actor block and keeps the reference's state indices
        self.student = ActorCritic(self.stu_num_obs, self.num_actions, self.model_cfg,
                                   cfg['add_proprio_obs'] * vec_env.num_obs['proprio_state']).to(self.device)
        print(self.student)
        flat = self.student.flatten_()
        tensors, offsets = list(flat.actor.parameters()), flat.actor_offs[:-1]
        self.optimizer = FlatAdam(flat.actor_flat[:flat.actor_n_clip], tensors, offsets, [len(tensors)], self.lr, 0, 0.0, index_offset=1,
                                  n_unstepped_tail=len(list(flat.critic.parameters())))
        self._grads = [self.optimizer.grad[o:o + t.numel()].view(t.shape) for t, o in zip(tensors, offsets)]
        self._stats = torch.zeros(2, device=self.device)
        self._acc = torch.zeros(2, device=self.device)
        self._inputs = {}
        self.total_time = 0
        self.curr_iter = 0
        self.resume(cfg['resume'])

    # ------------------------------------------------------------------ checkpoints (bc.py:79-108)
    def save(self, it):
        os.makedirs(self.save_ckpt_dir, exist_ok=True)
        save_path = pjoin(self.save_ckpt_dir, f'model_{it}.pth')
        weights = {name: t.detach().clone() for name, t in self.student.state_dict().items()}
        torch.save({'iteration': it, 'model_state_dict': weights, 'optimizer_state_dict': self.optimizer.state_dict(),
                    'obs_mode': self.stu_obs_mode, 'total_steps': 0, 'tricks': {'use_state_norm': False}, 'teacher': 0}, save_path)
        print(f'save ckpt to {save_path}!')

    def resume(self, ckpt_path):
        if ckpt_path is None:
            return
        print(f'load student ckpt from {ckpt_path}!')
        assert os.path.exists(ckpt_path)
        ckpt_dict = torch.load(ckpt_path, map_location=self.device, weights_only=False)
        self.student.load_state_dict(ckpt_dict["model_state_dict"])
        self.optimizer.load_state_dict(ckpt_dict["optimizer_state_dict"])
        self.curr_iter = ckpt_dict["iteration"]
        assert ckpt_dict['obs_mode'] == self.stu_obs_mode

    # ------------------------------------------------------------------ training (bc.py:110-177)
    def _model_input(self, tsdfs, states):
        """bc.py:126-133: host batch -> device; with add_proprio_obs the flattened volume and the proprio columns share one row."""
        B = tsdfs.shape[0]
        D = tsdfs[0].numel()
        P = states.shape[-1] if self.add_proprio_obs else 0
        buf = self._inputs.get(B)
        if buf is None:
            buf = self._inputs[B] = torch.empty(B, D + P, device=self.device, dtype=torch.float32)
        buf[:, :D].copy_(tsdfs.reshape(B, D), non_blocking=True)
        if P:
            buf[:, D:].copy_(states.reshape(B, P), non_blocking=True)
        return buf

    def run(self):
        if self.test_only:
            raise NotImplementedError
        train_dataset = Tsdf_Dataset(self.data_path)
        batch_size = len(train_dataset) // self.n_minibatches
        train_loader = torch.utils.data.DataLoader(train_dataset, batch_size=batch_size, shuffle=True, num_workers=self.num_workers)
        stu = self.student
        squash = stu.action_activate == 'tanh'
        dmu = {}
        while self.curr_iter < self.max_iter:
            self.curr_iter += 1
            self.log_dict = {}
            self._acc.zero_()
            count = 0
            for tsdfs, actions, states in train_loader:
                B = actions.shape[0]
                x = self._model_input(tsdfs.float(), states.float())
                act = actions.to(self.device, torch.float32).contiguous()
                if B not in dmu:
                    dmu[B] = torch.empty(B, self.num_actions, device=self.device)
                mu = stu.actor.runner.forward(x)                                     # update_act (actor_critic.py:67-69)
                ops.dagger_loss(mu, act, stu.max_action, squash, 1.0 / (B * self.num_actions), self._stats, dmu[B])
                ops.accumulate(self._stats, 1.0, self._acc, 0)
                stu.actor.runner.backward(x, dmu[B], self._grads)
                self.optimizer.step(None)
                count += 1
            mean_loss = self._acc[0].item() / count
            ops.check_tc_errors()
            if self.lr_schedule == 'linear_decay':
                self.optimizer.set_lr(self.lr * (1 - self.curr_iter / self.max_iter))
            elif self.lr_schedule == 'step_decay':
                self.optimizer.set_lr(self.lr if self.curr_iter < self.max_iter / 2 else self.lr * 0.1)
            elif self.lr_schedule != 'fixed':
                raise NotImplementedError
            self.log_dict['Train/learning_rate'] = self.optimizer.param_groups[0]['lr']
            self.log_dict['Train/bc_loss'] = mean_loss
            self.log_dict['Progress/total_steps'] = self.curr_iter
            if self.curr_iter % self.save_freq == 0:
                self.save(self.curr_iter)
            self.logger.info(self.log_dict, self.curr_iter)
