"""`dagger` — mirror of the reference DAgger runner (algorithms/dagger.py:14-337): a vision student acts, a frozen
state teacher labels, a ring buffer (storage.py:20-27, 84-91) feeds an MSE update of the student's actor.

Same constructor `(vec_env, cfg, logger)`, cfg keys, checkpoint keys and log keys.  The student forward/backward,
the teacher forward, the loss and the Adam step run in libpartmanip_b200.so; minibatches of the random sampler are
gathered with pm_gather_rows.  Differences, on purpose: `teacher_reward.npy` is only required when
`reward_reset` is on (the reference loads it unconditionally, dagger.py:33); the per-step debug prints are dropped.
"""
from __future__ import annotations

import os
import time
from copy import deepcopy
from os.path import join as pjoin

import numpy as np
import torch

from .. import ops
from .algo_utils import ActorCritic, RolloutStorage
from .ppo import FlatAdam


class dagger:
    def __init__(self, vec_env, cfg, logger):
        self.vec_env = vec_env
        self.num_envs = cfg['num_envs']
        self.stu_obs_mode = cfg['obs_mode']
        self.stu_num_obs = vec_env.num_obs[self.stu_obs_mode]
        self.stu_input_obs = self.stu_num_obs
        self.num_actions = vec_env.num_actions
        self.max_episode_length = vec_env.max_episode_length
        self.model_cfg = cfg['model']
        self.max_iter = cfg['max_iterations']
        self.n_steps = cfg['n_steps']
        self.n_updates = cfg['n_updates']
        self.num_mini_batches = cfg['n_minibatches']
        self.device = cfg['device']
        self.buf_size = cfg['buf_size']
        self.reward_reset = cfg['reward_reset']
        if self.reward_reset:
            self.tea_rew = torch.tensor(np.load('teacher_reward.npy')).to(self.device)
        self.add_proprio_obs = cfg['add_proprio_obs']
        self.offline_data_pth = cfg['offline_data_pth']
        if self.offline_data_pth is not None:
            raise NotImplementedError("offline TSDF replay (storage.add_transitions_offline) is outside the hot path")
        self.eval_round = cfg['eval_round']
        self.eval_freq = cfg['eval_frequence']
        self.save_freq = cfg['save_frequence']
        self.test_only = cfg['test_only']
        self.save_pose = cfg['save_pose']
        self.save_video = cfg['save_video']
        self.save_ckpt_dir = logger.save_ckpt_dir
        self.lr_schedule = cfg['lr_schedule']
        self.lr = cfg['lr']
        # student (dagger.py:53-56): one Adam over all student parameters; only the actor ever receives a gradient,
        # and torch's Adam skips grad-less tensors, so the flat optimiser covers exactly the actor block
        self.student = ActorCritic(self.stu_input_obs, self.num_actions, self.model_cfg,
                                   cfg['add_proprio_obs'] * vec_env.num_obs['proprio_state']).to(self.device)
        st = self.student.flatten_()
        a_params = list(st.actor.parameters())
        n_actor = st.actor_n_clip
        self.optimizer = FlatAdam(st.actor_flat[:n_actor], a_params, st.actor_offs[:-1], [len(a_params)], self.lr, 0, 0.0)
        self._grads = [self.optimizer.grad[o:o + p.numel()].view(p.shape) for p, o in zip(a_params, st.actor_offs[:-1])]
        self.logger = logger
        self.total_envsteps = 0
        self.total_time = 0
        self.curr_iter = 0
        # teacher (dagger.py:64-73)
        self.teacher_path = cfg['teacher']
        assert self.teacher_path is not None and os.path.exists(self.teacher_path)
        print(f'load teacher ckpt from {self.teacher_path}!')
        tea_dict = torch.load(self.teacher_path, map_location=self.device, weights_only=False)
        self.tea_obs_mode = tea_dict['obs_mode']
        self.tea_num_obs = vec_env.num_obs[self.tea_obs_mode]
        self.teacher = ActorCritic(self.tea_num_obs, self.num_actions, tea_dict['model_cfg']).to(self.device)
        self.teacher.load_state_dict(tea_dict["model_state_dict"])
        assert tea_dict['tricks']['use_state_norm'] == False  # noqa: E712  (dagger.py:73)
        self.resume(cfg['resume'])
        self.load_pretrain(cfg['pretrain'])
        self.storage = RolloutStorage(self.num_envs, self.buf_size, self.stu_num_obs, self.num_actions, self.device,
                                      sampler=cfg['sampler'], tea_obs_shape=self.tea_num_obs,
                                      max_length=self.max_episode_length)
        self._stats = torch.zeros(2, device=self.device)
        self._acc = torch.zeros(2, device=self.device)
        self._mb = {}

    def save(self, it):
        os.makedirs(self.save_ckpt_dir, exist_ok=True)
        save_path = pjoin(self.save_ckpt_dir, f'model_{it}.pth')
        torch.save({'iteration': it,
                    'model_state_dict': {k: v.detach().clone() for k, v in self.student.state_dict().items()},
                    'optimizer_state_dict': self.optimizer.state_dict(), 'total_steps': self.total_envsteps,
                    'obs_mode': self.stu_obs_mode, 'teacher': self.teacher_path}, save_path)
        print(f'save ckpt to {save_path}!')

    def load_pretrain(self, ckpt_path):
        if ckpt_path is not None:
            assert os.path.exists(ckpt_path)
            ckpt_dict = torch.load(ckpt_path, map_location=self.device, weights_only=False)
            ckpt_dict['model_state_dict'].pop('log_std')
            self.student.load_state_dict(ckpt_dict["model_state_dict"], strict=False)

    def resume(self, ckpt_path):
        if ckpt_path is not None:
            assert os.path.exists(ckpt_path)
            ckpt_dict = torch.load(ckpt_path, map_location=self.device, weights_only=False)
            self.student.load_state_dict(ckpt_dict["model_state_dict"])
            self.optimizer.load_state_dict(ckpt_dict["optimizer_state_dict"])
            self.curr_iter = ckpt_dict["iteration"]
            self.total_envsteps = ckpt_dict["total_steps"]

    def eval(self):
        self.student.eval()
        if self.test_only:
            self.log_dict = {}
        for r in range(self.eval_round):
            ep_infos = []
            all_curr_obs = self.vec_env.reset()
            stu_curr_obs = all_curr_obs[self.stu_obs_mode]
            for i in range(self.max_episode_length):
                actions = self.student.act(stu_curr_obs)
                save_image_path = (pjoin(self.logger.save_video_dir, f"Iter{self.curr_iter}", f"{i}.png")
                                   if self.save_video else None)
                next_obs, rews, dones, infos = self.vec_env.step(actions, save_image_path=save_image_path)
                infos['action_t'] = actions[:, :3].mean(dim=-1)
                infos['action_r'] = actions[:, 3:6].mean(dim=-1)
                infos['action_gripper'] = actions[:, -1]
                infos['reward'] = rews
                ep_infos.append(deepcopy(infos))
                stu_curr_obs = next_obs[self.stu_obs_mode]
            self.use_info_update_logdict(ep_infos, 'Test' if self.test_only else 'Val')

    def run(self):
        if self.test_only:
            self.eval()
            self.logger.info(self.log_dict, self.curr_iter)
            return
        all_curr_obs = self.vec_env.reset()
        tea_curr_obs = all_curr_obs[self.tea_obs_mode]
        stu_curr_obs = all_curr_obs[self.stu_obs_mode]
        while self.curr_iter < self.max_iter:
            self.curr_iter += 1
            self.student.train()
            self.teacher.eval()
            self.log_dict = {}
            ep_infos = []
            start = time.time()
            for i in range(self.n_steps):
                actions = self.student.random_act(stu_curr_obs)
                next_obs, rews, dones, infos = self.vec_env.step(actions)
                self.storage.add_transitions_dagger(stu_curr_obs, tea_curr_obs)
                infos['action_t'] = actions[:, :3].mean(dim=-1)
                infos['action_r'] = actions[:, 3:6].mean(dim=-1)
                infos['action_gripper'] = actions[:, -1]
                tea_curr_obs = next_obs[self.tea_obs_mode]
                stu_curr_obs = next_obs[self.stu_obs_mode]
                ep_infos.append(deepcopy(infos))
                if self.reward_reset:   # dagger.py:228-233
                    delta_step = 10
                    self.vec_env.dagger_reward_reset = (self.vec_env.progress_buf > delta_step) & (
                        rews < self.tea_rew[self.vec_env.progress_buf - delta_step])
            torch.cuda.synchronize()
            collection_time = time.time() - start
            start = time.time()
            self.update(self.curr_iter)
            torch.cuda.synchronize()
            learn_time = time.time() - start
            self.total_envsteps += self.n_steps * self.vec_env.num_envs
            self.total_time += collection_time + learn_time
            self.log_dict['Progress/total_steps'] = self.curr_iter
            self.log_dict['Progress/collection_time'] = collection_time
            self.log_dict['Progress/learn_time'] = learn_time
            self.log_dict['Progress/FPS'] = int(self.n_steps * self.vec_env.num_envs / (collection_time + learn_time))
            self.log_dict['Train/mean_action_noise_std'] = self.student.log_std.detach().exp().mean().item()
            self.log_dict['Train/cur_buf_size'] = self.storage.cur_buf_size
            self.log_dict['Train/succ_buf_ind'] = self.storage.succ_buf_ind
            self.log_dict['Train/mix_buf_ind'] = self.storage.mix_buf_ind
            self.use_info_update_logdict(ep_infos, 'Train')
            if self.curr_iter % self.eval_freq == 0:
                self.eval()
                all_curr_obs = self.vec_env.reset()
                tea_curr_obs = all_curr_obs[self.tea_obs_mode]
                stu_curr_obs = all_curr_obs[self.stu_obs_mode]
            if self.curr_iter % self.save_freq == 0:
                self.save(self.curr_iter)
            self.logger.info(self.log_dict, self.curr_iter)

    def use_info_update_logdict(self, info_lst, mode):
        """dagger.py:280-297."""
        for key in info_lst[0].keys():
            assert len(info_lst[0][key].shape) == 1, f"{key}: {info_lst[0][key].shape}"
            all_info = torch.stack([info[key].float() for info in info_lst], dim=-1)
            if mode != 'Train':
                self.log_dict.setdefault(f'{mode}/{key}_mean', 0)
                self.log_dict.setdefault(f'{mode}/{key}_max', 0)
                self.log_dict[f'{mode}/{key}_mean'] += torch.mean(all_info) / self.eval_round
                self.log_dict[f'{mode}/{key}_max'] += torch.mean(all_info.max(dim=-1)[0]) / self.eval_round
            else:
                self.log_dict[f'{mode}/{key}_mean'] = torch.mean(all_info)
                self.log_dict[f'{mode}/{key}_max'] = torch.mean(all_info.max(dim=-1)[0])

    def _gather(self, src, indices, tag):
        if hasattr(indices, 'start'):
            return src[indices.start:indices.stop]
        key = (tag, indices.numel())
        out = self._mb.get(key)
        if out is None:
            out = torch.empty(indices.numel(), src.shape[1], device=src.device)
            self._mb[key] = out
        return ops.gather_rows(src, indices, out)

    def update(self, it):
        """dagger.py:299-337."""
        if self.storage.cur_buf_size < 16:
            return
        st = self.storage
        stu = self.student
        squash = stu.action_activate == 'tanh'
        self._acc.zero_()
        count = 0
        dmu = None
        for epoch in range(self.n_updates):
            for indices in st.mini_batch_generator(self.num_mini_batches):
                stu_obs = self._gather(st.observations, indices, 's')
                tea_obs = self._gather(st.tea_obs, indices, 't')
                B = stu_obs.shape[0]
                if dmu is None or dmu.shape[0] != B:
                    dmu = torch.empty(B, self.num_actions, device=self.device)
                tea_act = self.teacher.act(tea_obs)
                mu = stu.actor.runner.forward(stu_obs)
                ops.dagger_loss(mu, tea_act, stu.max_action, squash, 1.0 / (B * self.num_actions), self._stats, dmu)
                ops.accumulate(self._stats, 1.0, self._acc, 0)
                stu.actor.runner.backward(stu_obs, dmu, self._grads)
                self.optimizer.step(None)
                count += 1
        mean_loss = self._acc[0].item() / max(count, 1)
        if self.lr_schedule == 'linear_decay':
            self.optimizer.set_lr(self.lr * max(1 - it / self.max_iter * 1.8, 0.1))
        elif self.lr_schedule != 'fixed':
            raise NotImplementedError
        if not hasattr(self, 'log_dict'):
            self.log_dict = {}
        self.log_dict['Train/learning_rate'] = self.optimizer.param_groups[0]['lr']
        self.log_dict['Train/dagger_loss'] = mean_loss
