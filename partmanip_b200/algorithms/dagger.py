"""`dagger` — mirror of the reference DAgger runner (algorithms/dagger.py:14-337): a vision student acts, a frozen
state teacher labels, a ring buffer (storage.py:20-27, 84-91) feeds an MSE update of the student's actor.

Same constructor `(vec_env, cfg, logger)`, cfg keys, checkpoint keys and log keys.  The student forward/backward,
the teacher forward, the loss and the Adam step run in libpartmanip_b200.so; minibatches of the random sampler are
gathered with pm_gather_rows; `offline_data_pth` prefills the ring from recorded steps (storage.py:58-82).  Differences, on purpose: `teacher_reward.npy` is only required when
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
from .ppo import FlatAdam, _path2video

# attribute <- cfg key (dagger.py:18-51); the attribute names are the reference's, other code reads them
_CFG_ATTRS = (("num_envs", "num_envs"), ("stu_obs_mode", "obs_mode"), ("model_cfg", "model"), ("max_iter", "max_iterations"),
              ("n_steps", "n_steps"), ("n_updates", "n_updates"), ("num_mini_batches", "n_minibatches"), ("device", "device"),
              ("buf_size", "buf_size"), ("reward_reset", "reward_reset"), ("add_proprio_obs", "add_proprio_obs"),
              ("offline_data_pth", "offline_data_pth"), ("eval_round", "eval_round"), ("eval_freq", "eval_frequence"),
              ("save_freq", "save_frequence"), ("test_only", "test_only"), ("save_pose", "save_pose"), ("save_video", "save_video"),
              ("lr_schedule", "lr_schedule"), ("lr", "lr"), ("teacher_path", "teacher"))
_REWARD_RESET_LAG = 10                                              # dagger.py:229 `delta_step`


def _action_summaries(infos, actions):
    """The three action statistics both loops log (dagger.py:159-161, 222-224)."""
    infos['action_t'] = actions[:, :3].mean(dim=-1)
    infos['action_r'] = actions[:, 3:6].mean(dim=-1)
    infos['action_gripper'] = actions[:, -1]
    return infos


class dagger:
    def __init__(self, vec_env, cfg, logger):
        self.vec_env, self.logger = vec_env, logger
        for attr, key in _CFG_ATTRS:
            setattr(self, attr, cfg[key])
        if self.reward_reset:
            self.tea_rew = torch.tensor(np.load('teacher_reward.npy')).to(self.device)
        self.stu_num_obs = self.stu_input_obs = vec_env.num_obs[self.stu_obs_mode]
        self.num_actions = vec_env.num_actions
        self.max_episode_length = vec_env.max_episode_length
        self.save_ckpt_dir = logger.save_ckpt_dir
        self.total_envsteps = self.total_time = self.curr_iter = 0
        self._build_student(vec_env.num_obs['proprio_state'] * self.add_proprio_obs)
        self._load_teacher()
        self.resume(cfg['resume'])
        self.load_pretrain(cfg['pretrain'])
        self.storage = RolloutStorage(self.num_envs, self.buf_size, self.stu_num_obs, self.num_actions, self.device,
                                      sampler=cfg['sampler'], tea_obs_shape=self.tea_num_obs,
                                      max_length=self.max_episode_length)
        self._stats = torch.zeros(2, device=self.device)
        self._acc = torch.zeros(2, device=self.device)
        self._mb = {}
        # The teacher is frozen and `act` is deterministic row by row, so its label for a ring row never changes: compute it once
        # when the row enters the ring instead of once per minibatch visit (n_updates * n_minibatches teacher forwards per
        # iteration in dagger.py:311).  Bit-identical to recomputing (tests/test_gpu_dagger.py); cfg['cache_teacher_actions'] = False
        # restores the per-minibatch teacher forward.
        self.cache_teacher_actions = bool(cfg.get('cache_teacher_actions', True))
        self._tea_act = torch.zeros(self.buf_size * self.num_envs, self.num_actions, device=self.device) if self.cache_teacher_actions else None
        self._rows_labelled = 0

    def _build_student(self, proprio_dim):
        """dagger.py:53-56: one Adam over all student parameters, in `student.parameters()` order [log_std, actor.*, critic.*].
        Only the actor ever receives a gradient and torch's Adam skips grad-less tensors, so the flat optimiser steps exactly
        the actor block; its state_dict keeps the reference's indices (actor tensors at 1..n, one group over all tensors)."""
        self.student = ActorCritic(self.stu_input_obs, self.num_actions, self.model_cfg, proprio_dim).to(self.device)
        flat = self.student.flatten_()
        tensors = list(flat.actor.parameters())
        offsets = flat.actor_offs[:-1]
        self.optimizer = FlatAdam(flat.actor_flat[:flat.actor_n_clip], tensors, offsets, [len(tensors)], self.lr, 0, 0.0,
                                  index_offset=1, n_unstepped_tail=len(list(flat.critic.parameters())))
        self._grads = [self.optimizer.grad[o:o + t.numel()].view(t.shape) for t, o in zip(tensors, offsets)]

    def _load_teacher(self):
        """dagger.py:64-73: a frozen state-based ActorCritic trained without observation normalisation."""
        assert self.teacher_path is not None and os.path.exists(self.teacher_path)
        print(f'load teacher ckpt from {self.teacher_path}!')
        ckpt = torch.load(self.teacher_path, map_location=self.device, weights_only=False)
        assert ckpt['tricks']['use_state_norm'] == False  # noqa: E712  (dagger.py:73)
        self.tea_obs_mode = ckpt['obs_mode']
        self.tea_num_obs = self.vec_env.num_obs[self.tea_obs_mode]
        self.teacher = ActorCritic(self.tea_num_obs, self.num_actions, ckpt['model_cfg']).to(self.device)
        self.teacher.load_state_dict(ckpt["model_state_dict"])

    # ------------------------------------------------------------------ checkpoints (dagger.py:85-112)
    def save(self, it):
        os.makedirs(self.save_ckpt_dir, exist_ok=True)
        target = pjoin(self.save_ckpt_dir, f'model_{it}.pth')
        weights = {name: t.detach().clone() for name, t in self.student.state_dict().items()}
        torch.save(dict(iteration=it, model_state_dict=weights, optimizer_state_dict=self.optimizer.state_dict(),
                        total_steps=self.total_envsteps, obs_mode=self.stu_obs_mode, teacher=self.teacher_path,
                        b200_rng=self.student.rng_state()), target)
        print(f'save ckpt to {target}!')

    def _read_ckpt(self, ckpt_path):
        assert os.path.exists(ckpt_path)
        return torch.load(ckpt_path, map_location=self.device, weights_only=False)

    def load_pretrain(self, ckpt_path):
        if ckpt_path is None:
            return
        weights = self._read_ckpt(ckpt_path)['model_state_dict']
        weights.pop('log_std')
        self.student.load_state_dict(weights, strict=False)

    def resume(self, ckpt_path):
        if ckpt_path is None:
            return
        ckpt = self._read_ckpt(ckpt_path)
        self.student.load_state_dict(ckpt["model_state_dict"])
        self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        self.curr_iter, self.total_envsteps = ckpt["iteration"], ckpt["total_steps"]
        if 'b200_rng' in ckpt:
            self.student.set_rng_state(ckpt['b200_rng'])

    # ------------------------------------------------------------------ evaluation (dagger.py:114-178)
    def eval(self):
        self.student.eval()
        if self.test_only:
            self.log_dict = {}
        mode = 'Test' if self.test_only else 'Val'
        for r in range(self.eval_round):
            episode, poses = [], []
            obs = self.vec_env.reset()[self.stu_obs_mode]
            for t in range(self.max_episode_length):
                actions = self.student.act(obs)
                frame = pjoin(self.logger.save_video_dir, f"Iter{self.curr_iter}", f"{t}.png") if self.save_video else None
                nxt, rews, _, infos = self.vec_env.step(actions, save_image_path=frame)
                _action_summaries(infos, actions)['reward'] = rews
                episode.append(deepcopy(infos))
                if self.save_pose:                                  # dagger.py:155-159
                    rec = self.vec_env.save_scene_pose(pjoin(self.logger.save_pose_dir, f"Iter{self.curr_iter}", f"{t}.npy"))
                    rec['state'], rec['action'] = obs.cpu().numpy(), actions.cpu().numpy()
                    poses.append(deepcopy(rec))
                obs = nxt[self.stu_obs_mode]
            if self.save_pose:                                      # dagger.py:163-167
                for t, rec in enumerate(poses):
                    rec['success'] = episode[-1]['obj_up_flag'].cpu().numpy()
                    np.save(pjoin(self.logger.save_pose_dir, f"Iter{self.curr_iter}", f"{t}.npy"), rec)
            if self.save_video and r == self.eval_round - 1:        # dagger.py:169-171
                _path2video()(pjoin(self.logger.save_video_dir, f"Iter{self.curr_iter}"))
            self.use_info_update_logdict(episode, mode)

    # ------------------------------------------------------------------ training loop (dagger.py:180-278)
    def _reset_env(self):
        first = self.vec_env.reset()
        return first[self.stu_obs_mode], first[self.tea_obs_mode]

    def _collect(self, stu_obs, tea_obs):
        """n_steps of the student acting with exploration noise; every visited (student obs, teacher obs) pair enters the ring."""
        episode = []
        for _ in range(self.n_steps):
            actions = self.student.random_act(stu_obs)
            nxt, rews, _, infos = self.vec_env.step(actions)
            slot = self.storage.mix_buf_ind
            self.storage.add_transitions_dagger(stu_obs, tea_obs)
            if self.cache_teacher_actions:
                ops.copy_rows(self.teacher.act(tea_obs), self._tea_act[slot:slot + self.num_envs])
                self._rows_labelled += self.num_envs
            episode.append(deepcopy(_action_summaries(infos, actions)))
            stu_obs, tea_obs = nxt[self.stu_obs_mode], nxt[self.tea_obs_mode]
            if self.reward_reset:                                   # dagger.py:228-233: restart envs that fall behind the teacher
                prog = self.vec_env.progress_buf
                self.vec_env.dagger_reward_reset = (prog > _REWARD_RESET_LAG) & (rews < self.tea_rew[prog - _REWARD_RESET_LAG])
        return stu_obs, tea_obs, episode

    def run(self):
        if self.test_only:
            self.eval()
            self.logger.info(self.log_dict, self.curr_iter)
            return
        if self.offline_data_pth is not None:                        # dagger.py:186-187: recorded steps first
            self.storage.add_transitions_offline(self.offline_data_pth, self.device, self.add_proprio_obs)
        stu_obs, tea_obs = self._reset_env()
        while self.curr_iter < self.max_iter:
            self.curr_iter += 1
            self.student.train()
            self.teacher.eval()
            self.log_dict = {}
            t0 = time.time()
            stu_obs, tea_obs, episode = self._collect(stu_obs, tea_obs)
            torch.cuda.synchronize()
            t1 = time.time()
            self.update(self.curr_iter)
            torch.cuda.synchronize()
            collection_time, learn_time = t1 - t0, time.time() - t1
            steps = self.n_steps * self.vec_env.num_envs
            self.total_envsteps += steps
            self.total_time += collection_time + learn_time
            self.log_dict.update({
                'Progress/total_steps': self.curr_iter,
                'Progress/collection_time': collection_time,
                'Progress/learn_time': learn_time,
                'Progress/FPS': int(steps / (collection_time + learn_time)),
                'Train/mean_action_noise_std': self.student.log_std.detach().exp().mean().item(),
                'Train/cur_buf_size': self.storage.cur_buf_size,
                'Train/succ_buf_ind': self.storage.succ_buf_ind,
                'Train/mix_buf_ind': self.storage.mix_buf_ind,
            })
            self.use_info_update_logdict(episode, 'Train')
            if self.curr_iter % self.eval_freq == 0:
                self.eval()
                stu_obs, tea_obs = self._reset_env()
            if self.curr_iter % self.save_freq == 0:
                self.save(self.curr_iter)
            self.logger.info(self.log_dict, self.curr_iter)

    def use_info_update_logdict(self, info_lst, mode):
        """dagger.py:280-297: per-key mean and mean-of-per-env-max over the episode; evaluation rounds are averaged."""
        running = mode != 'Train'
        for key, first in info_lst[0].items():
            assert first.dim() == 1, f"{key}: {first.shape}"
            series = torch.stack([step[key].float() for step in info_lst], dim=-1)          # (num_envs, steps)
            for suffix, value in (('mean', series.mean()), ('max', series.max(dim=-1)[0].mean())):
                name = f'{mode}/{key}_{suffix}'
                self.log_dict[name] = self.log_dict.get(name, 0) + value / self.eval_round if running else value

    # ------------------------------------------------------------------ the update (dagger.py:299-337)
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
        if self.cache_teacher_actions and self._rows_labelled != st.rows_added:
            # rows entered the ring behind the cache's back (storage used directly): label the whole live buffer once
            n = st.cur_buf_size
            ops.copy_rows(self.teacher.act(st.tea_obs[:n]), self._tea_act[:n])
            self._rows_labelled = st.rows_added
        for epoch in range(self.n_updates):
            for indices in st.mini_batch_generator(self.num_mini_batches):
                stu_obs = self._gather(st.observations, indices, 's')
                B = stu_obs.shape[0]
                if dmu is None or dmu.shape[0] != B:
                    dmu = torch.empty(B, self.num_actions, device=self.device)
                if self.cache_teacher_actions:
                    tea_act = self._gather(self._tea_act, indices, 'a')
                else:
                    tea_act = self.teacher.act(self._gather(st.tea_obs, indices, 't'))
                mu = stu.actor.runner.forward(stu_obs)
                ops.dagger_loss(mu, tea_act, stu.max_action, squash, 1.0 / (B * self.num_actions), self._stats, dmu)
                ops.accumulate(self._stats, 1.0, self._acc, 0)
                stu.actor.runner.backward(stu_obs, dmu, self._grads)
                self.optimizer.step(None)
                count += 1
        mean_loss = self._acc[0].item() / max(count, 1)
        ops.check_tc_errors()
        if self.lr_schedule == 'linear_decay':
            self.optimizer.set_lr(self.lr * max(1 - it / self.max_iter * 1.8, 0.1))
        elif self.lr_schedule != 'fixed':
            raise NotImplementedError
        if not hasattr(self, 'log_dict'):
            self.log_dict = {}
        self.log_dict['Train/learning_rate'] = self.optimizer.param_groups[0]['lr']
        self.log_dict['Train/dagger_loss'] = mean_loss
