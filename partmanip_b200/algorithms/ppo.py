"""`ppo` — mirror of the reference PPO runner (algorithms/ppo.py:17-411): same constructor `(vec_env, cfg, logger)`,
same `run / update / eval / save / resume` methods, same cfg keys, log keys and checkpoint format — with the
whole learner (rollout-side network forwards, sampling, GAE, both update phases, grad-clip + Adam) executed by
the CUDA kernels of libpartmanip_b200.so.

What differs from the reference, on purpose (results unchanged):
  * no host syncs inside the update: the KL-skip (`continue`, ppo.py:337-338) is a device predicate consumed by
    the Adam kernel, loss / KL sums accumulate on the device and are read back once per iteration;
  * each phase evaluates only the network it needs (the reference's update_act_cri runs both every time);
  * sequential-sampler minibatches are slices of the rollout buffer (no gather);
  * the per-env-step `print` with its four `.item()` syncs (ppo.py:229) is dropped.
Multi-GPU: when torch.distributed is initialised every rank owns `num_envs` envs; gradients, the
KL/loss sums and the observation statistics are all-reduced so all ranks take identical optimiser steps.
"""
from __future__ import annotations

import os
import time
from copy import deepcopy
from os.path import join as pjoin

import numpy as np
import torch

from .. import ops, parallel
from .algo_utils import ActorCritic, Normalization, RolloutStorage


def _path2video():
    """The reference's frame-folder -> video helper (`from utils import path2video`, ppo.py:1; cv2 + ffmpeg, outside the hot
    path).  Resolved lazily from the host project's `utils` package so that `save_video: True` keeps working when this
    class is dropped into the reference tree, and fails loudly (instead of silently skipping the step) elsewhere."""
    try:
        from utils import path2video
    except Exception as e:  # pragma: no cover - depends on the host project
        raise NotImplementedError("save_video needs the host project's utils.path2video (cv2/ffmpeg): %s" % (e,))
    return path2video


class FlatAdam:
    """torch.optim.Adam (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay) over one flat buffer, stepped by
    pm_adam_step.  state_dict()/load_state_dict() speak torch.optim's format so reference checkpoints round-trip
    (ppo.py:89-90, 121-122)."""

    def __init__(self, flat, params, offsets, group_sizes, lr, n_clip, max_norm, index_offset=0, n_unstepped_tail=0, fused=False):
        """`index_offset` / `n_unstepped_tail`: parameters of the reference optimiser that sit before / after this buffer's
        tensors in `module.parameters()` order and never receive a gradient (DAgger: log_std before, the critic after —
        dagger.py:56); they shift the state indices and widen param_groups[0]['params'] so the dict is torch.optim.Adam's."""
        self.flat, self.params, self.offsets, self.group_sizes = flat, params, offsets, group_sizes
        self.index_offset, self.n_unstepped_tail = int(index_offset), int(n_unstepped_tail)
        # gradient buffer with a 4-float tail: per-step scalars that must be summed over ranks (sum surrogate, sum KL) ride
        # in the SAME reduction as the gradients.  Multi-rank on NCCL: the buffer lives in symmetric memory so that the fused
        # step kernel (csrc/fused_step.cu) can pull the peers' gradients over NVLink itself — one launch per optimiser step
        # instead of all-reduce + finalize + 3 Adam launches; without symmetric memory the NCCL all-reduce path remains.
        self.sym = parallel.SymmetricGrad.create(flat.numel() + 4, flat.device) if (fused and flat.is_cuda) else None
        self.fused = bool(fused and flat.is_cuda and (parallel.world() == 1 or self.sym is not None))
        if self.sym is not None:
            self.grad_ext = self.sym.buf
        else:
            self.grad_ext = torch.zeros(flat.numel() + 4, device=flat.device, dtype=torch.float32)
        self._fused_ws = ops.fused_step_workspace(flat.numel(), 4, flat.device) if self.fused else None
        self.grad = self.grad_ext[:flat.numel()]
        self.tail = self.grad_ext[flat.numel():]
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.opt_state = torch.zeros(8, device=flat.device, dtype=torch.float32)   # [0]=step [1]=lr (device scalars)
        self.n_clip, self.max_norm = n_clip, max_norm
        self.param_groups = [{'lr': lr} for _ in group_sizes]
        self._lr_on_device = None
        self.set_lr(lr)

    def set_lr(self, lr):
        for g in self.param_groups:
            g['lr'] = lr
        if self._lr_on_device != lr:
            self.opt_state[1:2].fill_(float(lr))
            self._lr_on_device = lr

    def step(self, skip_flag=None):
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.n_clip, self.max_norm, self.opt_state,
                      skip_flag)

    def step_fused(self, finalize=None):
        """[all-reduce over ranks] + [KL-skip decision: finalize = (inv_batch, desired_kl, acc, skip_flag)] + clip + Adam, one launch."""
        ops.fused_step(self.flat, self.grad_ext, self.exp_avg, self.exp_avg_sq, self.n_clip, 4, self.max_norm, self.opt_state,
                       self._fused_ws, peers=self.sym.peers() if self.sym is not None else None, finalize=finalize)

    @property
    def step_count(self) -> int:
        return int(self.opt_state[0].item())

    def _views(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def state_dict(self):
        step = self.opt_state[0].detach().cpu().clone()
        state = {}
        if float(step) > 0:
            for i, (m, v) in enumerate(zip(self._views(self.exp_avg), self._views(self.exp_avg_sq))):
                state[i + self.index_offset] = {'step': step.clone(), 'exp_avg': m.clone(), 'exp_avg_sq': v.clone()}
        groups, k = [], 0
        sizes = list(self.group_sizes)
        sizes[0] += self.index_offset
        sizes[-1] += self.n_unstepped_tail
        for g, n in zip(self.param_groups, sizes):
            groups.append({'lr': g['lr'], 'betas': (0.9, 0.999), 'eps': 1e-08, 'weight_decay': 0, 'amsgrad': False,
                           'maximize': False, 'foreach': None, 'capturable': False, 'differentiable': False,
                           'fused': None, 'params': list(range(k, k + n))})
            k += n
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        st = sd['state']
        steps = []
        n_ref = sum(len(g['params']) for g in sd['param_groups'])
        n_here = len(self.params) + self.index_offset + self.n_unstepped_tail
        if n_ref != n_here:
            raise ValueError(f"optimizer state_dict covers {n_ref} parameters, this optimiser {n_here}")
        for i, (m, v) in enumerate(zip(self._views(self.exp_avg), self._views(self.exp_avg_sq))):
            j = i + self.index_offset
            if j in st:
                if tuple(st[j]['exp_avg'].shape) != tuple(m.shape):
                    raise ValueError(f"optimizer state {j}: shape {tuple(st[j]['exp_avg'].shape)} != {tuple(m.shape)}")
                m.copy_(st[j]['exp_avg'])
                v.copy_(st[j]['exp_avg_sq'])
                steps.append(float(st[j]['step']))
        # one step counter per flat buffer: the reference's two param groups always step together
        self.opt_state[0:1].fill_(max(steps) if steps else 0.0)
        self._lr_on_device = None
        self.set_lr(sd['param_groups'][0]['lr'])


class ppo:
    def __init__(self, vec_env, cfg, logger):
        # env infos (ppo.py:19-25)
        self.vec_env = vec_env
        self.num_envs = cfg['num_envs']
        self.obs_mode = cfg['obs_mode']
        self.num_obs = vec_env.num_obs[self.obs_mode]
        self.num_actions = vec_env.num_actions
        self.max_episode_length = vec_env.max_episode_length
        self.default_succ_value = cfg['succ_value']
        # training (ppo.py:27-33)
        self.model_cfg = cfg['model']
        self.max_iter = cfg['max_iterations']
        self.n_steps = cfg['n_steps']
        self.n_updates = cfg['n_updates']
        self.num_mini_batches = cfg['n_minibatches']
        self.device = cfg['device']
        if not str(self.device).startswith('cuda'):
            raise RuntimeError("partmanip_b200 runs on CUDA devices only (no CPU fallback); got device=%r" % (self.device,))
        # eval and save (ppo.py:35-42)
        self.eval_round = cfg['eval_round']
        self.eval_freq = cfg['eval_frequence']
        self.save_freq = cfg['save_frequence']
        self.test_only = cfg['test_only']
        self.save_pose = cfg['save_pose']
        self.save_video = cfg['save_video']
        self.save_ckpt_dir = logger.save_ckpt_dir
        # learning rate (ppo.py:44-52)
        self.lr_schedule = cfg['lr_schedule']
        self.lr = cfg['lr']
        self.desired_kl = cfg['desired_kl']
        assert self.desired_kl > 0
        if self.lr_schedule not in ('fixed', 'linear_decay', 'step_decay'):
            raise NotImplementedError
        # parameters (ppo.py:54-57)
        self.epsilon_clip = cfg['epsilon_clip']
        self.gamma = cfg['gamma']
        self.lam = cfg['lam']
        # tricks (ppo.py:59-68)
        self.tricks = {}
        self.tricks_keys = ['mini_adv_norm', 'whole_adv_norm', 'use_state_norm', 'use_clipped_value_loss', 'use_grad_clip']
        for k in self.tricks_keys:
            self.tricks[k] = cfg['tricks'][k]
        self.max_grad_norm = cfg['tricks']['max_grad_norm'] if self.tricks['use_grad_clip'] else 0.0
        if self.tricks['use_state_norm']:
            self.state_norm = Normalization(shape=self.num_obs, device=self.device)
            self.update_RMS = True
        # network, buffer, optimisers (ppo.py:70-74)
        self.world = parallel.world()
        self.actor_critic = ActorCritic(self.num_obs, self.num_actions, self.model_cfg).to(self.device)
        ac = self.actor_critic.flatten_()
        parallel.broadcast_(ac.actor_flat)      # identical replicas
        parallel.broadcast_(ac.critic_flat)
        self.storage = RolloutStorage(self.num_envs, self.n_steps, self.num_obs, self.num_actions, self.device,
                                      self.default_succ_value, self.tricks['whole_adv_norm'], cfg['sampler'])
        a_params = list(ac.actor.parameters())
        fused = bool(cfg.get('fused_step', True))        # build extension: one kernel per optimiser step (A/B switch)
        self.optimizer_actor = FlatAdam(ac.actor_flat, a_params + [ac.log_std], ac.actor_offs, [len(a_params), 1], self.lr,
                                        ac.actor_n_clip, self.max_grad_norm, fused=fused)
        c_params = list(ac.critic.parameters())
        self.optimizer_critic = FlatAdam(ac.critic_flat, c_params, ac.critic_offs, [len(c_params)], self.lr,
                                         ac.critic_flat.numel(), self.max_grad_norm, fused=fused)
        self._actor_grads = ac.grad_views(self.optimizer_actor.grad, "actor")
        self._critic_grads = ac.grad_views(self.optimizer_critic.grad, "critic")
        # device-side bookkeeping for one update(): acc = [sum surrogate, sum kl, count, kl_max, sum value loss]
        dev = self.device
        self._acc = torch.zeros(8, device=dev)
        self._stats_a = self.optimizer_actor.tail[:2]     # [sum surrogate, sum KL] — reduced together with the actor gradient
        self._stats_v = torch.zeros(2, device=dev)
        self._skip = torch.zeros(1, device=dev, dtype=torch.int32)
        self._adv_stats = torch.zeros(2, device=dev)
        self._clip_delta = torch.zeros(1, device=dev)
        self._next_obs = torch.empty(self.num_envs, self.num_obs, device=dev)
        self._mb = {}
        # log
        self.logger = logger
        self.total_envsteps = 0
        self.total_time = 0
        self.curr_iter = 0
        self.verbose = bool(cfg.get('verbose', False))
        # build extension: replay the whole device-side update (160 minibatch steps at E=4096) as ONE CUDA graph.  Legal
        # because the update has no host round trip (device-side KL-skip / step counters / loss sums) and sequential
        # minibatches are fixed slices of persistent buffers.  First call runs eagerly (allocates workspaces), the second
        # captures, later ones replay.  Off for the random sampler (host-side permutation).  Multi-rank: the NCCL all-reduces
        # are captured too (thread_local capture mode); call release_graph() before destroying the process group.
        self.cuda_graph = bool(cfg.get('cuda_graph', True)) and cfg['sampler'] == 'sequential' and \
            (self.world == 1 or bool(cfg.get('cuda_graph_multi_rank', True)))
        self._graph, self._graph_calls, self._n_critic = None, 0, 0
        # build extension: issue the (independent) actor and critic phases of the update on two streams — see _update_body
        self.overlap_phases = bool(cfg.get('overlap_phases', True)) and cfg['sampler'] == 'sequential'
        self._phase_streams = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)) if self.overlap_phases else None
        self._dbuf = {}
        self.resume(cfg['resume'])

    # ------------------------------------------------------------------ checkpoints (ppo.py:83-137)
    def save(self, it):
        os.makedirs(self.save_ckpt_dir, exist_ok=True)
        save_path = pjoin(self.save_ckpt_dir, f'model_{it}.pth')
        save_dict = {
            'iteration': it,
            'model_state_dict': {k: v.detach().clone() for k, v in self.actor_critic.state_dict().items()},
            'optimizer_actor': self.optimizer_actor.state_dict(),
            'optimizer_critic': self.optimizer_critic.state_dict(),
            'total_steps': self.total_envsteps,
            'tricks': self.tricks,
            'obs_mode': self.obs_mode,
            'model_cfg': self.model_cfg,
            'b200_rng': self.actor_critic.rng_state(),      # extra key (the reference ignores unknown keys)
        }
        if self.tricks['use_state_norm']:
            save_dict['state_running_ms'] = self.state_norm.running_ms.save()
        torch.save(save_dict, save_path)
        print(f'save ckpt to {save_path}!')

    def resume(self, ckpt_path):
        self.ckpt_path = ckpt_path
        if ckpt_path is None:
            return
        print(f'load ckpt from {ckpt_path}!')
        assert os.path.exists(ckpt_path)
        ckpt_dict = torch.load(ckpt_path, map_location=self.device, weights_only=False)
        self.actor_critic.load_state_dict(ckpt_dict["model_state_dict"])
        self.optimizer_actor.load_state_dict(ckpt_dict["optimizer_actor"])
        self.optimizer_critic.load_state_dict(ckpt_dict["optimizer_critic"])
        self.curr_iter = ckpt_dict["iteration"]
        self.total_envsteps = ckpt_dict["total_steps"]
        if 'b200_rng' in ckpt_dict:                         # absent in reference-written checkpoints
            self.actor_critic.set_rng_state(ckpt_dict['b200_rng'])
        for k in self.tricks_keys:
            if self.tricks[k] != ckpt_dict['tricks'][k]:
                print(f"WARNING: trick {k} is not consistent with ckpt! saved: {ckpt_dict['tricks'][k]}, now: {self.tricks[k]}")
                if k == 'use_state_norm':
                    print('this is not allowed')
                    exit(1)
        if self.tricks['use_state_norm']:
            self.state_norm.running_ms.load(ckpt_dict['state_running_ms'])
        assert self.obs_mode == ckpt_dict['obs_mode']

    # ------------------------------------------------------------------ evaluation (ppo.py:139-203)
    def eval(self):
        self.actor_critic.eval()
        self.vec_env.train_test_flag = 'test'
        if self.test_only:
            self.log_dict = {}
        ep_infos = []
        for r in range(self.eval_round):
            save_dict_lst = []
            curr_obs = self.vec_env.reset()[self.obs_mode]
            for i in range(self.max_episode_length):
                if self.tricks['use_state_norm']:
                    curr_obs = self.state_norm(curr_obs, update=False)
                actions, value = self.actor_critic.act_cri(curr_obs)
                save_image_path = (pjoin(self.logger.save_video_dir, f"Iter{self.curr_iter}", f"{i}.png")
                                   if self.save_video else None)
                next_obs, rews, _, infos = self.vec_env.step(actions, save_image_path=save_image_path)
                infos['action_t'] = actions[:, :3].mean(dim=-1)
                infos['action_r'] = actions[:, 3:6].mean(dim=-1)
                infos['action_gripper'] = actions[:, -1]
                infos['succ_rate'] = self.vec_env.success
                ep_infos.append(deepcopy(infos))
                if self.save_pose:                                  # ppo.py:177-181 (host-side bookkeeping of the env's scene)
                    save_dict = self.vec_env.save_scene_pose(pjoin(self.logger.save_pose_dir, f"Iter{self.curr_iter}", f"{i}.npy"))
                    save_dict['state'] = curr_obs.cpu().numpy()
                    save_dict['action'] = actions.cpu().numpy()
                    save_dict_lst.append(deepcopy(save_dict))
                curr_obs = next_obs[self.obs_mode]
            if self.save_pose:                                      # ppo.py:185-189
                for i in range(self.max_episode_length):
                    save_dict_lst[i]['success'] = ep_infos[-1]['obj_up_flag'].cpu().numpy()
                    np.save(pjoin(self.logger.save_pose_dir, f"Iter{self.curr_iter}", f"{i}.npy"), save_dict_lst[i])
            if self.save_video:                                     # ppo.py:191-193
                _path2video()(pjoin(self.logger.save_video_dir, f"Iter{self.curr_iter}"))
        mode = 'Test' if self.test_only else 'Val'
        self.use_info_update_logdict(ep_infos, mode)
        if self.log_dict[f'{mode}/succ_rate_max'] > 0.5 and getattr(self, 'update_RMS', False):
            self.update_RMS = False
        ep_infos.clear()

    # ------------------------------------------------------------------ rollout + learn (ppo.py:205-293)
    def _ingest(self, env_obs, out):
        """state-norm (or plain copy) of the env's observation into `out` (a rollout-buffer slot)."""
        if self.tricks['use_state_norm']:
            return self.state_norm(env_obs, update=self.update_RMS, out=out)
        return ops.copy_rows(env_obs, out)

    def collect(self, curr_obs, ep_infos=None, eps=None):
        """ppo.py:225-248: n_steps of act -> env.step -> store; returns (obs after the last step, last_values)."""
        st = self.storage
        for t in range(self.n_steps):
            actions, logp, values, mu, sigma = self.actor_critic.random_act_cri(curr_obs, None if eps is None else eps[t])
            next_obs, rews, dones, infos = self.vec_env.step(actions)
            if self.verbose:   # the reference prints every step (4 host syncs); opt-in here
                print('TrainIter: %d, SuccRate: %.3f, Finish: %d, Rew: %.3f, maxRew: %.3f' % (
                    self.curr_iter, infos['succ_rate'].data, dones.int().sum(), self.vec_env.rew_buf.mean(),
                    self.vec_env.rew_buf.max()))
            st.add_transitions(curr_obs, actions, rews, dones, self.vec_env.reset_succ, values, logp, mu, sigma)
            if ep_infos is not None:
                infos['action_t'] = actions[:, :3].abs().mean(dim=-1)
                infos['action_r'] = actions[:, 3:6].abs().mean(dim=-1)
                infos['action_gripper'] = actions[:, -1].abs()
                infos['value_pred'] = values.squeeze(-1)
                if self.tricks['use_state_norm']:
                    ms = self.state_norm.running_ms
                    infos['RMS_state_mean'] = ms.mean.mean().unsqueeze(0)
                    infos['RMS_state_std'] = ms.std.mean().unsqueeze(0)
                    infos['RMS_state_1_std'] = ms.std[:, :14].mean(dim=-1)
                    infos['RMS_state_2_std'] = ms.std[:, 14:].mean(dim=-1)
                ep_infos.append(deepcopy(infos))
            out = st.obs_slot() if t + 1 < self.n_steps else self._next_obs
            curr_obs = self._ingest(next_obs[self.obs_mode], out)
        last_values = self.actor_critic.cri(curr_obs)
        return curr_obs, last_values

    def run(self):
        if self.test_only:
            self.eval()
            self.logger.info(self.log_dict, self.curr_iter)
            return
        curr_obs = self._ingest(self.vec_env.reset()[self.obs_mode], self.storage.obs_slot())
        while self.curr_iter < self.max_iter:
            self.curr_iter += 1
            self.actor_critic.train()
            self.vec_env.train_test_flag = 'train'
            self.log_dict = {}
            ep_infos = []
            start = time.time()
            last_obs, last_values = self.collect(curr_obs, ep_infos)
            torch.cuda.synchronize()
            collection_time = time.time() - start

            start = time.time()
            self.storage.compute_returns(last_values, self.gamma, self.lam)
            self.update(self.curr_iter)
            self.storage.clear()
            # the observation after the last step opens the next rollout: move it into slot 0
            curr_obs = ops.copy_rows(last_obs, self.storage.obs_slot())
            torch.cuda.synchronize()
            learn_time = time.time() - start

            self.total_envsteps += self.n_steps * self.vec_env.num_envs
            self.total_time += collection_time + learn_time
            action_std = self.actor_critic.log_std.detach().exp()
            fps = int(self.n_steps * self.vec_env.num_envs / (collection_time + learn_time))
            self.log_dict['Progress/total_steps'] = self.curr_iter
            self.log_dict['Progress/collection_time'] = collection_time
            self.log_dict['Progress/learn_time'] = learn_time
            self.log_dict['Progress/FPS'] = fps
            self.log_dict['Train/mean_action_noise_std'] = action_std.mean().item()
            self.log_dict['Train/mean_t_noise_std'] = action_std[:3].mean()
            self.log_dict['Train/mean_r_noise_std'] = action_std[3:-1].mean()
            self.log_dict['Train/mean_gripper_noise_std'] = action_std[-1]
            self.use_info_update_logdict(ep_infos, 'Train')

            if self.curr_iter % self.eval_freq == 0:
                self.eval()
                self.storage.clear()
                curr_obs = self._ingest(self.vec_env.reset()[self.obs_mode], self.storage.obs_slot())
            if self.curr_iter % self.save_freq == 0:
                self.save(self.curr_iter)
            self.logger.info(self.log_dict, self.curr_iter)
            ep_infos.clear()

    def use_info_update_logdict(self, info_lst, mode):
        """ppo.py:295-305 (logging only; plain torch)."""
        if not info_lst:
            return
        for key in info_lst[0]:
            assert len(info_lst[0][key].shape) == 1, f"{key}: {info_lst[0][key].shape}"
            all_info = torch.stack([info[key].float() for info in info_lst], dim=-1)
            self.log_dict[f'{mode}/{key}_mean'] = torch.mean(all_info)
            self.log_dict[f'{mode}/{key}_max'] = torch.mean(all_info.max(dim=-1)[0])

    # ------------------------------------------------------------------ the update (ppo.py:307-411)
    def _minibatch(self, indices):
        """Views (sequential sampler) or gathered copies (random sampler) of the eight per-sample tensors."""
        st = self.storage
        D, A = self.num_obs, self.num_actions
        flat = dict(obs=st.observations.view(-1, D), act=st.actions.view(-1, A), val=st.values.view(-1, 1),
                    ret=st.returns.view(-1, 1), logp=st.actions_log_prob.view(-1, 1), adv=st.advantages.view(-1, 1),
                    mu=st.mu.view(-1, A), sigma=st.sigma.view(-1, A))
        if hasattr(indices, 'start'):
            return {k: v[indices.start:indices.stop] for k, v in flat.items()}
        B = indices.numel()
        out = self._mb.get(B)
        if out is None:
            out = {k: torch.empty(B, v.shape[1], device=v.device) for k, v in flat.items()}
            self._mb[B] = out
        for k, v in flat.items():
            ops.gather_rows(v, indices, out[k])
        return out

    def release_graph(self):
        """Drop the captured update graph (it pins NCCL work objects: do this before dist.destroy_process_group())."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._graph_calls = 0

    def _actor_step(self, mb, dmu, squash):
        """One minibatch step of the actor phase (ppo.py:315-357)."""
        ac = self.actor_critic
        inv_b = 1.0 / (mb['obs'].shape[0] * self.world)
        mu = ac.actor.runner.forward(mb['obs'])
        adv_stats = None
        if self.tricks['mini_adv_norm']:
            adv_stats = ops.normalize_stats(mb['adv'].reshape(-1), self._adv_stats)
        ops.ppo_actor_loss(mu, ac.log_std.data, mb['act'], mb['logp'].reshape(-1), mb['mu'], mb['sigma'],
                           mb['adv'].reshape(-1), adv_stats, inv_b, self.epsilon_clip, ac.max_action, squash,
                           self._stats_a, dmu, self._actor_grads[-1])
        ac.actor.runner.backward(mb['obs'], dmu, self._actor_grads[:-1])
        if self.optimizer_actor.fused:     # all-reduce + rank-consistent KL-skip + clip + Adam: ONE launch
            self.optimizer_actor.step_fused((inv_b, self.desired_kl, self._acc, self._skip))
        else:
            parallel.all_reduce_sum_(self.optimizer_actor.grad_ext)   # gradients + [sum surrogate, sum KL] in one collective
            ops.ppo_actor_finalize(self._stats_a, inv_b, self.desired_kl, self._acc, self._skip)   # rank-consistent KL-skip
            self.optimizer_actor.step(self._skip)

    def _critic_step(self, mb, dv):
        """One minibatch step of the critic phase (ppo.py:359-384)."""
        ac = self.actor_critic
        inv_b = 1.0 / (mb['obs'].shape[0] * self.world)
        v = ac.critic.runner.forward(mb['obs'])
        clip_delta = None
        if self.tricks['use_clipped_value_loss']:
            ops.abs_sum(mb['val'].reshape(-1), self.epsilon_clip * inv_b, self._clip_delta)
            parallel.all_reduce_sum_(self._clip_delta)
            clip_delta = self._clip_delta
        ops.value_loss(v, mb['ret'].reshape(-1), mb['val'].reshape(-1), clip_delta, inv_b, self._stats_v, dv)
        ops.accumulate(self._stats_v, inv_b, self._acc, 4)
        ac.critic.runner.backward(mb['obs'], dv, self._critic_grads)
        if self.optimizer_critic.fused:
            self.optimizer_critic.step_fused(None)
        else:
            parallel.all_reduce_sum_(self.optimizer_critic.grad)
            self.optimizer_critic.step(None)

    def _step_buffers(self, B):
        buf = self._dbuf.get(B)
        if buf is None:
            buf = (torch.empty(B, self.num_actions, device=self.device), torch.empty(B, 1, device=self.device))
            self._dbuf[B] = buf
        return buf

    def _update_body(self):
        """The device-side part of update(): both phases over all epochs and minibatches (graph-capturable).

        The reference runs the actor phase, then the critic phase (ppo.py:315-384).  The two phases touch disjoint state (different
        networks, optimisers, accumulator slots; both only READ the rollout buffer), so with `overlap_phases` they are issued on two
        streams, minibatch step by minibatch step: while one network sits in its small latency-bound kernels (head, loss, reduce,
        fused step — and, multi-GPU, waits for its peers' gradients) the other network's encoder kernels fill the GPU.  Every
        network sees exactly the sequence of operations it sees in the sequential order: results are bit-identical."""
        squash = self.actor_critic.action_activate == 'tanh'
        self._acc.zero_()
        batch = self.storage.mini_batch_generator(self.num_mini_batches)
        n_critic = 0
        if self.overlap_phases and self.storage.sampler == "sequential":
            cur = torch.cuda.current_stream()
            sa, sc = self._phase_streams
            sa.wait_stream(cur)
            sc.wait_stream(cur)
            for epoch in range(self.n_updates):
                for indices in batch:
                    mb = self._minibatch(indices)
                    dmu, dv = self._step_buffers(mb['obs'].shape[0])
                    with torch.cuda.stream(sa), ops.scratch_ns("actor/"):
                        self._actor_step(mb, dmu, squash)
                    with torch.cuda.stream(sc), ops.scratch_ns("critic/"):
                        self._critic_step(mb, dv)
                    n_critic += 1
            cur.wait_stream(sa)
            cur.wait_stream(sc)
        else:
            for epoch in range(self.n_updates):          # ---- phase 1: actor
                for indices in batch:
                    mb = self._minibatch(indices)
                    self._actor_step(mb, self._step_buffers(mb['obs'].shape[0])[0], squash)
            for epoch in range(self.n_updates):          # ---- phase 2: critic
                for indices in batch:
                    mb = self._minibatch(indices)
                    self._critic_step(mb, self._step_buffers(mb['obs'].shape[0])[1])
                    n_critic += 1
        parallel.all_reduce_sum_(self._acc[4:5])
        self._n_critic = n_critic

    def update(self, it):
        if not self.cuda_graph:
            self._update_body()
        elif self._graph is not None and self._graph_scratch_gen != ops.scratch_generation():
            # a workspace the captured kernels point into was re-allocated (another runner / a larger batch grew it): the
            # graph holds dangling pointers — drop it and start over (eager now, re-capture on the next call)
            self._graph, self._graph_calls = None, 0
            self._update_body()
        elif self._graph is not None:
            self._graph.replay()
            ops.count_launches(self._graph_launches)      # the replay launches the same kernels the capture recorded
        elif self._graph_calls == 0:
            self._update_body()                       # eager warm-up: sizes every workspace, sets kernel attributes
        else:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            # thread_local: NCCL's watchdog thread polls CUDA events while this thread captures (multi-rank)
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self._update_body()
            self._graph, self._graph_launches = graph, ops.launch_count() - n0
            self._graph_scratch_gen = ops.scratch_generation()
            graph.replay()
        self._graph_calls += 1
        n_critic = self._n_critic
        # ---- one read-back per iteration
        acc = self._acc.tolist()
        ops.check_tc_errors()
        count = int(round(acc[2]))
        mean_value_loss = acc[4] / max(n_critic, 1)
        if count == 0:
            # the reference divides by count and raises ZeroDivisionError here (ppo.py:387); keep training instead
            mean_surrogate_loss = mean_kl_mean = float('nan')
        else:
            mean_surrogate_loss = acc[0] / count
            mean_kl_mean = acc[1] / count
        # LR schedule touches the actor optimiser only (ppo.py:390-400)
        if self.lr_schedule == 'linear_decay':
            self.optimizer_actor.set_lr(max(self.lr * (1 - it / self.max_iter), 1e-5))
        elif self.lr_schedule == 'step_decay':
            self.optimizer_actor.set_lr(1e-5 if it > self.max_iter // 2 else self.lr)
        if not hasattr(self, 'log_dict'):
            self.log_dict = {}
        self.log_dict['Train/value_gt_return_mean'] = self.storage.returns.mean()
        self.log_dict['Train/value_gt_return_max'] = self.storage.returns.max()
        self.log_dict['Train/learning_rate'] = self.optimizer_actor.param_groups[0]['lr']
        self.log_dict['Train/value_function_loss'] = mean_value_loss
        self.log_dict['Train/surrogate_loss'] = mean_surrogate_loss
        self.log_dict['Train/kl'] = mean_kl_mean
        self.log_dict['Train/kl_max'] = acc[3]
        self.log_dict['Train/kl_update_count'] = count
