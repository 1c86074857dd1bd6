"""RolloutStorage — mirror of algorithms/algo_utils/storage.py:7-138.

Same attributes and methods; GAE runs in the K5 kernel, the sequential sampler yields contiguous
ranges (a slice of the flattened (T*E, .) buffers, no gather) and the random sampler yields
device index tensors consumed by pm_gather_rows.
"""
from __future__ import annotations

import os
from os.path import join as pjoin

import numpy as np
import torch

from ... import ops


class _Range:
    """A contiguous minibatch [start, stop) of the flattened buffer; iterable/len like the reference's index list."""
    __slots__ = ("start", "stop")

    def __init__(self, start, stop):
        self.start, self.stop = start, stop

    def __len__(self):
        return self.stop - self.start

    def __iter__(self):
        return iter(range(self.start, self.stop))

    def __getitem__(self, i):
        return range(self.start, self.stop)[i]


# PPO rollout fields (storage.py:28-41): name -> (trailing width: 'obs' / 'act' / 1, dtype)
_PPO_FIELDS = {"observations": ("obs", torch.float32), "rewards": (1, torch.float32), "actions": ("act", torch.float32),
               "dones": (1, torch.bool), "succs": (1, torch.bool), "actions_log_prob": (1, torch.float32),
               "values": (1, torch.float32), "returns": (1, torch.float32), "advantages": (1, torch.float32),
               "mu": ("act", torch.float32), "sigma": ("act", torch.float32), "step_id": (1, torch.float32)}


class RolloutStorage:

    def __init__(self, num_envs, n_steps, obs_shape, actions_shape, device, default_succ_value=0, whole_adv_norm=False,
                 sampler='sequential', tea_obs_shape=None, max_length=None):
        self.device, self.sampler = device, sampler
        self.n_steps, self.num_envs = n_steps, num_envs
        self.whole_adv_norm, self.default_succ_value = whole_adv_norm, default_succ_value
        self.max_episode_length = max_length
        self.first_fill = True
        self.step = 0
        rows = n_steps * num_envs
        if tea_obs_shape is not None:   # DAgger ring buffer (storage.py:20-27): flat (rows, .) arrays + ring cursors
            for name, width in (("tea_obs", tea_obs_shape), ("observations", obs_shape), ("succ_flag", 1)):
                setattr(self, name, torch.zeros(rows, width, device=device))
            self.mix_buf_ind = self.cur_buf_size = self.last_episode_buf_ind = 0
            self.rows_added = 0         # total rows ever written (lets a label cache notice rows it has not seen)
            self.succ_buf_ind = (max_length or 0) * num_envs
        else:                           # PPO (storage.py:28-41): (T, E, .) arrays, always full
            widths = {"obs": obs_shape, "act": actions_shape, 1: 1}
            for name, (w, dtype) in _PPO_FIELDS.items():
                setattr(self, name, torch.zeros(n_steps, num_envs, widths[w], device=device, dtype=dtype))
            self.cur_buf_size = rows

    def _slot_index(self):
        if self.step >= self.n_steps:
            raise AssertionError("Rollout buffer overflow")
        return self.step

    def obs_slot(self):
        """The (E, D) view the next add_transitions will fill — producers may write into it directly."""
        return self.observations[self._slot_index()]

    def add_transitions(self, observations, actions, rewards, dones, succs, values, actions_log_prob, mu, sigma):
        t = self._slot_index()
        slot = self.observations[t]
        if observations.data_ptr() != slot.data_ptr():          # already produced in place: nothing to move
            ops.copy_rows(observations, slot)
        per_env = dict(rewards=rewards, dones=dones, succs=succs, values=values, actions_log_prob=actions_log_prob)
        for name, column in per_env.items():                    # (E,) or (E,1) producers -> the (E,1) slot
            getattr(self, name)[t].copy_(column.view(-1, 1))
        for name, block in (("actions", actions), ("mu", mu), ("sigma", sigma)):
            getattr(self, name)[t].copy_(block)
        self.step = t + 1

    def add_transitions_dagger(self, stu_obs, tea_obs):
        """storage.py:84-91."""
        i = self.mix_buf_ind
        ops.copy_rows(stu_obs, self.observations[i:i + self.num_envs])
        ops.copy_rows(tea_obs, self.tea_obs[i:i + self.num_envs])
        self.rows_added += self.num_envs
        max_buf_size = self.n_steps * self.num_envs
        self.mix_buf_ind = (self.mix_buf_ind + self.num_envs) % max_buf_size
        if self.cur_buf_size < max_buf_size:
            self.cur_buf_size += self.num_envs

    def add_transitions_offline(self, folder, device, add_proprio_obs=False, chunk=256):
        """storage.py:58-82: pre-fill the DAgger ring from recorded steps, <folder>/<scene>/<step>.npy (sorted), each a pickled dict
        (tsdf, proprio_state, tea_obs), one ring row per file, wrapping like the online path.  The reference copies row by row; here
        the rows are staged on the host and moved `chunk` at a time."""
        print('Read offline data from ', folder)
        scene_list = sorted(os.listdir(folder))
        step_list = sorted(os.listdir(pjoin(folder, scene_list[0])))
        max_buf_size = self.n_steps * self.num_envs
        stu_rows, tea_rows = [], []

        def flush():
            k = len(stu_rows)
            if not k:
                return
            stu = torch.from_numpy(np.stack(stu_rows)).to(device)
            tea = torch.from_numpy(np.stack(tea_rows)).to(device)
            done = 0
            while done < k:                                              # split where the ring wraps
                n = min(k - done, max_buf_size - self.mix_buf_ind)
                ops.copy_rows(stu[done:done + n], self.observations[self.mix_buf_ind:self.mix_buf_ind + n])
                ops.copy_rows(tea[done:done + n], self.tea_obs[self.mix_buf_ind:self.mix_buf_ind + n])
                self.mix_buf_ind = (self.mix_buf_ind + n) % max_buf_size
                self.cur_buf_size = min(self.cur_buf_size + n, max_buf_size)
                self.rows_added += n
                done += n
            self.last_episode_buf_ind = self.mix_buf_ind
            stu_rows.clear()
            tea_rows.clear()
        for scene in scene_list:
            for step in step_list:
                data = np.load(pjoin(folder, scene, step), allow_pickle=True).item()
                tsdf = np.asarray(data['tsdf'], dtype=np.float32).reshape(-1)
                stu_rows.append(np.concatenate((tsdf, np.asarray(data['proprio_state'], dtype=np.float32).reshape(-1))) if add_proprio_obs else tsdf)
                tea_rows.append(np.asarray(data['tea_obs'], dtype=np.float32).reshape(-1))
                if len(stu_rows) == chunk:
                    flush()
        flush()

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam):
        """storage.py:96-114 in one K5 launch (+ one for whole_adv_norm)."""
        sv = self.default_succ_value
        ops.gae(self.rewards, self.values, self.dones, self.succs if sv is not None else None,
                last_values.contiguous().view(-1), gamma, lam, sv, self.returns, self.advantages)
        if self.whole_adv_norm:
            ops.normalize_(self.advantages)

    def mini_batch_generator(self, num_mini_batches):
        """storage.py:125-138: size = min(buf // nmb, 2048), drop_last.  Returns a list (re-iterable like BatchSampler)."""
        batch_size = self.cur_buf_size
        mini_batch_size = min(int(batch_size // num_mini_batches), 2048)
        if mini_batch_size <= 0:
            return []
        count = batch_size // mini_batch_size
        if self.sampler == "sequential":
            return [_Range(k * mini_batch_size, (k + 1) * mini_batch_size) for k in range(count)]
        elif self.sampler == "random":
            return _RandomBatches(batch_size, mini_batch_size, count, self.device)
        raise NotImplementedError(self.sampler)


class _RandomBatches:
    """SubsetRandomSampler + BatchSampler(drop_last=True) (storage.py:133-137): a fresh permutation on every iteration.

    The permutation is drawn exactly where the reference draws it — `torch.randperm(n)` on the HOST default generator
    (torch.utils.data.SubsetRandomSampler.__iter__) — so that for a given `torch.manual_seed` every minibatch holds the rows
    the reference's holds; one pinned, asynchronous copy moves the n indices to the device."""

    def __init__(self, n, size, count, device):
        self.n, self.size, self.count, self.device = n, size, count, device
        self._pinned = torch.empty(n, dtype=torch.int64).pin_memory() if str(device).startswith("cuda") else None

    def __len__(self):
        return self.count

    def __iter__(self):
        perm = torch.randperm(self.n)
        if self._pinned is not None:
            torch.cuda.current_stream().synchronize()      # the previous epoch's copy out of the pinned buffer has finished
            self._pinned.copy_(perm)
            perm = self._pinned.to(self.device, non_blocking=True)
        for k in range(self.count):
            yield perm[k * self.size:(k + 1) * self.size]
