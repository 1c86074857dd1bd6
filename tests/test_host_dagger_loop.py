"""CPU: the host-side control flow of the `dagger` runner (algorithms/dagger.py:114-297 in the reference) with the CUDA pieces
replaced by fakes — rollout / ring-buffer bookkeeping, evaluation cadence and averaging, log keys, the reward-reset rule."""
import importlib

import torch

D = importlib.import_module("partmanip_b200.algorithms.dagger")


class _Env:
    def __init__(self):
        self.num_envs, self.k, self.resets = 4, 0, 0
        self.progress_buf = torch.zeros(4, dtype=torch.long)

    def _obs(self):
        g = torch.Generator().manual_seed(self.k)
        self.k += 1
        return {'pc': torch.rand(4, 6, generator=g), 'state': torch.rand(4, 3, generator=g)}

    def reset(self):
        self.resets += 1
        return self._obs()

    def step(self, actions, save_image_path=None):
        self.progress_buf = self.progress_buf + 1
        g = torch.Generator().manual_seed(1000 + self.k)
        return self._obs(), torch.rand(4, generator=g), torch.zeros(4), {'succ': torch.rand(4, generator=g)}


class _Student:
    log_std = torch.full((7,), -0.5)

    def __init__(self):
        self.calls = []

    def train(self):
        pass

    def eval(self):
        pass

    def act(self, o):
        self.calls.append('act')
        return torch.tanh(o.sum(1, keepdim=True).repeat(1, 7))

    def random_act(self, o):
        self.calls.append('rand')
        return torch.tanh(o.mean(1, keepdim=True).repeat(1, 7))


class _Storage:
    cur_buf_size, succ_buf_ind, mix_buf_ind = 0, 3, 5

    def __init__(self):
        self.pairs = []

    def add_transitions_dagger(self, stu, tea):
        assert stu.shape == (4, 6) and tea.shape == (4, 3)              # student obs first, teacher obs second
        self.pairs.append((stu, tea))
        self.cur_buf_size += 1


class _Logger:
    save_video_dir = '/tmp/unused'

    def __init__(self):
        self.rows = []

    def info(self, d, it):
        self.rows.append((it, dict(d)))


def _runner(monkeypatch, **over):
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    r = object.__new__(D.dagger)
    r.vec_env, r.logger, r.student, r.storage = _Env(), _Logger(), _Student(), _Storage()
    r.teacher = _Student()
    r.stu_obs_mode, r.tea_obs_mode = 'pc', 'state'
    r.n_steps, r.max_iter, r.curr_iter, r.eval_freq, r.save_freq, r.eval_round = 3, 4, 0, 2, 100, 2
    r.max_episode_length, r.test_only, r.save_video, r.save_pose, r.reward_reset = 5, False, False, False, False
    r.cache_teacher_actions = False
    r.offline_data_pth = None
    r.total_envsteps = r.total_time = 0
    r.update = lambda it: r.log_dict.update({'Train/learning_rate': 0.1, 'Train/dagger_loss': 1.0 / it})
    for k, v in over.items():
        setattr(r, k, v)
    return r


def test_training_loop_bookkeeping(monkeypatch):
    r = _runner(monkeypatch)
    r.run()
    assert [it for it, _ in r.logger.rows] == [1, 2, 3, 4] and r.curr_iter == 4
    assert len(r.storage.pairs) == 3 * 4 and r.total_envsteps == 3 * 4 * 4
    # one reset at the start + per evaluation (iterations 2 and 4): eval_round resets inside + one afterwards
    assert r.vec_env.resets == 1 + 2 * (2 + 1)
    assert r.student.calls.count('rand') == 12 and r.student.calls.count('act') == 2 * 2 * 5
    first, second = r.logger.rows[0][1], r.logger.rows[1][1]
    for key in ('Progress/total_steps', 'Progress/collection_time', 'Progress/learn_time', 'Progress/FPS', 'Train/learning_rate',
                'Train/dagger_loss', 'Train/mean_action_noise_std', 'Train/cur_buf_size', 'Train/succ_buf_ind', 'Train/mix_buf_ind',
                'Train/succ_mean', 'Train/succ_max', 'Train/action_t_mean', 'Train/action_r_max', 'Train/action_gripper_mean'):
        assert key in first, key
    assert not any(k.startswith('Val/') for k in first) and 'Val/succ_mean' in second and 'Val/reward_max' in second
    assert abs(first['Train/mean_action_noise_std'] - float(torch.exp(torch.tensor(-0.5)))) < 1e-6
    assert first['Train/cur_buf_size'] == 3 and second['Train/dagger_loss'] == 0.5


def test_eval_averages_over_rounds(monkeypatch):
    r = _runner(monkeypatch, test_only=True)
    r.run()
    (it, row), = r.logger.rows
    assert it == 0 and set(k.split('/')[0] for k in row) == {'Test'}
    # replay the two evaluation rounds by hand: mean over (envs, steps) per round, averaged over the rounds
    env, stu, want = _Env(), _Student(), 0.0
    for _ in range(2):
        obs = env.reset()['pc']
        vals = []
        for _ in range(5):
            nxt, rews, _, infos = env.step(stu.act(obs))
            vals.append(rews)
            obs = nxt['pc']
        want += float(torch.stack(vals, -1).mean()) / 2
    assert abs(float(row['Test/reward_mean']) - want) < 1e-6


def test_reward_reset_rule(monkeypatch):
    r = _runner(monkeypatch, reward_reset=True, max_iter=5, eval_freq=100, tea_rew=torch.linspace(0, 1, 50))
    r.run()
    prog = r.vec_env.progress_buf
    assert int(prog[0]) == 15 and r.vec_env.dagger_reward_reset.shape == (4,) and r.vec_env.dagger_reward_reset.dtype == torch.bool


def test_dagger_optimizer_state_dict_is_the_reference_adams():
    """dagger.py:56 builds Adam(student.parameters()): parameter order [log_std, actor.*, critic.*], only the actor tensors ever
    get a gradient (torch's Adam keeps no state for the others).  The flat optimiser must read and write exactly that dict:
    actor moments at indices 1..n, one param group over ALL tensors — checked against a real torch.optim.Adam in both
    directions (a reference checkpoint resumes here; a checkpoint written here resumes in the reference)."""
    import torch.nn as nn
    from partmanip_b200.algorithms.algo_utils import ActorCritic
    torch.manual_seed(0)
    model_cfg = dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="MLP", hid_dim=[16, 8], activation="tanh"))
    r = D.dagger.__new__(D.dagger)
    r.stu_input_obs, r.num_actions, r.model_cfg, r.device, r.lr = 6, 3, model_cfg, "cpu", 1e-3
    r._build_student(0)
    names = [k for k, _ in r.student.named_parameters()]
    assert names[0] == "log_std" and names[1].startswith("actor.") and names[-1].startswith("critic.")
    n_actor = sum(k.startswith("actor.") for k in names)
    # the reference side: a plain module with identical parameter order, two Adam steps on actor gradients only
    twin = ActorCritic(6, 3, model_cfg)
    twin.load_state_dict(r.student.state_dict())
    ref_opt = torch.optim.Adam(twin.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    for _ in range(2):
        for k, p in twin.named_parameters():
            p.grad = torch.randn(p.shape, generator=g) if k.startswith("actor.") else None
        ref_opt.step()
    ref_sd = ref_opt.state_dict()
    assert sorted(ref_sd["state"]) == list(range(1, 1 + n_actor)) and ref_sd["param_groups"][0]["params"] == list(range(len(names)))
    # reference checkpoint -> here
    r.optimizer.load_state_dict(ref_sd)
    assert r.optimizer.step_count == 2
    for (k, p), m in zip([kp for kp in twin.named_parameters() if kp[0].startswith("actor.")], r.optimizer._views(r.optimizer.exp_avg)):
        assert torch.equal(m, ref_sd["state"][names.index(k)]["exp_avg"]), k
    # here -> the reference's Adam accepts it and holds the same moments
    ours = r.optimizer.state_dict()
    assert sorted(ours["state"]) == sorted(ref_sd["state"]) and ours["param_groups"][0]["params"] == list(range(len(names)))
    ref2 = torch.optim.Adam(ActorCritic(6, 3, model_cfg).parameters(), lr=1.0)
    ref2.load_state_dict(ours)
    for i in ours["state"]:
        assert torch.equal(ref2.state_dict()["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"])
    # a dict with another layout (e.g. written for the actor alone) is refused instead of being misassigned
    import pytest
    bad = {"state": {}, "param_groups": [{"lr": 1e-3, "params": list(range(n_actor))}]}
    with pytest.raises(ValueError, match="covers"):
        r.optimizer.load_state_dict(bad)
