"""GPU parity of the DAgger path (BASELINE config 4 shapes at test scale): vision student (PointNet, 2048-pt clouds) acts, a
frozen state-MLP teacher labels, ring buffer + MSE update — against the CPU oracle's restatement of dagger.py:299-337."""
import pytest
import torch

from oracle import ppo_oracle as O
from tests.helpers import close, max_err, sub

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Logger:
    save_ckpt_dir = save_video_dir = save_pose_dir = "/tmp/pm_b200_test_dagger"

    def info(self, d, it):
        pass


def _cfg(E, net, teacher_path, **over):
    cfg = dict(num_envs=E, obs_mode="obs", max_iterations=10, n_steps=8, n_updates=2, n_minibatches=4, device=DEV, buf_size=16,
               reward_reset=False, add_proprio_obs=False, offline_data_pth=None, eval_round=1, eval_frequence=10 ** 9,
               save_frequence=10 ** 9, test_only=False, save_pose=False, save_video=False, lr_schedule="fixed", lr=1e-4,
               teacher=teacher_path, resume=None, pretrain=None, sampler="sequential",
               model=dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=net))
    cfg.update(over)
    return cfg


@pytest.mark.parametrize("point_num,precision,tol,E", [(2048, "fp32", 1e-4, 4), (1024, "bf16", 1e-2, 4), (1024, "bf16", 1e-2, 32),
                                                        (1024, "fp32", 1e-4, 32)])
def test_dagger_rollout_and_update_vs_oracle(tmp_path, point_num, precision, tol, E):
    from partmanip_b200.algorithms import dagger
    from partmanip_b200.envs import FakeVecEnv
    torch.manual_seed(3)
    A, Dt = 10, 53
    D = point_num * 3
    g = torch.Generator().manual_seed(21)
    # frozen teacher: state MLP trained without state-norm (dagger.py:73 asserts that)
    tea_cfg = dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="MLP", hid_dim=[64, 64], activation="elu"))
    tea_p = O.mlp_init(Dt, A, [64, 64], gen=g)
    tea_sd = {f"actor.{k}": v for k, v in tea_p.items()}
    tea_sd.update({f"critic.{k}": v for k, v in O.mlp_init(Dt, 1, [64, 64], gen=g).items()})
    tea_sd["log_std"] = torch.full((A,), -0.69)
    path = str(tmp_path / "teacher.pth")
    torch.save(dict(obs_mode="state", model_cfg=tea_cfg, model_state_dict=tea_sd, tricks=dict(use_state_norm=False)), path)
    net = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=point_num, precision=precision)
    env = FakeVecEnv(E, D, A, DEV, cloud=True, seed=5, extra_obs={"state": Dt})
    r = dagger(env, _cfg(E, net, path), _Logger())
    stu0 = {k[len("actor."):]: v.detach().cpu().clone() for k, v in r.student.state_dict().items() if k.startswith("actor.")}
    # ---- rollout (dagger.py:209-222): student acts, ring buffer stores both observations
    obs = env.reset()
    stu_obs, tea_obs = obs["obs"], obs["state"]
    want_stu, want_tea = [], []
    for i in range(8):
        actions = r.student.random_act(stu_obs)
        assert actions.shape == (E, A) and float(actions.abs().max()) <= 1.0
        want_stu.append(stu_obs.cpu().clone()); want_tea.append(tea_obs.cpu().clone())
        r.storage.add_transitions_dagger(stu_obs, tea_obs)
        nxt, _, _, _ = env.step(actions)
        stu_obs, tea_obs = nxt["obs"], nxt["state"]
    n = 8 * E
    assert r.storage.cur_buf_size == n and r.storage.mix_buf_ind == n
    assert torch.equal(r.storage.observations[:n].cpu(), torch.cat(want_stu)) and torch.equal(r.storage.tea_obs[:n].cpu(), torch.cat(want_tea))
    # ---- update vs oracle (sequential sampler: contiguous minibatches of cur_buf_size // n_minibatches rows)
    r.update(1)
    opt = O.AdamState({k: v.clone() for k, v in stu0.items()}, 1e-4)
    S, Tt = torch.cat(want_stu), torch.cat(want_tea)
    mbs, losses = n // 4, []
    for epoch in range(2):
        for k in range(4):
            sl = slice(k * mbs, (k + 1) * mbs)
            tea_act = O.action_activation(O.mlp_forward(tea_p, Tt[sl], "elu"), 1.0)
            losses.append(O.dagger_update_step(opt.params, opt, S[sl], tea_act, "PointNet", net, 1.0))
    assert r.optimizer.step_count == 8
    assert abs(r.log_dict["Train/dagger_loss"] - sum(losses) / len(losses)) <= tol * max(1.0, abs(sum(losses) / len(losses)))
    got = {k[len("actor."):]: v.detach().cpu() for k, v in r.student.state_dict().items() if k.startswith("actor.")}
    disp = 8 * 1e-4                                             # Adam moves every weight ~lr per early step
    for k, v in opt.params.items():
        d = (got[k] - v).abs()
        frac = float((d > 0.05 * disp).float().mean())
        assert frac <= (0.02 if precision == "fp32" else 0.25), (k, frac, float(d.max()) / disp)


def test_dagger_ring_buffer_wraps_and_small_buffer_skips_update(tmp_path):
    from partmanip_b200.algorithms.algo_utils import RolloutStorage
    st = RolloutStorage(4, 2, 6, 3, DEV, sampler="random", tea_obs_shape=5, max_length=10)     # capacity 8 rows
    for i in range(3):
        st.add_transitions_dagger(torch.full((4, 6), float(i), device=DEV), torch.full((4, 5), float(10 + i), device=DEV))
    assert st.cur_buf_size == 8 and st.mix_buf_ind == 4                                         # wrapped once
    assert float(st.observations[0, 0]) == 2.0 and float(st.observations[4, 0]) == 1.0 and float(st.tea_obs[0, 0]) == 12.0
    batches = list(st.mini_batch_generator(2))
    assert len(batches) == 2 and all(b.numel() == 4 for b in batches)
    assert sorted(torch.cat(batches).tolist()) == list(range(8))


def test_cached_teacher_labels_are_bit_identical_to_recomputing(tmp_path):
    """The frozen teacher's label of a ring row never changes: labelling rows once at insertion (default) and running the teacher on
    every minibatch visit (dagger.py:311, cfg cache_teacher_actions=False) give bit-identical students — random sampler, wrapped ring."""
    from partmanip_b200.algorithms import dagger
    from partmanip_b200.envs import FakeVecEnv
    E, A, Dt, D = 16, 10, 53, 1024 * 3
    g = torch.Generator().manual_seed(4)
    tea_cfg = dict(action_std=0.5, action_activate="tanh", clipAction=1.0, network=dict(name="MLP", hid_dim=[512, 512, 512], activation="tanh"))
    tea_sd = {f"actor.{k}": v for k, v in O.mlp_init(Dt, A, [512, 512, 512], gen=g).items()}
    tea_sd.update({f"critic.{k}": v for k, v in O.mlp_init(Dt, 1, [512, 512, 512], gen=g).items()})
    tea_sd["log_std"] = torch.full((A,), -0.69)
    path = str(tmp_path / "teacher.pth")
    torch.save(dict(obs_mode="state", model_cfg=tea_cfg, model_state_dict=tea_sd, tricks=dict(use_state_norm=False)), path)
    net = dict(name="PointNet", activation="tanh", max_mean=False, sub_mean=False, point_num=1024, precision="bf16")
    runs = []
    for cache in (True, False):
        torch.manual_seed(9)
        env = FakeVecEnv(E, D, A, DEV, cloud=True, seed=5, extra_obs={"state": Dt})
        r = dagger(env, _cfg(E, net, path, buf_size=12, n_steps=8, max_iterations=3, sampler="random", cache_teacher_actions=cache), _Logger())
        if runs:
            r.student.load_state_dict(runs[0][0])
        init = {k: v.detach().clone() for k, v in r.student.state_dict().items()}
        torch.manual_seed(10)
        r.run()                                                  # 3 iterations x 8 steps into a 12-step ring: wraps twice
        runs.append((init, {k: v.detach().clone() for k, v in r.student.state_dict().items()}, r.log_dict["Train/dagger_loss"]))
    assert runs[0][2] == runs[1][2]
    for k in runs[0][1]:
        assert torch.equal(runs[0][1][k], runs[1][1][k]), k
    assert any(not torch.equal(runs[0][1][k], runs[0][0][k]) for k in runs[0][1] if k.startswith("actor."))


def test_offline_prefill_of_the_ring(tmp_path):
    """storage.py:58-82 add_transitions_offline: recorded (tsdf [+ proprio], tea_obs) steps enter the ring in sorted scene / step order, one
    row per file, wrapping like the online path; the DAgger runner prefills before its first rollout (dagger.py:186-187)."""
    import os
    import numpy as np
    from partmanip_b200.algorithms.algo_utils import RolloutStorage
    from tests.helpers_bc import write_dataset
    R, P, T = 6, 5, 7
    n = write_dataset(str(tmp_path / "data"), seed=2, scenes=3, steps=4, R=R, A=10, P=P, T=T)        # 12 rows
    rows = []
    for sc in sorted(os.listdir(str(tmp_path / "data"))):
        for stp in sorted(os.listdir(str(tmp_path / "data" / sc))):
            d = np.load(str(tmp_path / "data" / sc / stp), allow_pickle=True).item()
            rows.append((np.concatenate((d["tsdf"].reshape(-1), d["proprio_state"].reshape(-1))), d["tea_obs"]))
    want_stu = torch.from_numpy(np.stack([r[0] for r in rows]))
    want_tea = torch.from_numpy(np.stack([r[1] for r in rows]))
    for cap_steps, chunk in ((8, 5), (2, 256), (2, 3)):                                  # capacity 4 * steps rows: 32 (no wrap) or 8 (wraps)
        st = RolloutStorage(4, cap_steps, R ** 3 + P, 10, DEV, sampler="random", tea_obs_shape=T, max_length=10)
        st.add_transitions_offline(str(tmp_path / "data"), DEV, add_proprio_obs=True, chunk=chunk)
        cap = 4 * cap_steps
        assert st.cur_buf_size == min(n, cap) and st.mix_buf_ind == n % cap and st.last_episode_buf_ind == st.mix_buf_ind and st.rows_added == n
        for i in range(n):                                                                # the last write to a slot wins
            if i >= n - cap:
                assert torch.equal(st.observations[i % cap].cpu(), want_stu[i]) and torch.equal(st.tea_obs[i % cap].cpu(), want_tea[i]), (cap, i)
    st = RolloutStorage(4, 8, R ** 3, 10, DEV, sampler="random", tea_obs_shape=T, max_length=10)
    st.add_transitions_offline(str(tmp_path / "data"), DEV, add_proprio_obs=False)
    assert torch.equal(st.observations[:n].cpu(), want_stu[:, :R ** 3])
