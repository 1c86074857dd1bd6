"""Host logic of the env-side mixins (partmanip_b200/tasks/step_kernels.py) on CPU: the kernels are replaced by the pinned oracle
(tests only), so what is checked here is the control flow of tasks/hand_base.py:363-392 as mirrored — which branch runs, what is
published under which attribute / extras key, error behaviour.  The kernels themselves are checked on the GPU (test_gpu_env_step.py)."""
import pytest
import torch

from oracle import env_oracle as EO
from tests.helpers_env import synth_state


@pytest.fixture
def task(monkeypatch):
    from partmanip_b200 import ops
    from partmanip_b200.tasks import FrankaKernels, OpenDrawerKernels
    s = synth_state(24, 5)

    class Plan:
        """Stand-in for ops.OpenDrawerPostPlan: same constructor / call contract, arithmetic by the oracle."""

        def __init__(self, dof_all, rb_all, root, obj_actor, dof_mask, rb_mask, lt, rt, lo, hi, bbox, axis, jl, ju, lstid, suc, prog, succ_obj, out):
            self.a = (dof_all, rb_all, root, obj_actor, dof_mask, rb_mask, lt, rt, lo, hi, bbox, axis, jl, ju, lstid, suc, prog, succ_obj, out)
            self._bound = dict(dof_state_all=dof_all, rigid_body_all=rb_all, root_tensor=root, progress_buf=prog, succ_objid=succ_obj)
            self.calls = []

        def bound_to(self, **t):
            return all(self._bound[k] is v for k, v in t.items())

        def __call__(self, do_obs=True, do_reward=True, advance_progress=False):
            dof_all, rb_all, root, obj_actor, dof_mask, rb_mask, lt, rt, lo, hi, bbox, axis, jl, ju, lstid, suc, prog, succ_obj, out = self.a
            self.calls.append((do_obs, do_reward, advance_progress))
            if advance_progress:
                prog += 1
            o = EO.compute_observations(dof_all, rb_all, root, dof_mask, rb_mask, obj_actor, bbox, axis, dof_mask.shape[1] - 1, lt, rt, lo, hi)
            if do_obs:
                out["obs"].copy_(o["obs"]); out["part_bbox"].copy_(o["part_bbox"]); out["dof_state_tensor"].copy_(o["dof_state_tensor"])
                out["rigid_body_tensor"].copy_(o["rigid_body_tensor"]); out["tip_rb_tensor"].copy_(o["robot"]["tip_rb_tensor"])
                out["tip_rot_9d"].copy_(o["robot"]["tip_rot_9d"]); out["gripper_length"].copy_(o["robot"]["gripper_length"])
                out["dof_qpos_normalized"].copy_(o["robot"]["dof_qpos_normalized"])
            if do_reward:
                r = EO.compute_reward(o["part_bbox"], o["robot"], o["dof_state_tensor"], jl, ju, suc, lstid, succ_obj)
                out["rew_buf"].copy_(r["rew_buf"]); out["success"].copy_(r["success"].bool()); succ_obj.copy_(r["succ_objid_lst"])
                out["extras_f"][0].copy_(r["reaching_reward"]); out["extras_f"][4].copy_(r["is_grasped"]); out["extras_f"][5].copy_(prog.float())
                out["extras_b"][2].copy_(r["is_reached"])

    monkeypatch.setattr(ops, "OpenDrawerPostPlan", Plan)

    def franka_control(raw, mode, mobile, qpos, nd, lo, hi, quat, dt, action, dof_state_mask=None, jacobian=None, ltip_rb_index=0, rtip_rb_index=0,
                       jacobian_sum=None, damping=0.05):
        q = qpos[dof_state_mask[:, :nd], 0] if dof_state_mask is not None else qpos
        root = torch.tensor([0.0, 0.0, 0.0] + list(quat)) if mobile else None
        action.copy_(EO.control(raw, mode, mobile, q, dt, root, lo, hi, jacobian, ltip_rb_index, rtip_rb_index))
        if jacobian_sum is not None:
            jacobian_sum.fill_(float(EO.solve_ik(jacobian, torch.zeros(raw.shape[0], 6, 1), ltip_rb_index, rtip_rb_index, mobile, nd)[1]))
        return action

    def episode_flags(train, rew, prog, succ, best, step, explore, max_len, reset_buf, reset_succ, counts, succ_rate):
        f = EO.episode_flags("train" if train else "test", rew, prog, succ, best, step, explore, max_len)
        reset_buf.copy_(f["reset_buf"])
        counts[1] = int(f["reset_buf"].sum())
        if train:
            step.copy_(f["epis_max_step"]); best.copy_(f["epis_max_rew"]); reset_succ.copy_(f["reset_succ"]); succ_rate.copy_(f["succ_rate"])
            counts[0] = int(succ.sum())

    def scatter(pos_act, mask, nd, pos_all):
        pos_all[mask[:, :nd]] = pos_act

    monkeypatch.setattr(ops, "franka_control", franka_control)
    monkeypatch.setattr(ops, "episode_flags", episode_flags)
    monkeypatch.setattr(ops, "scatter_dof_targets", scatter)

    class Robot(FrankaKernels):
        pass

    class Task(OpenDrawerKernels):
        def refresh_gym_tensor(self):
            self.refreshed += 1

        def reset_idx(self, buf):
            self.reset_calls.append(buf.clone())

        def _pm_set_targets(self):
            self.targets_set += 1

    E = 24
    rob = Robot()
    rob.driveMode, rob.mobile, rob.dt, rob.num_dofs = "ik", True, 1 / 60, 12
    rob.ltip_rb_index, rob.rtip_rb_index = s["ltip"], s["rtip"]
    rob.dof_lower_limits_tensor, rob.dof_upper_limits_tensor = s["dof_lower"], s["dof_upper"]
    rob.default_root = torch.tensor([0.5, 0.0, 0.05, 0.0, 0.0, 1.0, 0.0])
    rob.jacobian_tensor = s["jac"]
    t = Task()
    t.num_envs, t.robot, t.obj_actor = E, rob, 1
    t.dof_state_tensor_all, t.rigid_body_tensor_all, t.root_tensor = s["dof_all"], s["rb_all"], s["root"]
    t.dof_state_mask, t.rigid_body_mask = s["dof_mask"], s["rb_mask"]
    for k in ("part_bbox_init", "part_axis_dir_init", "part_joint_upper_limits", "part_joint_lower_limits"):
        setattr(t, k, s[k])
    t.obj_lstid_lst, t.suc_prop = s["obj_lstid"], 0.5
    t.success = torch.zeros(E).bool()
    t.succ_objid_lst = torch.zeros(s["num_objs"]).bool()
    t.obs_buf, t.extras = {}, {}
    t.progress_buf = torch.zeros(E, dtype=torch.long)
    t.rew_buf = torch.zeros(E)
    t.epis_max_rew, t.epis_max_step = -100 * torch.ones(E), torch.zeros(E, dtype=torch.long)
    t.explore_step, t.max_episode_length, t.train_test_flag = 40, 200, "train"
    t.pos_act_all = torch.zeros(s["dof_all"].shape[0])
    t.reset_calls, t.targets_set, t.refreshed = [], 0, 0
    return t, s


def test_post_physics_step_publishes_the_reference_attributes(task):
    t, s = task
    t.post_physics_step(None)
    assert t.refreshed == 1 and t._pm_plan.calls == [(True, True, True)] and bool((t.progress_buf == 1).all())
    rob = t.robot
    assert t.obs_buf["normal_state"].shape == (24, 53) and t.part_bbox.shape == (24, 8, 3)
    assert rob.tip_pos.data_ptr() == rob.tip_rb_tensor.data_ptr() and rob.tip_rot_9d.shape == (24, 3, 3)
    assert rob.dof_qpos_raw.shape == (24, 12) and torch.equal(rob.dof_qpos_raw, t.dof_state_tensor[:, :12, 0])
    assert torch.equal(rob.ltip_rb_tensor, t.rigid_body_tensor[:, s["ltip"]]) and torch.equal(t.obj_root_tensor, t.root_tensor[:, 1])
    for k in ("is_open", "is_open_notgrasp", "reaching_reward", "close_reward", "rot_reward", "is_reached", "joint_state_reward", "raw_reward",
              "is_grasped", "success_objnum", "step_id"):
        assert k in t.extras, k
    assert t.extras["raw_reward"] is t.rew_buf and t.extras["success_objnum"] is t.succ_objid_lst and t.success.dtype == torch.bool
    # the separate entry points launch their own halves; a replaced progress buffer forces a new plan
    t.compute_observations()
    t.compute_reward(None)
    assert t._pm_plan.calls[-2:] == [(True, False, False), (False, True, False)]
    old = t._pm_plan
    t.progress_buf = t.progress_buf.clone()
    t.compute_observations()
    assert t._pm_plan is not old


def test_pre_physics_step_branches(task):
    t, s = task
    t.post_physics_step(None)
    acts = torch.rand(24, 10) * 2 - 1
    t.success[:] = False
    t.progress_buf[:] = 1
    t.epis_max_step[:] = 0
    t.pre_physics_step(acts)                                      # nobody stalled, nobody succeeded: targets go to the simulator
    assert t.targets_set == 1 and not t.reset_calls and t.pos_act.shape == (24, 12) and t.pos_act is t.robot.action_tensor
    want = torch.zeros_like(t.pos_act_all)
    want[t.dof_state_mask[:, :12]] = t.pos_act
    assert torch.equal(t.pos_act_all, want) and float(t.extras["succ_rate"]) == 0.0
    t.success[3] = True
    t.pre_physics_step(acts)                                      # one success: reset_idx gets the flags, no target upload
    assert t.targets_set == 1 and len(t.reset_calls) == 1 and bool(t.reset_calls[0][3]) and int(t.reset_calls[0].sum()) == 1
    assert bool(t.reset_succ[3]) and float(t.extras["succ_rate"]) == 1.0
    t.train_test_flag, t.max_episode_length = "test", 1
    t.pre_physics_step(acts)                                      # test mode: everybody past the episode length
    assert len(t.reset_calls) == 2 and bool(t.reset_calls[1].all())
    t.train_test_flag = "validate"
    with pytest.raises(NotImplementedError):
        t.pre_physics_step(acts)


def test_franka_control_modes_and_jacobian_guard(task):
    t, s = task
    t.post_physics_step(None)
    rob = t.robot
    rob.driveMode = "ik_abs"
    with pytest.raises(NotImplementedError):
        rob.control(torch.zeros(24, 11))
    rob.driveMode = "ik"
    rob.jacobian_tensor = torch.zeros_like(rob.jacobian_tensor)   # the reference prints and exits on an all-zero Jacobian (load_robot.py:145-147)
    with pytest.raises(SystemExit):
        rob.control(torch.zeros(24, 10))
    rob.check_jacobian = False                                    # opting out of the per-step host read-back
    rob.control(torch.zeros(24, 10))
