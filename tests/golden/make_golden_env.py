"""Golden vectors for the env-side kernels: runs the UNMODIFIED reference methods
  tasks/open_drawer.py:170-238 (compute_reward), :240-281 (compute_observations),
  tasks/load_robot.py:96-151 (franka.control + solve_ik), :153-164 (update_state),
  tasks/hand_base.py:363-385 (pre_physics_step's episode bookkeeping)
on CPU over a synthetic simulator state and records inputs and outputs in tests/golden/env_open_drawer.npz.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_env.py
Isaac Gym itself is absent; tests/golden/isaacgym_stub restates the quaternion helpers of isaacgym.torch_utils that the
reference's arithmetic calls (quat_rotate through quat_axis, tensor_clamp)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "isaacgym_stub"))
sys.path.insert(0, "/root/reference")
_u = types.ModuleType("utils")
_u.__path__ = ["/root/reference/utils"]
_u.TSDFVolume = _u.gen_camera_pose = _u.TSDFfromMesh = None
sys.modules["utils"] = _u
import tasks  # noqa: E402,F401
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.helpers_env import synth_state, synth_state_cube  # noqa: E402
from tasks.load_robot import franka  # noqa: E402

open_drawer = sys.modules["tasks.open_drawer"].open_drawer
grasp_cube = sys.modules["tasks.grasp_cube"].grasp_cube
BaseTask = sys.modules["tasks.hand_base"].BaseTask


class _Gym:
    def set_dof_position_target_tensor(self, sim, t):
        pass


def build_reference_task(s, drive_mode="ik", mobile=True):
    """The reference classes, constructed WITHOUT their __init__ (which needs the simulator); attributes as their __init__ sets them."""
    E = s["E"]
    rob = franka.__new__(franka)
    rob.device, rob.num_envs, rob.dt = "cpu", E, 1.0 / 60.0
    rob.driveMode, rob.mobile = drive_mode, mobile
    rob.num_dofs, rob.num_rigid_body = s["num_dofs"], s["nb_robot"]
    rob.ltip_rb_index, rob.rtip_rb_index = s["ltip"], s["rtip"]
    rob.dof_lower_limits_tensor, rob.dof_upper_limits_tensor = s["dof_lower"].clone(), s["dof_upper"].clone()
    rob.default_root = torch.tensor([0.5, 0.0, 0.05, 0.0, 0.0, 1.0, 0.0])
    rob.action_tensor = torch.zeros(E, s["num_dofs"])
    rob.jacobian_tensor = s["jac"].clone()
    t = open_drawer.__new__(open_drawer)
    t.num_envs, t.device, t.robot = E, "cpu", rob
    t.gym, t.sim = _Gym(), None
    t.dof_state_tensor_all, t.rigid_body_tensor_all, t.root_tensor = s["dof_all"].clone(), s["rb_all"].clone(), s["root"].clone()
    t.dof_state_mask, t.rigid_body_mask = s["dof_mask"], s["rb_mask"]
    t.obj_actor = 1
    for k in ("part_bbox_init", "part_axis_dir_init", "part_joint_upper_limits", "part_joint_lower_limits"):
        setattr(t, k, s[k].clone())
    t.obj_lstid_lst = s["obj_lstid"].clone()
    t.suc_prop = 0.5
    t.success = torch.zeros(E).bool()
    t.succ_objid_lst = torch.zeros(s["num_objs"]).bool()
    t.obs_buf, t.extras = {}, {}
    t.progress_buf = torch.zeros(E, dtype=torch.long)
    t.rew_buf = torch.zeros(E)
    t.reset_buf = torch.zeros(E, dtype=torch.long)
    t.epis_max_rew = -100 * torch.ones(E)
    t.epis_max_step = torch.zeros(E, dtype=torch.long)
    t.explore_step, t.max_episode_length = 40, 200
    t.train_test_flag = "train"
    t.pos_act_all = torch.zeros(s["dof_all"].shape[0])
    t.reset_calls = []
    t.reset_idx = lambda buf: t.reset_calls.append(buf.clone())
    return t


def run_reference(s, seed):
    out = {}
    g = torch.Generator().manual_seed(seed + 1)
    t = build_reference_task(s)
    t.progress_buf += 7
    t.compute_observations()
    out["obs"] = t.obs_buf["normal_state"].clone()
    out["part_bbox"] = t.part_bbox.clone()
    out["dof_state_tensor"], out["rigid_body_tensor"] = t.dof_state_tensor.clone(), t.rigid_body_tensor.clone()
    for k in ("tip_rb_tensor", "tip_rot_9d", "gripper_length", "dof_qpos_normalized", "dof_qpos_raw", "dof_qvel_raw"):
        out["robot_" + k] = getattr(t.robot, k).clone()
    actions = torch.rand(s["E"], 10, generator=g) * 2 - 1
    t.compute_reward(actions)
    out["rew_buf"], out["success"], out["succ_objid_lst"] = t.rew_buf.clone(), t.success.clone(), t.succ_objid_lst.clone()
    for k, v in t.extras.items():
        out["extras_" + k] = v.clone().float() if v.dtype == torch.bool else v.clone()
    # --- control: ik + mobile base (the shipped drive mode, cfg/tasks/open_drawer.yaml:31-32)
    out["actions"] = actions
    out["action_tensor_ik_mobile"] = t.robot.control(actions).clone()
    # --- pre_physics_step bookkeeping, train then test (robot.control runs again inside; same result)
    t.progress_buf = torch.randint(0, 120, (s["E"],), generator=g)
    t.epis_max_rew = torch.where(torch.rand(s["E"], generator=g) < 0.5, t.rew_buf + 0.3, t.rew_buf - 0.3)
    t.epis_max_step = torch.randint(0, 90, (s["E"],), generator=g)
    out["pre_progress_buf"], out["pre_epis_max_rew"], out["pre_epis_max_step"] = t.progress_buf.clone(), t.epis_max_rew.clone(), t.epis_max_step.clone()
    BaseTask.pre_physics_step(t, actions)
    out["train_epis_max_step"], out["train_epis_max_rew"] = t.epis_max_step.clone(), t.epis_max_rew.clone()
    out["train_reset_buf"], out["train_reset_succ"], out["train_succ_rate"] = t.reset_buf.clone(), t.reset_succ.clone(), t.extras["succ_rate"].clone()
    out["train_reset_called"] = torch.tensor(len(t.reset_calls))
    out["train_pos_act_all"] = t.pos_act_all.clone()
    t.train_test_flag = "test"
    t.max_episode_length = 60
    BaseTask.pre_physics_step(t, actions)
    out["test_reset_buf"] = t.reset_buf.clone()
    # --- control: pos drive, fixed base (9 dofs) and mobile
    for mobile in (False, True):
        nd = s["num_dofs"]
        t2 = build_reference_task(s, "pos", mobile)
        t2.compute_observations()
        a = torch.rand(s["E"], 8 + 3 * mobile, generator=g) * 2 - 1
        if not mobile:                                           # a fixed-base arm has 3 fewer dofs: reuse the state's last 9
            t2.robot.num_dofs = nd - 3
            t2.robot.dof_qpos_raw = t2.robot.dof_qpos_raw[:, 3:].contiguous()
            t2.robot.action_tensor = torch.zeros(s["E"], nd - 3)
            t2.robot.dof_lower_limits_tensor, t2.robot.dof_upper_limits_tensor = s["dof_lower"][3:].clone(), s["dof_upper"][3:].clone()
        out[f"actions_pos_{int(mobile)}"] = a
        out[f"action_tensor_pos_{int(mobile)}"] = t2.robot.control(a).clone()
    # --- control: ik, fixed base
    t3 = build_reference_task(s, "ik", False)
    t3.compute_observations()
    nd = s["num_dofs"]
    t3.robot.num_dofs = nd - 3
    t3.robot.dof_qpos_raw = t3.robot.dof_qpos_raw[:, 3:].contiguous()
    t3.robot.action_tensor = torch.zeros(s["E"], nd - 3)
    t3.robot.dof_lower_limits_tensor, t3.robot.dof_upper_limits_tensor = s["dof_lower"][3:].clone(), s["dof_upper"][3:].clone()
    t3.robot.jacobian_tensor = s["jac"][..., 3:].contiguous()
    a = torch.rand(s["E"], 7, generator=g) * 2 - 1
    out["actions_ik_fixed"] = a
    out["action_tensor_ik_fixed"] = t3.robot.control(a).clone()
    return out


def run_reference_cube(s, seed, add_proprio=False):
    """tasks/grasp_cube.py:66-138 on the unmodified class (constructed without its simulator-bound __init__)."""
    E = s["E"]
    rob = franka.__new__(franka)
    rob.device, rob.num_envs, rob.dt, rob.driveMode, rob.mobile = "cpu", E, 1.0 / 60.0, "ik", False
    rob.num_dofs, rob.ltip_rb_index, rob.rtip_rb_index = s["num_dofs"], s["ltip"], s["rtip"]
    rob.dof_lower_limits_tensor, rob.dof_upper_limits_tensor = s["dof_lower"].clone(), s["dof_upper"].clone()
    rob.default_root = torch.tensor([0.0, -0.5, 0.0, 0.0, 0.0, 0.707, 0.707])
    rob.action_tensor = torch.zeros(E, s["num_dofs"])
    rob.jacobian_tensor = s["jac"].clone()
    t = grasp_cube.__new__(grasp_cube)
    t.num_envs, t.device, t.robot, t.obj_actor = E, "cpu", rob, 1
    t.reset_range = 0.15
    t.pose_lower_limit = torch.tensor([-0.15, -0.15, 0.0, -1, -1, -1, -1], dtype=torch.float)
    t.pose_upper_limit = torch.tensor([0.15, 0.15, 0.4, 1, 1, 1, 1], dtype=torch.float)
    t.dof_state_tensor, t.rigid_body_tensor, t.root_tensor = s["dof"].clone(), s["rb"].clone(), s["root"].clone()
    t.goal_thresh = 0.025
    t.success_pos = torch.tensor([0, 0, 0.2])[None, :]
    t.obj_default_root = torch.tensor([0, 0, 0.025, 0, 0, 0, 1], dtype=torch.float)
    t.success = torch.zeros(E).bool()
    t.obs_buf, t.extras = {}, {}
    t.progress_buf = torch.zeros(E, dtype=torch.long) + 5
    t.learn_input_mode, t.add_proprio_obs = ("depth_pc" if add_proprio else "normal_state"), add_proprio
    out = {}
    if add_proprio:
        g = torch.Generator().manual_seed(seed + 9)
        t.obs_buf["depth_pc"] = torch.randn(E, 48, generator=g)
        out["vision_obs"] = t.obs_buf["depth_pc"].clone()
    t.compute_observations()
    out["obs"] = t.obs_buf["normal_state"].clone()
    if add_proprio:
        out["proprio_state"], out["vision_cat"] = t.obs_buf["proprio_state"].clone(), t.obs_buf["depth_pc"].clone()
    for k in ("tip_rb_tensor", "tip_rot_9d", "gripper_length", "dof_qpos_normalized", "dof_qpos_raw", "dof_qvel_raw"):
        out["robot_" + k] = getattr(t.robot, k).clone()
    t.compute_reward(None)
    out["rew_buf"], out["success"] = t.rew_buf.clone(), t.success.clone()
    for k, v in t.extras.items():
        out["extras_" + k] = v.clone().float() if v.dtype == torch.bool else v.clone()
    g = torch.Generator().manual_seed(seed + 2)
    a = torch.rand(E, 7, generator=g) * 2 - 1
    out["actions"] = a
    out["action_tensor_ik_fixed"] = t.robot.control(a).clone()
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)
    s = synth_state(96, 20260)
    out = run_reference(s, 20260)
    rec = {"in_" + k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in s.items()}
    rec.update({k: v.numpy() for k, v in out.items()})
    path = os.path.join(HERE, "env_open_drawer.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("extras_is_reached", "extras_is_grasped", "success", "extras_is_open", "train_reset_buf", "test_reset_buf"):
        print(k, float(out[k].float().mean()))
    print("rew", out["rew_buf"][:6], "succ_rate", out["train_succ_rate"], "reset calls", out["train_reset_called"])
    sc = synth_state_cube(96, 20261)
    rec = {"in_" + k: (v.numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in sc.items()}
    oc = run_reference_cube(sc, 20261)
    rec.update({k: v.numpy() for k, v in oc.items()})
    rec.update({"pp_" + k: v.numpy() for k, v in run_reference_cube(sc, 20261, add_proprio=True).items() if k in ("vision_obs", "proprio_state", "vision_cat")})
    path = os.path.join(HERE, "env_grasp_cube.npz")
    np.savez_compressed(path, **rec)
    print("wrote", path, os.path.getsize(path), "bytes")
    for k in ("extras_is_reached", "success", "extras_obj_up_flag"):
        print(k, float(oc[k].float().mean()))
