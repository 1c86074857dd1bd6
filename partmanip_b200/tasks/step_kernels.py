"""Mixins with the reference's method names and attribute contract; each method is one launch of partmanip_b200/csrc/env_step.cu.

    class franka(FrankaKernels, <reference franka>): ...            # tasks/load_robot.py:6
    class open_drawer(OpenDrawerKernels, BaseTask): ...             # tasks/open_drawer.py:15

The mixins read the attributes the reference's __init__ creates (dof_state_tensor_all, rigid_body_tensor_all, root_tensor,
dof_state_mask, rigid_body_mask, part_bbox_init, ... and robot.{num_dofs, ltip_rb_index, ...}) and publish the attributes its
methods publish (obs_buf['normal_state'], rew_buf, success, extras[...], part_bbox, robot.tip_rb_tensor, ...).  Result tensors
live in buffers allocated once and overwritten every step (the reference allocates new ones each step; its consumers copy,
storage.py:46).  No CPU fallback: every tensor must be on the GPU.
"""
from __future__ import annotations

import sys

import torch

from .. import ops


class FrankaKernels:
    """tasks/load_robot.py: franka.control (:96-118, 'pos' and 'ik' drive modes, fixed or mobile base) + solve_ik (:142-151)."""

    #: the reference tests `j_eef.sum().abs() < 1e-5` on the host every step (a device sync) and exits; False skips the read-back
    check_jacobian = True

    def control(self, raw_output):
        if self.driveMode not in ("pos", "ik"):
            raise NotImplementedError                      # 'ik_abs' / 'heuristic' (debug modes) stay on the reference's code
        E = raw_output.shape[0]
        if getattr(self, "_pm_action", None) is None or self._pm_action.shape[0] != E:
            self._pm_action = torch.zeros(E, self.num_dofs, device=raw_output.device)
            self._pm_jsum = torch.zeros(1, device=raw_output.device)
        ik = self.driveMode == "ik"
        mask = getattr(self, "_pm_dof_state_mask", None)  # set by OpenDrawerKernels: read the simulator tensor in place
        qpos = self._pm_dof_state_all if mask is not None else self.dof_qpos_raw
        ops.franka_control(raw_output.contiguous(), self.driveMode, self.mobile, qpos, self.num_dofs, self.dof_lower_limits_tensor,
                           self.dof_upper_limits_tensor, self._root_quat(), self.dt, self._pm_action, dof_state_mask=mask,
                           jacobian=self.jacobian_tensor if ik else None, ltip_rb_index=self.ltip_rb_index, rtip_rb_index=self.rtip_rb_index,
                           jacobian_sum=self._pm_jsum if ik else None)
        if ik and self.check_jacobian and abs(float(self._pm_jsum)) < 1e-5:
            print('Jacobian has problem!')
            sys.exit(1)
        self.action_tensor = self._pm_action
        return self.action_tensor

    def _root_quat(self):
        if getattr(self, "_pm_root_quat", None) is None:
            self._pm_root_quat = [float(v) for v in self.default_root[3:7].tolist()] if self.mobile else None
        return self._pm_root_quat


class BaseTaskKernels:
    """tasks/hand_base.py: pre_physics_step (:363-385).  Needs `_pm_flag_buffers()` and `_pm_dof_mask()` from the task mixin."""

    def _pm_flag_buffers(self):
        fb = getattr(self, "_pm_flags", None)
        if fb is None:
            dev, E = self.progress_buf.device, self.num_envs
            fb = self._pm_flags = dict(reset_buf=torch.zeros(E, device=dev, dtype=torch.bool), reset_succ=torch.zeros(E, device=dev, dtype=torch.bool),
                                       counts=torch.zeros(4, device=dev, dtype=torch.int32), succ_rate=torch.zeros(1, device=dev))
        return fb

    def pre_physics_step(self, actions):
        """hand_base.py:363-385."""
        out = self._pm_flag_buffers()
        self.pos_act = self.robot.control(actions)
        if self.train_test_flag not in ('train', 'test'):
            raise NotImplementedError
        train = self.train_test_flag == 'train'
        if self.success.dtype != torch.bool:
            self.success = self.success.bool()
        ops.episode_flags(train, self.rew_buf, self.progress_buf, self.success, self.epis_max_rew, self.epis_max_step, self.explore_step,
                          self.max_episode_length, out["reset_buf"], out["reset_succ"], out["counts"], out["succ_rate"])
        self.reset_buf = out["reset_buf"]
        if train:
            self.reset_succ = out["reset_succ"]
            self.extras['succ_rate'] = out["succ_rate"]
        if int(out["counts"][1]) > 0:                       # the reference's `if self.reset_buf.sum() > 0` (same one host sync)
            self.reset_idx(self.reset_buf)
        else:
            ops.scatter_dof_targets(self.pos_act, self._pm_dof_mask(), self.robot.num_dofs, self.pos_act_all)
            self._pm_set_targets()

    def _pm_set_targets(self):
        from isaacgym import gymtorch                      # the simulator binding stays the reference's (hand_base.py:383)
        self.gym.set_dof_position_target_tensor(self.sim, gymtorch.unwrap_tensor(self.pos_act_all))


class OpenDrawerKernels(BaseTaskKernels):
    """tasks/open_drawer.py: compute_observations (:240-281), compute_reward (:170-238); tasks/hand_base.py: pre_physics_step
    (:363-385) and post_physics_step (:387-392)."""

    # ------------------------------------------------------------------ buffers
    def _pm_buffers(self):
        out = getattr(self, "_pm_out", None)
        if out is not None:
            return out
        dev = self.dof_state_tensor_all.device
        E, nd, nb = self.num_envs, self.robot.num_dofs, self.rigid_body_mask.shape[1] - 2

        def f(*shape):
            return torch.zeros(*shape, device=dev, dtype=torch.float32)
        out = dict(obs=f(E, 29 + 2 * nd), part_bbox=f(E, 8, 3), dof_state_tensor=f(E, nd + 1, 2), rigid_body_tensor=f(E, nb + 2, 13),
                   tip_rb_tensor=f(E, 13), tip_rot_9d=f(E, 3, 3), gripper_length=f(E), dof_qpos_normalized=f(E, nd), rew_buf=f(E),
                   success=torch.zeros(E, device=dev, dtype=torch.bool), extras_f=f(6, E), extras_b=torch.zeros(3, E, device=dev, dtype=torch.bool))
        self._pm_out = out
        self._pm_const = dict(
            dof_mask=self.dof_state_mask.to(dev, torch.int64).contiguous(), rb_mask=self.rigid_body_mask.to(dev, torch.int64).contiguous(),
            bbox=self.part_bbox_init.to(dev, torch.float32).contiguous(), axis=self.part_axis_dir_init.to(dev, torch.float32).contiguous(),
            jl=self.part_joint_lower_limits.to(dev, torch.float32).contiguous(), ju=self.part_joint_upper_limits.to(dev, torch.float32).contiguous(),
            lstid=self.obj_lstid_lst.to(dev, torch.int64).contiguous())
        # the robot's control() reads the joint positions of the SIMULATOR tensor in place
        self.robot._pm_dof_state_mask = self._pm_const["dof_mask"]
        self.robot._pm_dof_state_all = self.dof_state_tensor_all
        return out

    def _pm_dof_mask(self):
        self._pm_buffers()
        return self._pm_const["dof_mask"]

    def _pm_launch(self, do_obs, do_reward, advance):
        out = self._pm_buffers()
        plan = getattr(self, "_pm_plan", None)
        if plan is None or not plan.bound_to(dof_state_all=self.dof_state_tensor_all, rigid_body_all=self.rigid_body_tensor_all,
                                             root_tensor=self.root_tensor, progress_buf=self.progress_buf, succ_objid=self.succ_objid_lst):
            c, rob = self._pm_const, self.robot
            plan = self._pm_plan = ops.OpenDrawerPostPlan(
                self.dof_state_tensor_all, self.rigid_body_tensor_all, self.root_tensor, self.obj_actor, c["dof_mask"], c["rb_mask"], rob.ltip_rb_index,
                rob.rtip_rb_index, rob.dof_lower_limits_tensor, rob.dof_upper_limits_tensor, c["bbox"], c["axis"], c["jl"], c["ju"], c["lstid"],
                self.suc_prop, self.progress_buf, self.succ_objid_lst, out)
        plan(do_obs, do_reward, advance)
        return out

    def _pm_publish_obs(self, out):
        rob, nd = self.robot, self.robot.num_dofs
        self.dof_state_tensor, self.rigid_body_tensor = out["dof_state_tensor"], out["rigid_body_tensor"]
        self.obj_root_tensor = self.root_tensor[:, self.obj_actor, :]
        rob.ltip_rb_tensor = self.rigid_body_tensor[:, rob.ltip_rb_index, :]          # load_robot.py:153-164
        rob.rtip_rb_tensor = self.rigid_body_tensor[:, rob.rtip_rb_index, :]
        rob.tip_rb_tensor, rob.tip_pos, rob.tip_rot_9d = out["tip_rb_tensor"], out["tip_rb_tensor"][:, :3], out["tip_rot_9d"]
        rob.gripper_length, rob.dof_qpos_normalized = out["gripper_length"], out["dof_qpos_normalized"]
        rob.dof_qpos_raw, rob.dof_qvel_raw = self.dof_state_tensor[:, :nd, 0], self.dof_state_tensor[:, :nd, 1]
        self.part_bbox = out["part_bbox"]
        self.obs_buf['normal_state'] = out["obs"]

    def _pm_publish_reward(self, out):
        self.rew_buf, self.success = out["rew_buf"], out["success"]
        ef, eb, ex = out["extras_f"], out["extras_b"], self.extras
        ex['is_open'], ex['is_open_notgrasp'], ex["is_reached"] = eb[0], eb[1], eb[2]
        ex['reaching_reward'], ex["close_reward"], ex["rot_reward"], ex["joint_state_reward"] = ef[0], ef[1], ef[2], ef[3]
        ex["raw_reward"], ex["is_grasped"], ex["success_objnum"], ex["step_id"] = self.rew_buf, ef[4], self.succ_objid_lst, ef[5]

    # ------------------------------------------------------------------ the reference's methods
    def compute_observations(self, type="step"):
        self._pm_publish_obs(self._pm_launch(True, False, False))

    def compute_reward(self, action):
        self._pm_publish_reward(self._pm_launch(False, True, False))

    def post_physics_step(self, actions):
        """hand_base.py:387-392: progress_buf += 1, refresh, observations and reward — one launch after the refresh."""
        self.refresh_gym_tensor()
        out = self._pm_launch(True, True, True)
        self._pm_publish_obs(out)
        self._pm_publish_reward(out)


class GraspCubeKernels(BaseTaskKernels):
    """tasks/grasp_cube.py: compute_observations (:118-138), compute_reward (:66-115); post_physics_step of tasks/hand_base.py."""

    def _pm_buffers(self):
        out = getattr(self, "_pm_out", None)
        if out is not None:
            return out
        dev = self.dof_state_tensor.device
        E, nd = self.num_envs, self.robot.num_dofs

        def f(*shape):
            return torch.zeros(*shape, device=dev, dtype=torch.float32)
        out = self._pm_out = dict(obs=f(E, 19 + 2 * nd), proprio=f(E, 7 + 2 * nd), tip_rb_tensor=f(E, 13), tip_rot_9d=f(E, 3, 3), gripper_length=f(E),
                                  dof_qpos_normalized=f(E, nd), rew_buf=f(E), success=torch.zeros(E, device=dev, dtype=torch.bool), extras_f=f(7, E),
                                  extras_b=torch.zeros(2, E, device=dev, dtype=torch.bool))
        self.robot._pm_dof_state_mask = None            # control() reads franka.dof_qpos_raw (a strided view of the simulator tensor)
        return out

    def _pm_dof_mask(self):
        return self.dof_state_mask

    def _pm_launch(self, do_obs, do_reward, advance):
        out = self._pm_buffers()
        plan = getattr(self, "_pm_plan", None)
        if plan is None or not plan.bound_to(dof_state=self.dof_state_tensor, rigid_body=self.rigid_body_tensor, root_tensor=self.root_tensor,
                                             progress_buf=self.progress_buf):
            rob = self.robot
            plan = self._pm_plan = ops.GraspCubePostPlan(
                self.dof_state_tensor, self.rigid_body_tensor, self.root_tensor, self.obj_actor, rob.num_dofs, rob.ltip_rb_index, rob.rtip_rb_index,
                rob.dof_lower_limits_tensor, rob.dof_upper_limits_tensor, self.pose_lower_limit.tolist(), self.pose_upper_limit.tolist(),
                self.success_pos.reshape(-1).tolist(), self.obj_default_root[:3].tolist(), self.goal_thresh, self.progress_buf, out)
        plan(do_obs, do_reward, advance)
        return out

    def _pm_publish_obs(self, out, type="step"):
        rob, nd = self.robot, self.robot.num_dofs
        self.obj_root_tensor = self.root_tensor[:, self.obj_actor, :]
        rob.ltip_rb_tensor = self.rigid_body_tensor[:, rob.ltip_rb_index, :]
        rob.rtip_rb_tensor = self.rigid_body_tensor[:, rob.rtip_rb_index, :]
        rob.tip_rb_tensor, rob.tip_pos, rob.tip_rot_9d = out["tip_rb_tensor"], out["tip_rb_tensor"][:, :3], out["tip_rot_9d"]
        rob.gripper_length, rob.dof_qpos_normalized = out["gripper_length"], out["dof_qpos_normalized"]
        rob.dof_qpos_raw, rob.dof_qvel_raw = self.dof_state_tensor[:, :nd, 0], self.dof_state_tensor[:, :nd, 1]
        self.obs_buf['normal_state'] = out["obs"]
        if self.learn_input_mode == 'mesh_tsdf':        # grasp_cube.py:128-131 (compute_scene_pose is the reference's own; it exits)
            rot, pos = self.compute_scene_pose()
            self.obs_buf['mesh_tsdf'] = self.mesh2TSDF.query_tsdf(rot, pos).reshape(self.num_envs, -1)
        if self.add_proprio_obs and type != 'init':     # grasp_cube.py:133-136: the vision observation gets the proprio columns appended
            self.obs_buf['proprio_state'] = out["proprio"]
            vis = self.obs_buf[self.learn_input_mode]
            D, P = vis.shape[1], out["proprio"].shape[1]
            cat = getattr(self, "_pm_cat", None)
            if cat is None or cat.shape != (self.num_envs, D + P):
                cat = self._pm_cat = torch.empty(self.num_envs, D + P, device=vis.device, dtype=torch.float32)
            ops.copy_rows(vis, cat[:, :D])
            ops.copy_rows(out["proprio"], cat[:, D:])
            self.obs_buf[self.learn_input_mode] = cat

    def _pm_publish_reward(self, out):
        self.rew_buf, self.success = out["rew_buf"], out["success"]
        ef, eb, ex = out["extras_f"], out["extras_b"], self.extras
        ex['reaching_reward'], ex["close_reward"], ex["rot_reward"], ex["reaching_goal_reward"] = ef[0], ef[1], ef[2], ef[3]
        ex["is_reached"], ex["obj_movement"], ex["raw_reward"], ex["obj_height"], ex["obj_up_flag"], ex["step_id"] = eb[0], ef[4], self.rew_buf, ef[5], eb[1], ef[6]

    def compute_observations(self, type="step"):
        self._pm_publish_obs(self._pm_launch(True, False, False), type)

    def compute_reward(self, action):
        self._pm_publish_reward(self._pm_launch(False, True, False))

    def post_physics_step(self, actions):
        self.refresh_gym_tensor()
        out = self._pm_launch(True, True, True)
        self._pm_publish_obs(out)
        self._pm_publish_reward(out)
