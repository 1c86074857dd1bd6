"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch, fp32) of the reference's env-side arithmetic for open_drawer
(SURVEY §8(f) rank 4).  The product (partmanip_b200/) never imports this file.

Pinned: tests/test_oracle_env.py checks every function here against tests/golden/env_open_drawer.npz, which
tests/golden/make_golden_env.py recorded by executing the UNMODIFIED reference methods on CPU.

Third-party arithmetic: `quat_rotate` / `tensor_clamp` come from isaacgym.torch_utils (Isaac Gym Preview 4, an NVIDIA binary
distribution that the reference does not pin or vendor); their published formulas are restated in quat_rotate() below."""
import torch


def quat_to_mat(q):
    """utils/torch_jit_utils.py:375-403 (x, y, z, w; NOT normalised first: two_s = 2 / |q|^2)."""
    i, j, k, r = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def quat_rotate(q, v):
    """isaacgym.torch_utils.quat_rotate: v (2 w^2 - 1) + 2 w (q_v x v) + 2 q_v (q_v . v); q need not be unit."""
    w, qv = q[:, 3:], q[:, :3]
    return v * (2.0 * w ** 2 - 1.0) + torch.cross(qv, v, dim=-1) * w * 2.0 + qv * (qv * v).sum(-1, keepdim=True) * 2.0


def quat_axis(q, axis):
    """utils/torch_jit_utils.py:65-69."""
    b = torch.zeros(q.shape[0], 3)
    b[:, axis] = 1
    return quat_rotate(q, b)


def update_state(rigid_body_tensor, dof_state_tensor, num_dofs, ltip, rtip, dof_lower, dof_upper):
    """tasks/load_robot.py:153-164."""
    lt, rt = rigid_body_tensor[:, ltip], rigid_body_tensor[:, rtip]
    tip = (lt + rt) / 2
    raw = dof_state_tensor[:, :num_dofs]
    return dict(ltip_rb_tensor=lt, rtip_rb_tensor=rt, tip_rb_tensor=tip, tip_rot_9d=quat_to_mat(tip[:, 3:7]),
                gripper_length=(lt[:, :3] - rt[:, :3]).norm(dim=-1),
                dof_qpos_normalized=2 * (raw[:, :, 0] - dof_lower) / (dof_upper - dof_lower) - 1,
                dof_qpos_raw=raw[:, :, 0], dof_qvel_raw=raw[:, :, 1])


def _handle(part_bbox):
    out, lng, sht = part_bbox[:, 0] - part_bbox[:, 4], part_bbox[:, 1] - part_bbox[:, 0], part_bbox[:, 3] - part_bbox[:, 0]
    mid = (part_bbox[:, 0] + part_bbox[:, 6]) / 2
    lo, ll, ls = out.norm(dim=-1), lng.norm(dim=-1), sht.norm(dim=-1)
    return out / lo[:, None], lng / ll[:, None], sht / ls[:, None], mid, lo, ll, ls


def compute_observations(dof_all, rb_all, root, dof_mask, rb_mask, obj_actor, part_bbox_init, part_axis_dir_init, num_dofs, ltip,
                         rtip, dof_lower, dof_upper):
    """tasks/open_drawer.py:240-281: masked gathers, robot state, articulated handle box in the world frame, 53-d state."""
    dof = dof_all[dof_mask]
    rb = rb_all[rb_mask]
    obj_root = root[:, obj_actor]
    rob = update_state(rb, dof, num_dofs, ltip, rtip, dof_lower, dof_upper)
    box_obj = part_bbox_init + dof[:, -1:, 0:1] * part_axis_dir_init.reshape(-1, 1, 3)
    part_bbox = torch.matmul(box_obj, quat_to_mat(obj_root[:, 3:7]).transpose(-1, -2)) + obj_root[:, None, :3]
    out, lng, sht, mid, lo, ll, ls = _handle(part_bbox)
    obs = torch.cat([rob["tip_rb_tensor"], mid, out, sht, lng, lo[:, None], ll[:, None], ls[:, None], rob["dof_qpos_normalized"],
                     rob["dof_qvel_raw"], dof[:, -1:, 0]], dim=-1)
    return dict(obs=obs, part_bbox=part_bbox, dof_state_tensor=dof, rigid_body_tensor=rb, robot=rob)


def compute_reward(part_bbox, rob, dof_state_tensor, part_joint_lower_limits, part_joint_upper_limits, suc_prop, obj_lstid, succ_objid_lst):
    """tasks/open_drawer.py:170-238."""
    out, lng, sht, mid, lo, ll, ls = _handle(part_bbox)
    tip = rob["tip_rb_tensor"]
    delta = tip[:, :3] - mid
    dist = delta.norm(dim=-1)
    r_out = (delta * out).sum(-1).abs() < lo / 2
    s_l = ((rob["ltip_rb_tensor"][:, :3] - mid) * sht).sum(-1)
    s_r = ((rob["rtip_rb_tensor"][:, :3] - mid) * sht).sum(-1)
    r_short = (s_l * s_r) < 0
    r_long = (delta * lng).sum(-1).abs() < ll / 2
    is_reached = r_out & r_short & r_long
    # bool + bool stays bool in torch: the bonus is 0.1 * (out OR short OR long), not 0.1 per satisfied axis (open_drawer.py:193)
    reaching = -dist + 0.1 * (r_out + r_short + r_long)
    q = tip[:, 3:7]
    grip, sep, down = quat_axis(q, 2), quat_axis(q, 1), quat_axis(q, 0)
    dot1 = (-grip * out).sum(-1)
    dot2 = torch.max((sep * sht).sum(-1), (-sep * sht).sum(-1))
    dot3 = torch.max((down * lng).sum(-1), (-down * lng).sum(-1))
    rot = dot1 + dot2 + dot3 - 3
    gl = rob["gripper_length"]
    close = (0.1 - gl) * is_reached + 0.1 * (gl - 0.1) * (~is_reached)
    grasp = is_reached & (gl < ls + 0.01) & (rot > -0.2)
    frac = (dof_state_tensor[:, -1, 0] - part_joint_lower_limits) / part_joint_upper_limits
    joint = grasp * (0.1 + torch.clamp(frac, max=suc_prop))
    is_open = grasp * (frac > 0.1)
    rew = reaching + 0.5 * rot + 5 * close + 5 * joint
    rew = rew + rew.abs() * rot
    success = grasp * ((dof_state_tensor[:, -1, 0] - part_joint_lower_limits) >= suc_prop * part_joint_upper_limits)
    succ_objid_lst = succ_objid_lst.clone()
    succ_objid_lst[obj_lstid[success == 1]] = True
    rew = rew + 2 * success
    return dict(rew_buf=rew, success=success, succ_objid_lst=succ_objid_lst, is_open=is_open, is_open_notgrasp=frac > 0.1,
                reaching_reward=reaching, close_reward=close, rot_reward=rot, is_reached=is_reached, joint_state_reward=joint,
                raw_reward=rew, is_grasped=grasp.float())


def solve_ik(jac, dpose, ltip, rtip, mobile, num_dofs, damping=0.05):
    """tasks/load_robot.py:142-151: damped least squares on the mean of the two finger-tip Jacobians."""
    lo = 3 if mobile else 0
    j = (jac[:, ltip - 1, :, lo:num_dofs - 2] + jac[:, rtip - 1, :, lo:num_dofs - 2]) / 2
    jt = j.transpose(1, 2)
    lam = torch.eye(6) * damping ** 2
    return (jt @ torch.inverse(j @ jt + lam) @ dpose).view(-1, num_dofs - 2 - lo), j.sum()


def control(raw_output, drive_mode, mobile, dof_qpos_raw, dt, default_root, dof_lower, dof_upper, jac=None, ltip=0, rtip=0):
    """tasks/load_robot.py:96-118 ('pos' and 'ik' drive modes, with / without the mobile base)."""
    num_dofs = dof_qpos_raw.shape[1]
    act = torch.zeros_like(dof_qpos_raw)
    lo = 3 if mobile else 0
    if mobile:
        dpose_base = raw_output[..., :3].unsqueeze(-1) * 0.005
        root_r = quat_to_mat(default_root[None, 3:7]).repeat(raw_output.shape[0], 1, 1)
        act[..., :3] = dof_qpos_raw[..., :3] + torch.bmm(root_r.transpose(-1, -2), dpose_base).squeeze(-1)
        raw_output = raw_output[..., 3:]
    if drive_mode == "pos":
        act[..., lo:-2] = dof_qpos_raw[..., lo:-2] + raw_output[:, :-1] * dt * 20
        act[..., -2:-1] = dof_qpos_raw[..., -2:-1] + raw_output[:, -1:] * dt
        act[..., -1:] = dof_qpos_raw[..., -1:] + raw_output[:, -1:] * dt
    elif drive_mode == "ik":
        dpose = torch.cat([raw_output[..., :3] * 0.005, raw_output[..., 3:6] * 0.005], -1).unsqueeze(-1)
        if mobile:
            dpose[:, :3] -= dpose_base
        act[:, lo:-2] = dof_qpos_raw[..., lo:-2] + solve_ik(jac, dpose, ltip, rtip, mobile, num_dofs)[0]
        act[:, -2:-1] = dof_qpos_raw[..., -2:-1] + raw_output[..., -1:] * dt / 5
        act[:, -1:] = dof_qpos_raw[..., -1:] + raw_output[..., -1:] * dt / 5
    else:
        raise NotImplementedError
    return torch.max(torch.min(act, dof_upper), dof_lower)


def episode_flags(train_test_flag, rew_buf, progress_buf, success, epis_max_rew, epis_max_step, explore_step, max_episode_length):
    """tasks/hand_base.py:367-377."""
    if train_test_flag == "train":
        epis_max_step = torch.where(rew_buf < epis_max_rew, epis_max_step, progress_buf)
        epis_max_rew = torch.maximum(rew_buf, epis_max_rew)
        reset_buf = (progress_buf >= epis_max_step + explore_step) | success
        succ_rate = success.int().sum(dim=-1, keepdim=True) / torch.clamp(reset_buf.int().sum(), min=1)
        return dict(epis_max_step=epis_max_step, epis_max_rew=epis_max_rew, reset_buf=reset_buf, reset_succ=success.clone(), succ_rate=succ_rate)
    if train_test_flag == "test":
        return dict(reset_buf=progress_buf >= max_episode_length)
    raise NotImplementedError


# ------------------------------------------------------------------------------------------------------------------ grasp_cube
def mat_diff_rad(m1, m2):
    """utils/torch_jit_utils.py:406-410."""
    d = torch.matmul(m1.transpose(-1, -2), m2)
    return torch.acos(torch.clamp((d[..., 0, 0] + d[..., 1, 1] + d[..., 2, 2] - 1) / 2, -1, 1))


def deambiguity_rotation(old_r):
    """utils/torch_jit_utils.py:412-425.  Note the sign flips index dim 2 of the (N, 24, 3, 2) stack: they negate ROW 0 (first 12
    candidates) and ROW 1 (candidates 6..17) of both selected columns."""
    R = quat_to_mat(old_r)
    ind = torch.tensor([[0, 1], [0, 2], [1, 2], [1, 0], [2, 0], [2, 1]])
    ind = torch.cat([ind, ind, ind, ind], dim=0)
    m12 = R[:, :, ind].transpose(-2, -3).clone()
    m12[:, :12, 0] = -m12[:, :12, 0]
    m12[:, 6:18, 1] = -m12[:, 6:18, 1]
    m3 = torch.cross(m12[..., 0], m12[..., 1], dim=-1).unsqueeze(-1)
    allm = torch.cat([m12, m3], dim=-1)
    rad = mat_diff_rad(allm, torch.eye(3)[None, None])
    return allm[torch.arange(allm.shape[0]), rad.argmin(dim=1)]


def cube_observations(dof, rb, root, obj_actor, num_dofs, ltip, rtip, dof_lower, dof_upper, pose_lo, pose_hi):
    """tasks/grasp_cube.py:118-126 (+ :132-135 proprio_state)."""
    rob = update_state(rb, dof, num_dofs, ltip, rtip, dof_lower, dof_upper)
    obj = root[:, obj_actor]
    tip_pose = 2 * (rob["tip_rb_tensor"][:, :7] - pose_lo) / (pose_hi - pose_lo) - 1
    obj_pos = 2 * (obj[:, :3] - pose_lo[:3]) / (pose_hi[:3] - pose_lo[:3]) - 1
    obj_pose = torch.cat([obj_pos, deambiguity_rotation(obj[:, 3:7]).reshape(obj.shape[0], -1)], dim=-1)
    obs = torch.cat([tip_pose, obj_pose, rob["dof_qpos_normalized"], rob["dof_qvel_raw"]], dim=-1)
    proprio = torch.cat([tip_pose, rob["dof_qpos_normalized"], rob["dof_qvel_raw"]], dim=-1)
    return dict(obs=obs, proprio=proprio, robot=rob, obj_root=obj)


def cube_reward(rob, obj_root, success_pos, goal_thresh, obj_default_pos):
    """tasks/grasp_cube.py:66-115."""
    pos = obj_root[:, :3]
    dist = (rob["tip_rb_tensor"][:, :3] - pos).norm(dim=-1)
    reached = dist < 0.02
    gl = rob["gripper_length"]
    close = (0.1 - gl) * reached + 0.1 * (gl - 0.1) * (~reached)
    orot = deambiguity_rotation(obj_root[:, 3:7])
    H = quat_to_mat(rob["tip_rb_tensor"][..., 3:7])
    down = -H[:, -1, -1]
    p1 = ((H[:, :, 0] * orot[:, :, 0]).abs() + (H[:, :, 1] * orot[:, :, 1]).abs()).sum(dim=-1)
    p2 = ((H[:, :, 0] * orot[:, :, 1]).abs() + (H[:, :, 1] * orot[:, :, 0]).abs()).sum(dim=-1)
    rot = down + torch.max(p1, p2) - 3
    gdist = (pos - success_pos).norm(dim=-1)
    goal = torch.max(0.2 - gdist, torch.zeros_like(gdist)) * reached
    rew = -dist + 0.5 * rot + 5 * close + 20 * goal
    success = (gdist <= goal_thresh) * reached
    rew = rew + 3 * success
    return dict(rew_buf=rew, success=success, reaching_reward=-dist, close_reward=close, rot_reward=rot, is_reached=reached,
                reaching_goal_reward=goal, obj_movement=(pos - obj_default_pos).norm(dim=-1), raw_reward=rew, obj_height=obj_root[..., 2],
                obj_up_flag=obj_root[..., 2] > 0.1)
