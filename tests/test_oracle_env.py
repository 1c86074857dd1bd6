"""Pins oracle/env_oracle.py against the recordings of the unmodified reference methods (tests/golden/env_open_drawer.npz,
made by tests/golden/make_golden_env.py): tasks/open_drawer.py:170-281, tasks/load_robot.py:96-164, tasks/hand_base.py:367-377."""
import torch

from oracle import env_oracle as EO
from tests.helpers import load_golden


def _g():
    return load_golden("env_open_drawer.npz")


def _obs(g):
    return EO.compute_observations(g["in_dof_all"], g["in_rb_all"], g["in_root"], g["in_dof_mask"], g["in_rb_mask"], 1, g["in_part_bbox_init"],
                                   g["in_part_axis_dir_init"], int(g["in_num_dofs"]), int(g["in_ltip"]), int(g["in_rtip"]), g["in_dof_lower"],
                                   g["in_dof_upper"])


def test_observations_bit_exact():
    g = _g()
    o = _obs(g)
    assert torch.equal(o["obs"], g["obs"]) and o["obs"].shape[1] == 53
    assert torch.equal(o["part_bbox"], g["part_bbox"])
    assert torch.equal(o["dof_state_tensor"], g["dof_state_tensor"]) and torch.equal(o["rigid_body_tensor"], g["rigid_body_tensor"])
    for k in ("tip_rb_tensor", "tip_rot_9d", "gripper_length", "dof_qpos_normalized", "dof_qpos_raw", "dof_qvel_raw"):
        assert torch.equal(o["robot"][k], g["robot_" + k]), k


def test_reward_bit_exact():
    g = _g()
    o = _obs(g)
    r = EO.compute_reward(o["part_bbox"], o["robot"], o["dof_state_tensor"], g["in_part_joint_lower_limits"], g["in_part_joint_upper_limits"], 0.5,
                          g["in_obj_lstid"], torch.zeros(int(g["in_num_objs"]), dtype=torch.bool))
    assert torch.equal(r["rew_buf"], g["rew_buf"])
    assert torch.equal(r["success"].bool(), g["success"].bool()) and torch.equal(r["succ_objid_lst"], g["succ_objid_lst"])
    for k in ("is_open", "is_open_notgrasp", "reaching_reward", "close_reward", "rot_reward", "is_reached", "joint_state_reward", "raw_reward", "is_grasped"):
        assert torch.equal(r[k].float(), g["extras_" + k].float()), k
    assert 0.05 < float(r["success"].float().mean()) < 0.5 and 0.3 < float(r["is_reached"].float().mean()) < 0.9   # a real mix of cases


def test_control_bit_exact():
    g = _g()
    o = _obs(g)
    nd, lt, rt = int(g["in_num_dofs"]), int(g["in_ltip"]), int(g["in_rtip"])
    root = torch.tensor([0.5, 0.0, 0.05, 0.0, 0.0, 1.0, 0.0])
    q = o["robot"]["dof_qpos_raw"]
    a = EO.control(g["actions"], "ik", True, q, 1 / 60, root, g["in_dof_lower"], g["in_dof_upper"], g["in_jac"], lt, rt)
    assert torch.equal(a, g["action_tensor_ik_mobile"])
    a = EO.control(g["actions_pos_1"], "pos", True, q, 1 / 60, root, g["in_dof_lower"], g["in_dof_upper"])
    assert torch.equal(a, g["action_tensor_pos_1"])
    a = EO.control(g["actions_pos_0"], "pos", False, q[:, 3:], 1 / 60, root, g["in_dof_lower"][3:], g["in_dof_upper"][3:])
    assert torch.equal(a, g["action_tensor_pos_0"])
    a = EO.control(g["actions_ik_fixed"], "ik", False, q[:, 3:], 1 / 60, root, g["in_dof_lower"][3:], g["in_dof_upper"][3:],
                   g["in_jac"][..., 3:].contiguous(), lt, rt)
    assert torch.equal(a, g["action_tensor_ik_fixed"])


def test_episode_flags_bit_exact():
    g = _g()
    f = EO.episode_flags("train", g["rew_buf"], g["pre_progress_buf"], g["success"].bool(), g["pre_epis_max_rew"], g["pre_epis_max_step"], 40, 200)
    assert torch.equal(f["epis_max_step"], g["train_epis_max_step"]) and torch.equal(f["epis_max_rew"], g["train_epis_max_rew"])
    assert torch.equal(f["reset_buf"], g["train_reset_buf"]) and torch.equal(f["reset_succ"], g["train_reset_succ"])
    assert torch.equal(f["succ_rate"], g["train_succ_rate"])
    f = EO.episode_flags("test", g["rew_buf"], g["pre_progress_buf"], g["success"].bool(), f["epis_max_rew"], f["epis_max_step"], 40, 60)
    assert torch.equal(f["reset_buf"], g["test_reset_buf"])


def test_grasp_cube_bit_exact():
    """tasks/grasp_cube.py:66-138 (incl. deambiguity_rotation, utils/torch_jit_utils.py:412-425) against the recordings."""
    g = load_golden("env_grasp_cube.npz")
    lo = torch.tensor([-0.15, -0.15, 0.0, -1, -1, -1, -1], dtype=torch.float)
    hi = torch.tensor([0.15, 0.15, 0.4, 1, 1, 1, 1], dtype=torch.float)
    o = EO.cube_observations(g["in_dof"], g["in_rb"], g["in_root"], 1, int(g["in_num_dofs"]), int(g["in_ltip"]), int(g["in_rtip"]), g["in_dof_lower"],
                             g["in_dof_upper"], lo, hi)
    assert torch.equal(o["obs"], g["obs"]) and o["obs"].shape[1] == 37 and torch.equal(o["proprio"], g["pp_proprio_state"])
    assert torch.equal(torch.cat((g["pp_vision_obs"], o["proprio"]), dim=-1), g["pp_vision_cat"])
    r = EO.cube_reward(o["robot"], o["obj_root"], torch.tensor([0, 0, 0.2])[None, :], 0.025, torch.tensor([0.0, 0.0, 0.025]))
    assert torch.equal(r["rew_buf"], g["rew_buf"]) and torch.equal(r["success"].bool(), g["success"].bool())
    for k in ("reaching_reward", "close_reward", "rot_reward", "is_reached", "reaching_goal_reward", "obj_movement", "raw_reward", "obj_height", "obj_up_flag"):
        assert torch.equal(r[k].float(), g["extras_" + k].float()), k
    assert 0.05 < float(r["success"].float().mean()) < 0.6 and 0.2 < float(r["is_reached"].float().mean()) < 0.9
    a = EO.control(g["actions"], "ik", False, o["robot"]["dof_qpos_raw"], 1 / 60, None, g["in_dof_lower"], g["in_dof_upper"], g["in_jac"],
                   int(g["in_ltip"]), int(g["in_rtip"]))
    assert torch.equal(a, g["action_tensor_ik_fixed"])


def test_deambiguity_rotation_properties():
    """Size-independent properties of utils/torch_jit_utils.py:412-425 as restated: the result is one of the 24 candidate frames built
    from R's columns, it is orthonormal with det +1 for a unit quaternion, and no other candidate is closer to the identity."""
    g = torch.Generator().manual_seed(3)
    q = torch.randn(257, 4, generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q[0] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    out = EO.deambiguity_rotation(q)
    eye = torch.eye(3).expand_as(out)
    assert float((out.transpose(-1, -2) @ out - eye).abs().max()) < 1e-5
    assert float((torch.linalg.det(out) - 1).abs().max()) < 1e-5
    R = EO.quat_to_mat(q)
    tr = out.diagonal(dim1=-2, dim2=-1).sum(-1)
    # every signed column permutation of R that is a proper rotation obtained by the reference's construction has trace <= the winner's
    ind = [(0, 1), (0, 2), (1, 2), (1, 0), (2, 0), (2, 1)]
    best = torch.full((q.shape[0],), -10.0)
    for k in range(24):
        a, b = R[:, :, ind[k % 6][0]].clone(), R[:, :, ind[k % 6][1]].clone()
        if k < 12:
            a[:, 0], b[:, 0] = -a[:, 0], -b[:, 0]
        if 6 <= k < 18:
            a[:, 1], b[:, 1] = -a[:, 1], -b[:, 1]
        c = torch.cross(a, b, dim=-1)
        best = torch.maximum(best, a[:, 0] + b[:, 1] + c[:, 2])
    assert float((tr - best).abs().max()) < 1e-5
    assert torch.allclose(out[0], torch.eye(3), atol=1e-6)                  # the identity stays the identity


def test_episode_flags_properties():
    """hand_base.py:367-377: an env resets iff it stalled for explore_step steps since its best reward or succeeded; the best-reward
    step only moves when the reward does not fall below the record."""
    g = torch.Generator().manual_seed(4)
    E = 1000
    rew, best = torch.randn(E, generator=g), torch.randn(E, generator=g)
    prog, step = torch.randint(0, 200, (E,), generator=g), torch.randint(0, 200, (E,), generator=g)
    succ = torch.rand(E, generator=g) < 0.1
    f = EO.episode_flags("train", rew, prog, succ, best, step, 40, 200)
    improved = ~(rew < best)
    assert torch.equal(f["epis_max_step"][improved], prog[improved]) and torch.equal(f["epis_max_step"][~improved], step[~improved])
    assert torch.equal(f["epis_max_rew"], torch.maximum(rew, best))
    assert bool(f["reset_buf"][succ].all()) and bool((f["reset_buf"] == ((prog >= f["epis_max_step"] + 40) | succ)).all())
    assert abs(float(f["succ_rate"]) - float(succ.sum()) / max(1, int(f["reset_buf"].sum()))) < 1e-7
    assert torch.equal(EO.episode_flags("test", rew, prog, succ, best, step, 40, 150)["reset_buf"], prog >= 150)
