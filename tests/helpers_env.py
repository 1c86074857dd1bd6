"""Synthetic simulator state for the env-side kernels (shared by tests/golden/make_golden_env.py and the GPU parity tests)."""
import torch


def rand_quat(g, n, spread=1.0):
    q = torch.randn(n, 4, generator=g) * spread + torch.tensor([0.0, 0.0, 0.0, 1.0]) * (1 - spread)
    return q / q.norm(dim=-1, keepdim=True)


def synth_state(E, seed, num_dofs=12, nb_robot=14, ltip=12, rtip=13, num_objs=5):
    """A simulator state with irregular per-env object sizes (so the gather masks are not a stride) and a mix of envs whose
    gripper is around the handle / away from it, drawers closed / half open / past the success threshold."""
    g = torch.Generator().manual_seed(seed)
    obj_ndof = torch.randint(1, 5, (num_objs,), generator=g)
    obj_nrb = obj_ndof + 1 + torch.randint(1, 3, (num_objs,), generator=g)
    obj_target_dof = torch.stack([torch.randint(0, int(n), (1,), generator=g)[0] for n in obj_ndof])
    obj_target_link = torch.stack([torch.randint(0, int(n) - 1, (1,), generator=g)[0] for n in obj_nrb])
    obj_target_handle = obj_target_link + 1
    obj_lstid = torch.randint(0, num_objs, (E,), generator=g)
    dof_mask = torch.zeros(E, num_dofs + 1, dtype=torch.long)
    rb_mask = torch.zeros(E, nb_robot + 2, dtype=torch.long)
    dc = rc = 0
    for i in range(E):                                         # tasks/open_drawer.py:62-70
        o = int(obj_lstid[i])
        dof_mask[i, :num_dofs] = torch.arange(dc, dc + num_dofs)
        dof_mask[i, -1] = dc + num_dofs + obj_target_dof[o]
        rb_mask[i, :nb_robot] = torch.arange(rc, rc + nb_robot)
        rb_mask[i, -2] = rc + nb_robot + obj_target_link[o]
        rb_mask[i, -1] = rc + nb_robot + obj_target_handle[o]
        dc += num_dofs + int(obj_ndof[o])
        rc += nb_robot + int(obj_nrb[o])
    # per-object handle boxes in the object frame (8 corners: 0 origin, 1 +long, 2 +long+short, 3 +short, 4..7 = 0..3 - out)
    boxes, axes, upper, lower = [], [], [], []
    for o in range(num_objs):
        R = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
        lo_, ll_, ls_ = 0.02 + 0.03 * torch.rand(3, generator=g) * torch.tensor([1.0, 4.0, 0.7])
        c = torch.tensor([-0.25, 0.0, 0.1]) + 0.1 * torch.randn(3, generator=g)
        b0 = c + (lo_ * R[:, 0] - ll_ * R[:, 1] - ls_ * R[:, 2]) / 2
        top = torch.stack([b0, b0 + ll_ * R[:, 1], b0 + ll_ * R[:, 1] + ls_ * R[:, 2], b0 + ls_ * R[:, 2]])
        boxes.append(torch.cat([top, top - lo_ * R[:, 0]]))
        axes.append(R[:, 0] + 0.05 * torch.randn(3, generator=g))
        upper.append(0.2 + 0.3 * torch.rand((), generator=g))
        lower.append(0.02 * torch.rand((), generator=g))
    part_bbox_init = torch.stack(boxes)[obj_lstid]
    part_axis_dir_init = torch.stack(axes)[obj_lstid]
    part_joint_upper_limits = torch.stack(upper)[obj_lstid] * 0.5           # * obj_scale, open_drawer.py:79
    part_joint_lower_limits = torch.stack(lower)[obj_lstid]

    dof_all = torch.randn(dc, 2, generator=g)
    rb_all = torch.randn(rc, 13, generator=g)
    root = torch.randn(E, 2, 13, generator=g)
    root[:, 1, :3] = torch.tensor([-0.6, 0.0, 0.5]) + 0.05 * torch.randn(E, 3, generator=g)
    root[:, 1, 3:7] = rand_quat(g, E, 0.3)
    root[:, 0, 3:7] = rand_quat(g, E, 0.2)
    # drawer joint: closed / partly open / past suc_prop
    frac = torch.rand(E, generator=g) * 1.2
    dof_all[dof_mask[:, -1], 0] = part_joint_lower_limits + frac * part_joint_upper_limits
    dof_lower = -2.5 + 0.5 * torch.rand(num_dofs, generator=g)
    dof_upper = 2.5 + 0.5 * torch.rand(num_dofs, generator=g)
    dof_all[dof_mask[:, :num_dofs].reshape(-1), 0] = (dof_lower + (dof_upper - dof_lower) * torch.rand(E, num_dofs, generator=g)).reshape(-1)
    # world-frame handle (same arithmetic as the env, only to PLACE the finger tips)
    i_, j_, k_, r_ = root[:, 1, 3:7].unbind(-1)
    two_s = 2.0 / (root[:, 1, 3:7] ** 2).sum(-1)
    Rm = torch.stack([1 - two_s * (j_ * j_ + k_ * k_), two_s * (i_ * j_ - k_ * r_), two_s * (i_ * k_ + j_ * r_),
                      two_s * (i_ * j_ + k_ * r_), 1 - two_s * (i_ * i_ + k_ * k_), two_s * (j_ * k_ - i_ * r_),
                      two_s * (i_ * k_ - j_ * r_), two_s * (j_ * k_ + i_ * r_), 1 - two_s * (i_ * i_ + j_ * j_)], -1).reshape(E, 3, 3)
    box_w = (part_bbox_init + dof_all[dof_mask[:, -1], 0][:, None, None] * part_axis_dir_init[:, None]) @ Rm.transpose(-1, -2) + root[:, 1, None, :3]
    mid = (box_w[:, 0] + box_w[:, 6]) / 2
    u_out, u_long, u_short = box_w[:, 0] - box_w[:, 4], box_w[:, 1] - box_w[:, 0], box_w[:, 3] - box_w[:, 0]
    l_short = u_short.norm(dim=-1, keepdim=True)
    u_short_n = u_short / l_short
    near = torch.rand(E, generator=g) < 0.6
    jitter = torch.where(near[:, None], 0.004 * torch.randn(E, 3, generator=g), 0.3 * torch.randn(E, 3, generator=g))
    half_open = torch.where(torch.rand(E, 1, generator=g) < 0.7, l_short * 0.5 + 0.003, l_short * 0.5 + 0.05)
    lt, rt = ltip, rtip
    rb_all[rb_mask[:, lt], :3] = mid + jitter + half_open * u_short_n
    rb_all[rb_mask[:, rt], :3] = mid + jitter - half_open * u_short_n
    # finger frames: for "near" envs align grip (z) with -out, separation (y) with short, down (x) with long, plus noise
    Rg = torch.stack([u_long / u_long.norm(dim=-1, keepdim=True), u_short_n, -u_out / u_out.norm(dim=-1, keepdim=True)], -1)
    Rg = torch.linalg.qr(Rg + 0.05 * torch.randn(E, 3, 3, generator=g))[0]
    Rg = Rg * torch.sign(torch.linalg.det(Rg))[:, None, None]
    w = torch.sqrt(torch.clamp(1 + Rg[:, 0, 0] + Rg[:, 1, 1] + Rg[:, 2, 2], min=1e-6)) / 2
    qg = torch.stack([(Rg[:, 2, 1] - Rg[:, 1, 2]) / (4 * w), (Rg[:, 0, 2] - Rg[:, 2, 0]) / (4 * w), (Rg[:, 1, 0] - Rg[:, 0, 1]) / (4 * w), w], -1)
    qg = qg / qg.norm(dim=-1, keepdim=True)
    q_far = rand_quat(g, E)
    q = torch.where(near[:, None], qg, q_far)
    rb_all[rb_mask[:, lt], 3:7] = q + 0.01 * torch.randn(E, 4, generator=g)
    rb_all[rb_mask[:, rt], 3:7] = q + 0.01 * torch.randn(E, 4, generator=g)
    jac = torch.randn(E, nb_robot - 1, 6, num_dofs, generator=g)
    return dict(E=E, num_dofs=num_dofs, nb_robot=nb_robot, ltip=ltip, rtip=rtip, num_objs=num_objs, obj_lstid=obj_lstid,
                dof_mask=dof_mask, rb_mask=rb_mask, dof_all=dof_all, rb_all=rb_all, root=root, part_bbox_init=part_bbox_init,
                part_axis_dir_init=part_axis_dir_init, part_joint_upper_limits=part_joint_upper_limits,
                part_joint_lower_limits=part_joint_lower_limits, dof_lower=dof_lower, dof_upper=dof_upper, jac=jac)


def synth_state_cube(E, seed, num_dofs=9, nb=13, ltip=10, rtip=11):
    """grasp_cube simulator state (regular per-env tensors, tasks/grasp_cube.py:25-33): a mix of grippers on / off the cube and
    cubes at / away from the goal; one env sits at the identity rotation (a 24-way tie of deambiguity_rotation is impossible,
    but several candidates coincide in angle there)."""
    g = torch.Generator().manual_seed(seed)
    dof = torch.randn(E, num_dofs, 2, generator=g)
    rb = torch.randn(E, nb, 13, generator=g)
    root = torch.randn(E, 2, 13, generator=g)
    dof_lower = -2.5 + 0.5 * torch.rand(num_dofs, generator=g)
    dof_upper = 2.5 + 0.5 * torch.rand(num_dofs, generator=g)
    at_goal = torch.rand(E, generator=g) < 0.35
    pos = torch.where(at_goal[:, None], torch.tensor([0.0, 0.0, 0.2]) + 0.015 * torch.randn(E, 3, generator=g),
                      torch.tensor([0.0, 0.0, 0.06]) + torch.tensor([0.15, 0.15, 0.08]) * (torch.rand(E, 3, generator=g) * 2 - 1))
    root[:, 1, :3] = pos
    root[:, 1, 3:7] = rand_quat(g, E, 1.0)
    root[0, 1, 3:7] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    near = torch.rand(E, generator=g) < 0.6
    off = torch.where(near[:, None], 0.008 * torch.randn(E, 3, generator=g), 0.2 * torch.randn(E, 3, generator=g))
    half = torch.tensor([0.0, 1.0, 0.0]) * (0.02 + 0.03 * torch.rand(E, 1, generator=g))
    rb[:, ltip, :3] = pos + off + half
    rb[:, rtip, :3] = pos + off - half
    q = rand_quat(g, E, 0.5)
    rb[:, ltip, 3:7] = q + 0.01 * torch.randn(E, 4, generator=g)
    rb[:, rtip, 3:7] = q + 0.01 * torch.randn(E, 4, generator=g)
    jac = torch.randn(E, nb - 2, 6, num_dofs, generator=g)
    return dict(E=E, num_dofs=num_dofs, nb=nb, ltip=ltip, rtip=rtip, dof=dof, rb=rb, root=root, dof_lower=dof_lower, dof_upper=dof_upper, jac=jac)
