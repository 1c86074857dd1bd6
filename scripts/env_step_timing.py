"""Dev tool: the env-side kernels (env_step.cu) at 4096 envs beside the same arithmetic as eager torch ops on the same GPU
(the oracle's restatement of tasks/open_drawer.py:170-281, load_robot.py:96-164, hand_base.py:367-377 — what the reference
executes every env step), CUDA-event timed; also reports how many outputs are bit-identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import env_oracle as EO
from partmanip_b200 import ops
from tests.helpers_env import synth_state
from tests.test_gpu_env_step import make_task, ROOT

dev = "cuda:0"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
s = synth_state(E, 5)
sg = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in s.items()}
t = make_task(s)
t.robot.check_jacobian = False
acts = torch.rand(E, 10, device=dev) * 2 - 1
t.post_physics_step(None)


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def torch_post():
    o = EO.compute_observations(sg["dof_all"], sg["rb_all"], sg["root"], sg["dof_mask"], sg["rb_mask"], 1, sg["part_bbox_init"], sg["part_axis_dir_init"],
                                s["num_dofs"], s["ltip"], s["rtip"], sg["dof_lower"], sg["dof_upper"])
    r = EO.compute_reward(o["part_bbox"], o["robot"], o["dof_state_tensor"], sg["part_joint_lower_limits"], sg["part_joint_upper_limits"], 0.5,
                          sg["obj_lstid"], torch.zeros(s["num_objs"], dtype=torch.bool, device=dev))
    return o, r


_z = torch.zeros
_orig_zeros = torch.zeros
torch.zeros = lambda *a, **k: _orig_zeros(*a, **{**k, "device": k.get("device", dev)})    # the oracle allocates its basis vectors with torch.zeros(...)
torch.eye_ = torch.eye
torch.eye = lambda *a, **k: torch.eye_(*a, **{**k, "device": k.get("device", dev)})
root = torch.tensor(ROOT, device=dev)
o, r = torch_post()
q = o["robot"]["dof_qpos_raw"]


def torch_control():
    return EO.control(acts, "ik", True, q, 1 / 60, root, sg["dof_lower"], sg["dof_upper"], sg["jac"], s["ltip"], s["rtip"])


def torch_flags():
    return EO.episode_flags("train", r["rew_buf"], t.progress_buf, r["success"].bool(), t.epis_max_rew, t.epis_max_step, 40, 200)


def graph_us(fn, n=20):
    """pure GPU time: n launches captured in one CUDA graph, replayed"""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n):
                fn()
    torch.cuda.current_stream().wait_stream(st)
    return timeit(g.replay, 20) / n


n0 = ops.launch_count()
t.post_physics_step(None)
k_post = ops.launch_count() - n0
print(f"E = {E}")
print(f"post_physics_step (obs + reward): kernel {timeit(lambda: t._pm_launch(True, True, False)):7.1f} us ({k_post} launch; {graph_us(lambda: t._pm_launch(True, True, False)):.1f} us on the GPU)   eager torch {timeit(torch_post):8.1f} us")
print(f"franka.control (ik, mobile)     : kernel {timeit(lambda: t.robot.control(acts)):7.1f} us (1 launch + memset; {graph_us(lambda: t.robot.control(acts)):.1f} us on the GPU)   eager torch {timeit(torch_control):8.1f} us")
out = t._pm_buffers()
print(f"episode flags (train)           : kernel {timeit(lambda: ops.episode_flags(True, t.rew_buf, t.progress_buf, t.success, t.epis_max_rew, t.epis_max_step, 40, 200, out['reset_buf'], out['reset_succ'], out['counts'], out['succ_rate'])):7.1f} us (1 launch + memset)   eager torch {timeit(torch_flags):8.1f} us")
# bit-identity against the same expressions evaluated by torch on the GPU
same = lambda a, b: float((a == b).float().mean())
print("bit-identical fraction vs eager torch on the GPU: obs %.4f  part_bbox %.4f  rew %.4f  rot_reward %.4f  is_reached %.4f  success %.4f" % (
    same(t.obs_buf["normal_state"], o["obs"]), same(t.part_bbox, o["part_bbox"]), same(t.rew_buf, r["rew_buf"]), same(t.extras["rot_reward"], r["rot_reward"]),
    same(t.extras["is_reached"], r["is_reached"]), same(t.success, r["success"].bool())))
print("max |obs err| %.2e   max |rew err| %.2e   control max err vs eager torch %.2e" % (
    float((t.obs_buf["normal_state"] - o["obs"]).abs().max()), float((t.rew_buf - r["rew_buf"]).abs().max()),
    float((t.robot.control(acts) - torch_control()).abs().max())))
bytes_rw = E * 4 * ((13 * 2 + 16 * 13 + 13 + 24 + 3 + 2) + (53 + 24 + 26 + 16 * 13 + 13 + 9 + 1 + 12 + 1 + 6)) + E * (16 + 13 + 1) * 8
print(f"algorithmic bytes per post-physics launch ~ {bytes_rw / 1e6:.2f} MB")
