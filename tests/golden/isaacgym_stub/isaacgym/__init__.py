"""Golden-generation stub for Isaac Gym (Preview 4; NVIDIA binary distribution, absent from /root/reference and from this
image).  Only tests/golden/make_golden_env.py imports it, to let the UNMODIFIED reference task classes (tasks/open_drawer.py,
tasks/load_robot.py, tasks/hand_base.py) be imported and their arithmetic executed on CPU.  `gymapi` / `gymtorch` are empty
shells (no simulator); `torch_utils` restates the published quaternion helpers the reference's arithmetic calls."""
