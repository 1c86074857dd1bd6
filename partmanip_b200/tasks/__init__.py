"""Env-side drop-ins: mixins that replace the torch arithmetic of the reference's task classes
(`tasks/open_drawer.py`, `tasks/load_robot.py`, `tasks/hand_base.py`) around the physics step with one CUDA launch per phase.
The simulator (Isaac Gym) stays the reference's; see INTEGRATION.md for the two-line binding."""
from .step_kernels import BaseTaskKernels, FrankaKernels, GraspCubeKernels, OpenDrawerKernels

__all__ = ["BaseTaskKernels", "FrankaKernels", "GraspCubeKernels", "OpenDrawerKernels"]
