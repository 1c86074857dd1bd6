"""ctypes binding of libpartmanip_b200.so (C-ABI declared in include/partmanip_b200.h).

The product path has NO fallback: if the shared library is missing or a symbol is absent the import
fails loudly.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C partmanip_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpartmanip_b200.so")

P, I, L, F, U64, SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_size_t


class EncoderParams(C.Structure):
    """pm_encoder_params / pm_encoder_grads (six device pointers W1,b1,W2,b2,W3,b3)."""
    _fields_ = [(n, P) for n in ("W1", "b1", "W2", "b2", "W3", "b3")]


EP = C.POINTER(EncoderParams)

# name -> (restype, argtypes) ; mirrors include/partmanip_b200.h one to one
SIGNATURES = {
    "pm_last_error": (C.c_char_p, []),
    "pm_version": (I, []),
    "pm_has_tcgen05": (I, []),
    "pm_tc_sticky_error": (I, [I]),
    "pm_colreduce_ws_bytes": (SZ, [I, I]),
    "pm_rms_colsum": (I, [P, L, I, I, P, P, P]),
    "pm_rms_colsqdev": (I, [P, L, I, I, P, F, P, P, P]),
    "pm_rms_update": (I, [P, P, P, P, P, F, I, I, P]),
    "pm_rms_normalize": (I, [P, L, P, L, I, I, P, P, P]),
    "pm_rms_forward_ws_bytes": (SZ, [I, I]),
    "pm_rms_forward": (I, [P, L, P, L, I, I, P, P, P, I, I, P, P]),
    "pm_gae": (I, [P, P, P, P, P, P, P, I, I, F, F, I, F, P]),
    "pm_normalize_ws_bytes": (SZ, [L]),
    "pm_normalize_inplace": (I, [P, L, P, P]),
    "pm_normalize": (I, [P, P, L, P, P]),
    "pm_randn": (I, [P, L, U64, U64, P]),
    "pm_policy_sample": (I, [P, P, P, I, I, F, I, P, P, P, P]),
    "pm_action_activation": (I, [P, P, L, F, I, P]),
    "pm_policy_logprob": (I, [P, L, P, P, I, I, F, I, P, P, P]),
    "pm_ppo_actor_loss_ws_bytes": (SZ, [I, I]),
    "pm_ppo_actor_loss": (I, [P, L, P, P, P, P, P, P, P, I, I, F, F, F, I, P, P, L, P, P, P, P]),
    "pm_ppo_actor_finalize": (I, [P, F, F, P, P, P]),
    "pm_value_loss": (I, [P, L, P, P, P, I, F, P, P, L, P, P]),
    "pm_dagger_loss": (I, [P, L, P, I, I, F, I, F, P, P, L, P, P]),
    "pm_abs_sum": (I, [P, L, F, P, P, P]),
    "pm_accumulate": (I, [P, F, P, I, P]),
    "pm_linear_forward": (I, [P, L, P, P, P, L, I, I, I, I, P, P]),
    "pm_linear_backward_ws_bytes": (SZ, [I, I, I]),
    "pm_linear_backward": (I, [P, L, P, P, L, P, P, P, L, I, I, I, I, P, P, P]),
    "pm_linear_forward_tc": (I, [P, L, P, P, P, L, I, I, I, I, I, P, P]),
    "pm_linear_backward_tc_ws_bytes": (SZ, [I, I, I]),
    "pm_linear_backward_tc": (I, [P, L, P, P, L, P, P, P, L, I, I, I, I, I, P, P, P]),
    "pm_pointnet_head_forward": (I, [P, L, I, I, EP, I, I, I, P, P, P, L, P]),
    "pm_pointnet_head_backward_ws_bytes": (SZ, [I, I]),
    "pm_pointnet_head_backward": (I, [P, L, I, I, EP, I, I, I, P, P, P, L, EP, P, L, I, P, SZ, P]),
    "pm_pointnet_center": (I, [P, L, I, I, I, P]),
    "pm_pointnet_encode_forward": (I, [P, L, I, I, I, EP, I, I, P, P, L, P, P, P, SZ, P]),
    "pm_pointnet_encode_forward_ws_bytes": (SZ, [I, I, I, I]),
    "pm_pointnet_tc_last_error": (I, [P, P]),
    "pm_pointnet_tc3_last_error": (I, [P, P]),
    "pm_pointnet_encode_backward_ws_bytes": (SZ, [I, I, I, I, I]),
    "pm_pointnet_encode_backward": (I, [P, L, I, I, I, EP, I, I, P, P, L, P, P, EP, P, SZ, P]),
    "pm_pointnet_bwd_tc_last_error": (I, [P, P]),
    "pm_adam_ws_bytes": (SZ, [L]),
    "pm_adam_step": (I, [P, P, P, P, L, L, F, F, F, F, P, P, P, P]),
    "pm_fused_step_ws_bytes": (SZ, [L, I]),
    "pm_fused_step": (I, [P, P, P, L, L, I, F, F, F, F, P, P, P, P, P, I, I, I, F, F, P, P, P, P]),
    "pm_depth2pc_backproject": (I, [P, I, I, I, I, C.POINTER(F), P, C.POINTER(F), F, P, P]),
    "pm_depth2pc_backproject_views": (I, [P, I, I, I, I, I, I, F, C.POINTER(F), P, C.POINTER(F), F, P, P]),
    "pm_fps_ws_bytes": (SZ, [I, I]),
    "pm_fps_cluster_max_active": (I, []),
    "pm_tsdf_voxel_tables": (I, [P, I, C.POINTER(F), I, I, F, I, C.POINTER(F), P, P, P]),
    "pm_tsdf_integrate": (I, [P, I, I, I, I, P, P, F, I, F, P, P]),
    "pm_tsdf_sparse_voxel_ws_bytes": (SZ, [I, I, I]),
    "pm_tsdf_sparse_voxel": (I, [P, I, I, F, F, I, P, P, SZ, P]),
    "pm_farthest_point_sample": (I, [P, I, I, I, I, P, P, P, SZ, P]),
    "pm_conv3d_out_dim": (I, [I, I, I]),
    "pm_conv3d_im2col": (I, [P, L, L, I, I, I, I, I, P, I, P]),
    "pm_conv3d_col2im": (I, [P, I, I, I, I, I, I, P, I, P, P]),
    "pm_conv3d_flatten": (I, [P, P, I, I, I, L, I, P]),
    "pm_conv3d_first_forward": (I, [P, L, I, I, I, P, P, I, P, P]),
    "pm_conv3d_first_backward_ws_bytes": (SZ, []),
    "pm_conv3d_first_backward": (I, [P, L, I, I, I, P, P, P, P]),
    "pm_conv3d_weight_permute": (I, [P, I, I, I, I, P, P]),
    "pm_maxpool3d_forward": (I, [P, I, I, I, I, P, P, P]),
    "pm_maxpool3d_backward": (I, [P, P, P, I, I, I, I, I, P, P]),
    "pm_mesh2sdf_query": (I, [P, L, P, P, P, I, I, I, P, P, P, I, I, C.POINTER(F), F, P, P]),
    "pm_open_drawer_obs_dim": (I, [I]),
    "pm_open_drawer_post_physics": (I, [P, P, P, I, I, P, P, I, I, I, I, I, P, P, P, P, P, P, P, F, I, I, I, P,
                                        P, P, P, P, P, P, P, P, P, P, P, P, P, P]),
    "pm_grasp_cube_obs_dim": (I, [I]),
    "pm_grasp_cube_post_physics": (I, [P, I, P, I, P, I, I, I, I, I, I, P, P, C.POINTER(F), C.POINTER(F), C.POINTER(F), C.POINTER(F), F,
                                       I, I, I, P, P, P, P, P, P, P, P, P, P, P, P]),
    "pm_franka_control": (I, [P, I, I, I, I, P, L, L, P, I, P, I, I, I, P, P, C.POINTER(F), F, F, P, P, P]),
    "pm_episode_flags": (I, [I, I, P, P, P, P, P, L, L, P, P, P, P, P]),
    "pm_scatter_dof_targets": (I, [P, P, I, I, I, P, P]),
    "pm_gather_rows": (I, [P, L, P, P, L, L, I, P]),
    "pm_copy_rows": (I, [P, L, P, L, L, I, P]),
}

PM_ACT = {None: 0, "none": 0, "tanh": 1, "relu": 2, "crelu": 2, "elu": 3, "selu": 4, "lrelu": 5, "sigmoid": 6}
PM_PREC = {"fp32": 0, "bf16": 1, "fp32_ffma": 2}


class PMError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. There is no CPU/PyTorch fallback — "
            "run `python -c 'import __graft_entry__ as g; g.build()'` (or `make -C partmanip_b200/csrc`).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so is stale: loud by design
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


# kernels launched per C-ABI call (typical path) — feeds bench.py's gpu_launches claim
KERNELS_PER_CALL = {
    "pm_rms_colsum": 2, "pm_rms_colsqdev": 2, "pm_rms_update": 1, "pm_rms_normalize": 1, "pm_rms_forward": 6, "pm_gae": 1,
    "pm_normalize": 1, "pm_normalize_inplace": 1, "pm_randn": 1, "pm_policy_sample": 1, "pm_action_activation": 1,
    "pm_policy_logprob": 1, "pm_ppo_actor_loss": 2, "pm_ppo_actor_finalize": 1, "pm_value_loss": 2, "pm_abs_sum": 2,
    "pm_accumulate": 1, "pm_dagger_loss": 2, "pm_linear_forward": 1, "pm_linear_backward": 4, "pm_linear_forward_tc": 1, "pm_linear_backward_tc": 5, "pm_pointnet_center": 1,
    "pm_pointnet_encode_forward": 2, "pm_pointnet_encode_backward": 3, "pm_pointnet_head_forward": 1,
    "pm_pointnet_head_backward": 3, "pm_adam_step": 3, "pm_fused_step": 1, "pm_gather_rows": 1,
    "pm_copy_rows": 1, "pm_conv3d_first_backward": 2,
}
LAUNCHES = [0]


def check(rc: int, what: str = ""):
    LAUNCHES[0] += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        raise PMError(f"{what or 'libpartmanip_b200'} failed (rc={rc}): {lib.pm_last_error().decode()}")
