"""ctypes binding of libmagical_b200.so (the C ABI in include/magical_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or does not
export the full ABI, loading fails loudly.
"""
import ctypes
import os

import numpy as np

from magical_b200 import scene as sc

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmagical_b200.so')

# every symbol include/magical_b200.h declares
ABI_SYMBOLS = [
    'mg_version', 'mg_last_error', 'mg_sizeof_scene', 'mg_sizeof_state',
    'mg_create', 'mg_destroy', 'mg_bind_obs', 'mg_obs_nbytes', 'mg_reset',
    'mg_step', 'mg_step_physics', 'mg_render', 'mg_score', 'mg_get_state',
    'mg_set_state', 'mg_set_pose', 'mg_launch_count', 'mg_synchronize', 'mg_update_scenes',
    'mg_set_draw_range', 'mg_overflow_count', 'mg_bind_obs_planes',
    'mg_newest_nbytes', 'mg_bind_newest', 'mg_stack_push', 'mg_step_render',
    'mg_comm_create', 'mg_comm_export', 'mg_comm_connect', 'mg_comm_scalar_ptr',
    'mg_comm_frame_ptr', 'mg_comm_barrier', 'mg_comm_gather_scalars',
    'mg_comm_stack_push', 'mg_comm_error', 'mg_comm_destroy',
    'mg_sizeof_placement', 'mg_set_placement', 'mg_get_env_scene',
    'mg_sampler_failures', 'mg_get_poses',
]
COMM_HANDLE_BYTES = 64

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """dlopen the library, bind signatures and verify the struct layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a). There is no CPU fallback for the hot path.")
    L = ctypes.CDLL(LIB_PATH)
    missing = [s for s in ABI_SYMBOLS if not hasattr(L, s)]
    if missing:
        raise NativeError(f"libmagical_b200.so lacks ABI symbols: {missing}")
    vp, i32, i64, f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64,
                         ctypes.c_double)
    L.mg_version.restype = ctypes.c_int
    L.mg_last_error.restype = ctypes.c_char_p
    L.mg_sizeof_scene.restype = i64
    L.mg_sizeof_state.restype = i64
    L.mg_create.argtypes = [vp, vp, vp, ctypes.POINTER(vp)]
    L.mg_destroy.argtypes = [vp]
    L.mg_bind_obs.argtypes = [vp, vp, i64]
    L.mg_obs_nbytes.restype = i64
    L.mg_obs_nbytes.argtypes = [vp]
    L.mg_reset.argtypes = [vp, vp, i32, vp]
    L.mg_step.argtypes = [vp, vp, vp, vp, vp]
    L.mg_step_physics.argtypes = [vp, vp, vp, vp, vp]
    L.mg_render.argtypes = [vp]
    L.mg_step_render.argtypes = [vp]
    L.mg_bind_obs_planes.argtypes = [vp, vp, i64, i64]
    L.mg_newest_nbytes.restype = i64
    L.mg_newest_nbytes.argtypes = [vp]
    L.mg_bind_newest.argtypes = [vp, vp, i64]
    L.mg_stack_push.argtypes = [vp, vp, vp, i64, i64, i32, i64, i32, vp]
    L.mg_sizeof_placement.restype = i64
    L.mg_set_placement.argtypes = [vp, i32, i32, vp]
    L.mg_get_env_scene.argtypes = [vp, i32, vp]
    L.mg_sampler_failures.argtypes = [vp, vp]
    L.mg_get_poses.argtypes = [vp, vp]
    L.mg_comm_create.argtypes = [i32, i32, i32, i64, i64, ctypes.POINTER(vp)]
    L.mg_comm_export.argtypes = [vp, vp]
    L.mg_comm_connect.argtypes = [vp, vp]
    L.mg_comm_scalar_ptr.restype = vp
    L.mg_comm_scalar_ptr.argtypes = [vp, i32]
    L.mg_comm_frame_ptr.restype = vp
    L.mg_comm_frame_ptr.argtypes = [vp, i32]
    L.mg_comm_barrier.argtypes = [vp, vp]
    L.mg_comm_gather_scalars.argtypes = [vp, i32, vp, vp]
    L.mg_comm_stack_push.argtypes = [vp, i32, i64, vp, vp, i64, i64, i64, i32, i32, vp]
    L.mg_comm_error.argtypes = [vp, vp]
    L.mg_comm_destroy.argtypes = [vp]
    L.mg_score.argtypes = [vp, vp]
    L.mg_get_state.argtypes = [vp, i32, vp]
    L.mg_set_state.argtypes = [vp, i32, vp]
    L.mg_set_pose.argtypes = [vp, i32, i32, f64, f64, f64]
    L.mg_launch_count.restype = i64
    L.mg_launch_count.argtypes = [vp]
    L.mg_synchronize.argtypes = [vp]
    L.mg_update_scenes.argtypes = [vp, i32, i32, vp]
    L.mg_set_draw_range.argtypes = [vp, i32, i32]
    L.mg_overflow_count.argtypes = [vp, vp]
    if L.mg_version() != sc.ABI_VERSION:
        raise NativeError("ABI version mismatch between Python and library")
    if L.mg_sizeof_placement() != sc.placement_dt.itemsize:
        raise NativeError(
            f"struct layout mismatch: mg_placement_t is {L.mg_sizeof_placement()} "
            f"bytes in C, {sc.placement_dt.itemsize} in numpy")
    if L.mg_sizeof_scene() != sc.scene_dt.itemsize \
            or L.mg_sizeof_state() != sc.state_dt.itemsize:
        raise NativeError(
            "struct layout mismatch: C has "
            f"{L.mg_sizeof_scene()}/{L.mg_sizeof_state()}, numpy has "
            f"{sc.scene_dt.itemsize}/{sc.state_dt.itemsize}")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load().mg_last_error().decode('utf-8', 'replace')
        raise NativeError(f"libmagical_b200 error {rc}: {msg}")


def make_config(**kwargs):
    cfg = np.zeros((), dtype=sc.config_dt)
    for k, v in kwargs.items():
        cfg[k] = v
    return cfg
