"""Loader of libptg_b200.so (the CUDA product).  There is NO CPU fallback: a missing or stale library raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("PTG_B200_SO") or os.path.join(_HERE, "csrc", "libptg_b200.so")   # override: kernel experiments
_LIB = None

# every symbol include/ptg_b200.h declares
EXPORTS = (
    "ptg_create", "ptg_destroy", "ptg_reset", "ptg_step", "ptg_step_many", "ptg_set_noise_tape", "ptg_get_state",
    "ptg_set_state", "ptg_episode_stats", "ptg_stats_combine", "ptg_poll_error", "ptg_obs_dim", "ptg_obs_layout",
    "ptg_obs_elems", "ptg_num_envs", "ptg_bytes_per_env_step", "ptg_kernel_launches", "ptg_host_standard_normal",
    "ptg_host_seed_state", "ptg_last_error",
    "ptg_abi_version", "ptg_vecnorm_moments", "ptg_vecnorm_apply", "ptg_features_dim", "ptg_features", "ptg_gae",
    "ptg_calculate_optimum", "ptg_allreduce_stats", "ptg_nccl_unique_id", "ptg_nccl_comm_create",
    "ptg_nccl_comm_destroy", "ptg_last_step_serial", "ptg_probe_box", "ptg_clock_uniform",
)


class PtgError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{_abi.STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


def load(build_if_missing: bool = False):
    """dlopen the in-tree library and declare the prototypes.  Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        if build_if_missing:
            from .csrc import build as _build
            _build.build()
        else:
            raise ImportError(f"{SO_PATH} is not built; run `python -m rl_ptg_b200.csrc.build` "
                              "(or __graft_entry__.build()).  There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise ImportError(f"{SO_PATH} does not export {name}; rebuild it")
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.ptg_create.argtypes = [C.POINTER(_abi.PtgConfig), C.POINTER(_abi.PtgTables), i64, i64, i64, i32, C.POINTER(vp)]
    L.ptg_destroy.argtypes = [vp]
    L.ptg_destroy.restype = None
    L.ptg_reset.argtypes = [vp, vp, vp, C.POINTER(_abi.PtgIO), vp]
    L.ptg_step.argtypes = [vp, vp, i32, C.POINTER(_abi.PtgIO), vp]
    L.ptg_step_many.argtypes = [vp, vp, i32, C.c_int32, C.POINTER(_abi.PtgIO), vp]
    L.ptg_set_noise_tape.argtypes = [vp, vp, i64]
    L.ptg_get_state.argtypes = [vp, C.POINTER(_abi.PtgStateSoA)]
    L.ptg_set_state.argtypes = [vp, C.POINTER(_abi.PtgStateSoA)]
    L.ptg_episode_stats.argtypes = [vp, vp, i32, vp]
    L.ptg_stats_combine.argtypes = [C.POINTER(_abi.PtgEpisodeStats), i32, C.POINTER(_abi.PtgEpisodeStats)]
    L.ptg_stats_combine.restype = None
    L.ptg_poll_error.argtypes = [vp, vp]
    L.ptg_allreduce_stats.argtypes = [vp, vp, vp, vp]
    L.ptg_nccl_unique_id.argtypes = [C.c_char_p]
    L.ptg_nccl_comm_create.argtypes = [C.c_char_p, i32, i32, C.POINTER(vp)]
    L.ptg_nccl_comm_destroy.argtypes = [vp]
    L.ptg_last_step_serial.argtypes = [vp]
    L.ptg_last_step_serial.restype = C.c_uint32
    L.ptg_probe_box.argtypes = [i32, C.POINTER(C.c_double * 4)]
    L.ptg_clock_uniform.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    f64 = C.c_double
    L.ptg_vecnorm_moments.argtypes = [vp, vp, vp, f64, vp, vp, vp]
    L.ptg_vecnorm_apply.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, f64, f64, vp, vp]
    L.ptg_features_dim.argtypes = [vp]
    L.ptg_features.argtypes = [vp, vp, vp, vp]
    L.ptg_gae.argtypes = [i64, C.c_int32, vp, vp, vp, vp, vp, f64, f64, vp, vp, vp]
    L.ptg_calculate_optimum.argtypes = [vp, i64, vp, vp, i64, vp, vp, vp]
    L.ptg_obs_dim.argtypes = [vp]
    L.ptg_obs_layout.argtypes = [vp, C.POINTER(_abi.PtgObsKey), i32]
    L.ptg_obs_elems.argtypes = [vp]
    L.ptg_obs_elems.restype = i64
    L.ptg_num_envs.argtypes = [vp]
    L.ptg_num_envs.restype = i64
    L.ptg_bytes_per_env_step.argtypes = [vp, i32]
    L.ptg_bytes_per_env_step.restype = i64
    L.ptg_kernel_launches.argtypes = [vp, C.POINTER(i64)]
    L.ptg_host_standard_normal.argtypes = [C.c_uint64, i64, vp]
    L.ptg_host_seed_state.argtypes = [C.c_uint64, vp]
    L.ptg_last_error.restype = C.c_char_p
    if L.ptg_abi_version() != _abi.PTG_ABI_VERSION:
        raise ImportError("libptg_b200.so ABI version mismatch; rebuild it")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise PtgError(rc, load().ptg_last_error().decode())
