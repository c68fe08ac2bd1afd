"""ctypes wrapper of ``oracle/ptg_oracle.c`` -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module.  The product package ``rl_ptg_b200`` never does (tests/test_no_oracle_in_product.py).

``OracleVecEnv`` steps N independent restated ``PTGEnv`` instances with DummyVecEnv semantics (env order,
auto-reset, module-global ``ep_index``).  Noise comes from a per-env tape pre-drawn with numpy:
``default_rng(seed).normal(0, noise, size=L)`` equals L successive ``np_random.normal(0, noise, size=1)[0]``
calls of the reference bit for bit (SURVEY.md hard part 1).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rl_ptg_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libptg_oracle.so")
    src = os.path.join(_HERE, "ptg_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ptg_b200.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libptg_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ptg_oracle_create.restype = C.c_void_p
        L.ptg_oracle_create.argtypes = [C.POINTER(_abi.PtgConfig), C.POINTER(_abi.PtgTables), C.c_int64]
        L.ptg_oracle_destroy.argtypes = [C.c_void_p]
        L.ptg_oracle_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ptg_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [
            C.c_int, C.c_int]
        L.ptg_oracle_get_state.argtypes = [C.c_void_p, C.POINTER(_abi.PtgStateSoA)]
        L.ptg_oracle_obs_dim.argtypes = [C.POINTER(_abi.PtgConfig)]
        L.ptg_oracle_pairwise_sum.restype = C.c_double
        L.ptg_oracle_pairwise_sum.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.ptg_oracle_get_index.restype = C.c_int64
        L.ptg_oracle_get_index.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _LIB = L
    return _LIB


def draw_noise_tape(seeds, noise: float, length: int) -> np.ndarray:
    """[n_envs, length] fp64: what env e's ``np_random.normal(0, noise, size=1)[0]`` calls return, in order,
    after ``reset(seed=seeds[e])`` (gymnasium: Generator(PCG64(SeedSequence(seed))))."""
    seeds = np.asarray(seeds, dtype=np.int64).reshape(-1)
    tape = np.empty((len(seeds), length), dtype=np.float64)
    for e, s in enumerate(seeds):
        tape[e] = np.random.default_rng(int(s)).normal(0, noise, size=length)
    return tape


_ACT_DTYPES = {np.dtype(np.int64): _abi.ACT_I64, np.dtype(np.int32): _abi.ACT_I32, np.dtype(np.uint8): _abi.ACT_U8,
               np.dtype(np.float32): _abi.ACT_F32}


class OracleVecEnv:
    """N restated PTGEnv instances behind a DummyVecEnv-like interface (fp64 outputs, row-major per env)."""

    def __init__(self, dict_input: dict, n_envs: int, train_or_eval: str = "train", noise_tape=None,
                 threads: int = 1):
        self.n_envs = int(n_envs)
        noise_mode = _abi.NOISE_OFF if noise_tape is None else _abi.NOISE_TAPE
        self.cfg = _abi.config_from_kwargs(dict_input, train_or_eval, noise_mode, _abi.SCHED_DUMMY)
        self.tables, self._keep = _abi.tables_from_kwargs(dict_input, self.cfg.price_ahead)
        self.L = lib()
        self.obs_dim = self.L.ptg_oracle_obs_dim(C.byref(self.cfg))
        self.h = self.L.ptg_oracle_create(C.byref(self.cfg), C.byref(self.tables), self.n_envs)
        self.tape = None if noise_tape is None else np.ascontiguousarray(noise_tape, dtype=np.float64)
        if self.tape is not None:
            assert self.tape.shape[0] == self.n_envs
        self.threads = threads
        self.eval = train_or_eval == "eval"
        n = self.n_envs
        self.obs = np.zeros((n, self.obs_dim))
        self.reward = np.zeros(n)
        self.done = np.zeros(n, dtype=np.uint8)
        self.terminal_obs = np.zeros((n, self.obs_dim))
        self.info = np.zeros((n, _abi.PTG_N_INFO))
        self.episode_return = np.zeros(n)
        self.episode_length = np.zeros(n, dtype=np.int32)

    def close(self):
        if self.h:
            self.L.ptg_oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.ptg_oracle_reset(self.h, None if m is None else m.ctypes.data, self.obs.ctypes.data,
                                self.info.ctypes.data)
        return self.obs

    def step(self, actions, auto_reset: bool = True, want_info: bool | None = None):
        a = np.ascontiguousarray(actions).reshape(-1)
        assert a.shape[0] == self.n_envs
        want_info = self.eval if want_info is None else want_info
        tape_ptr, tape_len = (None, 0) if self.tape is None else (self.tape.ctypes.data, self.tape.shape[1])
        rc = self.L.ptg_oracle_step(
            self.h, a.ctypes.data, _ACT_DTYPES[a.dtype], tape_ptr, tape_len, self.obs.ctypes.data,
            self.reward.ctypes.data, self.done.ctypes.data, self.terminal_obs.ctypes.data,
            self.info.ctypes.data if want_info else None, self.episode_return.ctypes.data,
            self.episode_length.ctypes.data, int(auto_reset), int(self.threads))
        if rc != 0:
            raise RuntimeError(f"oracle step failed: {_abi.STATUS_NAMES.get(rc, rc)}")
        return self.obs, self.reward, self.done

    def get_state(self) -> dict:
        s, arrays = _abi.alloc_state(self.n_envs)
        self.L.ptg_oracle_get_state(self.h, C.byref(s))
        return arrays

    def get_index(self, ds: int, t_cat: float) -> int:
        return int(self.L.ptg_oracle_get_index(self.h, ds, float(t_cat)))
