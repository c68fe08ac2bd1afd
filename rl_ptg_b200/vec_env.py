"""``PtGVecEnv`` -- the batched PtG environment behind the Stable-Baselines3 ``VecEnv`` interface.

Drop-in for what ``create_vec_envs`` builds in the reference (``src/rl_utils.py:448-500``):
``VecNormalize(make_vec_env('PtGEnv-v0', n_envs, seed, vec_env_cls, env_kwargs=dict(dict_input=...,
train_or_eval=..., render_mode="None")), norm_obs=False)`` minus the reward normalisation -- i.e. N ``PTGEnv``
instances (``env/ptg_gym_env.py:23``) + ``Monitor`` + ``DummyVecEnv`` auto-reset, all stepped by ONE CUDA
kernel launch per ``step()``.

* constructor: the same flat ``dict_input`` the reference env takes (``src/rl_utils.py:337-405``)
* ``reset() / step_async() / step_wait() / step() / seed() / close() / get_attr() / set_attr() / env_method() /
  env_is_wrapped() / get_images() / render()`` with SB3 semantics (numpy in, numpy out; on ``done`` the
  returned obs is the reset obs, ``infos[i]["terminal_observation"]`` the last obs, ``infos[i]["episode"]``
  the Monitor record, ``infos[i]["TimeLimit.truncated"] = False``)
* ``step_tensor() / reset_tensor() / rollout_tensor()``: the same on CUDA tensors with zero host copies
* ``episode_stats()``: finished-episode statistics, combined across ranks when torch.distributed is initialised

There is no CPU fallback: constructing this class needs a CUDA device and the built ``libptg_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import time
from collections.abc import Sequence as _SequenceABC
from typing import Any, Sequence

import numpy as np
import torch

from . import _abi, _lib, spaces

try:  # pragma: no cover - SB3 is not installed in the build container
    from stable_baselines3.common.vec_env.base_vec_env import VecEnv as _SB3VecEnv  # type: ignore
except ModuleNotFoundError:
    _SB3VecEnv = None

_NOISE = {"numpy": _abi.NOISE_NUMPY, "tape": _abi.NOISE_TAPE, "off": _abi.NOISE_OFF}
_TORCH_ACT = {torch.int64: _abi.ACT_I64, torch.int32: _abi.ACT_I32, torch.uint8: _abi.ACT_U8,
              torch.float32: _abi.ACT_F32}
_TORCH_FROM_NUMPY_OK = (np.dtype(np.int64), np.dtype(np.int32), np.dtype(np.uint8), np.dtype(np.int16),
                        np.dtype(np.int8), np.dtype(np.float32))


def shard_range(n_envs_global: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous env-id range ``[lo, hi)`` owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_envs_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class LazyInfos(_SequenceABC):
    """``infos`` of a large batch: behaves like the list of per-env dicts SB3 expects (``len``, indexing, iteration)
    but only materialises the dicts that are looked at.  Train mode: ``{}`` everywhere except the envs whose
    episode just ended; eval mode: the 24-field dict of ``ptg_gym_env.py:253-278`` built on demand from the host
    copy of the info array."""

    __slots__ = ("_n", "_over", "_make")

    def __init__(self, n: int, overrides: dict | None = None, make=None):
        self._n, self._over, self._make = n, overrides or {}, make

    def __len__(self):
        return self._n

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return [self[q] for q in range(*idx.indices(self._n))]
        idx = int(idx)
        if idx < 0:
            idx += self._n
        if not 0 <= idx < self._n:
            raise IndexError(idx)
        d = self._over.get(idx)
        if d is None:
            d = self._make(idx) if self._make is not None else {}
            self._over[idx] = d          # identity is stable: callers may annotate the dict they were handed
        return d


class _VecEnvBase:
    """Minimal SB3 ``VecEnv`` contract, used when stable_baselines3 is not importable."""

    def __init__(self, num_envs, observation_space, action_space):
        self.num_envs = num_envs
        self.observation_space = observation_space
        self.action_space = action_space
        self.reset_infos = [{} for _ in range(num_envs)]
        self._seeds = [None for _ in range(num_envs)]
        self._options = [{} for _ in range(num_envs)]

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def _reset_seeds(self):
        self._seeds = [None for _ in range(self.num_envs)]

    def _reset_options(self):
        self._options = [{} for _ in range(self.num_envs)]

    def set_options(self, options=None):
        pass

    @property
    def unwrapped(self):
        return self

    def _get_indices(self, indices):
        if indices is None:
            return range(self.num_envs)
        if isinstance(indices, int):
            return [indices]
        return indices


_Base = _SB3VecEnv if _SB3VecEnv is not None else _VecEnvBase


class StepGraph:
    """Steps captured by ``PtGVecEnv.capture_steps``: ``replay()`` issues them with one host call.  The replay
    invalidates the numpy API's host mirror of the market-window blocks (the steps bypass ``step_wait``), so a
    ``step()`` that follows re-transfers them."""

    def __init__(self, env: "PtGVecEnv", graph: "torch.cuda.CUDAGraph"):
        self.env, self.graph = env, graph

    def replay(self) -> None:
        self.env._win_valid = False
        self.env._clock_ok = False          # (replayed steps do not pass through ptg_step's host-side clock tracking)
        self.graph.replay()


class PtGVecEnv(_Base):
    metadata = {"render_modes": ["None"]}

    def __init__(self, dict_input: dict, n_envs: int, train_or_eval: str = "train", render_mode: str = "None",
                 seed: int | None = None, device: str | torch.device = "cuda:0", noise: str = "numpy",
                 env_id_offset: int = 0, n_envs_global: int | None = None, obs_dtype=np.float32,
                 info_limit: int = 4096, obs_layout: str = "dict", auto_reset: bool = True):
        if noise not in _NOISE:
            raise ValueError(f"noise must be one of {sorted(_NOISE)}")
        if obs_layout not in ("dict", "flat"):
            raise ValueError('obs_layout must be "dict" (key-major blocks) or "flat" (one feature row per env)')
        self.obs_layout = obs_layout
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("PtGVecEnv needs a CUDA device; there is no CPU fallback")
        self._L = _lib.load()
        self.train_or_eval = train_or_eval
        self.render_mode = render_mode
        self.obs_dtype = np.dtype(obs_dtype)
        self.info_limit = info_limit
        self.n_envs_global = int(n_envs_global if n_envs_global is not None else n_envs)
        self.env_id_offset = int(env_id_offset)
        self.cfg = _abi.config_from_kwargs(dict_input, train_or_eval, _NOISE[noise],
                                           obs_layout=_abi.OBS_FLAT if obs_layout == "flat" else _abi.OBS_KEY_MAJOR,
                                           auto_reset=auto_reset)
        self.auto_reset = bool(auto_reset)
        tables, keep = _abi.tables_from_kwargs(dict_input, self.cfg.price_ahead)
        self.raw_modified = dict_input["raw_modified"]
        self.action_type = dict_input["action_type"]
        obs_space = spaces.observation_space(self.raw_modified, self.cfg.price_ahead)
        act_space = spaces.action_space(self.action_type)
        if _SB3VecEnv is not None:   # pragma: no cover
            super().__init__(int(n_envs), obs_space, act_space)
        else:
            _VecEnvBase.__init__(self, int(n_envs), obs_space, act_space)

        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev_index):
            _lib.check(self._L.ptg_create(C.byref(self.cfg), C.byref(tables), self.num_envs, self.env_id_offset,
                                          self.n_envs_global, dev_index, C.byref(h)))
        self._h = h
        del keep
        self._dev_index = dev_index
        self.obs_dim = self._L.ptg_obs_dim(self._h)
        self.feature_dim = int(self._L.ptg_features_dim(self._h))
        self.obs_elems = int(self._L.ptg_obs_elems(self._h))
        keys = (_abi.PtgObsKey * 16)()
        nk = self._L.ptg_obs_layout(self._h, keys, 16)
        self.obs_keys = [(keys[q].name.decode(), int(keys[q].dim), bool(keys[q].is_int32), int(keys[q].offset))
                         for q in range(nk)]

        n, dev = self.num_envs, self.device
        self._obs = torch.zeros(self.obs_elems, dtype=torch.float32, device=dev)
        self._term_obs = torch.zeros(self.obs_elems, dtype=torch.float32, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._info = torch.zeros((_abi.PTG_N_INFO, n), dtype=torch.float64, device=dev)
        self._ep_ret = torch.zeros(n, dtype=torch.float64, device=dev)
        self._ep_len = torch.zeros(n, dtype=torch.int32, device=dev)
        self._stats_bufs = [torch.zeros(8, dtype=torch.float64, device=dev) for _ in range(2)]
        self._stats, self._stats_flip, self._stats_done, self._stats_ranks = self._stats_bufs[0], 0, [None, None], 1
        self.stats_stream = None                 # side stream of episode_stats_async(overlap=True), created on first use
        self._win_flag = torch.zeros(1, dtype=torch.int32, device=dev)     # PtgIO.windows_changed (step serial stamp)
        self._status_d = torch.zeros(n, dtype=torch.uint8, device=dev)     # PtgIO.status_u8: METH_STATUS, one byte per env
        self._status_h = torch.zeros(n, dtype=torch.uint8).pin_memory()
        # METH_STATUS as the Discrete(6) space's int64, handed out as a view of one of two reused buffers (valid until
        # the step after next, like every other observation block); filled by torch's multi-threaded converting copy
        self._status_i64 = [torch.zeros(n, dtype=torch.int64) for _ in range(2)]
        self._clock_ok = True                    # the library's shared-clock tracking saw every step (no graph replays)
        self._obs_dict = self._obs_views(self._obs)          # views are created once; buffers are reused
        self._io = self._make_io(self._obs, self._reward, self._done, self._term_obs,
                                 self._info if self.cfg.train_or_eval else None, self._ep_ret, self._ep_len)
        self._io_ref, self._ptg_step = C.byref(self._io), self._L.ptg_step
        # pinned host mirrors for the numpy API
        # two host observation buffers used alternately: the dict returned by step t stays valid until step
        # t + 2, which is what SB3's collect_rollouts needs (it stores the previous obs after the next step)
        self._obs_hh = [torch.zeros(self.obs_elems, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._flip = 0
        # The market-window blocks (3/4 of the observation bytes) only change when an env's clock crosses an hour or
        # its episode ends: the step kernel raises `windows_changed` then, and the host mirror re-transfers those
        # blocks only on such steps (the flag then carries the serial number of that step, so it never needs clearing).
        # They live in whichever of the two host buffers received them last (_win_buf);
        # a new version always goes to the OTHER buffer, so the dict handed out before stays intact.
        self._win_flag_h = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._win_buf = 0
        self._win_valid = False                  # host copy of the window blocks matches the device's
        self.d2h_bytes = 0                       # bytes the numpy API copied device -> host so far
        self._term_obs_h = torch.zeros(self.obs_elems, dtype=torch.float32).pin_memory()
        # rewards / dones: two pinned mirrors used alternately like the observation buffers, handed out as views
        self._reward_hh = [torch.zeros(n, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._done_hh = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._info_h = torch.zeros((_abi.PTG_N_INFO, n), dtype=torch.float64).pin_memory()
        self._ep_ret_h = torch.zeros(n, dtype=torch.float64).pin_memory()
        self._ep_len_h = torch.zeros(n, dtype=torch.int32).pin_memory()
        act_dtype = torch.float32 if self.action_type == "continuous" else torch.int64
        self._act_wire_u8 = self.action_type != "continuous"
        wire_dtype = torch.uint8 if self._act_wire_u8 else act_dtype
        self._act_h = torch.zeros(n, dtype=wire_dtype).pin_memory()
        self._act_d = torch.zeros(n, dtype=wire_dtype, device=dev)
        self.h2d_bytes_per_step = n * self._act_h.element_size()
        self._step_serial = 0
        self._comm = None                        # ncclComm_t of the statistics all-gather (created on first use)
        self._tape = None
        self._t_start = time.time()
        self._ev_small = torch.cuda.Event()
        self._ev_scalars = torch.cuda.Event()
        self._scalar_off = 0 if obs_layout == "flat" else min(off for name, _, _, off in self.obs_keys
                                                               if name == "METH_STATUS")
        self._status_elems = (n + 3) // 4 * 4     # padded length of one scalar block (16 B aligned block starts)
        self.bytes_per_env_step = int(self._L.ptg_bytes_per_env_step(self._h, _TORCH_ACT[act_dtype]))
        if seed is not None:
            self.seed(seed)

    # ------------------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _make_io(self, obs, reward, done, term_obs=None, info=None, ep_ret=None, ep_len=None) -> _abi.PtgIO:
        io = _abi.PtgIO()
        io.obs, io.reward, io.done = obs.data_ptr(), (reward.data_ptr() if reward is not None else None), (
            done.data_ptr() if done is not None else None)
        io.terminal_obs = term_obs.data_ptr() if term_obs is not None else None
        io.info = info.data_ptr() if info is not None else None
        io.episode_return = ep_ret.data_ptr() if ep_ret is not None else None
        io.episode_length = ep_len.data_ptr() if ep_len is not None else None
        io.windows_changed = self._win_flag.data_ptr() if (reward is not None and self.obs_layout != "flat") else None
        io.status_u8 = self._status_d.data_ptr() if reward is not None else None
        return io

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _obs_views(self, buf: torch.Tensor) -> dict:
        """dict key -> [n_envs, dim] view of one obs buffer (torch tensor on any device).  Flat layout: strided
        column views of the [n_envs, F] feature matrix (METH_STATUS = its six one-hot columns)."""
        n, out = self.num_envs, {}
        if self.obs_layout == "flat":
            rows = buf[:n * self.feature_dim].view(n, self.feature_dim)
            for name, dim, _, col in self.obs_keys:
                out[name] = rows[:, col:col + dim]
            return out
        for name, dim, is_int, off in self.obs_keys:
            v = buf[off:off + n * dim]
            out[name] = v.view(torch.int32) if is_int else v.view(n, dim)
        return out

    def features_view(self, buf: torch.Tensor | None = None) -> torch.Tensor:
        """Flat layout only: the [n_envs, F] feature matrix of an obs buffer (default: the env's current one)."""
        if self.obs_layout != "flat":
            raise RuntimeError('features_view() needs obs_layout="flat" (use vec_normalize.features_tensor otherwise)')
        buf = self._obs if buf is None else buf
        return buf[:self.num_envs * self.feature_dim].view(self.num_envs, self.feature_dim)

    def _obs_numpy(self, buf_h: torch.Tensor, rows=None, status=None, win_buf_h: torch.Tensor | None = None) -> dict:
        views = self._obs_views(buf_h)
        win_views = views if win_buf_h is None else self._obs_views(win_buf_h)
        out = {}
        for name, dim, is_int, off in self.obs_keys:
            if name == "METH_STATUS" and status is not None:
                out[name] = status
                continue
            is_window = self.obs_layout != "flat" and off < self._scalar_off
            a = (win_views if is_window else views)[name].numpy()
            if rows is not None:
                a = a[rows]
            if name == "METH_STATUS" and self.obs_layout == "flat":
                out[name] = a.argmax(axis=1).astype(np.int64)        # one-hot columns -> Discrete(6) value
            else:
                out[name] = a.astype(np.int64) if is_int else a.astype(self.obs_dtype, copy=False)
        return out

    def _next_obs_host(self) -> torch.Tensor:
        self._flip ^= 1
        return self._obs_hh[self._flip]

    def _check_open(self):
        if self._h is None:
            raise RuntimeError("PtGVecEnv is closed")

    def poll_error(self):
        """Raise if the device saw an invalid action / ran past the market tables / exhausted the noise tape."""
        _lib.check(self._L.ptg_poll_error(self._h, self._stream()))

    # ------------------------------------------------------------------------------------------------------
    # SB3 VecEnv interface (numpy)
    # ------------------------------------------------------------------------------------------------------
    def seed(self, seed: int | None = None) -> Sequence[int | None]:
        """SB3: env i is re-seeded with ``seed + i`` at the next ``reset()`` (global env id for sharded envs)."""
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self._seeds = [seed + self.env_id_offset + i for i in range(self.num_envs)]
        return list(self._seeds)

    def reset(self):
        self.reset_tensor(_return_views=False)
        obs_h = self._next_obs_host()
        obs_h.copy_(self._obs, non_blocking=True)
        self._win_buf, self._win_valid = self._flip, True
        self.d2h_bytes += obs_h.numel() * 4
        self._info_h.copy_(self._info, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.poll_error()
        if self.num_envs <= self.info_limit:      # reset() returns the full info dict even in train mode (:503-506)
            self.reset_infos = self._info_dicts(range(self.num_envs))
        else:
            snap = self._info_h.numpy().copy()
            self.reset_infos = LazyInfos(self.num_envs, None, lambda e, a=snap: self._info_dict(a[:, e]))
        return self._obs_numpy(obs_h)

    def step_async(self, actions) -> None:
        self._check_open()
        a = np.asarray(actions)
        if self.action_type == "continuous":
            a = a.astype(np.float32, copy=False).reshape(self.num_envs)
        else:
            if a.dtype.kind not in "iu":
                a = a.astype(np.int64)
            a = a.reshape(self.num_envs)
        # Discrete actions travel as uint8 (the ABI takes PTG_ACT_U8): 1 byte per env over PCIe instead of SB3's int64.
        # Values outside 0..4 map to 255, which the kernel reports like any other invalid action.  torch's converting
        # copy is multi-threaded for large arrays (torch.set_num_threads; under torchrun OMP_NUM_THREADS defaults to 1).
        if self._act_wire_u8 and a.dtype != np.uint8:
            lo, hi = (int(a.min()), int(a.max())) if a.size <= 4096 else \
                (int(v) for v in torch.aminmax(torch.from_numpy(np.ascontiguousarray(a))))
            if lo < 0 or hi > 4:                                 # rare: keep the invalid entries visible to the kernel
                a = np.where((a < 0) | (a > 4), 255, a).astype(np.uint8)
        if a.flags.c_contiguous and a.flags.writeable and a.dtype in _TORCH_FROM_NUMPY_OK:
            self._act_h.copy_(torch.from_numpy(a))               # (converting copy: int64 -> uint8)
        else:
            self._act_h.numpy()[:] = a
        self._act_d.copy_(self._act_h, non_blocking=True)
        _lib.check(self._L.ptg_step(self._h, self._ptr(self._act_d), _TORCH_ACT[self._act_d.dtype],
                                    C.byref(self._io), self._stream()))
        self._step_serial = int(self._L.ptg_last_step_serial(self._h))

    def step_wait(self):
        obs_h = self._next_obs_host()
        stream = torch.cuda.current_stream(self.device)
        # small results first: the host can look at `dones` while the observation block is still in flight
        done_h, reward_h = self._done_hh[self._flip], self._reward_hh[self._flip]
        done_h.copy_(self._done, non_blocking=True)
        reward_h.copy_(self._reward, non_blocking=True)
        self._win_flag_h.copy_(self._win_flag, non_blocking=True)
        self._ev_small.record(stream)
        # METH_STATUS travels as one byte per env (PtgIO.status_u8) ahead of everything else, so its widening to int64
        # overlaps the rest of the transfer; the 4-byte block of the obs buffer stays on the device
        self._status_h.copy_(self._status_d, non_blocking=True)
        self._ev_scalars.record(stream)
        cut = self._scalar_off
        sin_v, cos_v = C.c_float(), C.c_float()
        clock = None           # (sin, cos) shared by every env: those two blocks then need no transfer at all
        if self._clock_ok and self._L.ptg_clock_uniform(self._h, C.byref(sin_v), C.byref(cos_v)):
            clock = (np.float32(sin_v.value), np.float32(cos_v.value))
        if self.obs_layout == "flat":
            obs_h.copy_(self._obs, non_blocking=True)
            self.d2h_bytes += obs_h.numel() * 4 + self.num_envs * 6 + 4
        else:                                       # the six plant scalar blocks (+ the clock blocks if envs differ)
            mid = cut + int(self._status_elems)
            end = mid + (6 if clock is not None else 8) * int(self._status_elems)
            obs_h[mid:end].copy_(self._obs[mid:end], non_blocking=True)
            self.d2h_bytes += (end - mid) * 4 + self.num_envs * 6 + 4
        eval_mode = bool(self.cfg.train_or_eval)
        if eval_mode:
            self._info_h.copy_(self._info, non_blocking=True)
            self.d2h_bytes += self._info_h.numel() * 8
        self._ev_small.synchronize()
        # window blocks: only when the step moved them (or the host copy is not known to be current)
        if cut > 0 and (not self._win_valid or (int(self._win_flag_h[0]) & 0xffffffff) == self._step_serial):
            # a new version goes to the buffer that does not hold the version handed out last
            self._win_buf = (self._win_buf ^ 1) if self._win_valid else self._flip
            self._obs_hh[self._win_buf][:cut].copy_(self._obs[:cut], non_blocking=True)
            self._win_valid = True
            self.d2h_bytes += cut * 4
        dones = done_h.numpy().view(np.bool_)          # views: valid until the step after next, like the obs dict
        rewards = reward_h.numpy()
        any_done = bool(dones.any())
        self._ev_scalars.synchronize()
        status_t = self._status_i64[self._flip]
        status_t.copy_(self._status_h)                 # uint8 -> int64 (0.09 ms at 1 M envs; numpy's astype: 0.6 ms)
        status = status_t.numpy()
        if any_done:
            self._term_obs_h.copy_(self._term_obs, non_blocking=True)
            self._ep_ret_h.copy_(self._ep_ret, non_blocking=True)
            self._ep_len_h.copy_(self._ep_len, non_blocking=True)
        stream.synchronize()
        self.poll_error()
        lazy = self.num_envs > self.info_limit
        # snapshots of what the dicts are built from (the pinned mirrors are overwritten by the next step)
        info_a = None
        if eval_mode:
            info_a = self._info_h.numpy().copy() if lazy else self._info_h.numpy()
        term = ret = length = None
        if any_done:
            term = self._obs_numpy(self._term_obs_h)
            ret, length = self._ep_ret_h.numpy(), self._ep_len_h.numpy()
            if lazy:
                term = {k: v.copy() for k, v in term.items()}
                ret, length = ret.copy(), length.copy()
        t_now = round(time.time() - self._t_start, 6)
        done_snap = dones.copy() if any_done else None      # (the returned `dones` is a view of a reused mirror)

        def make(e):
            d = self._info_dict(info_a[:, e]) if info_a is not None else {}
            if any_done and done_snap[e]:
                d["episode"] = {"r": round(float(ret[e]), 6), "l": int(length[e]), "t": t_now}   # Monitor
                d["TimeLimit.truncated"] = False                     # the env only ever terminates (:478-481)
                d["terminal_observation"] = {k: np.array(v[e]) for k, v in term.items()}     # (own copy)
            return d

        if lazy:
            infos = LazyInfos(self.num_envs, None, make if (eval_mode or any_done) else None)
        else:
            infos = [make(e) for e in range(self.num_envs)]
        obs = self._obs_numpy(obs_h, status=status, win_buf_h=self._obs_hh[self._win_buf])
        if clock is not None and self.obs_layout != "flat":
            # one value per block: handed out as read-only broadcast views (the kernels wrote the same fp32 values)
            obs["Temp_hour_enc_sin"] = np.broadcast_to(clock[0].astype(self.obs_dtype), (self.num_envs, 1))
            obs["Temp_hour_enc_cos"] = np.broadcast_to(clock[1].astype(self.obs_dtype), (self.num_envs, 1))
        return obs, rewards, dones, infos

    def close(self) -> None:
        if getattr(self, "_h", None) is not None:
            if getattr(self, "_comm", None) is not None:
                self._L.ptg_nccl_comm_destroy(self._comm)
                self._comm = None
            self._L.ptg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _info_dict(col) -> dict:
        """One env's 24-field dict in the reference's key order (ptg_gym_env.py:253-278)."""
        d = {}
        for f, key in enumerate(_abi.INFO_KEYS):
            if key in ("step", "Meth_State", "Meth_Hot_Cold"):
                d[key] = int(col[f])
            elif key == "Meth_Action":
                d[key] = _abi.STATE_NAMES[int(col[f])]
            else:
                d[key] = float(col[f])
        return d

    def _info_dicts(self, indices):
        arr = self._info_h.numpy()
        return [self._info_dict(arr[:, e]) for e in indices]

    _STATE_ATTRS = {"Meth_State": "meth_state", "i": "i", "j": "j", "k": "k", "hot_cold": "hot_cold",
                    "Meth_T_cat": "t_cat", "act_ep_h": "act_ep_h", "act_ep_d": "act_ep_d", "cum_rew": "cum_reward"}

    def get_attr(self, attr_name: str, indices=None) -> list[Any]:
        idx = list(self._get_indices(indices))
        if attr_name in self._STATE_ATTRS:
            col = self.get_state()[self._STATE_ATTRS[attr_name]]
            return [col[i].item() for i in idx]
        if attr_name == "current_action":
            col = self.get_state()["current_action"]
            return [_abi.STATE_NAMES[int(col[i])] for i in idx]
        if hasattr(self, attr_name):
            return [getattr(self, attr_name) for _ in idx]
        raise AttributeError(attr_name)

    def set_attr(self, attr_name: str, value: Any, indices=None) -> None:
        if attr_name in ("render_mode", "info_limit"):
            setattr(self, attr_name, value)
            return
        raise AttributeError(f"attribute {attr_name!r} of the batched env cannot be set after construction")

    def env_method(self, method_name: str, *method_args, indices=None, **method_kwargs) -> list[Any]:
        raise NotImplementedError("the batched env has no per-env Python objects; use the VecEnv methods")

    def env_is_wrapped(self, wrapper_class, indices=None) -> list[bool]:
        """Monitor statistics are built in (infos[i]['episode']), so evaluate_policy may rely on them."""
        is_monitor = getattr(wrapper_class, "__name__", "") == "Monitor"
        return [is_monitor for _ in self._get_indices(indices)]

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def render(self, mode: str | None = None):
        return None

    # ------------------------------------------------------------------------------------------------------
    # device API (zero host copies)
    # ------------------------------------------------------------------------------------------------------
    def reset_tensor(self, seeds=None, mask=None, _return_views: bool = True):
        """Reset (all or the masked) envs; returns dict of CUDA tensor views into the env's obs buffer."""
        self._check_open()
        if seeds is None and any(s is not None for s in self._seeds):
            seeds = np.array([-1 if s is None else s for s in self._seeds], dtype=np.int64)
        seeds_p = None
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, dtype=np.int64)
            assert seeds.shape == (self.num_envs,)
            seeds_p = seeds.ctypes.data
        mask_p = None
        if mask is not None:
            mask = np.ascontiguousarray(mask, dtype=np.uint8)
            assert mask.shape == (self.num_envs,)
            mask_p = mask.ctypes.data
        io = self._make_io(self._obs, None, None, None, self._info)
        _lib.check(self._L.ptg_reset(self._h, seeds_p, mask_p, C.byref(io), self._stream()))
        if mask is None:
            self._clock_ok = True
        self._win_valid = False
        self._reset_seeds()
        return self._obs_dict if _return_views else None

    def step_tensor(self, actions: torch.Tensor):
        """One step on device: ``actions`` is a CUDA tensor [n_envs] (int64/int32/uint8, or float32 for
        continuous).  Returns (obs views, reward float32[n], done uint8[n]); buffers are reused every call.
        (Kept lean: at 131 072 envs a step is ~9 us of GPU time, so every microsecond of host work per call shows.)"""
        if self._h is None:
            raise RuntimeError("PtGVecEnv is closed")
        a = actions if actions.dim() == 1 else actions.reshape(-1)
        if a.device != self.device or a.numel() != self.num_envs or not a.is_contiguous():
            raise ValueError("actions must be a contiguous CUDA tensor with n_envs elements on the env's device")
        rc = self._ptg_step(self._h, a.data_ptr(), _TORCH_ACT[a.dtype], self._io_ref,
                            torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc)
        self._win_valid = False                 # (the host mirror of the numpy API did not see this step)
        return self._obs_dict, self._reward, self._done

    def rollout_tensor(self, actions: torch.Tensor, out: dict | None = None):
        """T steps in ONE launch (state stays in registers): ``actions`` is [T, n_envs] on device.  Returns a dict
        with ``obs`` [T, obs_elems] fp32 (key-major per step, see ``obs_views_of``), ``reward`` [T, n], ``done``
        [T, n]; pass ``out`` to reuse buffers."""
        self._check_open()
        T = int(actions.shape[0])
        if actions.device != self.device or actions.numel() != T * self.num_envs or not actions.is_contiguous():
            raise ValueError("actions must be a contiguous CUDA tensor [T, n_envs] on the env's device")
        if out is None:
            out = {"obs": torch.empty((T, self.obs_elems), dtype=torch.float32, device=self.device),
                   "reward": torch.empty((T, self.num_envs), dtype=torch.float32, device=self.device),
                   "done": torch.empty((T, self.num_envs), dtype=torch.uint8, device=self.device)}
        io = self._make_io(out["obs"], out["reward"], out["done"])
        _lib.check(self._L.ptg_step_many(self._h, self._ptr(actions), _TORCH_ACT[actions.dtype], T, C.byref(io),
                                         self._stream()))
        self._win_valid = False
        return out

    def capture_steps(self, action_buffers: Sequence[torch.Tensor]) -> "StepGraph":
        """Capture ``len(action_buffers)`` consecutive single steps into ONE CUDA graph (SURVEY.md build plan item 7):
        ``graph.replay()`` then issues them with a single host call -- for shards small enough that a step (a few us)
        costs less than a Python/driver launch.  The steps read ``action_buffers[q]`` (fill them before each
        replay) and leave observation / reward / done of the LAST step in the env's buffers; use
        ``rollout_tensor`` when every step's outputs are needed."""
        self._check_open()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                    # warm-up outside the capture (lazy module loading)
            self.step_tensor(action_buffers[0])
        torch.cuda.current_stream(self.device).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for a in action_buffers:
                self.step_tensor(a)
        return StepGraph(self, graph)

    def obs_views_of(self, buf: torch.Tensor) -> dict:
        """Key views of any single obs buffer with this env's layout (e.g. ``rollout['obs'][t]``)."""
        return self._obs_views(buf)

    @property
    def terminal_obs_tensor(self) -> dict:
        return self._obs_views(self._term_obs)

    def set_noise_tape(self, tape) -> None:
        """Tape-mode noise: [n_envs, L] fp64 values of ``normal(0, noise)`` consumed in order per env."""
        t = torch.as_tensor(np.ascontiguousarray(tape, dtype=np.float64)).to(self.device)
        assert t.shape[0] == self.num_envs
        self._tape = t
        _lib.check(self._L.ptg_set_noise_tape(self._h, self._ptr(t), t.shape[1]))

    # ------------------------------------------------------------------------------------------------------
    # state snapshot / statistics
    # ------------------------------------------------------------------------------------------------------
    def get_state(self) -> dict:
        self._check_open()
        s, arrays = _abi.alloc_state(self.num_envs)
        _lib.check(self._L.ptg_get_state(self._h, C.byref(s)))
        return arrays

    def set_state(self, arrays: dict) -> None:
        self._check_open()
        s = _abi.PtgStateSoA()
        keep = {}
        for name, dt in _abi.STATE_FIELDS:
            keep[name] = np.ascontiguousarray(arrays[name], dtype=dt)
            assert keep[name].shape == _abi.state_shape(name, self.num_envs), name
            setattr(s, name, keep[name].ctypes.data)
        _lib.check(self._L.ptg_set_state(self._h, C.byref(s)))
        self._win_valid = False

    def kernel_launches(self) -> int:
        v = C.c_int64()
        _lib.check(self._L.ptg_kernel_launches(self._h, C.byref(v)))
        return int(v.value)

    def _nccl_comm(self):
        """ncclComm_t for ``ptg_allreduce_stats``, created once: rank 0's unique id travels over torch.distributed."""
        import torch.distributed as dist
        if self._comm is None:
            world, rank = dist.get_world_size(), dist.get_rank()
            buf = C.create_string_buffer(_abi.PTG_NCCL_UNIQUE_ID_BYTES)
            if rank == 0:
                _lib.check(self._L.ptg_nccl_unique_id(buf))
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(self.device)
            dist.broadcast(t, 0)
            ident = bytes(t.cpu().numpy().tobytes())
            comm = C.c_void_p()
            with torch.cuda.device(self._dev_index):
                _lib.check(self._L.ptg_nccl_comm_create(ident, world, rank, C.byref(comm)))
            self._comm = comm
        return self._comm

    def episode_stats_async(self, clear: bool = True, reduce: bool = True, overlap: bool = False) -> torch.Tensor:
        """Device-side part of ``episode_stats``: the reduction over this rank's envs and -- with NCCL initialised --
        the cross-rank all-gather + combine inside the library (``ptg_allreduce_stats``), without a host
        synchronisation.  Returns the 8 x fp64 ``PtgEpisodeStats`` record (device tensor, one of two reused buffers).

        ``overlap=False``: everything on the current stream.  ``overlap=True``: only the local reduction (which reads
        and clears the per-env accumulators) stays on the current stream; the latency-bound NCCL all-gather + combine go
        to ``self.stats_stream`` and overlap the next roll-out -- the record is then ordered on that stream (read it
        under ``torch.cuda.stream(env.stats_stream)`` or after a device synchronisation)."""
        import torch.distributed as dist
        self._check_open()
        idx = self._stats_flip
        self._stats_flip ^= 1
        rec = self._stats_bufs[idx]
        main = torch.cuda.current_stream(self.device)
        if self._stats_done[idx] is not None:
            main.wait_event(self._stats_done[idx])      # the side stream's last use of this buffer (two calls ago)
            self._stats_done[idx] = None
        _lib.check(self._L.ptg_episode_stats(self._h, self._ptr(rec), int(clear), C.c_void_p(main.cuda_stream)))
        if (reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
                and dist.get_backend() == "nccl"):
            st = main
            if overlap:
                if self.stats_stream is None:
                    self.stats_stream = torch.cuda.Stream(device=self.device)
                st = self.stats_stream
                st.wait_stream(main)
            _lib.check(self._L.ptg_allreduce_stats(self._h, self._nccl_comm(), self._ptr(rec), C.c_void_p(st.cuda_stream)))
            if overlap:
                ev = torch.cuda.Event()
                ev.record(st)
                self._stats_done[idx] = ev
            self._stats_ranks = dist.get_world_size()
        else:
            self._stats_ranks = 1
        self._stats = rec
        return rec

    def episode_stats(self, clear: bool = True, reduce: bool = True) -> dict:
        """Finished-episode statistics since the last clear: device reduction (warp shuffles, deterministic), then --
        if torch.distributed is initialised and ``reduce`` -- ONE all-gather of the 64-byte record over NCCL/NVLink
        inside the library (``ptg_allreduce_stats``; gloo falls back to ``combine_stats``) and a fixed-order combine."""
        import torch.distributed as dist
        stats = self.episode_stats_async(clear, reduce)
        if self._stats_ranks > 1:
            return stats_dict(stats.cpu().numpy(), self._stats_ranks)
        return combine_stats(stats, reduce and not (dist.is_available() and dist.is_initialized()
                                                   and dist.get_backend() == "nccl" and dist.get_world_size() > 1))


def stats_dict(rec: np.ndarray, ranks: int) -> dict:
    """The user-facing view of one (combined) 8 x fp64 ``PtgEpisodeStats`` record."""
    n, s1, s2, sl, mn, mx, steps = (float(v) for v in rec[:7])
    mean = s1 / n if n > 0 else float("nan")
    var = max(s2 / n - mean * mean, 0.0) if n > 0 else float("nan")
    return {"episodes": int(n), "return_mean": mean, "return_std": var ** 0.5,
            "length_mean": sl / n if n > 0 else float("nan"), "return_min": mn, "return_max": mx,
            "env_steps": int(steps), "ranks": int(ranks)}


def combine_stats(stats: torch.Tensor, reduce: bool = True) -> dict:
    """All-gather one rank's 8 x fp64 ``PtgEpisodeStats`` record and combine in rank order."""
    import torch.distributed as dist
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        gathered = [torch.empty_like(stats) for _ in range(dist.get_world_size())]
        dist.all_gather(gathered, stats)
        per_rank = torch.stack(gathered).cpu().numpy()
    else:
        per_rank = stats.detach().cpu().numpy()[None, :]
    L = _lib.load()
    arr = (_abi.PtgEpisodeStats * per_rank.shape[0])()
    for r in range(per_rank.shape[0]):
        for f, (name, _) in enumerate(_abi.PtgEpisodeStats._fields_):
            setattr(arr[r], name, float(per_rank[r, f]))
    out = _abi.PtgEpisodeStats()
    L.ptg_stats_combine(arr, per_rank.shape[0], C.byref(out))
    rec = np.array([getattr(out, name) for name, _ in _abi.PtgEpisodeStats._fields_])
    return stats_dict(rec, per_rank.shape[0])
