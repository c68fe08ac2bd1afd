"""``PTGEnv`` -- the single-environment Gymnasium interface of the reference (``env/ptg_gym_env.py:23``,
registered as ``'PtGEnv-v0'`` by ``rl_main.py:44-65``) on top of the batched CUDA path.

Same constructor (``dict_input, train_or_eval="train", render_mode="None"``), ``reset(seed=None, options=None) ->
(obs, info)`` and ``step(action) -> (obs, reward, terminated, truncated, info)`` as the reference class, same
``observation_space`` / ``action_space``.  It is a batch of ONE env: every call is a kernel launch plus a device
synchronisation (tens of microseconds), so it exists for API completeness, spot checks and ``gym.make``-style
callers -- throughput comes from ``PtGVecEnv``.  Unlike the ``VecEnv`` it does NOT auto-reset
(``PtgConfig.no_auto_reset``): the step that terminates returns the terminal observation and leaves the env as it
is; only ``reset()`` consumes the next ``eps_ind`` entry.  A lone env therefore walks ``eps_ind[0]`` (constructor),
``[1]``, ``[2]``, ... exactly like a lone reference env (``env/ptg_gym_env.py:59-62, 490-493``; pinned by the
train-mode golden case ``single_env_train_resets``).
"""
from __future__ import annotations

import numpy as np

from .vec_env import PtGVecEnv


class PTGEnv:
    metadata = {"render_modes": ["None"]}

    def __init__(self, dict_input: dict, train_or_eval: str = "train", render_mode: str = "None",
                 device="cuda:0", noise: str = "numpy"):
        # eval-mode kernel variant: the info block is written every step; train mode returns {} like the reference
        self._venv = PtGVecEnv(dict_input, 1, train_or_eval="eval", render_mode=render_mode, device=device, noise=noise,
                               auto_reset=False)
        self.train_or_eval = train_or_eval
        self.render_mode = render_mode
        self.observation_space = self._venv.observation_space
        self.action_space = self._venv.action_space
        self._terminated = False

    @staticmethod
    def _single(obs: dict) -> dict:
        out = {}
        for k, v in obs.items():
            out[k] = int(v[0]) if k == "METH_STATUS" else np.array(v[0], dtype=np.float64, copy=True)
        return out

    def reset(self, seed: int | None = None, options=None):
        if seed is not None:
            self._venv.seed(int(seed))              # gymnasium: reset(seed=...) re-creates np_random
        obs = self._venv.reset()
        self._terminated = False
        return self._single(obs), dict(self._venv.reset_infos[0])

    def step(self, action):
        if self._terminated:
            raise RuntimeError("step() after the episode terminated: call reset() first")
        a = np.asarray(action).reshape(1)
        obs, rew, done, infos = self._venv.step(a)
        info = dict(infos[0])
        terminated = bool(done[0])
        if terminated:                               # Monitor / VecEnv annotations are not part of the Gymnasium API
            for key in ("terminal_observation", "episode", "TimeLimit.truncated"):
                info.pop(key, None)
            self._terminated = True
        obs_single = self._single(obs)               # (no auto-reset: at `terminated` this IS the terminal observation)
        if self.train_or_eval != "eval":
            info = {}                                # :471-474
        return obs_single, float(rew[0]), terminated, False, info

    def render(self):
        return None

    def close(self):
        self._venv.close()

    # a few of the attributes callers of the reference class read
    def __getattr__(self, name):
        if name in PtGVecEnv._STATE_ATTRS or name == "current_action":
            return self._venv.get_attr(name)[0]
        raise AttributeError(name)
