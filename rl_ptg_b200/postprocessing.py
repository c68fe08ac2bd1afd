"""Deterministic evaluation roll-outs -- the role of ``Postprocessing.test_performance`` (``src/rl_utils.py:528-565``;
SURVEY.md 8(f) row 4).

The reference steps ONE eval-mode env for ``eps_sim_steps_test`` steps with ``model.predict(obs, deterministic=True)``
and copies ``info[0]`` positionally into a ``(steps, 24)`` array whose columns are ``EnvConfig.stats_names``
(``src/rl_config_env.py:44-49``); ``Meth_Action`` strings become 0..4, and the rows of steps that return
``terminated`` stay zero (``if not terminated``, ``:541``).  ``Postprocessing`` below does exactly that through the
numpy ``VecEnv`` API.  ``eval_stats_tensor`` is the batched form: every env of an eval-mode ``PtGVecEnv`` (e.g. many
seeds or policies at once) is rolled out on the device and the result is one ``[steps, 24, n_envs]`` fp64 tensor --
the per-step info block written by the step kernel is copied device-to-device, nothing touches the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _abi
from .config import STATS_NAMES
from .vec_env import PtGVecEnv

_ACTION_INDEX = {name: float(q) for q, name in enumerate(_abi.STATE_NAMES)}


class Postprocessing:
    """``model`` needs ``predict(obs, deterministic=True) -> (actions, state)`` like an SB3 model."""

    def __init__(self, env_test_post: PtGVecEnv, model, eps_sim_steps_test: int, stats_names=STATS_NAMES):
        if not env_test_post.cfg.train_or_eval:
            raise ValueError('the post-processing env must be built with train_or_eval="eval" (src/rl_utils.py:491)')
        self.env_test_post, self.model = env_test_post, model
        self.eps_sim_steps_test = int(eps_sim_steps_test)
        self.stats_names = list(stats_names)
        self.stats_dict_test: dict = {}

    def test_performance(self) -> None:
        stats = np.zeros((self.eps_sim_steps_test, len(self.stats_names)))
        obs = self.env_test_post.reset()
        for i in range(self.eps_sim_steps_test):
            action, _ = self.model.predict(obs, deterministic=True)
            obs, _, terminated, info = self.env_test_post.step(action)
            if not terminated[0]:
                for j, (key, val) in enumerate(info[0].items()):
                    if j < 24:
                        stats[i, j] = _ACTION_INDEX[val] if key == "Meth_Action" else val
        for m, name in enumerate(self.stats_names):
            self.stats_dict_test[name] = stats[:, m]
        return None


@torch.no_grad()
def eval_stats_tensor(env: PtGVecEnv, policy, steps: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Roll every env of an eval-mode ``PtGVecEnv`` out for ``steps`` steps on the device.

    ``policy(obs_views) -> actions`` maps the dict of CUDA observation views to a CUDA action tensor ``[n_envs]``.
    Returns ``[steps, 24, n_envs]`` fp64 (columns = ``STATS_NAMES``); rows of terminated steps are zero like the
    reference's."""
    if not env.cfg.train_or_eval:
        raise ValueError('eval_stats_tensor needs an env built with train_or_eval="eval"')
    n = env.num_envs
    if out is None:
        out = torch.zeros((steps, _abi.PTG_N_INFO, n), dtype=torch.float64, device=env.device)
    obs = env.reset_tensor()
    for t in range(steps):
        actions = policy(obs)
        obs, _, done = env.step_tensor(actions.contiguous())
        torch.mul(env._info, (done == 0).to(torch.float64).unsqueeze(0), out=out[t])
    return out
