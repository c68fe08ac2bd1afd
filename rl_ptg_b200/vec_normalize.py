"""Device-side mirror of what the reference wraps around every env: ``VecNormalize(env, norm_obs=False)``
(``src/rl_utils.py:453`` for training envs, ``:491`` for the post-processing env) plus the two roll-out pieces a
``MultiInputPolicy`` PPO needs next to the env: the flat feature rows of SB3's
``CombinedExtractor`` and ``RolloutBuffer.compute_returns_and_advantage``.

``VecNormalizeReward`` keeps SB3's attribute names (``ret_rms``, ``returns``, ``training``, ``norm_reward``,
``clip_reward``, ``gamma``, ``epsilon``, ``get_original_reward()``, ``normalize_reward()``) and semantics:

    returns = returns * gamma + reward ; ret_rms.update(returns)        (training only)
    reward  = clip(reward / sqrt(ret_rms.var + epsilon), -clip, clip)
    returns[dones] = 0

The batch moments are reduced on the device by a fixed tree (bit-reproducible); with ``reduce="global"`` and
torch.distributed initialised, ranks exchange their 24-byte moment records (one all-gather per step) so that the
statistics do not depend on how the envs are sharded.  ``reduce="rank"`` (default) is the reference's situation:
one VecNormalize per training process.

SB3 is not installed in the build container: these semantics are restated from SB3 2.0.0a13 (the reference's pin,
``requirements.txt``) and checked by the test-suite against a plain numpy restatement -- parity unpinned.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .vec_env import PtGVecEnv


class RunningMeanStdView:
    """Read-only view of the device statistics with SB3's ``RunningMeanStd`` attribute names."""

    def __init__(self, owner: "VecNormalizeReward"):
        self._o = owner

    def _get(self):
        return self._o._st[self._o._cur].cpu().numpy()

    @property
    def mean(self) -> float:
        return float(self._get()[0])

    @property
    def var(self) -> float:
        return float(self._get()[1])

    @property
    def count(self) -> float:
        return float(self._get()[2])


class VecNormalizeReward:
    def __init__(self, venv: PtGVecEnv, training: bool = True, norm_reward: bool = True, clip_reward: float = 10.0,
                 gamma: float = 0.99, epsilon: float = 1e-8, reduce: str = "rank"):
        if reduce not in ("rank", "global"):
            raise ValueError("reduce must be 'rank' or 'global'")
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space, self.action_space = venv.observation_space, venv.action_space
        self.training, self.norm_reward = training, norm_reward
        self.norm_obs = False                   # the reference never normalises observations here (they already are)
        self.clip_reward, self.gamma, self.epsilon = float(clip_reward), float(gamma), float(epsilon)
        self.reduce = reduce
        dev = venv.device
        self.device = dev
        self._L = _lib.load()
        # RunningMeanStd(shape=()) starts at mean 0, var 1, count 1e-4 (SB3 running_mean_std.py)
        self._st = [torch.tensor([0.0, 1.0, 1e-4, 0.0], dtype=torch.float64, device=dev) for _ in range(2)]
        self._cur = 0
        self.returns = torch.zeros(self.num_envs, dtype=torch.float64, device=dev)
        self._moments = torch.zeros(3, dtype=torch.float64, device=dev)
        self._norm_reward = torch.zeros(self.num_envs, dtype=torch.float32, device=dev)
        self.old_reward = venv._reward          # SB3: get_original_reward() returns the unnormalised rewards
        self.ret_rms = RunningMeanStdView(self)

    # -- device API ------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def normalize_step(self, reward: torch.Tensor, done: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """The VecNormalize part of ``step_wait`` for one batch of (reward, done) CUDA tensors."""
        if not self.norm_reward:
            return reward
        out = self._norm_reward if out is None else out
        h, p = self.venv._h, PtGVecEnv._ptr
        moments, n_batch = self._moments, 1
        if self.training:
            _lib.check(self._L.ptg_vecnorm_moments(h, p(reward), p(self.returns), self.gamma, p(self._st[self._cur]),
                                                   p(self._moments), self._stream()))
            if self.reduce == "global":
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    gathered = torch.empty((dist.get_world_size(), 3), dtype=torch.float64, device=self.device)
                    dist.all_gather_into_tensor(gathered, self._moments)
                    moments, n_batch = gathered, gathered.shape[0]
        st_in, st_out = self._st[self._cur], self._st[self._cur ^ 1]
        _lib.check(self._L.ptg_vecnorm_apply(h, p(reward), p(done), p(self.returns), p(st_in), p(st_out), p(moments),
                                             n_batch, int(self.training), self.epsilon, self.clip_reward, p(out),
                                             self._stream()))
        self._cur ^= 1
        return out

    def step_tensor(self, actions: torch.Tensor):
        obs, reward, done = self.venv.step_tensor(actions)
        return obs, self.normalize_step(reward, done), done

    def reset_tensor(self, **kw):
        self.returns.zero_()
        return self.venv.reset_tensor(**kw)

    # -- SB3 numpy API ---------------------------------------------------------------------------------------
    def reset(self):
        self.returns.zero_()
        return self.venv.reset()

    def step_async(self, actions) -> None:
        self.venv.step_async(actions)

    def step_wait(self):
        norm = self.normalize_step(self.venv._reward, self.venv._done)     # queued behind the step kernel
        obs, rewards, dones, infos = self.venv.step_wait()
        self._old_reward_np = rewards
        if self.norm_reward:
            rewards = norm.cpu().numpy()
        return obs, rewards, dones, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def get_original_reward(self):
        return self._old_reward_np.copy()

    def get_original_obs(self):
        raise NotImplementedError("observations are not normalised by this wrapper (norm_obs=False)")

    def normalize_reward(self, reward: np.ndarray) -> np.ndarray:
        """SB3 ``VecNormalize.normalize_reward`` with the current statistics (host, no update)."""
        if not self.norm_reward:
            return reward
        return np.clip(reward / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)

    def close(self):
        self.venv.close()

    def __getattr__(self, name):          # VecEnvWrapper: everything else is the wrapped env's
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.venv, name)


# ---------------------------------------------------------------------------------------------------------------
# flat policy features, GAE
# ---------------------------------------------------------------------------------------------------------------
def feature_dim(env: PtGVecEnv) -> int:
    return int(_lib.load().ptg_features_dim(env._h))


def feature_names(env: PtGVecEnv) -> list[str]:
    """Column names of ``features_tensor`` (gymnasium Dict order = sorted keys; METH_STATUS one-hot(6))."""
    pa = env.cfg.price_ahead
    one_hot = [f"METH_STATUS={s}" for s in range(6)]
    if env.raw_modified == "mod":
        return (["CH4_syn_MolarFlow", "Elec_Heating", "H2O_DE_MassFlow", "H2_in_MolarFlow", "H2_res_MolarFlow"] + one_hot
                + [f"Part_Full[{a}]" for a in range(pa)] + [f"Pot_Reward[{a}]" for a in range(pa)]
                + ["T_CAT", "Temp_hour_enc_cos", "Temp_hour_enc_sin"])
    return (["CH4_syn_MolarFlow", "EUA_Price[0]", "EUA_Price[1]", "Elec_Heating"] + [f"Elec_Price[{a}]" for a in range(pa)]
            + ["Gas_Price[0]", "Gas_Price[1]", "H2O_DE_MassFlow", "H2_in_MolarFlow", "H2_res_MolarFlow"] + one_hot
            + ["T_CAT", "Temp_hour_enc_cos", "Temp_hour_enc_sin"])


def features_tensor(env: PtGVecEnv, obs_buffer: torch.Tensor | None = None, out: torch.Tensor | None = None):
    """[n_envs, F] fp32 feature rows of the env's current observation (or of ``obs_buffer``, a flat obs buffer with
    the env's layout, e.g. ``rollout['obs'][t]``)."""
    if env.obs_layout == "flat":                      # the step kernel already wrote the rows: zero-copy view
        view = env.features_view(obs_buffer)
        if out is None or out.data_ptr() == view.data_ptr():
            return view
        out.copy_(view)
        return out
    L = _lib.load()
    F = feature_dim(env)
    src = env._obs if obs_buffer is None else obs_buffer
    if out is None:
        out = torch.empty((env.num_envs, F), dtype=torch.float32, device=env.device)
    stream = C.c_void_p(torch.cuda.current_stream(env.device).cuda_stream)
    _lib.check(L.ptg_features(env._h, PtGVecEnv._ptr(src), PtGVecEnv._ptr(out), stream))
    return out


def gae(rewards: torch.Tensor, values: torch.Tensor, episode_starts: torch.Tensor, last_values: torch.Tensor,
        last_dones: torch.Tensor, gamma: float, gae_lambda: float, advantages: torch.Tensor | None = None,
        returns: torch.Tensor | None = None):
    """SB3 ``RolloutBuffer.compute_returns_and_advantage`` on CUDA tensors: rewards/values [T, n] fp32,
    episode_starts [T, n] uint8, last_values [n] fp32, last_dones [n] uint8 -> (advantages, returns) [T, n] fp32."""
    T, n = rewards.shape
    for t in (rewards, values, episode_starts, last_values, last_dones):
        if not t.is_cuda or not t.is_contiguous():
            raise ValueError("gae() needs contiguous CUDA tensors")
    if values.shape != (T, n) or episode_starts.shape != (T, n) or last_values.numel() != n or last_dones.numel() != n:
        raise ValueError("shape mismatch")
    if rewards.dtype != torch.float32 or values.dtype != torch.float32 or last_values.dtype != torch.float32:
        raise ValueError("rewards / values / last_values must be float32")
    es = episode_starts.view(torch.uint8) if episode_starts.dtype == torch.bool else episode_starts
    ld = last_dones.view(torch.uint8) if last_dones.dtype == torch.bool else last_dones
    if es.dtype != torch.uint8 or ld.dtype != torch.uint8:
        raise ValueError("episode_starts / last_dones must be uint8 or bool")
    advantages = torch.empty_like(rewards) if advantages is None else advantages
    returns = torch.empty_like(rewards) if returns is None else returns
    p = PtGVecEnv._ptr
    stream = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
    _lib.check(_lib.load().ptg_gae(n, T, p(rewards), p(values), p(es), p(last_values), p(ld), float(gamma),
                                   float(gae_lambda), p(advantages), p(returns), stream))
    return advantages, returns
