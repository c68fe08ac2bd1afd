"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the Stable-Baselines3 pieces the reference puts either side of
the env path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

PARITY UNPINNED: stable_baselines3 / gymnasium are not installed in the build container (no network), so these are
restated from the algorithm of SB3 2.0.0a13 (the reference's pin, requirements.txt:9) -- not validated against the
package itself.  Call sites in the reference: src/rl_utils.py:453,491 (VecNormalize(norm_obs=False)),
src/rl_config_agent.py:132-149 (PPO("MultiInputPolicy", gamma, gae_lambda, ...)).

  RunningMeanStd            stable_baselines3/common/running_mean_std.py
  VecNormalizeRewardRef     stable_baselines3/common/vec_env/vec_normalize.py (step_wait, reward path only)
  gae_ref                   stable_baselines3/common/buffers.py RolloutBuffer.compute_returns_and_advantage
  combined_extractor_ref    stable_baselines3/common/torch_layers.py CombinedExtractor + preprocessing.preprocess_obs
                            (Dict keys in gymnasium 0.28's sorted order, Discrete -> one-hot, Box -> flatten)
"""
from __future__ import annotations

import numpy as np


class RunningMeanStd:
    def __init__(self, epsilon: float = 1e-4, shape=()):
        self.mean = np.zeros(shape, np.float64)
        self.var = np.ones(shape, np.float64)
        self.count = epsilon

    def update(self, arr: np.ndarray) -> None:
        batch_mean = np.mean(arr, axis=0)
        batch_var = np.var(arr, axis=0)
        batch_count = arr.shape[0]
        self.update_from_moments(batch_mean, batch_var, batch_count)

    def update_from_moments(self, batch_mean, batch_var, batch_count) -> None:
        delta = batch_mean - self.mean
        tot_count = self.count + batch_count
        new_mean = self.mean + delta * batch_count / tot_count
        m_a = self.var * self.count
        m_b = batch_var * batch_count
        m_2 = m_a + m_b + np.square(delta) * self.count * batch_count / (self.count + batch_count)
        new_var = m_2 / (self.count + batch_count)
        self.mean, self.var, self.count = new_mean, new_var, tot_count


class VecNormalizeRewardRef:
    """VecNormalize(norm_obs=False, norm_reward=True): the reward path of step_wait()."""

    def __init__(self, num_envs: int, training=True, clip_reward=10.0, gamma=0.99, epsilon=1e-8):
        self.ret_rms = RunningMeanStd(shape=())
        self.returns = np.zeros(num_envs)
        self.training, self.clip_reward, self.gamma, self.epsilon = training, clip_reward, gamma, epsilon

    def step(self, rewards: np.ndarray, dones: np.ndarray) -> np.ndarray:
        if self.training:
            self.returns = self.returns * self.gamma + rewards
            self.ret_rms.update(self.returns)
        out = np.clip(rewards / np.sqrt(self.ret_rms.var + self.epsilon), -self.clip_reward, self.clip_reward)
        self.returns[dones] = 0
        return out


def gae_ref(rewards, values, episode_starts, last_values, dones, gamma, gae_lambda):
    """rewards/values/episode_starts: float32 [T, n]; last_values float32 [n]; dones bool [n]."""
    T = rewards.shape[0]
    advantages = np.zeros_like(rewards, dtype=np.float32)
    last_gae_lam = 0
    for step in reversed(range(T)):
        if step == T - 1:
            next_non_terminal = 1.0 - dones
            next_values = last_values
        else:
            next_non_terminal = 1.0 - episode_starts[step + 1]
            next_values = values[step + 1]
        delta = rewards[step] + gamma * next_values * next_non_terminal - values[step]
        last_gae_lam = delta + gamma * gae_lambda * next_non_terminal * last_gae_lam
        advantages[step] = last_gae_lam
    returns = advantages + values
    return advantages, returns


def combined_extractor_ref(obs: dict, n_status: int = 6) -> np.ndarray:
    """[n_envs, F] float32: keys in sorted order, METH_STATUS one-hot, everything else flattened."""
    cols = []
    for key in sorted(obs.keys()):
        v = np.asarray(obs[key])
        if key == "METH_STATUS":
            cols.append(np.eye(n_status, dtype=np.float32)[v.astype(np.int64).reshape(-1)])
        else:
            cols.append(v.reshape(v.shape[0], -1).astype(np.float32))
    return np.concatenate(cols, axis=1)
