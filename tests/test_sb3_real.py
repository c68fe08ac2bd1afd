"""Opportunistic pin of oracle/sb3_restated.py (and with it SURVEY.md 8(f) rows 1-2) against the REAL
stable_baselines3 -- skipped where the package is not installed (the build container and the GPU image have neither
stable_baselines3 nor gymnasium; there is no network).  Until one run of this file is green the f1/f2 rows stay
"parity unpinned".  The GPU kernels are compared with the restatement in test_gpu_train_side.py."""
import numpy as np
import pytest

sb3 = pytest.importorskip("stable_baselines3", reason="stable_baselines3 not installed: f1/f2 parity stays unpinned")


def test_running_mean_std_and_vecnormalize_reward_path():
    from stable_baselines3.common.running_mean_std import RunningMeanStd as RealRMS
    from oracle.sb3_restated import RunningMeanStd, VecNormalizeRewardRef
    rng = np.random.default_rng(0)
    a, b = RealRMS(shape=()), RunningMeanStd(shape=())
    ref = VecNormalizeRewardRef(64)
    returns = np.zeros(64)
    for _ in range(30):
        rew = rng.normal(0, 3, 64).astype(np.float32)
        done = rng.random(64) < 0.1
        # stable_baselines3/common/vec_env/vec_normalize.py: step_wait -> _update_reward -> normalize_reward
        returns = returns * 0.99 + rew
        a.update(returns)
        want = np.clip(rew / np.sqrt(a.var + 1e-8), -10.0, 10.0)
        returns[done] = 0
        got = ref.step(rew, done)
        b = ref.ret_rms
        assert np.array_equal(got, want)
        assert a.mean == b.mean and a.var == b.var and a.count == b.count


def test_gae_against_rollout_buffer():
    import torch as th
    from gymnasium import spaces
    from stable_baselines3.common.buffers import RolloutBuffer
    from oracle.sb3_restated import gae_ref
    T, n, gamma, lam = 37, 11, 0.973, 0.8002
    rng = np.random.default_rng(1)
    buf = RolloutBuffer(T, spaces.Box(-1, 1, (3,), np.float32), spaces.Discrete(5), device="cpu", gamma=gamma,
                        gae_lambda=lam, n_envs=n)
    rewards = rng.normal(0, 1, (T, n)).astype(np.float32)
    values = rng.normal(0, 2, (T, n)).astype(np.float32)
    starts = (rng.random((T, n)) < 0.05).astype(np.float32)
    for t in range(T):
        buf.add(np.zeros((n, 3), np.float32), np.zeros((n, 1)), rewards[t], starts[t], th.as_tensor(values[t]),
                th.zeros(n))
    last_values = rng.normal(0, 2, n).astype(np.float32)
    dones = rng.random(n) < 0.1
    buf.compute_returns_and_advantage(th.as_tensor(last_values), dones)
    adv, ret = gae_ref(rewards, values, starts, last_values, dones, gamma, lam)
    assert np.array_equal(buf.advantages, adv) and np.array_equal(buf.returns, ret)


def test_combined_extractor_feature_order():
    import torch as th
    from gymnasium import spaces
    from stable_baselines3.common.preprocessing import preprocess_obs
    from stable_baselines3.common.torch_layers import CombinedExtractor
    from oracle.sb3_restated import combined_extractor_ref
    from rl_ptg_b200 import spaces as ptg_spaces
    pa, n = 13, 9
    space = ptg_spaces.observation_space("mod", pa)
    if not isinstance(space, spaces.Dict):           # (the local duck-typed spaces are used when gymnasium is absent)
        pytest.skip("needs gymnasium spaces")
    rng = np.random.default_rng(2)
    obs = {k: (rng.integers(0, 6, n) if k == "METH_STATUS" else rng.random((n,) + s.shape).astype(np.float32))
           for k, s in space.spaces.items()}
    ext = CombinedExtractor(space)
    t_obs = preprocess_obs({k: th.as_tensor(v) for k, v in obs.items()}, space)
    want = ext(t_obs).numpy()
    assert np.array_equal(combined_extractor_ref(obs), want)
