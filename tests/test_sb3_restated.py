"""CPU checks of the numpy restatement of the SB3 pieces (oracle/sb3_restated.py) against independent formulations:
the restatement is test infrastructure and "parity unpinned" (SB3 is not installed), so at least its algebra is
pinned here."""
import numpy as np

from oracle.sb3_restated import RunningMeanStd, VecNormalizeRewardRef, combined_extractor_ref, gae_ref


def test_running_mean_std_equals_pooled_moments():
    rng = np.random.default_rng(0)
    rms = RunningMeanStd(epsilon=1e-4)
    chunks = [rng.normal(3, 2, size=rng.integers(1, 50)) for _ in range(30)]
    for c in chunks:
        rms.update(c)
    allv = np.concatenate(chunks)
    # the epsilon pseudo-count (mean 0, var 1) is part of SB3's definition: fold it in analytically
    n, eps = allv.size, 1e-4
    mean = allv.sum() / (n + eps)
    var = (eps * (1.0 + mean ** 2) + np.sum((allv - mean) ** 2)) / (n + eps)
    assert np.isclose(rms.mean, mean, rtol=1e-12)
    assert np.isclose(rms.var, var, rtol=1e-10)
    assert np.isclose(rms.count, n + eps)


def test_vecnormalize_reward_path():
    vn = VecNormalizeRewardRef(4, gamma=0.5, clip_reward=1.5)
    r = np.array([1.0, -2.0, 0.0, 4.0], dtype=np.float32)
    out = vn.step(r, np.array([False, True, False, False]))
    assert np.array_equal(vn.returns, [1.0, 0.0, 0.0, 4.0])             # returns[dones] = 0 after the update
    assert np.all(np.abs(out) <= 1.5) and out[2] == 0.0
    out2 = vn.step(r, np.zeros(4, bool))
    assert np.allclose(vn.returns, [1.5, -2.0, 0.0, 6.0])
    assert np.allclose(out2, np.clip(r / np.sqrt(vn.ret_rms.var + 1e-8), -1.5, 1.5))


def test_gae_against_plain_recursion():
    rng = np.random.default_rng(1)
    T, n, gamma, lam = 17, 9, 0.973, 0.8002
    rewards = rng.normal(size=(T, n)).astype(np.float32)
    values = rng.normal(size=(T, n)).astype(np.float32)
    starts = (rng.random((T, n)) < 0.2).astype(np.float32)
    last_values = rng.normal(size=n).astype(np.float32)
    dones = rng.random(n) < 0.3
    adv, ret = gae_ref(rewards, values, starts, last_values, dones, gamma, lam)
    want = np.zeros((T, n))
    for e in range(n):
        last = 0.0
        for t in reversed(range(T)):
            nnt = 1.0 - (float(dones[e]) if t == T - 1 else float(starts[t + 1, e]))
            nv = float(last_values[e]) if t == T - 1 else float(values[t + 1, e])
            delta = float(rewards[t, e]) + gamma * nv * nnt - float(values[t, e])
            last = delta + gamma * lam * nnt * last
            want[t, e] = last
    assert np.allclose(adv, want, rtol=2e-5, atol=2e-6)
    assert np.allclose(ret, want + values, rtol=2e-5, atol=2e-6)
    assert adv.dtype == np.float32 and ret.dtype == np.float32


def test_combined_extractor_order_and_one_hot():
    obs = {"T_CAT": np.array([[0.5], [0.25]]), "METH_STATUS": np.array([1, 4]),
           "Part_Full": np.array([[1., 0.], [-1., 1.]]), "CH4_syn_MolarFlow": np.array([[0.1], [0.2]])}
    f = combined_extractor_ref(obs)
    # sorted keys: CH4_syn_MolarFlow, METH_STATUS(one-hot 6), Part_Full, T_CAT
    assert f.shape == (2, 1 + 6 + 2 + 1) and f.dtype == np.float32
    assert np.array_equal(f[0], np.float32([0.1, 0, 1, 0, 0, 0, 0, 1, 0, 0.5]))
    assert np.array_equal(f[1], np.float32([0.2, 0, 0, 0, 0, 1, 0, -1, 1, 0.25]))
