"""Pins the CPU oracle (oracle/ptg_oracle.c) to vectors recorded from the UNMODIFIED reference env.

The oracle gets its kwargs from the product's preprocessing on the real reference data and its noise from a
numpy tape -- so this also pins preprocessing + tape semantics end to end.
"""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, golden_kwargs, load_golden
from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape

FP64_TOL = 1e-12   # the oracle restates the reference op for op; observed difference is exactly 0


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_oracle_matches_reference(case):
    g = load_golden(case)
    m = g["meta"]
    kw = golden_kwargs(case)
    n, steps = m["n_envs"], m["steps"]
    tape = draw_noise_tape(m["seed"] + np.arange(n), kw["noise"], steps)
    env = OracleVecEnv(kw, n, train_or_eval=m["mode"], noise_tape=tape)
    obs = env.reset()
    assert np.array_equal(obs, g["reset_obs"])
    assert np.allclose(env.info, g["reset_info"], rtol=FP64_TOL, atol=0)

    keep = {int(t): q for q, t in enumerate(g["obs_steps"])}
    term = {(int(t), int(e)): q for q, (t, e) in enumerate(g["term_steps"])}
    actions = g["actions"]
    for t in range(steps):
        obs, rew, done = env.step(actions[t])
        st = env.get_state()
        got = np.stack([st["meth_state"], st["i"], st["j"], st["hot_cold"], done.astype(np.int32), st["k"],
                        st["act_ep_h"], st["act_ep_d"], st["partial_ds"], st["full_ds"]], axis=1)
        assert np.array_equal(got, g["ints"][t]), f"integer state diverged at step {t}"
        assert np.allclose(rew, g["rewards"][t], rtol=FP64_TOL, atol=0), f"reward diverged at step {t}"
        if t in keep:
            assert np.allclose(obs, g["obs"][keep[t]], rtol=FP64_TOL, atol=1e-15), f"obs diverged at step {t}"
        if m["mode"] == "eval":
            assert np.allclose(env.info, g["infos"][t], rtol=FP64_TOL, atol=1e-15), f"info diverged at step {t}"
        for e in np.nonzero(done)[0]:
            q = term[(t, int(e))]
            assert np.allclose(env.terminal_obs[e], g["term_obs"][q], rtol=FP64_TOL, atol=1e-15)
            assert env.episode_return[e] == pytest.approx(g["episode_return"][q], rel=1e-12)
    assert len(term) == int(g["ints"][:, :, 4].sum())


def test_zero_reward_in_cold_cooldown_is_exact():
    g = load_golden("bs2_op2_mod")
    assert g["rewards"][0, 0] == 0.0 and g["rewards"][1, 0] == 0.0
