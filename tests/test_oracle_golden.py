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


def load_single_env_golden():
    import json
    import os
    from helpers import GOLDEN_DIR
    with np.load(os.path.join(GOLDEN_DIR, "single_env_train_resets.npz"), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    g["meta"] = json.loads(str(g["meta"]))
    return g


def test_oracle_single_env_gymnasium_semantics_on_the_training_split():
    """ONE env, no auto-reset, three terminated -> reset() cycles on the training split: the constructor uses
    eps_ind[0], the resets [1], [2], [3] (env/ptg_gym_env.py:59-62, 490-493; golden from the unmodified reference)."""
    from helpers import real_kwargs
    g = load_single_env_golden()
    m = g["meta"]
    kw = real_kwargs(m["overrides"], m["split"], m["action_type"], m["seed_train"])
    assert np.array_equal(kw["eps_ind"][:8], g["eps_ind_head"])
    tape = draw_noise_tape([m["seed"]], kw["noise"], len(g["actions"]))
    env = OracleVecEnv(kw, 1, noise_tape=tape)
    st = env.get_state()
    assert (st["act_ep_h"][0], st["act_ep_d"][0]) == tuple(g["offsets"][0])
    keep = {int(t): q for q, t in enumerate(g["obs_steps"])}
    t = 0
    for ep in range(m["episodes"]):
        obs = env.reset()
        st = env.get_state()
        assert (st["act_ep_h"][0], st["act_ep_d"][0]) == tuple(g["offsets"][ep + 1]), f"episode {ep} schedule"
        assert np.array_equal(obs[0], g["reset_obs"][ep])
        assert np.allclose(env.info[0], g["reset_info"][ep], rtol=FP64_TOL, atol=0)
        for _ in range(m["ep_len"]):
            obs, rew, done = env.step(g["actions"][t:t + 1].astype(np.int64), auto_reset=False)
            st = env.get_state()
            got = (st["meth_state"][0], st["i"][0], st["j"][0], st["hot_cold"][0], int(done[0]), st["k"][0])
            assert got == tuple(g["ints"][t]), f"integer state diverged at step {t}"
            assert rew[0] == pytest.approx(g["rewards"][t], rel=FP64_TOL, abs=0)
            if t in keep:
                assert np.allclose(obs[0], g["obs"][keep[t]], rtol=FP64_TOL, atol=1e-15)
            t += 1
        assert done[0] == 1 and np.allclose(obs[0], g["term_obs"][ep], rtol=FP64_TOL, atol=1e-15)
