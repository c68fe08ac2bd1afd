"""-m gpu parity tests: the CUDA path (through the C ABI) against the golden vectors of the UNMODIFIED reference
and against the CPU oracle on identical actions / seeds.

Bar (BASELINE.json north_star): bit-exact plant state, indices, episode offsets and termination flags;
observations and rewards within 1e-5 relative in fp32 (tolerance written below); exact zeros stay exact."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, golden_kwargs, load_golden, real_kwargs, synthetic_kwargs

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5        # north_star tolerance for fp32 observations / rewards
ABS_ZERO = 0.0        # exact zeros must be exact


def assert_close_fp32(got, want, what):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    zero = want == 0.0
    assert np.all(got[zero] == 0.0), f"{what}: exact zeros differ"
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    # values whose magnitude is below fp32 resolution of O(1) quantities (e.g. sin(2*pi*n) ~ 1e-13 in the
    # reference) are compared absolutely
    tiny = np.abs(want) < 1e-9
    assert np.all(err[~tiny & ~zero] <= REL_TOL), f"{what}: rel err {err[~tiny & ~zero].max()}"
    assert np.all(np.abs(got[tiny]) < 1e-6), f"{what}: tiny values differ"


def flat_obs(obs: dict, keys) -> np.ndarray:
    return np.concatenate([np.asarray(obs[k], dtype=np.float64).reshape(len(obs[k]), -1) for k in keys], axis=1)


def state_ints(env, done):
    st = env.get_state()
    return np.stack([st["meth_state"], st["i"], st["j"], st["hot_cold"], np.asarray(done).astype(np.int32), st["k"],
                     st["act_ep_h"], st["act_ep_d"], st["partial_ds"], st["full_ds"]], axis=1)


def make_env(kw, n, **kwargs):
    from rl_ptg_b200.vec_env import PtGVecEnv
    return PtGVecEnv(kw, n, **kwargs)


@pytest.mark.parametrize("noise", ["tape", "numpy"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_cuda_matches_reference_golden(case, noise):
    """CUDA env vs vectors recorded from the unmodified reference; `numpy` = on-device PCG64+ziggurat."""
    from oracle.ptg_oracle import draw_noise_tape
    g = load_golden(case)
    m = g["meta"]
    kw = golden_kwargs(case)
    n, steps = m["n_envs"], m["steps"]
    env = make_env(kw, n, train_or_eval=m["mode"], noise=noise)
    if noise == "tape":
        env.set_noise_tape(draw_noise_tape(m["seed"] + np.arange(n), kw["noise"], steps))
    env.seed(m["seed"])
    obs = env.reset()
    keys = m["obs_keys"]
    assert list(obs.keys()) == keys
    assert_close_fp32(flat_obs(obs, keys), g["reset_obs"], "reset obs")
    if m["mode"] == "eval":
        got = np.array([[_info_val(d[k]) for k in d] for d in env.reset_infos])
        assert np.allclose(got, g["reset_info"], rtol=1e-12, atol=0)

    keep = {int(t): q for q, t in enumerate(g["obs_steps"])}
    term = {(int(t), int(e)): q for q, (t, e) in enumerate(g["term_steps"])}
    check_every = 1 if steps <= 3000 else 7
    for t in range(steps):
        obs, rew, done, infos = env.step(g["actions"][t])
        assert np.array_equal(done, g["ints"][t, :, 4].astype(bool)), f"done flags diverged at step {t}"
        assert_close_fp32(rew, g["rewards"][t], f"reward step {t}")
        if t % check_every == 0 or done.any() or t in keep:
            assert np.array_equal(state_ints(env, done), g["ints"][t]), f"integer state diverged at step {t}"
        if t in keep:
            assert_close_fp32(flat_obs(obs, keys), g["obs"][keep[t]], f"obs step {t}")
        if m["mode"] == "eval":
            got = np.array([[_info_val(d[k]) for k in list(d)[:24]] for d in infos])
            assert np.allclose(got, g["infos"][t], rtol=1e-9, atol=1e-12), f"info diverged at step {t}"
        for e in np.nonzero(done)[0]:
            q = term[(t, int(e))]
            tobs = infos[e]["terminal_observation"]
            assert_close_fp32(np.concatenate([np.atleast_1d(tobs[k]).astype(np.float64).ravel() for k in keys]),
                              g["term_obs"][q], "terminal obs")
            assert infos[e]["episode"]["l"] == kw["eps_sim_steps"] - 5
            assert infos[e]["episode"]["r"] == pytest.approx(g["episode_return"][q], abs=1e-6)   # Monitor rounds to 6 digits
            assert infos[e]["TimeLimit.truncated"] is False
    st = env.get_state()
    assert np.array_equal(g["ints"][-1][:, 0], st["meth_state"])
    env.close()


def _info_val(v):
    from rl_ptg_b200._abi import STATE_NAMES
    return float(STATE_NAMES.index(v)) if isinstance(v, str) else float(v)


@pytest.mark.parametrize("n_envs,overrides,steps", [
    (4096, dict(scenario=1, operation="OP1"), 400),      # BASELINE config 2
    (65536, dict(scenario=3, operation="OP2"), 60),      # BASELINE config 3 (CHP/EEG reward path)
])
def test_cuda_matches_oracle_at_baseline_sizes(n_envs, overrides, steps):
    """Bit-exact trajectory check vs the CPU oracle on real data, device RNG vs numpy tape."""
    import os
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    kw = real_kwargs(overrides)
    seeds = 3654 + np.arange(n_envs)
    tape = draw_noise_tape(seeds, kw["noise"], steps)
    ora = OracleVecEnv(kw, n_envs, noise_tape=tape, threads=len(os.sched_getaffinity(0)))
    env = make_env(kw, n_envs, seed=3654, noise="numpy")
    o_obs = ora.reset().copy()
    obs = env.reset()
    keys = list(obs.keys())
    assert_close_fp32(flat_obs(obs, keys), o_obs, "reset obs")
    rng = np.random.default_rng(1)
    for t in range(steps):
        a = rng.choice(5, size=n_envs, p=[0.15, 0.15, 0.3, 0.2, 0.2]).astype(np.int64)
        o_obs, o_rew, o_done = ora.step(a)
        obs, rew, done, _ = env.step(a)
        assert np.array_equal(done, o_done.astype(bool))
        assert_close_fp32(rew, o_rew, f"reward step {t}")
        if t % 10 == 0 or t == steps - 1:
            so, sg = ora.get_state(), env.get_state()
            for f in ("meth_state", "i", "j", "k", "hot_cold", "standby_ds", "startup_ds", "partial_ds", "full_ds",
                      "current_action", "act_ep_h", "act_ep_d", "episode_count", "draws"):
                assert np.array_equal(so[f], sg[f]), f"{f} diverged at step {t}"
            assert np.array_equal(so["t_cat"], sg["t_cat"])
            assert np.allclose(so["cum_reward"], sg["cum_reward"], rtol=1e-12, atol=1e-12)
            assert_close_fp32(flat_obs(obs, keys), o_obs, f"obs step {t}")
    env.close()
    ora.close()


def test_rollout_kernel_equals_single_steps():
    """ptg_step_many (T steps per launch, state in registers) == T x ptg_step, bit for bit."""
    import torch
    kw = synthetic_kwargs(dict(scenario=2, operation="OP2"))
    n, T = 5000, 64          # not a multiple of the block / warp size on purpose (ragged tail)
    a = torch.from_numpy(np.random.default_rng(3).integers(0, 5, size=(T, n))).cuda()
    e1 = make_env(kw, n, seed=11)
    e2 = make_env(kw, n, seed=11)
    e1.reset(); e2.reset()
    out = e2.rollout_tensor(a)
    for t in range(T):
        obs, rew, done = e1.step_tensor(a[t])
        assert torch.equal(e1._obs, out["obs"][t]), f"obs differ at t={t}"
        assert torch.equal(rew, out["reward"][t]) and torch.equal(done, out["done"][t])
    s1, s2 = e1.get_state(), e2.get_state()
    for f in s1:
        assert np.array_equal(s1[f], s2[f]), f
    e1.close(); e2.close()


def test_episode_boundaries_stats_and_state_roundtrip():
    """Short synthetic episodes: auto-reset order == DummyVecEnv order of the oracle, Monitor records, device
    statistics reduction, get_state/set_state round trip."""
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    kw = dict(synthetic_kwargs(dict(scenario=1, operation="OP2")))
    kw["eps_sim_steps"] = 40          # episodes of 35 steps
    n, steps = 300, 150
    seeds = 100 + np.arange(n)
    tape = draw_noise_tape(seeds, kw["noise"], steps)
    ora = OracleVecEnv(kw, n, noise_tape=tape)
    env = make_env(kw, n, seed=100)
    ora.reset(); env.reset()
    rng = np.random.default_rng(5)
    rets = []
    for t in range(steps):
        a = rng.integers(0, 5, size=n)
        o_obs, o_rew, o_done = ora.step(a)
        obs, rew, done, infos = env.step(a)
        assert np.array_equal(done, o_done.astype(bool))
        for e in np.nonzero(done)[0]:
            assert infos[e]["episode"]["l"] == 35
            assert infos[e]["episode"]["r"] == pytest.approx(ora.episode_return[e], abs=1e-6)   # Monitor rounds to 6 digits
            rets.append(ora.episode_return[e])
        so, sg = ora.get_state(), env.get_state()
        for f in ("meth_state", "i", "j", "k", "act_ep_h", "act_ep_d", "episode_count"):
            assert np.array_equal(so[f], sg[f]), f"{f} diverged at step {t}"
    stats = env.episode_stats(clear=True)
    assert stats["episodes"] == len(rets) == n * (steps // 35)
    assert stats["return_mean"] == pytest.approx(np.mean(rets), rel=1e-9)
    assert stats["return_min"] == pytest.approx(np.min(rets), rel=1e-12)
    assert stats["return_max"] == pytest.approx(np.max(rets), rel=1e-12)
    assert stats["length_mean"] == 35 and stats["env_steps"] == n * steps
    assert env.episode_stats()["episodes"] == 0
    # state round trip into a fresh env
    snap = env.get_state()
    env2 = make_env(kw, n, seed=100)
    env2.reset()
    env2.set_state(snap)
    back = env2.get_state()
    for f in snap:
        assert np.array_equal(snap[f], back[f]), f
    env.close(); env2.close(); ora.close()


def test_invalid_action_and_closed_env_fail_loudly():
    from rl_ptg_b200._lib import PtgError
    kw = synthetic_kwargs()
    env = make_env(kw, 64, seed=1)
    env.reset()
    bad = np.zeros(64, dtype=np.int64)
    bad[7] = 5
    with pytest.raises(PtgError) as ei:
        env.step(bad)
    assert ei.value.code == -4
    env.step(np.zeros(64, dtype=np.int64))      # error word was cleared; env still usable
    env.close()
    with pytest.raises(RuntimeError):
        env.step(np.zeros(64, dtype=np.int64))


def _oracle_features(obs: np.ndarray, pa: int) -> np.ndarray:
    """Oracle observation rows (reference key order) -> SB3 CombinedExtractor feature rows (sorted keys, one-hot
    METH_STATUS), the layout obs_layout="flat" writes."""
    n = obs.shape[0]
    f = np.zeros((n, 14 + 2 * pa))
    o_st = 2 * pa
    o_T, o_h2, o_ch4, o_h2res, o_h2o, o_heat, o_sin, o_cos = (2 * pa + 1 + q for q in range(8))
    f[:, 0], f[:, 1], f[:, 2], f[:, 3], f[:, 4] = obs[:, o_ch4], obs[:, o_heat], obs[:, o_h2o], obs[:, o_h2], obs[:, o_h2res]
    f[np.arange(n), 5 + obs[:, o_st].astype(np.int64)] = 1.0
    f[:, 11:11 + pa], f[:, 11 + pa:11 + 2 * pa] = obs[:, pa:2 * pa], obs[:, :pa]
    f[:, 11 + 2 * pa], f[:, 12 + 2 * pa], f[:, 13 + 2 * pa] = obs[:, o_T], obs[:, o_cos], obs[:, o_sin]
    return f


@pytest.mark.parametrize("layout", ["dict", "flat"])
def test_benchmarked_workload_matches_oracle(layout):
    """The workload bench.py times -- synthetic BS2/OP2 `mod`, uniform random discrete actions, seeds 3654 + i, numpy
    noise on the device -- at 65 536 envs against the CPU oracle: every step's rewards, dones and FULL observations
    across eight hour crossings, one episode end (eps_sim_steps shortened to 46), terminal observations and the
    auto-reset rows.  `flat`: the same against the oracle rows re-ordered to the feature layout (not against the
    dict layout)."""
    import torch
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    kw = dict(synthetic_kwargs(dict(scenario=2, operation="OP2")))
    kw["eps_sim_steps"] = 46                      # episodes of 41 steps
    n, steps, pa = 65536, 50, int(kw["price_ahead"])
    seeds = 3654 + np.arange(n)
    ora = OracleVecEnv(kw, n, noise_tape=draw_noise_tape(seeds, kw["noise"], steps), threads=len(os.sched_getaffinity(0)))
    env = make_env(kw, n, seed=3654, obs_layout=layout)
    o_obs = ora.reset().copy()
    rng = np.random.default_rng(0)
    if layout == "dict":
        obs = env.reset()
        keys = list(obs.keys())
        assert_close_fp32(flat_obs(obs, keys), o_obs, "reset obs")
    else:
        env.reset_tensor()
        assert_close_fp32(env.features_view().cpu().numpy(), _oracle_features(o_obs, pa), "reset rows")
    n_done = 0
    for t in range(steps):
        a = rng.integers(0, 5, size=n)
        o_obs, o_rew, o_done = ora.step(a)
        if layout == "dict":
            obs, rew, done, infos = env.step(a)
            got = flat_obs(obs, keys)
            want = o_obs
        else:
            _, rew_t, done_t = env.step_tensor(torch.as_tensor(a, device=env.device))
            rew, done = rew_t.cpu().numpy(), done_t.cpu().numpy().astype(bool)
            got, want = env.features_view().cpu().numpy(), _oracle_features(o_obs, pa)
        assert np.array_equal(done, o_done.astype(bool)), f"done step {t}"
        assert_close_fp32(rew, o_rew, f"reward step {t}")
        assert_close_fp32(got, want, f"obs step {t}")
        if o_done.any():
            n_done += int(o_done.sum())
            idx = np.nonzero(o_done)[0]
            if layout == "dict":
                term = np.stack([np.concatenate([np.atleast_1d(np.asarray(infos[e]["terminal_observation"][k], np.float64)).ravel()
                                                 for k in keys]) for e in idx[:64]])
                assert_close_fp32(term, ora.terminal_obs[idx[:64]], "terminal obs")
                assert infos[int(idx[0])]["episode"]["r"] == pytest.approx(ora.episode_return[idx[0]], abs=1e-6)
            else:
                term = env._term_obs[:n * env.feature_dim].view(n, env.feature_dim).cpu().numpy()
                assert_close_fp32(term[idx], _oracle_features(ora.terminal_obs[idx], pa), "terminal rows")
    assert n_done == n                            # every env ended exactly one episode
    so, sg = ora.get_state(), env.get_state()
    for f in ("meth_state", "i", "j", "k", "hot_cold", "partial_ds", "full_ds", "act_ep_h", "act_ep_d", "episode_count", "draws"):
        assert np.array_equal(so[f], sg[f]), f
    env.close(); ora.close()


@pytest.mark.parametrize("layout", ["dict", "flat"])
def test_sticky_policy_matches_oracle(layout):
    """A policy that holds its action for tens of steps (what a trained agent does): most warps consist of envs that ALL
    continue their current table, which the step kernels serve on a fast path (step-table entry requested before the
    transition, cont_entry / cont_advance).  8 192 envs x 260 steps against the CPU oracle with full observations:
    tables run to their ends (start-up hands over to partial load, the last row repeats), the load-change chains see
    long dwell times, one episode end (201 steps) with its auto-reset; single steps and the roll-out kernel."""
    import torch
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    kw = dict(synthetic_kwargs(dict(scenario=2, operation="OP2")))
    kw["eps_sim_steps"] = 206                     # episodes of 201 steps
    n, steps, pa = 8192, 260, int(kw["price_ahead"])
    seeds = 3654 + np.arange(n)
    rng = np.random.default_rng(7)
    # actions: every BLOCK of 64 envs (two warps) switches at its own random times, envs inside a block mostly agree on
    # when to hold (so that whole warps continue) but not on what: a new action per env at a switch
    acts = np.empty((steps, n), dtype=np.int64)
    cur = rng.integers(0, 5, size=n)
    next_switch = rng.integers(1, 60, size=n // 64)
    for t in range(steps):
        sw = np.repeat(next_switch == t, 64)
        lone = rng.random(n) < 0.002              # a few envs switch on their own (mixed warps)
        new = rng.integers(0, 5, size=n)
        cur = np.where(sw | lone, new, cur)
        next_switch = np.where(next_switch == t, t + rng.integers(5, 90, size=n // 64), next_switch)
        acts[t] = cur
    ora = OracleVecEnv(kw, n, noise_tape=draw_noise_tape(seeds, kw["noise"], steps), threads=len(os.sched_getaffinity(0)))
    env = make_env(kw, n, seed=3654, obs_layout=layout)
    twin = make_env(kw, n, seed=3654, obs_layout=layout)       # the same actions through the roll-out kernel
    ora.reset()
    if layout == "dict":
        obs = env.reset(); keys = list(obs.keys())
    else:
        env.reset_tensor()
    twin.reset_tensor()
    T = 20
    n_done = 0
    for t in range(steps):
        o_obs, o_rew, o_done = ora.step(acts[t])
        if layout == "dict":
            obs, rew, done, infos = env.step(acts[t])
            got, want = flat_obs(obs, keys), o_obs
        else:
            _, rew_t, done_t = env.step_tensor(torch.as_tensor(acts[t], device=env.device))
            rew, done = rew_t.cpu().numpy(), done_t.cpu().numpy().astype(bool)
            got, want = env.features_view().cpu().numpy(), _oracle_features(o_obs, pa)
        assert np.array_equal(done, o_done.astype(bool)), f"done step {t}"
        assert_close_fp32(rew, o_rew, f"reward step {t}")
        assert_close_fp32(got, want, f"obs step {t}")
        n_done += int(o_done.sum())
        if t % T == 0:                                         # roll-out twin: T steps in one launch
            out = twin.rollout_tensor(torch.as_tensor(acts[t:t + T], device=twin.device))
            roll_rew, roll_done = out["reward"].cpu().numpy(), out["done"].cpu().numpy().astype(bool)
        assert np.array_equal(roll_done[t % T], o_done.astype(bool)), f"roll-out done step {t}"
        assert_close_fp32(roll_rew[t % T], o_rew, f"roll-out reward step {t}")
    assert n_done == n
    so = ora.get_state()
    for e_ in (env, twin):
        sg = e_.get_state()
        for f in ("meth_state", "i", "j", "k", "hot_cold", "partial_ds", "full_ds", "act_ep_h", "act_ep_d", "episode_count", "draws"):
            assert np.array_equal(so[f], sg[f]), f
    env.close(); twin.close(); ora.close()
