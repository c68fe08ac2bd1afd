"""SURVEY.md A.9 "quirk checklist (each needs a test)" -- one test per quirk of the reference's step(), asserted on
vectors recorded from the UNMODIFIED reference (tests/golden/gen_golden.py, cases ``quirks_s300`` and ``quirks_s50``:
closed-loop scripts drive the reference env to each quirk and leave a (label, step) mark per env in the fixture).

What each test does: (1) states the quirk as an assertion on the REFERENCE's recorded integers / observations, so the
fixture provably contains the situation; (2) where the quirk involves something the fixture does not record (the bound
start-up / standby table, the catalyst temperature, the draw counter), replays the tape through the CPU oracle -- which
``test_oracle_golden.py`` pins to the same fixture step by step -- and asserts it there.  The CUDA path meets the same
two fixtures bit for bit in ``test_gpu_parity.py::test_cuda_matches_reference_golden`` (both noise modes).

Reference lines: env/ptg_gym_env.py (the number in each docstring is the item of SURVEY.md A.9)."""
import functools

import numpy as np
import pytest

from helpers import golden_kwargs, load_golden, real_kwargs
from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
from rl_ptg_b200 import _abi

DS = {name: idx for idx, name in enumerate(_abi.DATASET_NAMES)}
STATE, I, J, HOT, DONE, K, EP_H, EP_D, PART, FULL = range(10)          # columns of the fixture's `ints`
STANDBY, COOLDOWN, STARTUP, PARTIAL, FULL_LOAD = range(5)
# A.2: bit (5 * action + state) set <=> the step continues the current table (no handler, no noise)
CONT = {(0, 0), (1, 1), (2, 2), (2, 3), (2, 4), (3, 0), (3, 1), (3, 2), (3, 3), (4, 0), (4, 1), (4, 2), (4, 4)}


def marks_of(case):
    """[{label: [steps]} per env]"""
    out = []
    for marks in load_golden(case)["meta"]["marks"]:
        d = {}
        for label, t in marks:
            d.setdefault(label, []).append(int(t))
        out.append(d)
    return out


@functools.lru_cache(maxsize=None)
def oracle_replay(case):
    """Replays the fixture's tape through the oracle; per step: the snapshot AFTER the step (incl. t_cat, draws,
    standby_ds / startup_ds, which the fixture does not record).  Integers are checked against the fixture on the way."""
    g = load_golden(case)
    m = g["meta"]
    kw = golden_kwargs(case)
    n, steps = m["n_envs"], m["steps"]
    tape = draw_noise_tape(m["seed"] + np.arange(n), kw["noise"], steps)
    env = OracleVecEnv(kw, n, train_or_eval=m["mode"], noise_tape=tape)
    env.reset()
    snaps = [env.get_state()]                      # snaps[t] = state BEFORE step t; snaps[t + 1] = after it
    for t in range(steps):
        _, _, done = env.step(g["actions"][t])
        st = env.get_state()
        got = np.stack([st["meth_state"], st["i"], st["j"], st["hot_cold"]], axis=1)
        assert np.array_equal(got, g["ints"][t][:, :4]), f"oracle left the reference's trajectory at step {t}"
        snaps.append(st)
    return g, kw, tape, snaps, env


def op_table(kw, ds_name):
    """The [rows, 7] array of an operation table in the env kwargs (whatever prefix the kwargs use)."""
    for k, v in kw.items():
        if isinstance(v, np.ndarray) and v.ndim == 2 and v.shape[1] == 7 and k.endswith(ds_name):
            return v
    raise KeyError(ds_name)


def test_q1_reset_returns_full_info_even_in_train_mode():
    """(1) reset() returns the 24-field info in train mode (:503-506); the fields show the reset state."""
    g = load_golden("quirks_s300")
    info = g["reset_info"]
    assert info.shape == (3, 24)
    assert np.all(info[:, 0] == 0)                           # step
    assert np.all(info[:, 4] == COOLDOWN)                    # Meth_State
    assert np.all(info[:, 7] == 16.0)                        # Meth_T_cat (:117)
    assert np.all(info[:, 12:22] == 0.0)                     # no reward constituents yet
    assert np.all(info[:, 1] != 0.0)                         # ... but the market data of the episode start is there


def test_q2_first_cont_cooldown_after_reset_reads_rows_i_to_i_plus_S():
    """(2) after reset j == 0; the first cont(cooldown) makes it 1 and simulates rows [i, i + S) of the cooldown table,
    i = argmin |cooldown.T - 16| (:118)."""
    g, kw, _, snaps, env = oracle_replay("quirks_s300")
    cooldown = op_table(kw, "cooldown")
    i0 = int(np.abs(cooldown[:, 1] - 16.0).argmin())
    assert i0 == 41218
    S = kw["sim_step"] // kw["time_step_op"]
    for e, marks in enumerate(marks_of("quirks_s300")):
        (t,) = marks["first_cont_cooldown"]
        assert t == 0 and g["actions"][t, e] == COOLDOWN
        assert snaps[0]["j"][e] == 0 and snaps[0]["i"][e] == i0
        assert tuple(g["ints"][t, e][[STATE, I, J]]) == (COOLDOWN, i0, 1)
        assert snaps[1]["t_cat"][e] == cooldown[i0 + S - 1, 1]            # last row of the window [i, i + S)


def test_q3_time_op_equal_to_a_threshold_falls_through_to_the_final_else():
    """(3) the chains use strict `<` on both sides (:656-683): time_op == time2_p_f_p matches neither
    `time1 < t < time2` nor `time2 < t < time_p_f`, nor any later branch -> op8_f_p, i = 0, j = 1."""
    g = load_golden("quirks_s50")
    kw = golden_kwargs("quirks_s50")
    S = kw["sim_step"] // kw["time_step_op"]
    assert S == 50 and kw["time2_p_f_p"] == 150
    for e, marks in enumerate(marks_of("quirks_s50")):
        (t,) = marks["partial_at_threshold"]
        pre, post = g["ints"][t - 1, e], g["ints"][t, e]
        assert g["actions"][t, e] == PARTIAL and pre[STATE] == FULL_LOAD and pre[FULL] == DS["op3_p_f"]
        assert pre[I] + pre[J] * S == kw["time2_p_f_p"]
        assert tuple(post[[STATE, I, J, PART]]) == (PARTIAL, 0, 1, DS["op8_f_p"])


def test_q4_j_plus_one_branch_keeps_i_of_the_previous_table():
    """(4) time1_p_f_p < time_op < time2_p_f_p (:656-659): partial := op4_p_f_p_5, j += 1, i is NOT reset."""
    g = load_golden("quirks_s50")
    kw = golden_kwargs("quirks_s50")
    S = kw["sim_step"] // kw["time_step_op"]
    for e, marks in enumerate(marks_of("quirks_s50")):
        (t,) = marks["partial_keep_i"]
        pre, post = g["ints"][t - 1, e], g["ints"][t, e]
        assert pre[STATE] == FULL_LOAD and pre[FULL] == DS["op3_p_f"]
        assert kw["time1_p_f_p"] < pre[I] + pre[J] * S < kw["time2_p_f_p"]
        assert tuple(post[[STATE, I, J, PART]]) == (PARTIAL, pre[I], pre[J] + 1, DS["op4_p_f_p_5"])


def test_q5_fully_developed_indices_point_past_the_table_end():
    """(5) time_op < time1_p_f_p (:650-654) / < time1_f_p_f (:716-720): i, j := 12 000, 100 -- far past the end of
    op8_f_p / op3_p_f -- so every following step is S copies of the table's last row, and the T_cat assignment
    next to it is overwritten by the window's last row (the same value)."""
    g, kw, _, snaps, _ = oracle_replay("quirks_s50")
    for label, tab, col in (("partial_fully_developed", "op8_f_p", PART), ("full_fully_developed", "op3_p_f", FULL)):
        last = op_table(kw, tab)[-1]
        assert kw["i_fully_developed"] > len(op_table(kw, tab))
        for e, marks in enumerate(marks_of("quirks_s50")):
            (t,) = marks[label]
            post = g["ints"][t, e]
            assert tuple(post[[I, J]]) == (kw["i_fully_developed"], kw["j_fully_developed"]) and post[col] == DS[tab]
            assert snaps[t + 1]["t_cat"][e] == last[1]
            # observation = the constant last row: plant scalars (T, H2, CH4, H2_res, H2O, heating) repeat exactly
            keys = g["meta"]["obs_keys"]
            q = {int(s): n for n, s in enumerate(g["obs_steps"])}
            sc = slice(-9 + 1, -2) if keys[-2:] == ["Temp_hour_enc_sin", "Temp_hour_enc_cos"] else None
            assert sc is not None
            if label == "partial_fully_developed":             # the script stays two more steps on that row
                assert np.array_equal(g["obs"][q[t], e][sc], g["obs"][q[t + 1], e][sc])


def test_q6_noise_is_drawn_only_by_standby_cooldown_startup_handlers():
    """(6) one np_random.normal draw per _standby/_cooldown/_startup call (:585, :599, :621) and none anywhere else --
    in particular not for the argmin of _partial (:641)."""
    for case in ("quirks_s300", "quirks_s50"):
        g, kw, _, snaps, env = oracle_replay(case)
        n = g["meta"]["n_envs"]
        expect = np.zeros(n, dtype=np.int64)
        for t in range(g["meta"]["steps"]):
            for e in range(n):
                a, s = int(g["actions"][t, e]), int(snaps[t]["meth_state"][e])
                expect[e] += (a <= STARTUP) and ((a, s) not in CONT)
            assert np.array_equal(snaps[t + 1]["draws"], expect), f"{case}: draw counters at step {t}"
        assert expect.min() >= 5
    g, kw, _, snaps, env = oracle_replay("quirks_s50")
    op1 = op_table(kw, "op1_start_p")
    for e, marks in enumerate(marks_of("quirks_s50")):
        (t,) = marks["partial_argmin_no_noise"]
        pre, post = g["ints"][t - 1, e], g["ints"][t, e]
        assert pre[STATE] == FULL_LOAD and pre[FULL] == DS["op2_start_f"]
        assert pre[I] + pre[J] * 50 < kw["time2_start_f_p"]
        assert snaps[t + 1]["draws"][e] == snaps[t]["draws"][e]
        want = int(np.abs(op1[:, 1] - snaps[t]["t_cat"][e]).argmin())
        assert tuple(post[[STATE, I, J, PART]]) == (PARTIAL, want, 1, DS["op1_start_p"])


def test_q7_hot_cold_hysteresis():
    """(7) hot_cold := 0 at T <= 160, := 1 at T >= 350, unchanged in between (:339-342; evaluated on the temperature
    BEFORE the step); reset gives 0.  _startup binds startup_hot / startup_cold by that flag, not by T (:616-619)."""
    g, kw, _, snaps, _ = oracle_replay("quirks_s300")
    lo, hi = kw["t_cat_startup_cold"], kw["t_cat_startup_hot"]
    assert (lo, hi) == (160, 350)
    n = g["meta"]["n_envs"]
    for t in range(g["meta"]["steps"]):
        T, before, after = snaps[t]["t_cat"], snaps[t]["hot_cold"], g["ints"][t, :, HOT]
        want = np.where(T <= lo, 0, np.where(T >= hi, 1, before))
        assert np.array_equal(after, want), f"step {t}"
    assert np.all(snaps[0]["hot_cold"] == 0)
    for e, marks in enumerate(marks_of("quirks_s300")):
        (t,) = marks["startup_in_hysteresis_band"]          # cooled down from operation: still "hot" at 160 < T < 350
        assert lo < snaps[t]["t_cat"][e] < hi and g["ints"][t, e, HOT] == 1
        assert snaps[t + 1]["startup_ds"][e] == DS["startup_hot"]
        (t,) = marks["startup_cold_again"]                  # warmed up from below 160: still "cold" in the same band
        assert lo < snaps[t]["t_cat"][e] < hi and g["ints"][t, e, HOT] == 0
        assert snaps[t + 1]["startup_ds"][e] == DS["startup_cold"]
        (t,) = marks["standby_from_hot"]                    # T > t_cat_standby: standby_down; below: standby_up (:579-582)
        assert snaps[t]["t_cat"][e] > kw["t_cat_standby"] and snaps[t + 1]["standby_ds"][e] == DS["standby_down"]
        (t,) = marks["standby_below_188"]
        assert snaps[t]["t_cat"][e] <= kw["t_cat_standby"] and snaps[t + 1]["standby_ds"][e] == DS["standby_up"]


def test_q8_argmin_returns_the_first_minimum():
    """(8) np.abs(..).argmin() returns the FIRST index of the minimum; the temperature columns are non-monotone with
    many repeats (startup_hot: ~1 250 distinct values in 2 048 rows).  The oracle scans literally (the CUDA path looks
    the answer up in a table built with the same tie rule, k_build_argmin)."""
    kw = real_kwargs(dict(scenario=2, operation="OP2"))
    env = OracleVecEnv(kw, 1)
    hot = op_table(kw, "startup_hot")[:, 1]
    assert len(np.unique(hot)) < 0.7 * len(hot)
    rng = np.random.default_rng(0)
    for name in ("cooldown", "standby_up", "standby_down", "startup_cold", "startup_hot", "op1_start_p"):
        T = op_table(kw, name)[:, 1]
        probes = np.concatenate([rng.choice(T, size=60), np.round(rng.uniform(0, 600, size=60), 1), [16.0, 15.2, 160.1]])
        for t in probes:
            want = int(np.abs(T - t).argmin())
            assert env.get_index(DS[name], float(t)) == want
            hits = np.nonzero(np.abs(T - t) == np.abs(T - t).min())[0]
            assert want == hits[0]
    dup = op_table(kw, "standby_up")[:, 1]
    assert np.all(dup[:5] == 15.2) and env.get_index(DS["standby_up"], 15.2) == 0
    env.close()


def test_q9_reward_and_info_use_the_market_data_of_the_new_clock():
    """(9) step() advances the clock first (:442-447), so reward and info of step k carry the prices of hour
    floor((k + 1) * sim_step / 3600) and of day floor(hours / 24) -- the hour changes on every 6th step."""
    g = load_golden("bs2_op2_eval_test")                     # eval mode: info rows recorded; test split: offsets 0
    kw = golden_kwargs("bs2_op2_eval_test")
    infos = g["infos"][:, 0]
    for t in (0, 4, 5, 6, 11, 143, 144, 145, 1000):
        h = (t + 1) * kw["sim_step"] // 3600
        assert infos[t, 0] == t                              # "step" is k before the increment
        assert infos[t, 1] == kw["e_r_b"][0, 0, h]           # el price of the NEW hour
        assert infos[t, 2] == kw["g_e"][0, 0, h // 24] and infos[t, 3] == kw["g_e"][1, 0, h // 24]
    assert kw["e_r_b"][0, 0, 0] != kw["e_r_b"][0, 0, 1] and infos[5, 1] != infos[4, 1]


def test_q10_business_scenarios_overwrite_the_tables_not_the_formula():
    """(10) BS2 / BS3 change the market tables the env receives (src/rl_utils.py:119-126): constant gas price in BS2,
    zero gas and EUA prices in BS3; the reward formula only switches the CHP terms on b_s3 (:76-77, :291-297)."""
    bs1 = real_kwargs(dict(scenario=1, operation="OP2"))
    bs2 = real_kwargs(dict(scenario=2, operation="OP2"))
    bs3 = real_kwargs(dict(scenario=3, operation="OP2"))
    assert len(np.unique(bs1["g_e"][0])) > 100
    assert len(np.unique(bs2["g_e"][0])) == 1 and bs2["g_e"][0, 0, 0] > 0
    assert np.array_equal(bs2["g_e"][1], bs1["g_e"][1])
    assert np.all(bs3["g_e"] == 0.0)
    assert np.array_equal(bs1["e_r_b"][0], bs3["e_r_b"][0])          # the electricity prices are the same series


def test_q11_info_has_exactly_the_24_reference_fields():
    """(11) Q_chp only exists as an attribute after the first _get_reward and is not part of info (:251-278)."""
    assert len(_abi.INFO_KEYS) == _abi.PTG_N_INFO == 24
    assert not any("chp" in k.lower() and "rev" not in k.lower() for k in _abi.INFO_KEYS)
    assert load_golden("bs2_op2_eval_test")["infos"].shape[-1] == 24


def test_q12_noisy_index_is_clamped_at_zero_and_truncated():
    """(12) i = int(max(argmin + normal(0, noise), 0)) (:584-585): negative sums give 0, positive ones are truncated
    toward zero; and an index past the END of the table is kept (A.3: S copies of the last row, i and j unchanged,
    the state still moves on)."""
    g, kw, tape, snaps, env = oracle_replay("quirks_s50")
    up, hot = op_table(kw, "standby_up")[:, 1], op_table(kw, "startup_hot")
    clamped = 0
    for e, marks in enumerate(marks_of("quirks_s50")):
        for t in marks["standby_from_cold"]:
            T = snaps[t]["t_cat"][e]
            assert T <= 15.2 and snaps[t + 1]["standby_ds"][e] == DS["standby_up"]
            v = int(np.abs(up - T).argmin()) + tape[e, snaps[t]["draws"][e]]
            assert g["ints"][t, e, I] == int(max(v, 0)) and g["ints"][t, e, J] == 1
            clamped += v < 0
        (t,) = marks["standby_clamped_at_0"]
        assert g["ints"][t, e, I] == 0
        (t,) = marks["startup_hot_past_end"]
        T = snaps[t]["t_cat"][e]
        assert T > 400 and int(np.abs(hot[:, 1] - T).argmin()) == len(hot) - 1
        v = len(hot) - 1 + tape[e, snaps[t]["draws"][e]]
        pre, post = g["ints"][t - 1, e], g["ints"][t, e]
        assert post[I] == int(v) >= len(hot) and post[J] == 1              # (i, j) of the handler survive the step
        assert pre[STATE] == STANDBY and post[STATE] == PARTIAL            # ... but the state is already partial load
        assert snaps[t + 1]["t_cat"][e] == hot[-1, 1]
    assert clamped >= 3
