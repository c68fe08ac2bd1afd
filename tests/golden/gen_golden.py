"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (run where it is mounted).

    python tests/golden/gen_golden.py            # needs /root/reference (or PTG_REFERENCE_ROOT)

Outputs (committed):
  ref_data.npz          raw content of the reference's CSV inputs (market series + OP1/OP2 tables), so the same
                        inputs exist on boxes without the reference checkout
  golden_<case>.npz     per case: config overrides, seeds, the action sequence, and what the reference env
                        produced under DummyVecEnv semantics (make_vec_env order, reset(seed+i), lock-step,
                        auto-reset): per-step integer plant state, rewards (fp64), a subset of observations
                        (fp64), terminal observations, eval-mode info rows, and the preprocessing scalars.

The cases mirror BASELINE.json's configs (BS2/OP2, BS1/OP1, BS3/OP2) plus the edge cases of SURVEY.md A.9.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import REF_ROOT, ReferenceSession  # noqa: E402
from rl_ptg_b200._abi import DATASET_NAMES  # noqa: E402
from rl_ptg_b200.data import save_raw_npz  # noqa: E402

ACTION_NAMES = ["standby", "cooldown", "startup", "partial_load", "full_load"]

CASES = {
    # name: (config_env overrides, options)
    "bs2_op2_mod": (dict(scenario=2, operation="OP2"), dict(n_envs=3, steps=5323 + 250, actions="uniform")),
    "bs1_op1_mod": (dict(scenario=1, operation="OP1"), dict(n_envs=2, steps=5323 + 40, actions="uniform")),
    "bs3_op2_mod_penalty": (dict(scenario=3, operation="OP2", state_change_penalty=0.5),
                            dict(n_envs=2, steps=2500, actions="load")),
    "bs1_op2_raw": (dict(scenario=1, operation="OP2", raw_modified="raw"),
                    dict(n_envs=2, steps=2500, actions="load")),
    "bs2_op2_continuous": (dict(scenario=2, operation="OP2"),
                           dict(n_envs=2, steps=2000, actions="continuous", action_type="continuous")),
    "bs2_op2_eval_test": (dict(scenario=2, operation="OP2"),
                          dict(n_envs=1, steps=8635 + 30, actions="load", split="test", mode="eval", seed=605)),
    "bs1_op1_fast": (dict(scenario=1, operation="OP1", sim_step=120, eps_len_d=37),
                     dict(n_envs=2, steps=6000, actions="load2")),
    "bs2_op2_s50": (dict(scenario=2, operation="OP2", sim_step=100),
                    dict(n_envs=2, steps=6000, actions="load2")),
    # narrower / wider hour rows than the default 13 (kernel instantiations NV = 3 and NV = 5), raw and mod
    "bs3_op1_raw_pa6": (dict(scenario=3, operation="OP1", raw_modified="raw", price_ahead=6),
                        dict(n_envs=2, steps=1500, actions="uniform")),
    "bs1_op2_mod_pa16": (dict(scenario=1, operation="OP2", price_ahead=16),
                         dict(n_envs=2, steps=1500, actions="load")),
    # validation split in eval mode (EvalCallback env, src/rl_utils.py:456-469), continuous + raw + penalty
    "bs2_op1_val_eval": (dict(scenario=2, operation="OP1"),
                         dict(n_envs=2, steps=1200, actions="uniform", split="val", mode="eval", seed=605)),
    "bs3_op2_raw_continuous_penalty": (dict(scenario=3, operation="OP2", raw_modified="raw", state_change_penalty=0.25),
                                       dict(n_envs=2, steps=1500, actions="continuous", action_type="continuous")),
    # scripted tours through the quirks of SURVEY.md A.9 (closed loop on the reference env, recorded as a tape;
    # tests/test_quirks.py asserts what each marked step must show).  quirks_s50 moves time2_p_f_p onto a time_op the
    # 50-row steps can reach (150), so that the strict `<` chains are met with equality.
    "quirks_s300": (dict(scenario=2, operation="OP2"), dict(n_envs=3, steps=420, actions="tour_s300")),
    "quirks_s50": (dict(scenario=2, operation="OP2", sim_step=100, time2_p_f_p=150),
                   dict(n_envs=3, steps=420, actions="tour_s50")),
}


# ---- closed-loop scripts (policy(env, mem) -> action); `mem["marks"]` collects (label, step) pairs -------------
def _phase(mem):
    mem.setdefault("ph", 0); mem.setdefault("n", 0); mem.setdefault("marks", []); mem.setdefault("tries", 0)
    mem["n"] += 1
    return mem["ph"], mem["n"]


def _go(mem, ph):
    mem["ph"], mem["n"] = ph, 0


def tour_s300(env, mem, t):
    """Default thresholds, S = 300: first step after reset, cold start-up with hand-over, the op1 -> op3 -> op7 chain,
    an exhausted table, hot/cold hysteresis, both standby tables, index clamp at 0."""
    ph, n = _phase(mem)
    st, mark = env.Meth_State, lambda label: mem["marks"].append((label, t))
    if ph == 0:                      # cont(cooldown) right after reset: j 0 -> 1; on until T_cat <= 15.0 degC
        if n == 1: mark("first_cont_cooldown")
        if env.Meth_T_cat <= 15.0: _go(mem, 1)
        return 1
    if ph == 1:                      # _standby below the first row of standby_up (15.2): argmin = 0, clamped if noise < 0
        mark("standby_from_cold"); _go(mem, 2); return 0
    if ph == 2:                      # cold start-up until the hand-over to partial load
        if n == 1: mark("startup_cold")
        if st == 3: mark("handover_done"); _go(mem, 3); return 3
        return 2
    if ph == 3:                      # dwell in op1_start_p beyond time1_start_p_f = 1201 rows
        if env.i + env.j * env.step_size >= 1201 + 300: mark("full_from_op1_late"); _go(mem, 4); return 4
        return 3
    if ph == 4:                      # one more full-load step: time_op = 600
        _go(mem, 5); return 4
    if ph == 5:                      # time45_p_f_p < 600 < time5_p_f_p: op7_p_f_p_22, i = time5_p_f_p
        mark("partial_563_675"); _go(mem, 6); return 3
    if ph == 6:                      # exhaust op7_p_f_p_22 (3 481 rows): padded, then constant last row
        if n >= 14: mark("full_from_catch_all"); _go(mem, 7); return 4
        return 3
    if ph == 7:                      # full load (T_cat >= 350: hot), then standby_down
        if n >= 3: mark("standby_from_hot"); _go(mem, 8); return 0
        return 4
    if ph == 8:                      # standby_down until 160 < T_cat < 350, then a start-up: hot_cold kept at 1
        if env.Meth_T_cat < 300: mark("startup_in_hysteresis_band"); _go(mem, 9); return 2
        return 0
    if ph == 9:
        if st == 3: mark("cooldown_from_load"); _go(mem, 10); return 1
        return 2
    if ph == 10:                     # cooldown below 160 degC: hot_cold -> 0
        if env.Meth_T_cat < 150: mark("standby_below_188"); _go(mem, 11); return 0
        return 1
    if ph == 11:                     # standby_up for a few steps, then a COLD start-up although T_cat > 160
        if n >= 3: mark("startup_cold_again"); _go(mem, 12); return 2
        return 0
    if ph == 12:
        if st == 3: _go(mem, 13); return 4
        return 2
    return 4 if (n // 7) % 2 == 0 else 3


def tour_s50(env, mem, t):
    """S = 50 (sim_step = 100), time2_p_f_p = 150: "fully developed" indices, `j += 1` with i kept, a threshold met with
    equality, a hot start-up whose noisy index lands past the end of the table."""
    ph, n = _phase(mem)
    st, mark = env.Meth_State, lambda label: mem["marks"].append((label, t))
    t_op = env.i + env.j * env.step_size
    if ph == 0 and "cold" not in mem:                # start with the cold probes below
        mem["cold"] = 0; _go(mem, 20); ph = 20
    if ph == 20:                     # cooldown (cont, or _cooldown after a probe) until T_cat <= 15.2 = standby_up[0:5, T]
        if env.Meth_T_cat <= 15.2 and st == 1: _go(mem, 21); mark("standby_from_cold"); mem["cold"] += 1; return 0
        return 1
    if ph == 21:                     # argmin + noise < 0 -> i = int(max(.., 0)) = 0; repeat until one probe was clamped
        if env.i == 0 and env.j == 1: mem["marks"].append(("standby_clamped_at_0", t - 1))
        _go(mem, 0 if ((env.i == 0 and env.j == 1) or mem["cold"] >= 6) else 20)
        return 1 if mem["ph"] == 20 else 2
    if ph == 0:                      # cold start-up; right at the hand-over: _full at time_op < time1_start_p_f
        if st == 3: mark("full_from_op1_early"); _go(mem, 30); return 4
        return 2
    if ph == 30:                     # op2_start_f at time_op = 50 < time2_start_f_p: _partial takes i from the argmin
        mark("partial_argmin_no_noise"); _go(mem, 1); return 3        # of op1_start_p WITHOUT a noise draw (:641)
    if ph == 1:                      # dwell in op1_start_p beyond time1_start_p_f, then _full -> op3_p_f (0, 1)
        if t_op >= 1201: mark("full_from_op1_late"); _go(mem, 2); return 4
        return 3
    if ph == 2:                      # time_op = 50 < time1_p_f_p = 51
        mark("partial_fully_developed"); _go(mem, 3); return 3
    if ph == 3:                      # two steps on the last row of op8_f_p, then _full at time_op = 17 000+
        if n >= 2: mark("full_from_op8_late"); _go(mem, 4); return 4
        return 3
    if ph == 4:                      # cont(full): time_op = 100
        _go(mem, 5); return 4
    if ph == 5:                      # 51 < 100 < 150: op4_p_f_p_5, j += 1, i from the previous table
        mark("partial_keep_i"); _go(mem, 6); return 3
    if ph == 6:                      # part_op = op4_p_f_p_5: catch-all branch of _full
        mark("full_from_catch_all"); _go(mem, 7); return 4
    if ph == 7:
        if t_op >= 150: mark("partial_at_threshold"); _go(mem, 8); return 3     # time_op == time2_p_f_p
        return 4
    if ph == 8:                      # _full from op8_f_p at time_op = 50 < time1_f_p_f = 51
        mark("full_fully_developed"); _go(mem, 9); return 4
    if ph == 9:                      # re-heat at full load, one standby step, then a start-up at T_cat > 400 degC:
        if n >= 6: _go(mem, 10); return 0            # argmin = last row of startup_hot (2 047 of 2 048)
        return 4
    if ph == 10:
        mark("startup_hot_probe"); mem["tries"] += 1; _go(mem, 11); return 2
    if ph == 11:                     # repeat the probe until the noisy index fell past the end (i >= L; i, j kept)
        hit = env.j == 1 and env.i >= len(env.startup_hot)
        if hit: mem["marks"].append(("startup_hot_past_end", t - 1))
        _go(mem, 12 if (hit or mem["tries"] >= 12) else 9)
        return 4
    return 4 if (n // 3) % 2 == 0 else 3


SCRIPTS = {"tour_s300": tour_s300, "tour_s50": tour_s50}


def make_actions(kind: str, steps: int, n_envs: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 5, size=(steps, n_envs)).astype(np.int64)
    if kind == "load":      # dwell in the load states so _partial/_full and table exhaustion are exercised
        return rng.choice(5, size=(steps, n_envs), p=[0.05, 0.05, 0.3, 0.3, 0.3]).astype(np.int64)
    if kind == "load2":     # sticky actions: long stays with occasional switches between partial and full
        a = np.empty((steps, n_envs), dtype=np.int64)
        cur = np.full(n_envs, 2)
        for t in range(steps):
            sw = rng.random(n_envs) < 0.25
            cur = np.where(sw, rng.choice(5, size=n_envs, p=[0.04, 0.04, 0.12, 0.4, 0.4]), cur)
            a[t] = cur
        return a
    if kind == "continuous":
        a = rng.uniform(-1.0, 1.0, size=(steps, n_envs)).astype(np.float32)
        edge = np.array([-1.0, 1.0, -0.6, -0.2, 0.2, 0.6, -0.19999999, 0.20000002, 0.99999994, 0.6000001],
                        dtype=np.float32)
        for t in range(0, steps, 37):   # sprinkle threshold / boundary values
            a[t, t % n_envs] = edge[(t // 37) % len(edge)]
        return a
    raise ValueError(kind)


def flat_obs(obs: dict) -> np.ndarray:
    return np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel() for v in obs.values()])


def info_row(info: dict) -> np.ndarray:
    row = np.zeros(24)
    for j, (k, v) in enumerate(info.items()):
        row[j] = ACTION_NAMES.index(v) if k == "Meth_Action" else float(v)
    return row


def run_case(name: str, overrides: dict, opt: dict) -> dict:
    n_envs, steps = opt["n_envs"], opt["steps"]
    split, mode = opt.get("split", "train"), opt.get("mode", "train")
    seed = opt.get("seed", 3654)
    action_type = opt.get("action_type", "discrete")
    sess = ReferenceSession(overrides, action_type=action_type)
    kw = sess.kwargs(split)
    script = SCRIPTS.get(opt["actions"])
    actions = np.zeros((steps, n_envs), dtype=np.int64) if script else make_actions(opt["actions"], steps, n_envs)
    mems = [dict() for _ in range(n_envs)]

    sess.pg.ep_index = 0                                   # fresh process
    envs = [sess.pg.PTGEnv(kw, mode) for _ in range(n_envs)]          # make_vec_env: constructors in order
    reset_obs, reset_info = [], []
    for e, env in enumerate(envs):                                    # VecEnv.seed(seed); VecEnv.reset()
        o, inf = env.reset(seed=seed + e)
        reset_obs.append(flat_obs(o))
        reset_info.append(info_row(inf))
    obs_dim = len(reset_obs[0])
    keys = list(o.keys())

    ints = np.zeros((steps, n_envs, 10), dtype=np.int32)  # state, i, j, hot_cold, done, k, act_ep_h, act_ep_d,
    #                                                       partial table id, full table id (enum PtgDataset)
    rewards = np.zeros((steps, n_envs))
    obs_all = np.zeros((steps, n_envs, obs_dim))
    infos = np.zeros((steps, n_envs, 24)) if mode == "eval" else None
    term_steps, term_obs, ep_ret, term_ints = [], [], [], []
    running = np.zeros(n_envs)
    for t in range(steps):
        for e, env in enumerate(envs):
            if script:
                actions[t, e] = script(env, mems[e], t)
            a = actions[t, e]
            o, r, term, trunc, inf = env.step(np.array([a], dtype=np.float32) if action_type == "continuous" else int(a))
            running[e] += r
            rewards[t, e] = r
            if infos is not None:
                infos[t, e] = info_row(inf)
            if term:
                term_steps.append((t, e))
                term_ints.append((env.Meth_State, env.i, env.j, env.hot_cold))     # plant state before the reset
                term_obs.append(flat_obs(o))
                ep_ret.append(running[e])
                running[e] = 0.0
                o, _ = env.reset()                                    # DummyVecEnv auto-reset
            # what a VecEnv exposes after step_wait(): post-auto-reset state for done envs
            ints[t, e, :6] = (env.Meth_State, env.i, env.j, env.hot_cold, int(term), env.k)
            ints[t, e, 6:10] = (env.act_ep_h, env.act_ep_d, DATASET_NAMES.index(env.part_op),
                                DATASET_NAMES.index(env.full_op))
            obs_all[t, e] = flat_obs(o)

    keep = np.unique(np.concatenate([np.arange(min(400, steps)), np.arange(0, steps, 16),
                                     np.arange(max(0, steps - 100), steps)] +
                                    [np.arange(max(0, t - 3), min(steps, t + 4)) for t, _ in term_steps]))
    out = dict(
        meta=json.dumps(dict(case=name, overrides=overrides, n_envs=n_envs, steps=steps, split=split, mode=mode,
                             seed=seed, action_type=action_type, obs_keys=keys, numpy=np.__version__,
                             seed_train=3654, seed_test=605, marks=[m.get("marks", []) for m in mems])),
        actions=actions, ints=ints, rewards=rewards, obs_steps=keep.astype(np.int32),
        obs=obs_all[keep], reset_obs=np.stack(reset_obs), reset_info=np.stack(reset_info),
        term_steps=np.array(term_steps, dtype=np.int32).reshape(-1, 2),
        term_obs=np.array(term_obs).reshape(-1, obs_dim), episode_return=np.array(ep_ret),
        term_ints=np.array(term_ints, dtype=np.int32).reshape(-1, 4),
        final_cum_rew=np.array([env.cum_rew for env in envs]),
        # preprocessing pins
        rew_l_b=kw["rew_l_b"], rew_u_b=kw["rew_u_b"], reward_level=np.asarray(kw["reward_level"]),
        max_h2_volumeflow=kw["max_h2_volumeflow"], eps_sim_steps=kw["eps_sim_steps"],
        n_eps_loops=kw["n_eps_loops"],
        eps_ind=(kw["eps_ind"] if kw["eps_ind"] is not None else np.zeros(0, dtype=np.int64)),
        e_r_b_sum=kw["e_r_b"].sum(axis=2), g_e_sum=kw["g_e"].sum(axis=2),
        e_r_b_probe=kw["e_r_b"][:, :, ::997].copy(),
        topt=np.array([sess.P.dict_pot_r_b[f"pot_rew_{s}"].sum() for s in ("train", "val", "test")]),
    )
    if infos is not None:
        out["infos"] = infos
    sess.close()
    return out


def main():
    save_raw_npz(os.path.join(HERE, "ref_data.npz"), REF_ROOT)
    print("wrote ref_data.npz", os.path.getsize(os.path.join(HERE, "ref_data.npz")) // 1024, "KiB")
    only = sys.argv[1:]
    for name, (overrides, opt) in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, overrides, opt)
        path = os.path.join(HERE, f"golden_{name}.npz")
        np.savez_compressed(path, **out)
        st = out["ints"]
        print(f"{name}: steps={st.shape[0]} envs={st.shape[1]} episodes_done={len(out['term_steps'])} "
              f"state_visits={np.bincount(st[:, :, 0].ravel(), minlength=5).tolist()} "
              f"part_tabs={sorted(set(st[:, :, 8].ravel().tolist()))} full_tabs={sorted(set(st[:, :, 9].ravel().tolist()))} "
              f"size={os.path.getsize(path) // 1024} KiB")


if __name__ == "__main__":
    main()
