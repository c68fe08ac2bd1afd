"""Differential fuzz of the CPU oracle (oracle/ptg_oracle.c, fed by the PRODUCT's preprocessing) against the UNMODIFIED
reference env -- TEST INFRASTRUCTURE, runs only where the reference checkout is mounted (/root/reference).

    python tools/fuzz_oracle_vs_reference.py [trials] [seed]      # log: profiles/oracle_fuzz_rNN.log

Every trial draws a configuration the 14 committed golden cases do not contain (business scenario, OP set, observation
design, sim_step, price_ahead, noise level, state-change penalty, perturbed load-change thresholds, action policy, env
seeds), records the reference under DummyVecEnv semantics with tests/golden/gen_golden.run_case, and replays the tape
through the oracle: integer plant state every step bit-exact, rewards / observations / info rows to 1e-12.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import gen_golden  # noqa: E402
from helpers import real_kwargs  # noqa: E402
from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape  # noqa: E402

TOL = 1e-12
THRESHOLDS = ["time1_start_p_f", "time2_start_f_p", "time_p_f", "time_f_p", "time1_p_f_p", "time2_p_f_p", "time23_p_f_p",
              "time34_p_f_p", "time45_p_f_p", "time5_p_f_p", "time1_f_p_f", "time23_f_p_f", "time34_f_p_f",
              "time45_f_p_f", "time5_f_p_f"]


def draw_trial(rng: np.random.Generator):
    ov = dict(scenario=int(rng.integers(1, 4)), operation=str(rng.choice(["OP1", "OP2"])),
              raw_modified=str(rng.choice(["raw", "mod"])), sim_step=int(rng.choice([100, 120, 200, 300, 600, 900])),
              price_ahead=int(rng.choice([1, 4, 6, 13, 16])), noise=int(rng.choice([0, 3, 10, 40])),
              state_change_penalty=float(rng.choice([0.0, 0.0, 0.25, 1.5])),
              eps_len_d=int(rng.choice([1, 37, 37, 41])))       # 1-day episodes: several auto-resets per trial (ep_index)
    base = real_kwargs(dict(scenario=1, operation="OP2"))
    for name in rng.choice(THRESHOLDS, size=int(rng.integers(0, 5)), replace=False):     # move a few thresholds
        S = ov["sim_step"] // 2
        ov[str(name)] = int(rng.choice([base[str(name)] + int(rng.integers(-40, 41)), S * int(rng.integers(1, 6))]))
    mode = str(rng.choice(["train", "train", "eval"]))
    split = "train" if mode == "train" else str(rng.choice(["val", "test"]))
    continuous = bool(rng.random() < 0.25)
    opt = dict(n_envs=int(rng.integers(1, 4)), steps=int(rng.integers(250, 700)), split=split, mode=mode,
               seed=int(rng.integers(0, 2 ** 31 - 1)),
               actions="continuous" if continuous else str(rng.choice(["uniform", "load", "load2"])))
    if continuous:
        opt["action_type"] = "continuous"
    return ov, opt


def replay_on_oracle(g: dict, ov: dict, opt: dict) -> str | None:
    import json
    m = json.loads(g["meta"])
    kw = real_kwargs(m["overrides"], m["split"], m["action_type"], m["seed_train"])
    n, steps = m["n_envs"], m["steps"]
    tape = draw_noise_tape(m["seed"] + np.arange(n), kw["noise"], steps + 1)
    env = OracleVecEnv(kw, n, train_or_eval=m["mode"], noise_tape=tape)
    obs = env.reset()
    if not np.array_equal(obs, g["reset_obs"]):
        return "reset obs"
    if not np.allclose(env.info, g["reset_info"], rtol=TOL, atol=0):
        return "reset info"
    keep = {int(t): q for q, t in enumerate(g["obs_steps"])}
    for t in range(steps):
        obs, rew, done = env.step(g["actions"][t])
        st = env.get_state()
        got = np.stack([st["meth_state"], st["i"], st["j"], st["hot_cold"], done.astype(np.int32), st["k"],
                        st["act_ep_h"], st["act_ep_d"], st["partial_ds"], st["full_ds"]], axis=1)
        if not np.array_equal(got, g["ints"][t]):
            return f"integer state, step {t}: {got.tolist()} vs {g['ints'][t].tolist()}"
        if not np.allclose(rew, g["rewards"][t], rtol=TOL, atol=0):
            return f"reward, step {t}: {rew} vs {g['rewards'][t]}"
        if t in keep and not np.allclose(obs, g["obs"][keep[t]], rtol=TOL, atol=1e-15):
            return f"obs, step {t}"
        if m["mode"] == "eval" and not np.allclose(env.info, g["infos"][t], rtol=TOL, atol=1e-15):
            return f"info, step {t}"
    env.close()
    return None


def main():
    trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 20261018)
    bad = 0
    t0 = time.time()
    for q in range(trials):
        ov, opt = draw_trial(rng)
        try:
            g = gen_golden.run_case(f"fuzz{q}", ov, opt)
        except Exception as exc:                      # a configuration the REFERENCE itself rejects / crashes on
            print(f"trial {q:3d} reference raised {type(exc).__name__}: {str(exc)[:90]} | {ov} {opt}", flush=True)
            try:
                real_kwargs(ov, opt["split"], opt.get("action_type", "discrete"))
                print("          (the product's preprocessing accepted it)", flush=True)
            except Exception as exc2:
                print(f"          product preprocessing raised {type(exc2).__name__} too", flush=True)
            continue
        err = replay_on_oracle(g, ov, opt)
        visits = np.bincount(g["ints"][:, :, 0].ravel(), minlength=5).tolist()
        print(f"trial {q:3d} {'OK  ' if err is None else 'FAIL'} {ov} {opt} state_visits={visits}"
              + ("" if err is None else f"  <-- {err}"), flush=True)
        bad += err is not None
    print(f"{trials} trials, {bad} mismatches, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
