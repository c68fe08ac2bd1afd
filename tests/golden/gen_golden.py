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
}


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
    actions = make_actions(opt["actions"], steps, n_envs)

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
                             seed_train=3654, seed_test=605)),
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
