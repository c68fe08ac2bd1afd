"""Golden vectors of ONE reference env driven through the Gymnasium API (no VecEnv, no auto-reset) on the TRAINING
split: constructor -> reset(seed) -> steps until `terminated` -> reset() -> ... for three episodes.  Pins the episode
schedule of a lone env (eps_ind[0] for the constructor, [1], [2], ... for the resets; env/ptg_gym_env.py:59-62,
490-493) for rl_ptg_b200.gym_env.PTGEnv.

    python tests/golden/gen_golden_single.py       # needs the reference checkout (see oracle/ref_harness.py)
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_harness import ReferenceSession  # noqa: E402
sys.path.insert(0, HERE)
from gen_golden import flat_obs, info_row  # noqa: E402

EPISODES = 3
SEED = 3654


def main():
    overrides = dict(scenario=2, operation="OP2")
    sess = ReferenceSession(overrides)
    kw = sess.kwargs("train")
    ep_len = int(kw["eps_sim_steps"]) - 5
    rng = np.random.default_rng(7)
    actions = rng.choice(5, size=EPISODES * ep_len, p=[0.1, 0.1, 0.2, 0.3, 0.3]).astype(np.int64)
    sess.pg.ep_index = 0
    env = sess.pg.PTGEnv(kw, "train")
    offsets = [(env.act_ep_h, env.act_ep_d)]                 # constructor: eps_ind[0]
    reset_obs, reset_info, term_obs = [], [], []
    ints = np.zeros((len(actions), 6), dtype=np.int32)       # state, i, j, hot_cold, terminated, k (after the step)
    rewards = np.zeros(len(actions))
    obs_rows = {}
    t = 0
    for ep in range(EPISODES):
        o, inf = env.reset(seed=SEED) if ep == 0 else env.reset()
        offsets.append((env.act_ep_h, env.act_ep_d))
        reset_obs.append(flat_obs(o))
        reset_info.append(info_row(inf))
        while True:
            o, r, term, trunc, inf = env.step(int(actions[t]))
            assert inf == {} and trunc is False
            ints[t] = (env.Meth_State, env.i, env.j, env.hot_cold, int(term), env.k)
            rewards[t] = r
            if t % 97 == 0 or term:
                obs_rows[t] = flat_obs(o)
            t += 1
            if term:
                term_obs.append(flat_obs(o))
                break
    assert t == len(actions)
    keep = np.array(sorted(obs_rows), dtype=np.int32)
    out = dict(meta=json.dumps(dict(overrides=overrides, split="train", seed=SEED, episodes=EPISODES, ep_len=ep_len,
                                    seed_train=3654, action_type="discrete", numpy=np.__version__)),
               actions=actions.astype(np.uint8), ints=ints, rewards=rewards, offsets=np.array(offsets, dtype=np.int32),
               reset_obs=np.stack(reset_obs), reset_info=np.stack(reset_info), term_obs=np.stack(term_obs),
               obs_steps=keep, obs=np.stack([obs_rows[int(q)] for q in keep]), eps_ind_head=kw["eps_ind"][:8])
    path = os.path.join(HERE, "single_env_train_resets.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(actions)} steps, offsets {offsets}, eps_ind[:4] {kw['eps_ind'][:4].tolist()}, "
          f"{os.path.getsize(path) // 1024} KiB")
    sess.close()


if __name__ == "__main__":
    main()
