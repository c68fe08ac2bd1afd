"""Full-length parity soak of the benchmarked workload (run on a GPU box; not part of the test suite):

    python tools/soak_parity.py [--envs 8192] [--steps 5400]

Synthetic BS2/OP2 `mod` with the REAL episode length (5 323 steps), seeds 3654 + i, uniform random actions, numpy-exact
noise on the device, against the CPU oracle on the same actions: every step's dones and rewards, the full observation
every 7th step and around the episode end, the integer plant state every 50th step, Monitor records at the episode
end.  Prints one summary line; exits non-zero on the first mismatch."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape  # noqa: E402
from rl_ptg_b200.vec_env import PtGVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=8192)
ap.add_argument("--steps", type=int, default=5400)
args = ap.parse_args()
kw = bench.make_kwargs()
n, steps = args.envs, args.steps
ep_len = int(kw["eps_sim_steps"]) - 5
t0 = time.time()
ora = OracleVecEnv(kw, n, noise_tape=draw_noise_tape(3654 + np.arange(n), kw["noise"], steps), threads=len(os.sched_getaffinity(0)))
env = PtGVecEnv(kw, n, seed=3654)
o_obs = ora.reset().copy()
obs = env.reset()
keys = list(obs.keys())
flat = lambda o: np.concatenate([np.asarray(o[k], dtype=np.float64).reshape(n, -1) for k in keys], axis=1)   # noqa: E731


def close(got, want, what):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    zero = want == 0.0
    if not np.all(got[zero] == 0.0):
        sys.exit(f"MISMATCH {what}: exact zeros differ")
    tiny = np.abs(want) < 1e-9
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    if np.any(err[~tiny & ~zero] > 1e-5):
        sys.exit(f"MISMATCH {what}: rel err {err[~tiny & ~zero].max():.3e}")


close(flat(obs), o_obs, "reset obs")
rng = np.random.default_rng(0)
n_done = obs_checked = 0
for t in range(steps):
    a = rng.integers(0, 5, size=n)
    o_obs, o_rew, o_done = ora.step(a)
    obs, rew, done, infos = env.step(a)
    if not np.array_equal(done, o_done.astype(bool)):
        sys.exit(f"MISMATCH done at step {t}")
    close(rew, o_rew, f"reward step {t}")
    near_end = abs((t + 1) % ep_len) <= 2 or (t + 1) % ep_len >= ep_len - 2
    if t % 7 == 0 or near_end:
        close(flat(obs), o_obs, f"obs step {t}")
        obs_checked += 1
    if o_done.any():
        n_done += int(o_done.sum())
        e = int(np.nonzero(o_done)[0][0])
        if infos[e]["episode"]["l"] != ep_len or abs(infos[e]["episode"]["r"] - ora.episode_return[e]) > 1e-6 * max(1.0, abs(ora.episode_return[e])):
            sys.exit(f"MISMATCH Monitor record at step {t}")
    if t % 50 == 0 or t == steps - 1:
        so, sg = ora.get_state(), env.get_state()
        for f in ("meth_state", "i", "j", "k", "hot_cold", "standby_ds", "startup_ds", "partial_ds", "full_ds",
                  "current_action", "act_ep_h", "act_ep_d", "episode_count", "draws"):
            if not np.array_equal(so[f], sg[f]):
                sys.exit(f"MISMATCH state field {f} at step {t}")
print(f"soak OK: {n} envs x {steps} steps of the benchmarked workload (episode length {ep_len}): dones and rewards every step, "
      f"{obs_checked} full observation checks, {n_done} episode ends, integer state every 50 steps -- all equal to the CPU "
      f"oracle (bit-exact integers, <= 1e-5 relative fp32) in {time.time() - t0:.0f} s")
