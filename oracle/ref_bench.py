"""Timing harness for the UNMODIFIED reference ``PTGEnv`` on host cores -- TEST / BENCH INFRASTRUCTURE ONLY.

``bench.py``'s ``cpu_baseline`` leg and ``--impl reference`` arm are the only callers.  The reference's own
``env/ptg_gym_env.py`` is imported as it lies in ``baseline/_ref/`` (staged by ``tools/stage_reference.sh``; the copy
travels to the GPU box, ``/root/reference`` does not) through the gymnasium stub of ``oracle/ref_harness.py`` (real
gymnasium / stable_baselines3 are used automatically when importable).  Two arrangements (BASELINE.md
section 3):

* ``single_env_episode``  -- BASELINE config 1: ONE reference env, DummyVecEnv semantics (``reset(seed=...)``, step,
  in-loop auto-reset), a fixed action tape, one training episode; also returns the SHA-256 of the per-step
  ``(Meth_State, i, j, hot_cold, terminated)`` record for the bit-exact comparison with the CUDA path.
* ``SubprocLockstep``     -- the reference's ``parallel: Multiprocessing`` arrangement
  (``src/rl_utils.py:484-486``: SB3 ``SubprocVecEnv``): one forked worker process per env, SB3's worker protocol over
  ``multiprocessing.Pipe`` restated (``("step", a)`` -> ``(obs, rew, done, info, reset_info)`` with Monitor bookkeeping
  and auto-reset inside the worker, dict observations stacked per key in the parent), lock-step gather.  SB3 itself is
  not installable here (no network); when it is importable the real ``SubprocVecEnv`` + ``Monitor`` are used instead.
"""
from __future__ import annotations

import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

from . import ref_harness


def reference_root() -> str | None:
    """Where the unmodified reference can be imported from on THIS machine (staged copy first: it is what travels)."""
    for cand in (os.environ.get("PTG_REFERENCE_ROOT"), ref_harness.STAGED_ROOT, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "env", "ptg_gym_env.py")):
            return cand
    return None


_PG = None


def load_reference_env_module(root: str | None = None):
    """Import the reference's ``env.ptg_gym_env`` unmodified (module object; its ``PTGEnv`` is the class under test)."""
    global _PG
    if _PG is not None:
        return _PG
    root = root or reference_root()
    if root is None:
        raise RuntimeError("reference not available: run tools/stage_reference.sh where /root/reference is mounted")
    ref_harness._install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    _PG = importlib.import_module("env.ptg_gym_env")
    assert os.path.abspath(_PG.__file__).startswith(os.path.abspath(root)), _PG.__file__
    return _PG


def single_env_episode(kw: dict, actions: np.ndarray, seed: int, record: bool = False, repeats: int = 1,
                       train_or_eval: str = "train"):
    """One reference env stepped through ``len(actions)`` steps with DummyVecEnv semantics.

    Returns ``(best_seconds, n_steps, sha or None, n_episode_ends)``.  Every repeat constructs a fresh env with the
    module-global episode counter at 0 (``env/ptg_gym_env.py:9``) and calls ``reset(seed=seed)`` like SB3's first
    ``VecEnv.reset()`` after ``make_vec_env(seed=...)``."""
    pg = load_reference_env_module()
    best, sha, ends = float("inf"), None, 0
    acts = [int(a) for a in np.asarray(actions).reshape(-1)] if kw["action_type"] == "discrete" else \
        [np.array([a], dtype=np.float32) for a in np.asarray(actions, dtype=np.float32).reshape(-1)]
    for rep in range(repeats):
        pg.ep_index = 0
        env = pg.PTGEnv(kw, train_or_eval)
        env.reset(seed=int(seed))
        rec = np.zeros((len(acts), 5), dtype=np.int64) if (record and rep == 0) else None
        ends = 0
        t0 = time.perf_counter()
        for t, a in enumerate(acts):
            _, _, terminated, _, _ = env.step(a)
            if rec is not None:
                rec[t] = (env.Meth_State, env.i, env.j, env.hot_cold, terminated)
            if terminated:
                ends += 1
                env.reset()
        dt = time.perf_counter() - t0
        best = min(best, dt)
        if rec is not None:
            sha = hashlib.sha256(rec.tobytes()).hexdigest()
    return best, len(acts), sha, ends


# ---------------------------------------------------------------------------------------------------------------
# SubprocVecEnv-style lock-step (SB3 2.0 `_worker` protocol restated)
# ---------------------------------------------------------------------------------------------------------------
def _worker(remote, parent_remote, root, kw, train_or_eval):
    parent_remote.close()
    pg = load_reference_env_module(root)
    env = pg.PTGEnv(kw, train_or_eval)
    ep_ret, ep_len, t_start = 0.0, 0, time.time()         # Monitor (make_vec_env wraps every env in one)
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                obs, reward, terminated, truncated, info = env.step(data)
                ep_ret += reward
                ep_len += 1
                done = terminated or truncated
                info["TimeLimit.truncated"] = truncated and not terminated
                reset_info = {}
                if done:
                    info["episode"] = {"r": round(ep_ret, 6), "l": ep_len, "t": round(time.time() - t_start, 6)}
                    info["terminal_observation"] = obs
                    obs, reset_info = env.reset()
                    ep_ret, ep_len = 0.0, 0
                remote.send((obs, reward, done, info, reset_info))
            elif cmd == "reset":
                obs, reset_info = env.reset(seed=data)
                ep_ret, ep_len = 0.0, 0
                remote.send((obs, reset_info))
            elif cmd == "close":
                remote.close()
                break
            else:
                raise NotImplementedError(cmd)
    except (EOFError, KeyboardInterrupt):
        pass


class SubprocLockstep:
    """``n`` reference envs, one forked process each, stepped in lock-step over pipes."""

    def __init__(self, kw: dict, n: int, seed: int = 3654, train_or_eval: str = "train"):
        self.n = n
        root = reference_root()
        load_reference_env_module(root)                       # import once in the parent: forked children inherit it
        kw = dict(kw, parallel="Multiprocessing")             # what TrainConfig.parallel is when SubprocVecEnv is used
        ctx = mp.get_context("fork")
        self.remotes, self.work_remotes = zip(*[ctx.Pipe() for _ in range(n)])
        self.procs = []
        for wr, r in zip(self.work_remotes, self.remotes):
            p = ctx.Process(target=_worker, args=(wr, r, root, kw, train_or_eval), daemon=True)
            p.start()
            self.procs.append(p)
            wr.close()
        for q, r in enumerate(self.remotes):
            r.send(("reset", seed + q))
        self.obs = self._stack([r.recv()[0] for r in self.remotes])

    @staticmethod
    def _stack(obs_list):
        return {k: np.stack([o[k] for o in obs_list]) for k in obs_list[0]}      # SB3 _flatten_obs for Dict spaces

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        results = [r.recv() for r in self.remotes]
        obs, rews, dones, infos, _ = zip(*results)
        return self._stack(obs), np.stack(rews), np.stack(dones), infos

    def close(self):
        for r in self.remotes:
            try:
                r.send(("close", None))
            except (BrokenPipeError, OSError):
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.kill()


def subproc_rate(kw: dict, n: int, lock_steps: int, warm: int = 20, seed: int = 3654):
    """Total env-steps/s of ``n`` forked reference envs over ``lock_steps`` lock-steps (uniform random actions)."""
    venv = SubprocLockstep(kw, n, seed)
    try:
        rng = np.random.default_rng(0)
        acts = rng.integers(0, 5, size=(warm + lock_steps, n))
        acts = [[int(a) for a in row] for row in acts]
        for t in range(warm):
            venv.step(acts[t])
        t0 = time.perf_counter()
        for t in range(warm, warm + lock_steps):
            venv.step(acts[t])
        dt = time.perf_counter() - t0
    finally:
        venv.close()
    return n * lock_steps / dt, dt


def dummy_rate(kw: dict, n: int, lock_steps: int, seed: int = 3654):
    """The reference's default arrangement (``parallel: Singleprocessing`` -> DummyVecEnv): ``n`` envs stepped in a
    Python loop in ONE process; total env-steps/s."""
    pg = load_reference_env_module()
    pg.ep_index = 0
    envs = [pg.PTGEnv(kw, "train") for _ in range(n)]
    for q, e in enumerate(envs):
        e.reset(seed=seed + q)
    rng = np.random.default_rng(0)
    acts = [[int(a) for a in row] for row in rng.integers(0, 5, size=(lock_steps, n))]
    t0 = time.perf_counter()
    for row in acts:
        obs_list = []
        for e, a in zip(envs, row):
            obs, _, terminated, _, _ = e.step(a)
            if terminated:
                obs, _ = e.reset()
            obs_list.append(obs)
        SubprocLockstep._stack(obs_list)
    dt = time.perf_counter() - t0
    return n * lock_steps / dt, dt
