#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched PtG environment on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE.json configs[3]'s environment -- BS2/OP2, "mod" observation design,
discrete actions, training episodes -- with 1,048,576 envs PER GPU (weak scaling: each rank owns its own shard of
the global env-id range, no data-path collective) on synthetic market/process data of the repo's shapes.
One "step" = one VecEnv step over every env (one k_step launch per rank).

  value     env-steps/s with the action tensor already resident in HBM (device API `step_tensor`)
  e2e       the same through the SB3-style numpy API `env.step(actions)`: pinned H2D of the actions and D2H of
            observations, rewards and dones inside the timed region
  roofline  algorithmic HBM bytes per env-step (ptg_bytes_per_env_step) x envs / kernel time (CUDA
            events around the timed region on the launching stream) vs the measured copy peak
  cpu_baseline  the UNMODIFIED Python reference PTGEnv (baseline/_ref, staged by tools/stage_reference.sh) on this
            box's host cores: SubprocVecEnv-style lock-step over all cores + one env / one episode in-process, with
            the SHA-256 of its (Meth_State, i, j, hot_cold, terminated) trajectory compared with the CUDA path's;
            the compiled C oracle port is reported beside it as cpu_baseline_port
  config.strong_scaling  BASELINE config 4 as written: 1,048,576 envs IN TOTAL over the ranks, the episode-statistics
            all-gather (ptg_allreduce_stats, NCCL) once per 32-step roll-out inside the timed loop

The timed region follows an untimed pre-roll (600 steps, stationary plant-state mix) and >= 10 ms of untimed single
steps, so a 20-step run measures the same steady state as a 4000-step run.

`--impl reference` times the unmodified Python reference under the SubprocVecEnv-style harness as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 1 << 20
# Algorithmic HBM bytes of one transition-noise draw (numpy-exact PCG64 state lives in HBM per env): PCG64 state
# 16 B read + 16 B written, increment 16 B read; the draw counter rides in spare bits of the per-step state word
# (DESIGN.md section 3).
RNG_BYTES_PER_DRAW = 48
WORKLOAD = "BS2/OP2 mod, discrete int64 actions, train episodes, synthetic data of repo shapes"
METRIC, UNIT = "env-steps/sec", "env-steps/s"
REF_LOCK_STEPS_PER_STEP = 100     # reference arm: one bench "step" = this many lock-steps of the Python reference


def make_kwargs(scenario=2, operation="OP2"):
    import rl_ptg_b200 as ptg
    E = ptg.EnvConfiguration(scenario=scenario, operation=operation)
    price, op = ptg.synthetic_data(E, seed=0)
    pp = ptg.Preprocessing(price, op, ptg.AgentConfiguration(), E, ptg.TrainConfiguration())
    return pp.dict_env_kwargs("train")


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[q] for r in self.rows if len(r) >= 6 for q in range(4) if r[2 + q] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_rate(kw, n_envs: int, lock_steps: int, threads: int, warm: int = 5):
    """env-steps/s of the CPU oracle (C restatement of PTGEnv, pthreads over envs) on this host."""
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    tape = draw_noise_tape(3654 + np.arange(n_envs), kw["noise"], lock_steps + warm + 1)
    env = OracleVecEnv(kw, n_envs, noise_tape=tape, threads=threads)
    env.reset()
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, size=(lock_steps + warm, n_envs))
    for t in range(warm):
        env.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warm, warm + lock_steps):
        env.step(acts[t])
    dt = time.perf_counter() - t0
    env.close()
    return n_envs * lock_steps / dt, dt


def episode_action_tape(kw) -> np.ndarray:
    """the fixed action tape of config 1 -- default_rng(0).integers(0, 5, eps_sim_steps), of which
    one training episode (eps_sim_steps - 5 steps) is used."""
    return np.random.default_rng(0).integers(0, 5, size=int(kw["eps_sim_steps"]))[:int(kw["eps_sim_steps"]) - 5]


def reference_python_baseline(kw, cores: int, lock_steps: int, repeats: int = 3, record: bool = True) -> dict:
    """The UNMODIFIED reference `PTGEnv` (env/ptg_gym_env.py, imported from baseline/_ref through the gymnasium stub)
    timed on this box's host cores in the two arrangements of SURVEY.md 8(d) / BASELINE.md section 3:
      (i)  config 1: ONE env, one training episode, fixed action tape, in-process, best of `repeats`
      (ii) SubprocVecEnv-style lock-step: `cores` forked workers over pipes, in-worker auto-reset, `lock_steps` steps
    plus the reference's default arrangement (DummyVecEnv, 6 envs in one process)."""
    from oracle import ref_bench
    acts = episode_action_tape(kw)
    best, n_steps, sha, ends = ref_bench.single_env_episode(kw, acts, 3654, record=record, repeats=repeats)
    sub_rate, sub_dt = ref_bench.subproc_rate(kw, cores, lock_steps)
    dummy_rate, _ = ref_bench.dummy_rate(kw, 6, max(50, lock_steps // 8))
    return {"single_env_steps_per_s": n_steps / best, "single_env_episode_s": best, "single_env_steps": n_steps,
            "single_env_episode_ends": ends, "trajectory_sha256": sha,
            "subproc_steps_per_s": sub_rate, "subproc_workers": cores, "subproc_lock_steps": lock_steps,
            "subproc_seconds": sub_dt, "dummy6_steps_per_s": dummy_rate, "root": ref_bench.reference_root()}


def reference_headline(ref: dict):
    """The reference figure the GPU path is compared with: the FASTEST of the reference's own arrangements on this box
    (SubprocVecEnv over all cores as the north star names it; DummyVecEnv with 6 envs, the reference's default; one env
    in-process) -- pipes and pickling make Subproc the slowest of the three on some hosts."""
    cands = {"SubprocVecEnv-style, one env per core": ref["subproc_steps_per_s"],
             "DummyVecEnv, 6 envs in one process": ref["dummy6_steps_per_s"],
             "one env in-process": ref["single_env_steps_per_s"]}
    which = max(cands, key=cands.get)
    return cands[which], which


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path -- the unmodified Python `PTGEnv` under a
    SubprocVecEnv-style harness on all host cores (`parallel: Multiprocessing`, src/rl_utils.py:484-486).  One "step"
    of this arm is a bounded sample of REF_LOCK_STEPS_PER_STEP lock-steps over `cores` envs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    kw = make_kwargs()
    from oracle import ref_bench
    if ref_bench.reference_root() is None:          # (never on a box that received baseline/_ref; kept loud, not silent)
        rate, dt = cpu_oracle_rate(kw, 4096 * max(1, cores // 4), 100, cores)
        _emit({"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": WORKLOAD, "note": "baseline/_ref missing (run tools/stage_reference.sh): CPU oracle port timed instead"},
               "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": "oracle port"},
               "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    K = max(1, min(args.steps, 60))
    W = max(1, min(args.warmup, 10))
    lock_steps = K * REF_LOCK_STEPS_PER_STEP
    ref = reference_python_baseline(kw, cores, lock_steps, repeats=3, record=True)
    port_rate, _ = cpu_oracle_rate(kw, 4096 * max(1, cores // 4), 60, cores)
    rate, which = reference_headline(ref)
    sample = (f"unmodified reference PTGEnv, {cores} forked workers (SubprocVecEnv protocol over pipes, in-worker "
              f"auto-reset), {lock_steps} lock-steps = {ref['subproc_seconds']:.1f} s, uniform random actions; value = "
              f"the fastest of the reference's arrangements on this box ({which})")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * ref["subproc_seconds"] / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": cores, "lock_steps_per_step": REF_LOCK_STEPS_PER_STEP,
                   "arrangement": "SubprocVecEnv-style (parallel: Multiprocessing), one reference env per host core; also "
                                  "DummyVecEnv with 6 envs (the reference's default) and one env in-process",
                   "arrangement_reported": which,
                   "reference": ref,
                   "cpu_oracle_port_steps_per_s": port_rate},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def cuda_trajectory_sha(kw, dev) -> str:
    """SHA-256 of the per-step (Meth_State, i, j, hot_cold, terminated) record of ONE env on the CUDA path for the
    config-1 episode (seed 3654, the fixed action tape, numpy-exact noise on the device) -- compared with the same
    record of the unmodified reference env (reference_python_baseline)."""
    import hashlib
    from rl_ptg_b200.vec_env import PtGVecEnv
    acts = episode_action_tape(kw)
    env = PtGVecEnv(kw, 1, seed=3654, device=dev, auto_reset=False)
    env.reset()
    rec = np.zeros((len(acts), 5), dtype=np.int64)
    for t, a in enumerate(acts):
        _, _, done, _ = env.step(np.array([a]))
        st = env.get_state()
        rec[t] = (st["meth_state"][0], st["i"][0], st["j"][0], st["hot_cold"][0], int(done[0]))
    env.close()
    return hashlib.sha256(rec.tobytes()).hexdigest()


def strong_scaling_leg(kw, dev, world, rank, peak, draws_per_step, n_total=1 << 20, T=32, rollouts=8):
    """BASELINE config 4 as written: 1 048 576 envs IN TOTAL, sharded over the ranks, with the episode-statistics
    reduction (ptg_episode_stats + ptg_allreduce_stats: one NCCL all-gather of the 64-byte record + combine, inside
    the library) issued once per T-step roll-out INSIDE the timed loop.  Episodes are shortened (257 steps) in this
    leg so that the gathered record is non-zero.  Three ways to issue the T steps of a roll-out."""
    import torch
    import torch.distributed as dist
    from rl_ptg_b200.vec_env import PtGVecEnv, shard_range, stats_dict
    lo, hi = shard_range(n_total, rank, world)
    n = hi - lo
    kws = dict(kw)
    kws["eps_sim_steps"] = 262
    env = PtGVecEnv(kws, n, seed=3654, device=dev, env_id_offset=lo, n_envs_global=n_total)
    env.reset_tensor()
    g = torch.Generator(device=dev)
    g.manual_seed(99 + rank)
    acts = torch.randint(0, 5, (T, n), generator=g, device=dev, dtype=torch.int64)
    out = env.rollout_tensor(acts)                      # also the pre-roll: 2 x T steps
    env.rollout_tensor(acts, out=out)
    graph = env.capture_steps([acts[t] for t in range(T)])
    stats_h = torch.zeros((rollouts, 8), dtype=torch.float64).pin_memory()
    bpe0 = env.bytes_per_env_step
    bpe_single = bpe0 + RNG_BYTES_PER_DRAW * draws_per_step
    bpe_many = bpe_single - 64 + 64 / T
    res = {}

    def eager():
        for t in range(T):
            env.step_tensor(acts[t])

    modes = (("eager_ptg_step", eager, bpe_single), ("cuda_graph_of_ptg_step", graph.replay, bpe_single),
             ("ptg_step_many", lambda: env.rollout_tensor(acts, out=out), bpe_many))
    for name, issue, bpe in modes:
        for _ in range(2):                              # warm-up roll-outs (and a first use of the communicator)
            issue()
            env.episode_stats_async(clear=True, overlap=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(rollouts):
            issue()
            # local reduction in stream; the NCCL all-gather + combine on the env's side stream, under the next roll-out
            rec = env.episode_stats_async(clear=True, overlap=True)
            with torch.cuda.stream(env.stats_stream or torch.cuda.current_stream(dev)):
                stats_h[r].copy_(rec, non_blocking=True)
        if env.stats_stream is not None:
            torch.cuda.current_stream(dev).wait_stream(env.stats_stream)     # the last collective is inside the region
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms_step = float(ms.item()) / (rollouts * T)
        tot = stats_h.numpy()
        res[name] = {"ms_per_step": ms_step, "env_steps_per_s_total": n_total / (ms_step * 1e-3),
                     "bytes_per_env_step": round(bpe, 2),
                     "roofline_frac": bpe * n / (ms_step * 1e-3) / 1e9 / peak,
                     "episodes_in_gathered_records": int(tot[:, 0].sum()),
                     "env_steps_in_gathered_records": int(tot[:, 6].sum())}
    last = stats_dict(stats_h[-1].numpy(), world)
    env.close()
    return {"envs_total": n_total, "envs_per_gpu": n, "T": T, "rollouts_timed": rollouts,
            "collective": "ptg_episode_stats (in stream) + ptg_allreduce_stats (ncclAllGather of 64 B + combine kernel, on "
                          "a side stream so that it overlaps the next roll-out) once per roll-out inside the timed loop; "
                          "the region ends after the last collective" if world > 1 else
                          "ptg_episode_stats once per roll-out inside the timed loop (one rank: nothing to gather)",
            "episode_length_in_this_leg": 257, "modes": res, "last_record": last}


def ppo_quick(kw, dev, n_envs=16384, n_steps=16, n_epochs=2, iters=2):
    """BASELINE config 5 in one small, time-boxed line (config.ppo of the default bench): the GPU VecEnv feeding the
    torch policy -- PPO with the reference's hyper-parameters except the roll-out geometry, env + VecNormalize + feature
    rows + GAE on the device, the policy update in torch (library code: the consumer of the path, not the path)."""
    import torch
    from rl_ptg_b200.ppo import PPO, reference_hyper_kwargs
    from rl_ptg_b200.vec_env import PtGVecEnv
    hyper = reference_hyper_kwargs()
    hyper.update(n_steps=n_steps, batch_size=n_envs, n_epochs=n_epochs, seed=3654)
    env = PtGVecEnv(kw, n_envs, seed=3654, device=dev, obs_layout="flat")
    model = PPO(env, **hyper)
    model.learn(n_envs * n_steps)                     # warm-up iteration (cuBLAS, allocator)
    torch.cuda.synchronize()
    t0, n0, tc = time.perf_counter(), model.num_timesteps, 0.0
    for _ in range(iters):
        c0 = time.perf_counter()
        model.collect_rollouts()
        torch.cuda.synchronize()
        tc += time.perf_counter() - c0
        model.train()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = model.num_timesteps - n0
    launches = env.kernel_launches()
    env.close()
    return {"workload": "PPO (MultiInputPolicy 2x358 ReLU) on BS2/OP2 mod: env + VecNormalize + feature rows + GAE on the "
                        "device, policy update in torch fp32", "n_envs": n_envs, "n_steps": n_steps, "batch_size": n_envs,
            "n_epochs": n_epochs, "iterations_timed": iters, "train_env_steps_per_s": steps / dt,
            "collect_env_steps_per_s": steps / tc, "collect_share_of_time": tc / dt, "env_kernel_launches": launches,
            "note": "the reference's shipped TensorBoard run logs time/fps = 166.6 env-steps/s (6 envs, unknown host); "
                    "`bench.py --workload ppo` is the full-size version of this line"}


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from rl_ptg_b200.vec_env import PtGVecEnv, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cores = len(os.sched_getaffinity(0))
    # torchrun exports OMP_NUM_THREADS=1: give every rank its share of the host cores for the converting copies
    host_threads = max(1, cores // world)
    torch.set_num_threads(host_threads)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_local = args.envs_per_gpu
    n_global = n_local * world
    lo, hi = shard_range(n_global, rank, world)
    kw = make_kwargs()
    env = PtGVecEnv(kw, hi - lo, seed=3654, device=dev, env_id_offset=lo, n_envs_global=n_global)
    env.reset_tensor()
    K, W = args.steps, args.warmup
    n_pool = 8
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    pool = torch.randint(0, 5, (n_pool, n_local), generator=g, device=dev, dtype=torch.int64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def total_draws():
        return int(env.get_state()["draws"].sum())

    # ---------------- untimed pre-roll to a stationary plant-state mix ----------------
    # A fresh batch has every env in cooldown at the same table row; the mix of plant states (and with it the gather
    # pattern and the draw rate) only becomes stationary after a few hundred random steps.  ptg_step_many, T = 8.
    T_pre = 8
    acts_pre = pool[:T_pre].contiguous()
    out_pre = env.rollout_tensor(acts_pre)
    for _ in range(max(0, args.preroll_steps // T_pre - 1)):
        env.rollout_tensor(acts_pre, out=out_pre)
    del out_pre
    preroll_done = max(1, args.preroll_steps // T_pre) * T_pre
    torch.cuda.synchronize()

    def timed(pool_, n_steps, pre):
        """`pre` untimed single steps (>= 10 ms: clocks and caches in their steady state), then n_steps timed ones."""
        d0 = total_draws()
        for t in range(pre):
            env.step_tensor(pool_[t % n_pool])
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for t in range(n_steps):
            env.step_tensor(pool_[(pre + t) % n_pool])
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / n_steps, (total_draws() - d0) / ((pre + n_steps) * n_local)

    # ---------------- headline: K single steps, device-resident inputs ----------------
    P = max(W, args.presteps)                            # untimed single steps right in front of the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    draws0 = total_draws()
    for t in range(P):
        env.step_tensor(pool[t % n_pool])
    barrier()
    # The K timed steps are bracketed by barrier + synchronize on both sides; LEAD untimed steps are queued in front of
    # the first event so that the timed launches find a busy GPU and a full launch queue -- otherwise the first timed
    # kernel waits for its own launch to cross PCIe and runs without the overlap with its predecessor's tail
    # (programmatic dependent launch), which a 20-step region would not amortise.
    LEAD = 64
    for t in range(LEAD):
        env.step_tensor(pool[(P + t) % n_pool])
    l0 = env.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(K):
        env.step_tensor(pool[(P + LEAD + t) % n_pool])
    ev1.record()
    barrier()
    launches = env.kernel_launches() - l0
    ms = ev0.elapsed_time(ev1)
    # the same K steps once more with the first event directly behind the synchronisation (no lead steps): reported
    # beside the headline as config.ms_per_step_without_lead_steps
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(K):
        env.step_tensor(pool[(P + LEAD + K + t) % n_pool])
    s1.record()
    barrier()
    ms_strict = s0.elapsed_time(s1)
    draws_per_step = (total_draws() - draws0) / ((P + LEAD + 2 * K) * n_local)
    # secondary action distributions: the load-biased mix, and an agent-like "sticky" policy (each
    # env repeats one action, so almost no transition noise is drawn after warm-up)
    Ks = max(200, K)
    probs = torch.tensor([.05, .05, .3, .3, .3], device=dev)
    biased = torch.multinomial(probs, n_pool * n_local, replacement=True, generator=g).view(n_pool, n_local)
    biased_ms, biased_draws = timed(biased, Ks, 400)
    del biased
    sticky = pool[:1].repeat(n_pool, 1).contiguous()
    sticky_ms, sticky_draws = timed(sticky, Ks, 400)
    del sticky
    long_ms, _ = timed(pool, max(1000, K), 200)          # the same uniform workload over a long timed region
    clocks = sampler.stop() if rank == 0 else None
    env.poll_error()
    stats = env.episode_stats(clear=True)
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())

    # ---------------- this box's own copy bandwidth (information only; the roofline peak stays MEASURED_PEAKS.json) ----
    copy_gbs = None
    if rank == 0:
        ca = torch.empty(1 << 28, dtype=torch.float32, device=dev)
        cb = torch.empty_like(ca)
        best = 1e9
        for _ in range(6):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            cb.copy_(ca)
            c1.record()
            torch.cuda.synchronize()
            best = min(best, c0.elapsed_time(c1))
        copy_gbs = 2 * ca.numel() * 4 / (best * 1e-3) / 1e9
        del ca, cb

    box_probe = None
    if rank == 0:
        import ctypes
        from rl_ptg_b200 import _lib
        out4 = (ctypes.c_double * 4)()
        _lib.check(_lib.load().ptg_probe_box(local_rank, ctypes.byref(out4)))
        box_probe = {"l2_dependent_load_ns": round(out4[0], 1), "dram_dependent_load_ns": round(out4[1], 1),
                     "sm_clock_mhz_seen_by_a_thread": round(out4[2], 1),
                     "note": "the step kernel is bound by L1TEX wavefronts and latency, not by HBM bandwidth: boxes of this "
                             "pool with the same copy bandwidth run the same binary up to 14 % apart (DESIGN.md section 7)"}

    # ---------------- rollout kernel (T steps per launch), reported in config ----------------
    T = 16
    acts_T = pool[:min(T, n_pool)].repeat((T + n_pool - 1) // n_pool, 1)[:T].contiguous()
    roll_ms = None
    if args.rollout:
        out = env.rollout_tensor(acts_T)
        for _ in range(8):
            env.rollout_tensor(acts_T, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        reps = max(8, K // T)
        for _ in range(reps):
            env.rollout_tensor(acts_T, out=out)
        e1.record()
        torch.cuda.synchronize()
        roll_ms = e0.elapsed_time(e1) / (reps * T)
        del out

    # ---------------- the flat observation layout (what a GPU policy consumes), reported in config ----------------
    flat = None
    if args.flat_line:
        fenv = PtGVecEnv(kw, hi - lo, seed=3654, device=dev, env_id_offset=lo, n_envs_global=n_global, obs_layout="flat")
        fenv.reset_tensor()
        fout = fenv.rollout_tensor(acts_pre)
        for _ in range(max(0, args.preroll_steps // T_pre - 1)):
            fenv.rollout_tensor(acts_pre, out=fout)
        del fout
        fres = {}
        for name, pl in (("uniform", pool), ("sticky", pool[:1].repeat(n_pool, 1).contiguous())):
            for t in range(400):
                fenv.step_tensor(pl[t % n_pool])
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for t in range(max(200, K)):
                fenv.step_tensor(pl[t % n_pool])
            f1.record()
            torch.cuda.synchronize()
            fres[name] = f0.elapsed_time(f1) / max(200, K)
        flat = {"bytes_per_env_step_without_rng": fenv.bytes_per_env_step, "feature_dim": fenv.feature_dim,
                "ms_per_step_uniform": fres["uniform"], "ms_per_step_sticky": fres["sticky"]}
        fenv.close()
        del fenv

    # ---------------- end to end through the numpy API ----------------
    acts_h = [pool[q].cpu().numpy() for q in range(n_pool)]
    Ke = max(3, min(K, args.e2e_steps))
    for t in range(6):
        env.step(acts_h[t % n_pool])
    barrier()
    d2h0 = env.d2h_bytes
    t0 = time.perf_counter()
    for t in range(Ke):
        obs, rew, done, infos = env.step(acts_h[t % n_pool])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    d2h_measured = (env.d2h_bytes - d2h0) / Ke
    t_e = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    h2d = env.h2d_bytes_per_step
    d2h = d2h_measured          # counted from the tensors copied: the window blocks travel only when they changed
    d2h_full = env.obs_elems * 4 + n_local * 4 + n_local
    bpe0 = env.bytes_per_env_step
    obs_dim = env.obs_dim
    env.close()
    del pool

    # ---------------- BASELINE config 4 as written: 1M envs in total over the ranks, collective in the timed loop
    peak, peak_src = measured_peak_gbs()
    strong = None
    if args.strong:
        strong = strong_scaling_leg(kw, dev, world, rank, peak, draws_per_step)

    ppo = None
    if world == 1 and args.ppo_line:
        ppo = ppo_quick(kw, dev)

    if rank == 0:
        bpe = bpe0 + RNG_BYTES_PER_DRAW * draws_per_step     # + the RNG state of the env-steps that draw noise
        kernel_ms = ms / K                                   # rank 0's own kernel time (CUDA events, same stream)
        achieved = bpe * n_local / (kernel_ms * 1e-3) / 1e9

        def frac_of(ms_, draws_):
            return (bpe0 + RNG_BYTES_PER_DRAW * draws_) * n_local / (ms_ * 1e-3) / 1e9 / peak

        traffic = None
        tps = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("traffic_r")) \
            if os.path.isdir(os.path.join(ROOT, "profiles")) else []
        if tps:
            with open(os.path.join(ROOT, "profiles", tps[-1])) as fh:
                traffic = json.load(fh).get("dram_bytes_per_launch")
        cpu = cpu_port = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_bench
            n_cpu = 4096 * max(1, cores // 4)
            port_rate, port_dt = cpu_oracle_rate(kw, n_cpu, 60, cores, warm=2)
            cpu_port = {"value": port_rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"CPU oracle (C restatement of PTGEnv, pthreads), {n_cpu} envs x 60 lock-steps = {port_dt:.1f} s"}
            if ref_bench.reference_root() is not None:
                ref = reference_python_baseline(kw, cores, args.cpu_lock_steps, repeats=3, record=True)
                sha_cuda = cuda_trajectory_sha(kw, dev)
                ref_value, ref_which = reference_headline(ref)
                cpu = {"value": ref_value, "unit": UNIT, "cores": cores, "kind": "reference",
                       "arrangement_reported": ref_which, "subproc_steps_per_s": ref["subproc_steps_per_s"],
                       "sample": f"unmodified reference PTGEnv (baseline/_ref): {cores} forked workers, SubprocVecEnv "
                                 f"protocol over pipes, {args.cpu_lock_steps} lock-steps = {ref['subproc_seconds']:.1f} s; "
                                 f"+ one env, one {ref['single_env_steps']}-step training episode, best of 3 = "
                                 f"{ref['single_env_episode_s']:.2f} s",
                       "single_env_steps_per_s": ref["single_env_steps_per_s"],
                       "dummy_vec_env_6_steps_per_s": ref["dummy6_steps_per_s"],
                       "trajectory_sha256_reference": ref["trajectory_sha256"], "trajectory_sha256_cuda": sha_cuda,
                       "trajectory_match": ref["trajectory_sha256"] == sha_cuda}
            else:
                cpu = dict(cpu_port, note="baseline/_ref missing on this box (tools/stage_reference.sh): port timed instead")
        line = {
            "metric": METRIC, "value": n_global * K / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": n_local, "envs_total": n_global,
                       "obs_dim": obs_dim, "noise": "on-device PCG64+ziggurat (numpy-exact)",
                       "preroll_steps": preroll_done, "untimed_single_steps_before_timed_region": P,
                       "untimed_steps_queued_ahead_of_the_first_event": LEAD,
                       "ms_per_step_without_lead_steps": ms_strict / K,
                       "copy_gbs_this_box": copy_gbs, "box_probe": box_probe,
                       "bytes_per_env_step": round(bpe, 2),
                       "bytes_per_env_step_breakdown": {"state_action_obs_reward_done": bpe0,
                                                        "rng_state_per_draw": RNG_BYTES_PER_DRAW,
                                                        "draws_per_env_step": round(draws_per_step, 4)},
                       "l2": f"per-step traffic {bpe * n_local / 1e6:.0f} MB > 126 MB L2 (inputs larger than L2); plant "
                             "state, RNG records and tables are kept L2-resident on purpose (evict_last), observations "
                             "/ rewards / actions stream (evict_first)",
                       "actions": "uniform random over the 5 actions (worst case: ~40% of env-steps redraw noise)",
                       "roofline_frac_excluding_rng_bytes": bpe0 * n_local / (kernel_ms * 1e-3) / 1e9 / peak,
                       "uniform_long_run": {"steps": max(1000, K), "ms_per_step": long_ms,
                                            "roofline_frac": frac_of(long_ms, draws_per_step)},
                       "load_biased_mix": {"p": [.05, .05, .3, .3, .3], "ms_per_step": biased_ms,
                                           "draws_per_env_step": round(biased_draws, 4),
                                           "roofline_frac": frac_of(biased_ms, biased_draws)},
                       "sticky_policy": {"ms_per_step": sticky_ms, "draws_per_env_step": round(sticky_draws, 4),
                                         "roofline_frac": frac_of(sticky_ms, sticky_draws)},
                       "rollout_kernel": None if roll_ms is None else {
                           "T": T, "ms_per_step": roll_ms,
                           "bytes_per_env_step": round(bpe - 64 + 64 / T, 2),     # state stays in registers between steps
                           "roofline_frac": (bpe - 64 + 64 / T) * n_local / (roll_ms * 1e-3) / 1e9 / peak},
                       "flat_layout": None if flat is None else dict(
                           flat, note="obs_layout='flat': the step kernel writes the [n_envs, 40] MultiInputPolicy feature rows "
                                      "(one bulk store per warp) instead of the key-major Dict blocks",
                           roofline_frac_uniform=(flat["bytes_per_env_step_without_rng"] + RNG_BYTES_PER_DRAW * draws_per_step)
                           * n_local / (flat["ms_per_step_uniform"] * 1e-3) / 1e9 / peak,
                           roofline_frac_sticky=flat["bytes_per_env_step_without_rng"] * n_local
                           / (flat["ms_per_step_sticky"] * 1e-3) / 1e9 / peak),
                       "strong_scaling": strong,
                       "ppo": ppo,
                       "host_threads_per_rank": host_threads,
                       "episodes_finished": stats["episodes"]},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full "
                         "(cold-cache replay; profiles/traffic_r*.json); below the algorithmic bytes because plant state and "
                         "RNG records are served from / absorbed by the L2 (DESIGN.md section 3)",
                         "peak_source": peak_src, "kernel": "k_step<4,true,false,false,13>",
                         "kernel_ms": kernel_ms},
            "cpu_baseline": cpu,
            "cpu_baseline_port": cpu_port,
            "e2e": {"value": n_global * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "d2h_bytes_per_step_if_every_block_travelled": d2h_full,
                    "steps": Ke, "note": "numpy VecEnv.step(): int64 action array in, uint8 on the wire; METH_STATUS "
                                         "comes back as one byte per env, the clock-encoding blocks not at all while all "
                                         "envs share one clock, the market-window blocks (109 of 147 MB) only on steps "
                                         "that move them (clock crosses an hour / episode end)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json config 5: end-to-end PPO training, GPU VecEnv feeding the torch policy vs the CPU env baseline
# ---------------------------------------------------------------------------------------------------------------
PPO_METRIC = "PPO training env-steps/sec"


def run_ppo_gpu(args):
    """PPO (SB3 algorithm, reference hyper-parameters except the roll-out geometry) with env, VecNormalize, feature
    rows and GAE on the device.  One "step" = one PPO iteration (collect n_steps x n_envs, then n_epochs of updates)."""
    import torch
    import torch.distributed as dist
    from rl_ptg_b200.ppo import PPO, reference_hyper_kwargs
    from rl_ptg_b200.vec_env import PtGVecEnv, shard_range
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:           # data-parallel PPO: env shard + policy replica per rank, gradients averaged over NCCL
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.ppo_tf32:                      # policy matmuls on the tensor cores (SB3's default is plain fp32)
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    kw = make_kwargs()
    hyper = reference_hyper_kwargs()
    hyper.update(n_steps=args.ppo_n_steps, batch_size=args.ppo_batch, n_epochs=args.ppo_epochs, seed=3654)
    n_global = args.ppo_envs * world
    lo, hi = shard_range(n_global, rank, world)
    env = PtGVecEnv(kw, hi - lo, seed=3654, device=dev, obs_layout=args.ppo_layout, env_id_offset=lo, n_envs_global=n_global)
    model = PPO(env, **hyper)
    per_iter = args.ppo_envs * args.ppo_n_steps
    model.learn(per_iter * max(1, args.warmup if args.warmup < 3 else 1))      # warm-up iteration(s): cuBLAS, allocator
    torch.cuda.synchronize()
    t0, n0 = time.perf_counter(), model.num_timesteps
    tc = 0.0
    iters = max(1, min(args.steps, args.ppo_iters))
    for _ in range(iters):
        c0 = time.perf_counter()
        model.collect_rollouts()
        torch.cuda.synchronize()
        tc += time.perf_counter() - c0
        model.train()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps_done = model.num_timesteps - n0
    ep = env.episode_stats(clear=True, reduce=False)
    if world > 1:
        t_all = torch.tensor([dt, tc], device=dev, dtype=torch.float64)
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dt, tc = float(t_all[0]), float(t_all[1])
        steps_done *= world
        if rank != 0:
            env.close()
            dist.destroy_process_group()
            return
    _emit({"metric": PPO_METRIC, "value": steps_done / dt, "unit": UNIT, "n_gpus": world, "steps": iters,
           "warmup": 1, "ms_per_step": 1e3 * dt / iters, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": steps_done / dt / 166.6, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "PPO (MultiInputPolicy 2x358 ReLU) on BS2/OP2 mod, GPU VecEnv + VecNormalize + "
                                  "feature rows + GAE on device", "n_envs": args.ppo_envs * world, "n_steps": args.ppo_n_steps,
                      "parallelism": f"dp{world}: env shard + policy replica per GPU, NCCL gradient all-reduce per mini-batch",
                      "batch_size": args.ppo_batch, "n_epochs": args.ppo_epochs,
                      "policy_matmul": "tf32" if args.ppo_tf32 else "fp32", "obs_layout": args.ppo_layout,
                      "collect_env_steps_per_s": steps_done / tc, "collect_share_of_time": tc / dt,
                      "baseline_note": "vs_baseline = value / 166.6 env-steps/s, the reference's shipped TensorBoard "
                                       "time/fps (BASELINE.md; unknown hardware, 6 envs)",
                      "episodes_finished": ep["episodes"], "ep_rew_mean": ep["return_mean"]},
           "gpu_launches": env.kernel_launches()})
    env.close()
    if world > 1:
        dist.destroy_process_group()


def run_ppo_reference(args):
    """The reference's arrangement: the env on the host cores (CPU oracle, one thread per env like SubprocVecEnv
    workers, without the pickling), observations to the torch policy and actions back every step, VecNormalize and
    GAE in numpy (SB3's own code path, restated), the same PPO update.  Reference hyper-parameters
    (config_agent.yaml: 6 envs, n_steps 4263, batch 203, 13 epochs) unless overridden."""
    import torch
    from oracle.ptg_oracle import OracleVecEnv, draw_noise_tape
    from oracle.sb3_restated import VecNormalizeRewardRef, gae_ref
    from rl_ptg_b200.ppo import PPOCore, reference_hyper_kwargs
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    kw = make_kwargs()
    n, T = args.ppo_ref_envs, args.ppo_ref_n_steps
    hyper = reference_hyper_kwargs()
    hyper.update(n_steps=T, seed=3654)
    pa = int(kw["price_ahead"])
    F = 14 + 2 * pa
    # oracle obs columns (reference key order) -> sorted-key feature columns
    o_pot, o_pf, o_st = np.arange(0, pa), np.arange(pa, 2 * pa), 2 * pa
    o_T, o_h2, o_ch4, o_h2res, o_h2o, o_heat, o_sin, o_cos = (2 * pa + 1 + q for q in range(8))
    iters = max(1, min(args.steps, args.ppo_iters))
    tape = draw_noise_tape(3654 + np.arange(n), kw["noise"], T * (iters + 1) + 8)
    cores = len(os.sched_getaffinity(0))
    env = OracleVecEnv(kw, n, noise_tape=tape, threads=min(cores, n))
    model = PPOCore(n, F, dev, **hyper)
    vn = VecNormalizeRewardRef(n)

    def features(obs):
        f = np.zeros((n, F), np.float32)
        f[:, 0], f[:, 1], f[:, 2], f[:, 3], f[:, 4] = obs[:, o_ch4], obs[:, o_heat], obs[:, o_h2o], obs[:, o_h2], obs[:, o_h2res]
        f[np.arange(n), 5 + obs[:, o_st].astype(np.int64)] = 1.0
        f[:, 11:11 + pa], f[:, 11 + pa:11 + 2 * pa] = obs[:, o_pf], obs[:, o_pot]
        f[:, 11 + 2 * pa], f[:, 12 + 2 * pa], f[:, 13 + 2 * pa] = obs[:, o_T], obs[:, o_cos], obs[:, o_sin]
        return f

    obs = env.reset().copy()
    starts = np.ones(n, np.float32)
    rew_buf, val_buf, st_buf = np.zeros((T, n), np.float32), np.zeros((T, n), np.float32), np.zeros((T, n), np.float32)
    t_all = t_col = 0.0
    for it in range(iters + 1):                     # iteration 0 is the warm-up
        c0 = time.perf_counter()
        with torch.no_grad():
            for t in range(T):
                feat = torch.from_numpy(features(obs)).to(dev)
                actions, values, logp = model.policy(feat)
                model.buf_feat[t].copy_(feat); model.buf_actions[t].copy_(actions); model.buf_logp[t].copy_(logp)
                val_buf[t] = values.cpu().numpy(); st_buf[t] = starts
                o, r, d = env.step(actions.cpu().numpy())
                rew_buf[t] = vn.step(r.astype(np.float32), d.astype(bool))
                starts = d.astype(np.float32)
                obs = o.copy()
            last_values = model.policy.value(torch.from_numpy(features(obs)).to(dev)).cpu().numpy()
        adv, ret = gae_ref(rew_buf, val_buf, st_buf, last_values, starts.astype(bool), model.gamma, model.gae_lambda)
        model.buf_values.copy_(torch.from_numpy(val_buf)); model.buf_adv.copy_(torch.from_numpy(adv))
        model.buf_ret.copy_(torch.from_numpy(ret))
        c1 = time.perf_counter()
        model.train()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if it > 0:
            t_col += c1 - c0
            t_all += time.perf_counter() - c0
    rate = iters * T * n / t_all
    _emit({"impl": "reference", "metric": PPO_METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": iters,
           "warmup": 1, "ms_per_step": 1e3 * t_all / iters, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": rate / 166.6, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "PPO (MultiInputPolicy 2x358 ReLU) on BS2/OP2 mod, env on host cores (CPU oracle), "
                                  "policy on " + dev.type, "n_envs": n, "n_steps": T, "batch_size": hyper["batch_size"],
                      "n_epochs": hyper["n_epochs"], "collect_env_steps_per_s": iters * T * n / t_col,
                      "collect_share_of_time": t_col / t_all},
           "cpu_baseline": {"value": rate, "unit": UNIT, "cores": min(cores, n), "kind": "port",
                            "sample": f"{iters} PPO iteration(s) of {T} x {n} env-steps"},
           "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
    env.close()


_REAL_STDOUT = None


def _quiet_stdout():
    """Route fd 1 to stderr for the whole run (NCCL / libraries print banners to stdout); the ONE JSON line is
    written to the saved original stdout by _emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--preroll-steps", type=int, default=600, help="untimed ptg_step_many steps to a stationary state mix")
    ap.add_argument("--presteps", type=int, default=1024, help="untimed single steps right before the timed region (>= 10 ms)")
    ap.add_argument("--strong", action="store_true", default=True, help="BASELINE config 4 leg: 1M envs in total")
    ap.add_argument("--no-strong", dest="strong", action="store_false")
    ap.add_argument("--no-flat-line", dest="flat_line", action="store_false", default=True,
                    help="skip config.flat_layout (the same step with the flat feature-row observation layout)")
    ap.add_argument("--no-ppo-line", dest="ppo_line", action="store_false", default=True,
                    help="skip config.ppo (BASELINE config 5 as one small time-boxed line)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--cpu-lock-steps", type=int, default=2000, help="lock-steps of the Python reference under the Subproc harness")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rollout", action="store_true", default=True)
    ap.add_argument("--no-rollout", dest="rollout", action="store_false")
    ap.add_argument("--workload", default="env", choices=["env", "ppo"],
                    help="env: the north-star env-steps/s bench (default) | ppo: BASELINE.json config 5")
    ap.add_argument("--ppo-envs", type=int, default=65536)
    ap.add_argument("--ppo-n-steps", type=int, default=32)
    ap.add_argument("--ppo-batch", type=int, default=65536)
    ap.add_argument("--ppo-epochs", type=int, default=13)
    ap.add_argument("--ppo-iters", type=int, default=4)
    ap.add_argument("--ppo-tf32", action="store_true")
    ap.add_argument("--ppo-layout", default="flat", choices=["flat", "dict"],
                    help="flat: the step kernel writes the policy's feature rows | dict: key-major obs + ptg_features")
    ap.add_argument("--ppo-ref-envs", type=int, default=6)
    ap.add_argument("--ppo-ref-n-steps", type=int, default=4263)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "ppo":
        run_ppo_reference(args) if args.impl == "reference" else run_ppo_gpu(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
