"""Kernel-only timing of ptg_step / ptg_step_many for kernel experiments (PTG_B200_SO selects a build variant)."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rl_ptg_b200.vec_env import PtGVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--T", type=int, default=16)
ap.add_argument("--policy", default="both", choices=["uniform", "sticky", "both"])
ap.add_argument("--no-rollout", action="store_true")
ap.add_argument("--noise", default="numpy", choices=["numpy", "off"])
ap.add_argument("--layout", default="dict", choices=["dict", "flat"])
args = ap.parse_args()
kw = bench.make_kwargs()
env = PtGVecEnv(kw, args.envs, seed=3654, noise=args.noise, obs_layout=args.layout)
env.reset_tensor()
dev = env.device
g = torch.Generator(device=dev); g.manual_seed(0)
upool = torch.randint(0, 5, (8, args.envs), generator=g, device=dev, dtype=torch.int64)
pre = env.rollout_tensor(upool)                      # pre-roll to a stationary plant-state mix (as bench.py does)
for _ in range(60):
    env.rollout_tensor(upool, out=pre)
del pre
bpe = env.bytes_per_env_step
tag = os.path.basename(os.environ.get('PTG_B200_SO', 'default'))
for policy in (["uniform", "sticky"] if args.policy == "both" else [args.policy]):
    # sticky = an agent-like policy: actions change rarely (each env repeats one action)
    pool = upool if policy == "uniform" else upool[:1].repeat(8, 1).contiguous()
    for t in range(400):
        env.step_tensor(pool[t % 8])
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(args.steps):
            env.step_tensor(pool[t % 8])
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / args.steps)
    ms = best
    print(f"so={tag} noise={args.noise} layout={args.layout} {policy:8s} step: {ms*1e3:.1f} us/step  {args.envs/ms/1e6:.2f} G env-steps/s  "
          f"{bpe*args.envs/ms/1e6:.0f} GB/s ({bpe} B/env-step)", flush=True)
    if args.no_rollout:
        continue
    acts = pool[:args.T % 9 or 8].repeat((args.T + 7) // 8, 1)[:args.T].contiguous()
    out = env.rollout_tensor(acts)
    torch.cuda.synchronize()
    e0.record()
    reps = max(1, args.steps // args.T)
    for _ in range(reps):
        env.rollout_tensor(acts, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * args.T)
    print(f"   {policy} rollout T={args.T}: {ms*1e3:.1f} us/step  {args.envs/ms/1e6:.2f} G env-steps/s", flush=True)
    del out
env.poll_error()
