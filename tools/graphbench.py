"""Single-step launches replayed from a CUDA graph (8 steps per graph): the launch-bound small-shard regime."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rl_ptg_b200.vec_env import PtGVecEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=131072)
ap.add_argument("--steps", type=int, default=4000)
args = ap.parse_args()
env = PtGVecEnv(bench.make_kwargs(), args.envs, seed=3654)
env.reset_tensor()
dev = env.device
g = torch.Generator(device=dev); g.manual_seed(0)
pool = torch.randint(0, 5, (8, args.envs), generator=g, device=dev, dtype=torch.int64)
for t in range(16):
    env.step_tensor(pool[t % 8])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for t in range(args.steps):
    env.step_tensor(pool[t % 8])
e1.record(); torch.cuda.synchronize()
print(f"envs={args.envs} eager : {e0.elapsed_time(e1) / args.steps * 1e3:.2f} us/step")
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for t in range(8):
        env.step_tensor(pool[t])
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for t in range(8):
        env.step_tensor(pool[t])
torch.cuda.synchronize()
for _ in range(4):
    graph.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(args.steps // 8):
    graph.replay()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (args.steps // 8 * 8)
print(f"envs={args.envs} graph : {ms * 1e3:.2f} us/step  {args.envs / ms / 1e6:.2f} G env-steps/s")
env.poll_error()
