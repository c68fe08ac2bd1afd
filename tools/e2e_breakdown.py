"""Where the time of the numpy VecEnv.step() goes at 1M envs (host phases, PCIe copies)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rl_ptg_b200.vec_env import PtGVecEnv
n = 1 << 20
env = PtGVecEnv(bench.make_kwargs(), n, seed=3654)
env.reset_tensor()
acts = [np.random.default_rng(q).integers(0, 5, n) for q in range(4)]
for t in range(5):
    env.step(acts[t % 4])
def T(f, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
print("step() total            %.2f ms" % T(lambda: env.step(acts[0])))
print("step_async              %.2f ms" % T(lambda: env.step_async(acts[0])))
h = env._obs_hh[0]
print("D2H obs 146.8 MB pinned %.2f ms" % T(lambda: h.copy_(env._obs, non_blocking=True)))
a = env._act_h.numpy()
print("host copy actions 8 MB  %.2f ms" % T(lambda: a.__setitem__(slice(None), acts[1])))
print("H2D actions             %.2f ms" % T(lambda: env._act_d.copy_(env._act_h, non_blocking=True)))
print("_obs_numpy              %.2f ms" % T(lambda: env._obs_numpy(h)))
print("poll_error              %.2f ms" % T(lambda: env.poll_error()))
big = torch.empty(1 << 28, dtype=torch.uint8, device=env.device); hb = torch.empty(1 << 28, dtype=torch.uint8).pin_memory()
ms = T(lambda: hb.copy_(big, non_blocking=True), 5)
print("PCIe D2H 256 MiB        %.2f ms = %.1f GB/s" % (ms, (1 << 28) / ms / 1e6))
ms = T(lambda: big.copy_(hb, non_blocking=True), 5)
print("PCIe H2D 256 MiB        %.2f ms = %.1f GB/s" % (ms, (1 << 28) / ms / 1e6))
