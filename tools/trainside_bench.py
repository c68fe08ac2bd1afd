"""Device time of the train-side kernels (VecNormalize, feature rows, GAE) at 1M envs, as achieved HBM GB/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rl_ptg_b200.vec_env import PtGVecEnv
from rl_ptg_b200.vec_normalize import VecNormalizeReward, features_tensor, gae

n = 1 << 20
env = PtGVecEnv(bench.make_kwargs(), n, seed=3654)
vn = VecNormalizeReward(env)
vn.reset_tensor()
dev = env.device
a = torch.randint(0, 5, (n,), device=dev)
env.step_tensor(a)
flush = torch.empty(1 << 28, dtype=torch.uint8, device=dev)


def timed(fn, reps=50):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                                   # 256 MiB > L2: the kernel under test starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(200):
    vn.normalize_step(env._reward, env._done)
e1.record(); torch.cuda.synchronize()
print(f"vecnorm pipelined (L2-warm, 38 MB working set): {e0.elapsed_time(e1)/200*1e3:.1f} us per call")
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    vn.normalize_step(env._reward, env._done); vn.normalize_step(env._reward, env._done)
torch.cuda.current_stream().wait_stream(side)
with torch.cuda.graph(g):
    vn.normalize_step(env._reward, env._done); vn.normalize_step(env._reward, env._done)   # 2 calls: ping-pong state
torch.cuda.synchronize(); e0.record()
for _ in range(100):
    g.replay()
e1.record(); torch.cuda.synchronize()
print(f"vecnorm from a CUDA graph (device time): {e0.elapsed_time(e1)/200*1e3:.1f} us per call")
ms = timed(lambda: vn.normalize_step(env._reward, env._done))
print(f"vecnorm (2 launches)  {ms*1e3:7.1f} us   {n*25/ms/1e6:7.0f} GB/s  (25 B/env: reward r, returns r/w, done r, out w)")
feat = features_tensor(env)
ms = timed(lambda: features_tensor(env, out=feat))
print(f"features [n,40]       {ms*1e3:7.1f} us   {n*300/ms/1e6:7.0f} GB/s  (140 B in + 160 B out per env)")
T = 32
rew = torch.randn(T, n, device=dev); val = torch.randn(T, n, device=dev)
st = (torch.rand(T, n, device=dev) < 0.01).to(torch.uint8)
lv = torch.randn(n, device=dev); ld = torch.zeros(n, dtype=torch.uint8, device=dev)
adv, ret = gae(rew, val, st, lv, ld, 0.973, 0.8002)
ms = timed(lambda: gae(rew, val, st, lv, ld, 0.973, 0.8002, adv, ret), reps=20)
print(f"gae T=32              {ms*1e3:7.1f} us   {n*T*17/ms/1e6:7.0f} GB/s  (17 B per element)")
