"""Where the time of the numpy VecEnv.step() goes at 1M envs per rank (host phases, PCIe copies), at N ranks:

    python tools/e2e_breakdown.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 \
        tools/e2e_breakdown.py

Every phase is timed on all ranks at once (barrier before each) and reported as the max over ranks."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rl_ptg_b200.vec_env import PtGVecEnv  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
cores = len(os.sched_getaffinity(0))
torch.set_num_threads(max(1, cores // world))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
env = PtGVecEnv(bench.make_kwargs(), n, seed=3654, device=dev, env_id_offset=rank * n, n_envs_global=world * n)
env.reset_tensor()
acts = [np.random.default_rng(q + 10 * rank).integers(0, 5, n) for q in range(4)]
for t in range(6):
    env.step(acts[t % 4])


def T(f, reps=12):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


rows = []
rows.append(("step() total", T(lambda: env.step(acts[0]))))
rows.append(("step_async (host copy + H2D + launch)", T(lambda: env.step_async(acts[0]))))
h = env._obs_hh[0]
cut, mid = env._scalar_off, env._scalar_off + env._status_elems
rows.append(("D2H scalar blocks (36 MB) pinned", T(lambda: h[cut:].copy_(env._obs[cut:], non_blocking=True))))
rows.append(("D2H window blocks (109 MB) pinned", T(lambda: h[:cut].copy_(env._obs[:cut], non_blocking=True))))
rows.append(("D2H rewards + dones (5 MB)", T(lambda: (env._reward_hh[0].copy_(env._reward, non_blocking=True),
                                                       env._done_hh[0].copy_(env._done, non_blocking=True)))))
a_t = torch.from_numpy(acts[1])
rows.append(("host int64 -> uint8 converting copy (8 MB in)", T(lambda: env._act_h.copy_(a_t))))
rows.append(("aminmax validity scan of the actions", T(lambda: torch.aminmax(a_t))))
st = env._obs_views(h)["METH_STATUS"].numpy()
rows.append(("METH_STATUS int32 -> int64 (numpy)", T(lambda: st.astype(np.int64))))
rows.append(("dones.any()", T(lambda: env._done_hh[0].numpy().view(np.bool_).any())))
rows.append(("_obs_numpy (views)", T(lambda: env._obs_numpy(h))))
rows.append(("poll_error (4 B D2H + sync)", T(lambda: env.poll_error())))
if rank == 0:
    print(f"ranks={world} host cores={cores} torch threads per rank={torch.get_num_threads()}  (ms, max over ranks)")
    for name, ms in rows:
        print(f"  {name:48s} {ms:8.3f}")
    print(f"  => e2e {world * n / (rows[0][1] * 1e-3):.3e} env-steps/s")
env.close()
if world > 1:
    dist.destroy_process_group()
