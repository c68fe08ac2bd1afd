"""torchrun check of ptg_allreduce_stats (the path's only collective, inside the C ABI) on N GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_allreduce_stats.py

Every rank steps its shard of one global batch through short episodes, reduces the finished-episode statistics on the
device and all-gathers + combines the 64-byte records over NCCL inside the library.  Checked: every rank holds the
bit-identical combined record, and it equals the fixed-order host combine (ptg_stats_combine) of the per-rank records
gathered through torch.distributed -- and the episode count / env-step count of the whole batch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rl_ptg_b200.vec_env import PtGVecEnv, combine_stats, shard_range, stats_dict  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
kw = dict(bench.make_kwargs())
kw["eps_sim_steps"] = 30                        # episodes of 25 steps
n_global, steps = 10007, 80                     # ragged shards
lo, hi = shard_range(n_global, rank, world)
env = PtGVecEnv(kw, hi - lo, seed=3654, device=dev, env_id_offset=lo, n_envs_global=n_global)
env.reset_tensor()
g = torch.Generator(device=dev)
g.manual_seed(5 + rank)
for t in range(steps):
    env.step_tensor(torch.randint(0, 5, (hi - lo,), generator=g, device=dev))
local = env.episode_stats_async(clear=False, reduce=False).clone()           # this rank's record
want = combine_stats(local, reduce=True)                                    # torch all_gather + ptg_stats_combine
rec = env.episode_stats_async(clear=True, reduce=True).clone()               # ptg_allreduce_stats (NCCL, in-library)
torch.cuda.synchronize()
got = stats_dict(rec.cpu().numpy(), world)
all_recs = [torch.empty_like(rec) for _ in range(world)]
dist.all_gather(all_recs, rec)
same = all(torch.equal(all_recs[0], r) for r in all_recs)
ok = same and got == want and got["episodes"] == n_global * (steps // 25) and got["env_steps"] == n_global * steps
# the overlapped form: local reduction in stream, all-gather + combine on the env's side stream, two records in flight
for t in range(50):
    env.step_tensor(torch.randint(0, 5, (hi - lo,), generator=g, device=dev))
r1 = env.episode_stats_async(clear=True, overlap=True)
for t in range(25):
    env.step_tensor(torch.randint(0, 5, (hi - lo,), generator=g, device=dev))
r2 = env.episode_stats_async(clear=True, overlap=True)
torch.cuda.synchronize()
d1, d2 = stats_dict(r1.cpu().numpy(), world), stats_dict(r2.cpu().numpy(), world)
ok = ok and d1["episodes"] == n_global * 2 and d1["env_steps"] == n_global * 50 \
    and d2["episodes"] == n_global and d2["env_steps"] == n_global * 25 and d1["ranks"] == world
if rank == 0:
    print(f"ptg_allreduce_stats on {world} GPUs: {'OK' if ok else 'MISMATCH'} {got} overlapped: {d1['episodes']} / {d2['episodes']} episodes")
env.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
