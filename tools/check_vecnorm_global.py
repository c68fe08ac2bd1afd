"""torchrun check (N GPUs): VecNormalizeReward(reduce="global") exchanges the 24-byte moment records over NCCL; every
rank must end up with the same reward statistics, equal to what ONE env over all shards computes."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rl_ptg_b200.vec_env import PtGVecEnv, shard_range
from rl_ptg_b200.vec_normalize import VecNormalizeReward

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
kw = bench.make_kwargs()
n_global = 20000
lo, hi = shard_range(n_global, rank, world)
env = PtGVecEnv(kw, hi - lo, seed=3654, device=dev, env_id_offset=lo, n_envs_global=n_global)
vn = VecNormalizeReward(env, reduce="global")
vn.reset_tensor()
whole = wvn = None
if rank == 0:
    whole = PtGVecEnv(kw, n_global, seed=3654, device=dev)
    wvn = VecNormalizeReward(whole)
    wvn.reset_tensor()
g = torch.Generator(device="cpu"); g.manual_seed(5)
for t in range(30):
    a = torch.randint(0, 5, (n_global,), generator=g)           # same stream on every rank
    _, r, _ = vn.step_tensor(a[lo:hi].to(dev))
    if rank == 0:
        _, rw, _ = wvn.step_tensor(a.to(dev))
        assert torch.allclose(r, rw[lo:hi], rtol=1e-6, atol=1e-9), t
mine = torch.tensor([vn.ret_rms.mean, vn.ret_rms.var, vn.ret_rms.count], dtype=torch.float64, device=dev)
allst = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allst, mine)
for s in allst:
    assert torch.equal(s, allst[0]), "ranks disagree on the reward statistics"
if rank == 0:
    ref = np.array([wvn.ret_rms.mean, wvn.ret_rms.var, wvn.ret_rms.count])
    assert np.allclose(mine.cpu().numpy(), ref, rtol=1e-11), (mine, ref)
    print(f"OK: {world} ranks agree bit for bit; statistics == single-batch run to 1e-11 (mean {ref[0]:.6f}, var {ref[1]:.6f})")
dist.destroy_process_group()
