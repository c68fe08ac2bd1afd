"""The box's host-transfer ceiling for the numpy API at N ranks (run under torchrun, or plainly for one GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tools/pcie_ceiling.py [--mb 59]

Every rank copies `mb` MB device -> pinned host (what one numpy `VecEnv.step()` of 1 M envs moves on average) and
1 MB host -> device, back to back, all ranks at once; prints the per-rank and aggregate GB/s and the env-steps/s this
bounds (`e2e` can not exceed 2^20 envs / copy time per rank).  Also the same with a single-threaded host pass over the
received bytes (the int64 widening of METH_STATUS etc. touches ~13 MB per step)."""
import argparse
import os
import time

import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=59.0)
ap.add_argument("--reps", type=int, default=40)
args = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(args.mb * 1e6)
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
a_h = torch.empty(1 << 20, dtype=torch.uint8).pin_memory()
a_d = torch.empty(1 << 20, dtype=torch.uint8, device=dev)


def run(touch: bool):
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        a_d.copy_(a_h, non_blocking=True)
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
        if touch:
            h.numpy()[:13_000_000].view(np.int32).astype(np.int64)
    dt = (time.perf_counter() - t0) / args.reps
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for touch in (False, True):
    dt = run(touch)
    if rank == 0:
        print(f"ranks={world} {'copy + 13 MB host pass' if touch else 'copy only            '}: {dt * 1e3:.3f} ms per step-sized "
              f"transfer ({args.mb:.0f} MB D2H + 1 MB H2D) = {n / dt / 1e9:.1f} GB/s per rank, {world * n / dt / 1e9:.1f} GB/s "
              f"aggregate -> e2e ceiling {world * (1 << 20) / dt:.3e} env-steps/s", flush=True)
if world > 1:
    dist.destroy_process_group()
