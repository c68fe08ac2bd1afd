"""N > 1 host logic on CPU (gloo, world_size 2): env-id sharding and the episode-statistics all-gather + combine.
The data path itself has no collective (envs are independent); this is the only cross-rank exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_ptg_b200.vec_env import combine_stats, shard_range


def test_shard_range_partitions_exactly():
    for n, w in [(1 << 20, 8), (1000, 3), (7, 8), (65536, 4), (5, 1)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank r finished (r + 1) episodes with returns 10*r + q
        rets = np.array([10.0 * rank + q for q in range(rank + 1)])
        stats = torch.tensor([len(rets), rets.sum(), (rets ** 2).sum(), 35.0 * len(rets), rets.min(), rets.max(),
                              1000.0 * (rank + 1), 0.0], dtype=torch.float64)
        out = combine_stats(stats, reduce=True)
        local = combine_stats(stats, reduce=False)
        ret[rank] = (out, local)
    finally:
        dist.destroy_process_group()


def test_stats_allgather_combine_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        results = dict(ret)
    all_rets = np.array([0.0, 10.0, 11.0])
    for rank in range(world):
        out, local = results[rank]
        assert out["ranks"] == 2 and out["episodes"] == 3 and out["env_steps"] == 3000
        assert out["return_mean"] == pytest.approx(all_rets.mean())
        assert out["return_std"] == pytest.approx(all_rets.std())
        assert out["return_min"] == 0.0 and out["return_max"] == 11.0 and out["length_mean"] == 35.0
        assert local["ranks"] == 1 and local["episodes"] == rank + 1
    assert results[0][0] == results[1][0]          # every rank computes the identical combined record
