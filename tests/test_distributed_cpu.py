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


def _ppo_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rl_ptg_b200.ppo import PPOCore
        m = PPOCore(4, 40, "cpu", n_steps=8, batch_size=32, n_epochs=1, seed=5)
        w0 = torch.cat([p.detach().reshape(-1) for p in m.policy.parameters()]).clone()
        g = torch.Generator().manual_seed(100 + rank)             # different roll-out data on every rank
        m.buf_feat.copy_(torch.randn(m.buf_feat.shape, generator=g))
        m.buf_actions.copy_(torch.randint(0, 5, m.buf_actions.shape, generator=g))
        m.buf_values.copy_(torch.randn(m.buf_values.shape, generator=g))
        m.buf_logp.fill_(-1.6)
        m.buf_adv.copy_(torch.randn(m.buf_adv.shape, generator=g))
        m.buf_ret.copy_(torch.randn(m.buf_ret.shape, generator=g))
        m.train()                                                 # one mini-batch: one averaged-gradient update
        w1 = torch.cat([p.detach().reshape(-1) for p in m.policy.parameters()])
        ret[rank] = (w0.numpy(), w1.numpy())
    finally:
        dist.destroy_process_group()


def test_data_parallel_ppo_keeps_replicas_identical():
    """Data-parallel PPO (BASELINE config 5 on N GPUs): same initial policy on every rank, gradients of each
    mini-batch averaged across ranks -> the replicas stay bit-identical although every rank sees different data."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ppo_worker, args=(world, port, ret), nprocs=world, join=True)
        results = dict(ret)
    assert np.array_equal(results[0][0], results[1][0])           # broadcast initial weights
    assert np.array_equal(results[0][1], results[1][1])           # identical after the averaged update
    assert not np.array_equal(results[0][0], results[0][1])       # and the update did move them


def _ppo_uneven_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rl_ptg_b200.ppo import PPOCore
        lo, hi = shard_range(9, rank, world)                      # 5 + 4 envs
        try:
            PPOCore(hi - lo, 40, "cpu", n_steps=8, batch_size=16, n_epochs=1, seed=5)
            ret[rank] = "constructed"
        except ValueError as e:
            ret[rank] = str(e)
    finally:
        dist.destroy_process_group()


def test_data_parallel_ppo_rejects_uneven_shards():
    """An env count that does not divide by the world size gives ranks different mini-batch counts, i.e. different
    numbers of gradient all-reduces: refused at construction on every rank instead of hanging later."""
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ppo_uneven_worker, args=(world, port, ret), nprocs=world, join=True)
        results = dict(ret)
    assert all("identical (n_envs" in results[r] for r in range(world)), results
