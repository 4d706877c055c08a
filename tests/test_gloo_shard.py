"""N>1 host logic on CPU: world_size-2 gloo run of the sharding + advantage-statistics exchange."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from settlers_of_catan_rl_b200.sharding import (adv_stats_from_tensor, allreduce_adv_stats, mean_std_from_stats,
                                                shard_range)


def test_shard_range_partitions_exactly():
    for n_total, world in [(524288, 8), (65536, 1), (10, 3), (7, 8), (131072, 2)]:
        got = [shard_range(n_total, r, world) for r in range(world)]
        assert got[0][0] == 0
        for (f0, n0), (f1, _) in zip(got, got[1:]):
            assert f0 + n0 == f1
        assert got[-1][0] + got[-1][1] == n_total
        assert max(n for _, n in got) - min(n for _, n in got) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        T = 20
        full = torch.randn(T, n_total, generator=g) * 37.0 + 5.0      # every rank can rebuild the global tensor
        first, n = shard_range(n_total, rank, world)
        local = full[:, first:first + n].contiguous()
        stats = allreduce_adv_stats(adv_stats_from_tensor(local))
        mean, std = mean_std_from_stats(stats)
        mine = (local - mean) / (std + 1e-5)
        want = ((full - full.mean()) / (full.std() + 1e-5))[:, first:first + n]      # process_batch.py:141-142, global
        ret[rank] = (float((mine - want).abs().max()), float(stats[0]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_global_advantage_statistics():
    world, n_total = 2, 1001
    ctx = mp.get_context("spawn")
    with ctx.Manager() as m:
        ret = m.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        for r in range(world):
            err, count = ret[r]
            assert count == 20 * n_total          # statistics are global on every rank
            assert err < 1e-4
