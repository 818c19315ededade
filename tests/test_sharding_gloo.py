"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path — contiguous env shards that cover every
env exactly once (Philox streams are keyed by the GLOBAL env index, so results cannot depend on the split) and the
one collective of the path, the sum of the episode counters."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_uav_collision_avoidance_b200 import sharding


def test_shard_range_partitions_the_env_axis():
    for total in (0, 1, 7, 65536, 1048576, 1000003):
        for world in (1, 2, 3, 4, 8):
            covered = 0
            for rank in range(world):
                base, n = sharding.shard_range(total, rank, world)
                assert base == covered and n >= 0
                covered += n
            assert covered == total
            sizes = [sharding.shard_range(total, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_reduce_is_identity_without_a_process_group():
    s = dict(episodes=3, reach=1, collisions=2, steps=99, live_steps=5)
    assert sharding.reduce_stats(s) == dict(episodes=3, reach=1, collisions=2, steps=99)
    assert sharding.max_over_ranks(1.5) == 1.5


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, n = sharding.shard_range(1001, rank, world)
        stats = dict(episodes=10 + rank, reach=rank, collisions=2 * rank + 1, steps=1000 * (rank + 1))
        total = sharding.reduce_stats(stats)
        slowest = sharding.max_over_ranks(1.0 + rank)
        # every rank contributes its shard size: the sum must be the whole env axis
        t = torch.tensor([n], dtype=torch.int64)
        dist.all_reduce(t)
        q.put((rank, base, n, total, slowest, int(t.item())))
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, n0, tot0, slow0, all0), (r1, b1, n1, tot1, slow1, all1) = res
    assert (b0, n0, b1, n1) == (0, 501, 501, 500)
    assert tot0 == tot1 == dict(episodes=21, reach=1, collisions=4, steps=3000)
    assert slow0 == slow1 == 2.0 and all0 == all1 == 1001
