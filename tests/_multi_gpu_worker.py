"""One rank of the multi-GPU parity run (launched by tests/test_multi_gpu.py through torchrun, one process per GPU).

BASELINE configs[3] in small: N=32 UAVs per env, the env axis cut into contiguous shards, no per-step collective.
Every rank steps its shard on ITS GPU and checks a window of it against the oracle keyed by the GLOBAL env index
(done flags / reset masks / positions bit-exact, rewards / observations within 1e-5); at the end the episode counters
are summed over NCCL and compared with the oracle windows' own sum (gathered over the same process group)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gym_uav_collision_avoidance_b200 as G  # noqa: E402
from gym_uav_collision_avoidance_b200 import sharding  # noqa: E402
from oracle import oracle as O  # noqa: E402

from _golden import obs_close  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    total, N, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    base, B = sharding.shard_range(total, rank, world)
    kw = dict(reset_mode=O.RESET_ON_DONE0, max_episode_steps=37, seed=0xC4)
    env = G.BatchedMultiUAVWorld2D(B, num_agents=N, env_index_base=base, device=dev, **kw)
    W = min(B, 256)
    start = (B - W) // 2
    orc = O.Oracle(O.multi_config(W, N, env_index_base=base + start, **kw), nthreads=4)
    env.reset()
    orc.reset()
    sl = slice(start, start + W)
    assert np.array_equal(env.state.pos[sl].cpu().numpy(), orc.state.pos)
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    for t in range(steps):
        a = torch.rand((B, N, 2), generator=gen, device=dev) * 20 - 10
        obs, rew, done, info = env.step(a)
        out = orc.step(a[sl].cpu().numpy())
        assert np.array_equal(done[sl].cpu().numpy(), out["done"]), f"rank {rank}: done flags differ at step {t}"
        assert np.array_equal(info["reset_mask"][sl].cpu().numpy(), out["reset_mask"]), f"rank {rank}: reset mask, step {t}"
        if t % 5 == 0 or t == steps - 1:
            assert np.array_equal(env.state.pos[sl].cpu().numpy(), orc.state.pos), f"rank {rank}: positions, step {t}"
            assert np.array_equal(env.state.vel[sl].cpu().numpy(), orc.state.vel)
            r = rew[sl].cpu().numpy()
            assert (np.abs(r - out["reward"]) <= 1e-5 * np.abs(out["reward"]) + 1e-6).all()
            assert obs_close(obs[sl].cpu().numpy(), out["obs"], 1e-5, 1e-6).all()
    # the one collective of the path: episode counters summed over the ranks (NCCL)
    s = env.stats()
    tot = sharding.reduce_stats(s, device=dev)
    mine = torch.tensor([s[k] for k in sharding.STAT_KEYS], dtype=torch.int64, device=dev)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    assert [int(v) for v in torch.stack(allv).sum(0).tolist()] == [tot[k] for k in sharding.STAT_KEYS]
    assert tot["episodes"] >= total * (steps // 37)
    dist.barrier()
    if rank == 0:
        print(f"MULTI_GPU_OK world={world} total_envs={total} N={N} steps={steps} episodes={tot['episodes']}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
