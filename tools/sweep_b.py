"""Kernel time vs batch size (multi-UAV step): separates the per-launch fixed cost from the streaming rate."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
res = []
for logb in range(8, 22):
    B = 1 << logb
    if B * N > 40e6: break
    per = B * N * 94
    ring = max(1, min(64, int(-(-3.2 * 126e6 // per))))
    envs = [G.BatchedMultiUAVWorld2D(B, num_agents=N, reset_mode=1, max_episode_steps=1500, seed=1, env_index_base=r * B) for r in range(ring)]
    gen = torch.Generator(device=dev).manual_seed(0)
    acts = [(torch.rand((B, N, 2), generator=gen, device=dev) * 20 - 10) for _ in range(ring)]
    for e in envs: e.reset()
    steps = max(ring, min(400, int(2e9 // (B * N)) // ring * ring)) or ring
    for k in range(ring): envs[k].step(acts[k])
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for k in range(steps): envs[k % ring].step(acts[k % ring])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(5): g.replay()
        e1.record(st)
    st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * steps)
    res.append((B, us))
    print(f"N={N} B={B:8d} ring={ring:3d} steps/graph={steps:4d}  {us:9.2f} us/step  {B*N/us/1e3:8.2f} G UAV-steps/s  {B*N*(107+24/N)/us/1e3/6515.7:6.3f} of roofline", flush=True)
    del envs, acts, g
    torch.cuda.empty_cache()
