"""Small workload touching every kernel and every rare path (auto-reset, counters, final_obs, pool resets, generic-N
kernel, ragged last warp, replay push, host step) — meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

torch.manual_seed(0)
for N in (1, 3, 8, 10, 13, 32):
    B = 131
    env = G.BatchedMultiUAVWorld2D(B, num_agents=N, x_size=12.0, y_size=12.0, reset_mode=G.RESET_ON_DONE0, max_episode_steps=7, seed=N)
    env.enable_final_obs()
    env.reset()
    rb = G.DeviceReplay(B * N * 3, 10, 2)
    for t in range(12):
        a = torch.rand((B, N, 2), device="cuda") * 2 - 1
        prev = env.obs.clone()
        env.step(a, action_mode="polar" if t % 2 else "scaled")
        rb.push(prev, a, env.reward, env.final_obs, env.done)
    env.reset(mask=(torch.arange(B, device="cuda") % 3 == 0))
    env.observe(); env.stats()
    print("multi N", N, "ok", env.stats()["episodes"])
env = G.BatchedUAVWorld2D(1000, reset_mode=G.RESET_ON_ANY_DONE, max_episode_steps=9, seed=3)
env.enable_final_obs(); env.reset()
for t in range(15):
    env.step(torch.rand((1000, 1, 2), device="cuda") * 24 - 12)
h = [torch.empty(s, dtype=d).pin_memory() for s, d in (((1000, 1, 2), torch.float32), ((1000, 1, 4), torch.float32), ((1000, 1), torch.float32), ((1000, 1), torch.uint8))]
h[0].uniform_(-12, 12)
env.step_host(*h)
p = [torch.empty(s, dtype=d) for s, d in (((1000, 1, 2), torch.float32), ((1000, 1, 4), torch.float32), ((1000, 1), torch.float32), ((1000, 1), torch.uint8))]
p[0].uniform_(-12, 12)
env.step_host(*p)
torch.cuda.synchronize()
print("single ok", env.stats()["episodes"])
