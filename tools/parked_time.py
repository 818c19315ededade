"""Step time when most UAVs are parked (the evaluation regime of a good policy: parked UAVs re-run finish() on every
step, multi_uav_world_2d.py:218-222) next to ordinary flight.   python tools/parked_time.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 262144


def run(parked_fraction):
    env = G.BatchedMultiUAVWorld2D(B, num_agents=N, seed=3)
    env.reset()
    if parked_fraction > 0:
        m = torch.rand(B, N, device="cuda") < parked_fraction
        env.state.pos[m] = env.state.tgt[m] + torch.tensor([0.1, 0.05], device="cuda")  # 11 cm from the target ...
        env.state.vel[m] = torch.tensor([0.0007, 0.0007], dtype=torch.float64, device="cuda")  # ... at parking speed
        env.state.flags[m] = 1
    a = torch.zeros(B, N, 2, device="cuda")
    for _ in range(3):
        env.step(a, evaluate=True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.graph(g, stream=st):
        for _ in range(50):
            env.step(a, evaluate=True)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(4):
            g.replay()
        e1.record(st)
    st.synchronize()
    parked = int((env.state.flags & 1).sum())
    return e0.elapsed_time(e1) * 1e3 / 200, parked / (B * N)


for frac in (0.0, 0.05, 0.5, 1.0):
    us, p = run(frac)
    print(f"N={N} B={B}: parked {p:5.1%}  {us:8.2f} us/step  {B * N / us / 1e3:6.2f} G UAV-steps/s")
