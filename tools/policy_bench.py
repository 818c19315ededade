"""Time of the fused tcgen05 acting kernel vs the eager PyTorch policy on [M, 10] observations.
    python tools/policy_bench.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

M = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
torch.manual_seed(0)
p = G.GaussianPolicy(10, 2).cuda()
f = G.FusedGaussianPolicy(p)
obs = torch.rand(M, 10, device="cuda") * 2 - 1
out = torch.empty(M, 2, device="cuda")
head = torch.zeros(M, 4, device="cuda")
noise = torch.randn(M, 2, device="cuda")
f.act(obs, out=out, noise=noise, head=head)
with torch.no_grad():
    mean, log_std = p(obs)
print("max |head - ref|", (head - torch.cat([mean, log_std], 1)).abs().max().item())


def timeit(fn, n=50):
    """us per call, replayed from a CUDA graph (device time, no host launch overhead)."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(4):
            g.replay()
        e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (4 * n)


t_f = timeit(lambda: f.act(obs, out=out))
t_e = timeit(lambda: p.act(obs))
with torch.autocast("cuda", dtype=torch.bfloat16):
    t_b = timeit(lambda: p.act(obs))
flops = 2.0 * M * (16 * 256 + 256 * 256 + 256 * 4)
print(f"M={M}: fused {t_f:.1f} us ({flops / t_f / 1e6:.1f} TFLOP/s), eager fp32 {t_e:.1f} us, eager bf16 autocast {t_b:.1f} us")
