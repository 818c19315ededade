"""BASELINE configs[4]: end-to-end SAC-style multi-agent rollout with the env on the device, B=16,384 envs x N=10.
One acting step = GaussianPolicy(10->256->256->2) forward over [B*N, 10] (random-init weights: the reference ships
none) + tanh-Gaussian sample -> step kernel with the polar action map fused -> all B*N transitions appended to the
device replay ring.  Prints env-steps/s and the split env / policy / replay, from CUDA-graph replays.
    python tools/rollout_bench.py [B] [N] [steps] [precisions, comma separated]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_uav_collision_avoidance_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False


def timed(fn, n_graph=50):
    """us per call of fn(), replayed from a CUDA graph of n_graph calls."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(n_graph):
            fn()
    g.replay()
    torch.cuda.synchronize()
    reps = max(1, steps // n_graph)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record(st)
            for _ in range(reps):
                g.replay()
            e1.record(st)
        st.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * n_graph))
    return best


only = sys.argv[4].split(",") if len(sys.argv) > 4 else None  # e.g. "fused"
for precision, fused_append in (("fp32", True), ("tf32", True), ("bf16", True), ("fused", True), ("fused", False)):
    if only and precision not in only:
        continue
    env = G.BatchedMultiUAVWorld2D(B, num_agents=N, reset_mode=G.RESET_ON_DONE0, max_episode_steps=1500, seed=0x5EED)
    policy = G.GaussianPolicy(10, 2).to(dev)
    replay = G.DeviceReplay(min(B * N * 16, 4_000_000), 10, 2, device=dev)
    ro = G.BatchedRollout(env, policy, replay, action_mode="polar", precision=precision, fused_append=fused_append)
    ro.reset()
    t_all = timed(ro.step)
    t_act = timed(ro._act)
    def env_only():  # with fused_append: step + append in one launch
        ro.state = env.obs
        ro._env_step()

    t_env = timed(env_only)
    t_push = None if ro.fused_append else timed(lambda: replay.push(ro.state, ro.action, env.reward, env.final_obs, env.done))
    sr, cr, eps = ro.success_collision_rates()
    print(json.dumps({
        "workload": f"SAC-style rollout, B={B} envs x N={N} UAVs, policy 10-256-256-2 {precision} (random init), polar map fused, "
                    f"device replay ({'appended by the step kernel' if ro.fused_append else 'separate append launch'})",
        "us_per_step": t_all, "env_steps_per_s": B / t_all * 1e6, "uav_steps_per_s": B * N / t_all * 1e6,
        "split_us": {"policy_forward_and_sample": t_act, "env_step_kernel": t_env, "replay_push_kernel": t_push},
        "episodes": eps, "SR": sr, "CR": cr}), flush=True)
