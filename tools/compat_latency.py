"""Latency of one `env.step` of the B=1 drop-in classes (float32 and float64 actions) next to the literal reference.
    python tools/compat_latency.py [N] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gym_uav_collision_avoidance_b200 import compat

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rng = np.random.default_rng(0)
for dtype in (np.float32, np.float64):
    env = compat.MultiUAVWorld2D(num_agents=N)
    env.reset()
    acts = [[rng.uniform(-10, 10, 2).astype(dtype) for _ in range(N)] for _ in range(K)]
    for a in acts[:50]:
        env.step(a)
    t0 = time.perf_counter()
    for a in acts:
        _, _, dones, _ = env.step(a)
        if dones[0]:
            env.reset()
    dt = time.perf_counter() - t0
    print(f"compat.MultiUAVWorld2D N={N} {np.dtype(dtype).name} actions: {dt / K * 1e6:.1f} us per env.step")
    env.close()
try:
    from oracle import ref_loader as R
    _, Ref = R.load_reference()
    env = Ref(num_agents=N)
    env.reset()
    acts = [[rng.uniform(-10, 10, 2) for _ in range(N)] for _ in range(300)]
    t0 = time.perf_counter()
    for a in acts:
        _, _, dones, _ = env.step(a)
        if dones[0]:
            env.reset()
    print(f"literal reference N={N}: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us per env.step")
except Exception as e:  # noqa: BLE001
    print("literal reference unavailable:", e)
