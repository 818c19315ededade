#!/usr/bin/env python
"""bench.py — UAV env-steps/s of the batched environment step on B200, next to the reference CPU step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--impl ours|reference]

A "step" is one pass of the hot path (one `uavca_step_*` launch) over one batch of B envs.  The default workload
is BASELINE.json configs[2] — multi-UAV, N=8 UAVs per env, B=65,536 envs per GPU, random cartesian actions,
auto-reset on dones[0] / 1,500 steps from the on-device Philox stream.  One batch is ~50 MB (L2-resident on a
B200), so the bench cycles through a ring of independent batches whose combined footprint exceeds L2 several
times over: every timed launch reads its state and actions from HBM.  Steps are replayed from a CUDA graph and
timed with CUDA events; multi-GPU runs are one process per GPU (torchrun), envs sharded with no per-step
collective, time = max over ranks.

The JSON line carries `value` (device-resident throughput), `e2e` (same metric through `uavca_step_host` with
pinned HOST buffers: H2D actions + step + D2H obs/reward/done inside the timed region), `roofline` (algorithmic
bytes per launch / measured launch time against MEASURED_PEAKS.json) and `cpu_baseline` (the oracle port of the
reference step timed on this host).  `--impl reference` times the reference's CPU algorithm (oracle port, the
Python reference cannot travel to the GPU box) with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: kind, N, B per GPU, algorithmic bytes per UAV-step (SURVEY.md §8d / DESIGN.md §5), default steps
    "c3": dict(kind="multi", N=8, B=65536, desc="multi-UAV N=8, B=65,536 envs/GPU (BASELINE configs[2])", steps=20000),
    "c2": dict(kind="single", N=1, B=65536, desc="single-UAV, B=65,536 envs/GPU (BASELINE configs[1])", steps=20000, streams=4),
    "c4": dict(kind="multi", N=32, B=1048576, desc="multi-UAV N=32, B=1,048,576 envs/GPU (BASELINE configs[3] shape)", steps=300),
    "c4s": dict(kind="multi", N=32, B=1048576, desc="multi-UAV N=32, B=1,048,576 envs IN TOTAL, sharded over the GPUs (BASELINE configs[3])",
                steps=300, shard_total=True),
    "c5": dict(kind="multi", N=10, B=16384, desc="multi-UAV N=10, B=16,384 envs/GPU (BASELINE configs[4] env part)", steps=20000, streams=4),
}
L2_BYTES = 126e6
GRAPH_STEPS = 200


def algorithmic_bytes_per_unit(kind: str, N: int) -> float:
    """multi 107 B per UAV-step + 24/N for the per-env counters; single 89 B per env-step."""
    return 89.0 if kind == "single" else 107.0 + 24.0 / N


def footprint_bytes_per_unit(kind: str) -> float:
    # resident bytes per UAV of one batch: state (pos 8, vel 16, tgt 8, init 4, prev 4, flags 1) + io (action 8, obs, reward 4, done 1)
    return 41.0 + 8 + (16 if kind == "single" else 40) + 4 + 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---- clocks sampled DURING the timed region ---------------------------------------------------------------------


class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index: int, period_s: float = 0.01):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for b, name in self.REASONS.items():
                    if bits & b and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        s = sorted(self.sm)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---- CPU side: the oracle port of the reference step ------------------------------------------------------------


def time_oracle(kind: str, N: int, envs: int, steps: int, nthreads: int, seed: int = 0):
    """UAV-steps/s of oracle/uav_oracle.c (the restated reference step) on `nthreads` host threads."""
    import numpy as np

    from oracle import oracle as O

    if kind == "single":
        cfg = O.single_config(envs, reset_mode=O.RESET_ON_ANY_DONE, max_episode_steps=1500, seed=seed)
        amax = 12.0
    else:
        cfg = O.multi_config(envs, N, reset_mode=O.RESET_ON_DONE0, max_episode_steps=1500, seed=seed)
        amax = 10.0
    orc = O.Oracle(cfg, nthreads=nthreads)
    orc.reset()
    rng = np.random.default_rng(seed)
    acts = [rng.uniform(-amax, amax, size=(envs, N, 2)).astype(np.float32) for _ in range(8)]
    orc.step(acts[0])
    t0 = time.perf_counter()
    for k in range(steps):
        orc.step(acts[k % 8])
    dt = time.perf_counter() - t0
    return envs * N * steps / dt, dt


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU algorithm for the path (oracle port; the Python reference itself
    cannot travel to the GPU box), all host threads, bounded sample of the same workload per step."""
    if rank != 0:
        return
    from oracle import oracle as O

    nthreads = O.max_threads()
    envs = 8192 if wl["N"] <= 8 else 1024
    envs = min(envs, wl["B"])
    steps, warm = max(1, args.steps), max(0, args.warmup)
    steps = min(steps, 400)  # bounded: a step here is one pass over `envs` envs
    time_oracle(wl["kind"], wl["N"], envs, max(1, min(warm, 5)), nthreads)
    ups, dt = time_oracle(wl["kind"], wl["N"], envs, steps, nthreads)
    unit = "UAV env-steps/s"
    sample = f"{envs} envs x {wl['N']} UAVs x {steps} steps of the {args.workload} workload (oracle/uav_oracle.c, pthreads)"
    line = {
        "impl": "reference", "metric": "UAV env-steps/sec", "value": ups, "unit": unit, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32 mixed (as the reference)", "data": "synthetic",
        "config": {"workload": wl["desc"], "envs_per_step": envs, "uavs_per_env": wl["N"], "actions": "uniform random cartesian",
                   "auto_reset": "dones[0] or 1500 steps", "host_cpu": cpu_model()},
        "cpu_baseline": {"value": ups, "unit": unit, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": ups, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- GPU side ---------------------------------------------------------------------------------------------------


def run_ours(args, wl, rank, world, local_rank):
    import torch

    import gym_uav_collision_avoidance_b200 as G

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
    kind, N, B = wl["kind"], wl["N"], wl["B"]
    if wl.get("shard_total"):  # strong scaling: the named total is cut into contiguous shards
        from gym_uav_collision_avoidance_b200 import sharding as _sh

        B = _sh.shard_range(wl["B"], rank, world)[1]
    K, W = args.steps, max(args.warmup, 3)
    units_per_step = B * N
    # ring of independent batches: combined footprint > 3x L2 so that every launch streams from HBM
    per_batch = units_per_step * footprint_bytes_per_unit(kind)
    ring = max(1, int(-(-3.2 * L2_BYTES // per_batch)))
    S = max(1, min(args.streams, ring))
    if ring > 1 and ring % S:
        ring += S - ring % S  # every stream serves the same number of batches (no idle stream at the end of a ring cycle)
    amax = 12.0 if kind == "single" else 10.0
    envs = []
    for r in range(ring):
        kw = dict(device=dev, seed=0x5EED, max_episode_steps=1500, env_index_base=(rank * ring + r) * B)
        if kind == "single":
            e = G.BatchedUAVWorld2D(B, reset_mode=G.RESET_ON_ANY_DONE, **kw)
        else:
            e = G.BatchedMultiUAVWorld2D(B, num_agents=N, reset_mode=G.RESET_ON_DONE0, **kw)
        e.reset()
        envs.append(e)
    n_act = max(ring, min(32, int(2e9 // (units_per_step * 8))))
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = [(torch.rand((B, N, 2), generator=gen, device=dev) * 2 - 1) * amax for _ in range(n_act)]

    def one_step(k):
        envs[k % ring].step(actions[(k + k // ring) % n_act])

    # The batches of the ring are independent environments: like a double-buffered env pool they are pipelined over
    # `S` streams, so that one batch's ramp-up overlaps the previous batch's tail.  A batch always runs on the same
    # stream (its own steps stay ordered).  S=1 (--streams 1) serialises every launch behind the previous one.

    # warm-up (eager), then capture graphs of GRAPH_STEPS steps and of the remainder
    for k in range(W):
        one_step(k)
    torch.cuda.synchronize(dev)
    launches_before = sum(e.launch_count for e in envs)
    q, rem = divmod(K, GRAPH_STEPS)
    stream = torch.cuda.Stream(device=dev)

    def capture(n, k0, n_streams):
        side = [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            fork = torch.cuda.Event()
            fork.record(stream)
            for s_ in side:
                s_.wait_event(fork)
            for k in range(n):
                lane = ((k0 + k) % ring) % n_streams
                with torch.cuda.stream(stream if lane == 0 else side[lane - 1]):
                    one_step(k0 + k)
            for s_ in side:
                join = torch.cuda.Event()
                join.record(s_)
                stream.wait_event(join)
        return g

    # the same steps once more on ONE stream (every launch waits for the previous one): reported beside the headline
    serial_us = None
    if S > 1 and q > 0:
        l0 = sum(e.launch_count for e in envs)
        g_ser = capture(GRAPH_STEPS, W, 1)
        g_ser.replay()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, min(q, 20))
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(reps):
                g_ser.replay()
            e1.record(stream)
        stream.synchronize()
        serial_us = e0.elapsed_time(e1) * 1e3 / (reps * GRAPH_STEPS)
        del g_ser
        launches_before += sum(e.launch_count for e in envs) - l0

    g_main = capture(GRAPH_STEPS, W, S) if q > 0 else None
    launches_per_graph = sum(e.launch_count for e in envs) - launches_before
    g_rem = capture(rem, W + GRAPH_STEPS, S) if rem > 0 else None
    launches_rem = sum(e.launch_count for e in envs) - launches_before - launches_per_graph
    if g_main is not None:
        g_main.replay()  # graph warm-up (upload)
    if g_rem is not None:
        g_rem.replay()
    torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(q):
                g_main.replay()
            if g_rem is not None:
                g_rem.replay()
            ev1.record(stream)
        stream.synchronize()
        if K * 1 < 2000:  # short runs: keep sampling a little so that at least one clock sample lands
            time.sleep(0.03)
    barrier()
    from gym_uav_collision_avoidance_b200 import sharding

    ms = sharding.max_over_ranks(ev0.elapsed_time(ev1), device=dev)  # device time, slowest rank
    gpu_launches = q * launches_per_graph + launches_rem
    value = world * units_per_step * K / (ms * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (uavca_step_host): pinned host buffers, H2D + step + D2H
    D = 4 if kind == "single" else 10
    h_act = [torch.empty((B, N, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    for h, a in zip(h_act, actions):
        h.copy_(a.cpu())
    h_obs = torch.empty((B, N, D), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((B, N), dtype=torch.float32).pin_memory()
    h_done = torch.empty((B, N), dtype=torch.uint8).pin_memory()
    e2e_steps = max(1, min(K, 200 if units_per_step < 4e6 else 10))
    e2e_ring = min(ring, 4)  # PCIe-bound: L2 residency is irrelevant here; keep the lazily created staging small
    for k in range(2 * e2e_ring):
        envs[k % e2e_ring].step_host(h_act[k % 2], h_obs, h_rew, h_done)
    barrier()
    launches_e2e0 = sum(e.launch_count for e in envs)
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        envs[k % e2e_ring].step_host(h_act[k % 2], h_obs, h_rew, h_done)
    torch.cuda.synchronize(dev)
    e2e_s = sharding.max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_value = world * units_per_step * e2e_steps / e2e_s
    gpu_launches += sum(e.launch_count for e in envs) - launches_e2e0

    # episode statistics: the only collective of the path (NCCL all-reduce of 4 counters, outside the timed region)
    local = dict.fromkeys(sharding.STAT_KEYS, 0)
    for e in envs:
        s = e.stats()
        for k in sharding.STAT_KEYS:
            local[k] += s[k]
    stats = sharding.reduce_stats(local, device=dev)

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg = algorithmic_bytes_per_unit(kind, N)
        launch_s = ms * 1e-3 / K
        achieved = units_per_step * alg / launch_s / 1e9
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # a reported baseline, timed on rank 0 at N=1 only
            c_envs = min(B, 4096 if N <= 8 else 512)
            c_steps = max(20, int(12e6 // (c_envs * N)))
            ups, dt = time_oracle(kind, N, c_envs, c_steps, 1)
            cpu = {"value": ups, "unit": "UAV env-steps/s", "cores": 1, "kind": "port",
                   "sample": f"{c_envs} envs x {N} UAVs x {c_steps} steps of the same workload, oracle/uav_oracle.c on 1 thread "
                             f"({dt:.1f} s; host: {cpu_model()}, {os.cpu_count()} logical cores)"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload)
        line = {
            "metric": "UAV env-steps/sec", "value": value, "unit": "UAV env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if wl.get("shard_total") else "weak", "vs_baseline": None,
            "dtype": "f64 velocity / f32 position+obs (as the reference)", "data": "synthetic",
            "config": {"workload": wl["desc"], "envs_per_gpu": B, "uavs_per_env": N, "env_steps_per_s": value / N,
                       "actions": "uniform random cartesian, resident in HBM", "auto_reset": "on-device Philox, dones[0] or 1500 steps",
                       "l2": f"ring of {ring} independent batches ({ring * per_batch / 1e6:.0f} MB > L2) so every launch streams from HBM",
                       "timing": f"CUDA events around CUDA-graph replays ({GRAPH_STEPS} steps per graph)", "parallelism": f"env-sharded x{world}, no per-step collective",
                       "streams": S, "pipelining": (f"independent batches of the ring pipelined over {S} streams (a batch's own steps stay ordered)" if S > 1 else "none: every launch waits for the previous one"),
                       "single_stream_us_per_step": serial_us},
            "e2e": {"value": e2e_value, "unit": "UAV env-steps/s", "h2d_bytes_per_step": units_per_step * 8,
                    "d2h_bytes_per_step": units_per_step * (D * 4 + 4 + 1), "steps": e2e_steps,
                    "path": "uavca_step_host, pinned host buffers: outputs >= 256 MB leave by DMA (chunked H2D/step/D2H pipeline over two "
                            "streams), smaller batches are zero-copy (the kernel reads/writes mapped host memory through PCIe)"
                            + (f" [forced: {os.environ['UAVCA_HOST_PATH']}]" if os.environ.get("UAVCA_HOST_PATH") else "")},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_unit": alg,
                         "units_per_launch": units_per_step, "launch_us": launch_s * 1e6,
                         "launch_us_is": "timed region / launches" + (f" ({S} independent launches in flight)" if S > 1 else ""),
                         "frac_single_stream": (units_per_step * alg / (serial_us * 1e-6) / 1e9 / peak) if serial_us else None},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
            "episode_stats": stats,
        }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def protect_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries print there too (NCCL announces its version on stdout at
    init): point fd 1 at stderr for the whole run and keep the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=None,
                    help="streams the independent batches of the ring are pipelined over (default: 2; 4 for the small c2/c5 batches)")
    args = ap.parse_args()
    protect_stdout()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.steps is None:
        args.steps = wl["steps"] if args.impl == "ours" else 100
    if args.streams is None:
        args.streams = wl.get("streams", 2)
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess

        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29513", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))  # the ranks inherit the REAL stdout
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
