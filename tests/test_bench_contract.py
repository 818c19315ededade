"""CPU: the reference arm of bench.py runs here (it times the oracle port, no GPU) and prints ONE JSON line with the
keys the driver reads; the roofline arithmetic of the main arm is the one SURVEY.md §8d / DESIGN.md §5 state."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "UAV env-steps/sec" and d["unit"] == "UAV env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_follow_the_survey():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.algorithmic_bytes_per_unit("multi", 8) == 110.0       # 41 read + 66 written + 24/N counters
    assert bench.algorithmic_bytes_per_unit("multi", 32) == 107.75
    assert bench.algorithmic_bytes_per_unit("single", 1) == 89.0
    assert set(bench.WORKLOADS) == {"c1", "c2", "c3", "c4", "c4s", "c5", "c5r"}
    assert bench.WORKLOADS["c3"]["B"] == 65536 and bench.WORKLOADS["c3"]["N"] == 8
    assert bench.WORKLOADS["c4s"]["B"] == 1048576 and bench.WORKLOADS["c4s"]["N"] == 32 and bench.WORKLOADS["c4s"]["shard_total"]


def test_reference_arm_runs_the_full_batch_and_the_literal_reference():
    """Same config as the GPU arm (BASELINE configs[2]: all 65,536 envs per step) and, where the literal reference is
    reachable (/root/reference here, oracle/_ref on the GPU box), its single-core number beside the port's."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["config"]["envs_per_step"] == 65536 and d["config"]["envs_per_gpu"] == 65536 and d["config"]["uavs_per_env"] == 8
    lit = d["cpu_baseline_literal"]
    assert "unavailable" in lit or (lit["kind"] == "reference" and lit["cores"] == 1 and 1e3 < lit["value"] < d["value"])


def test_multi_gpu_default_workload_is_configs3():
    """`--gpus N>1` with no --workload measures BASELINE configs[3]: N=32, 1,048,576 envs in total, sharded."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["config"]["uavs_per_env"] == 32 and d["config"]["envs_total"] == 1048576 and d["config"]["envs_per_gpu"] == 131072
    assert d["scaling"] == "strong"
