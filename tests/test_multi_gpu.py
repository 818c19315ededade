"""GPU, >= 2 devices: the sharded path on REAL GPUs — one process per GPU under torchrun, NCCL only for the episode
counters.  Skipped on single-GPU boxes (tests/test_sharding_gloo.py covers the host logic on CPU with world_size 2,
tests/test_cuda_parity.py::test_shard_invariance the shard invariance on one device)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs at least two CUDA devices")
def test_sharded_rollout_on_every_gpu_matches_the_oracle():
    world = min(torch.cuda.device_count(), 8)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "_multi_gpu_worker.py"), str(8192 * world + 3), "32", "120"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1200, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"MULTI_GPU_OK world={world}" in out.stdout
