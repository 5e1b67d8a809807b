"""N > 1 on real GPUs (runs only where >= 2 CUDA devices are visible)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_sharded_equals_unsharded(transport):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ)
    env.pop("SONAR_B200_NO_PEER", None)
    if transport == "nccl":
        env["SONAR_B200_NO_PEER"] = "1"
    port = 29600 + (0 if transport == "peer" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 8)}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(REPO / "tests" / "multi_gpu_check.py")]  # fmt: skip
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, check=False)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert proc.stdout.count("OK ") == 4
