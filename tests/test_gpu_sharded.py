"""NCCL-sharded simulation on >= 2 GPUs equals the single-GPU result (skipped on one GPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharding_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2); host logic is covered by test_sharding_gloo.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "check_sharded_cuda.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded over 2 GPUs == single GPU: True" in r.stdout
    assert "SBC sharded over 2 GPUs == single GPU: True" in r.stdout
    assert "potential sharded over 2 GPUs == single GPU: True" in r.stdout
    assert r.stdout.count("fused peer-store gather over 2 GPUs == single GPU: True") == 2
