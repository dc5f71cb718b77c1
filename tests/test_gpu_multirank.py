"""-m gpu: multi-rank byte equality over NCCL (tools/multirank_check.py under torchrun).  With one GPU the same script runs
as world size 1 (the sharded code paths with nobody to talk to); with two or more it runs one process per GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(ROOT, "tools", "multirank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    return r.stdout


def test_sharded_paths_world_size_1(cuda):
    out = _run(1)
    assert "world 1" in out


def test_image_and_tile_sharding_are_byte_identical_over_nccl(cuda):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run with gpurun --gpus 2)")
    out = _run(min(n, 4))
    assert "MULTIRANK_OK" in out
