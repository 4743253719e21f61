"""
Multi-GPU correctness on real devices: self-launches ``torch.distributed.run`` over NCCL when the
box has at least two GPUs (skipped otherwise).  scripts/dist_check.py asserts, for FP32 and FP64
mode and an uneven split, that a ShardedMuscleBatch equals the unsharded MuscleBatch BITWISE for a
single step, a fused window, the asynchronous NCCL gather and the gather fused into the kernel
(peer stores through symmetric memory over NVLink).
"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_equals_unsharded_on_real_gpus(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count() if torch.cuda.is_available() else 0}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world),
           os.path.join(ROOT, "scripts", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("sharded == unsharded (bitwise)") == 2, r.stdout[-3000:]
