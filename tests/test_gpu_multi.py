"""Multi-GPU path (runs only on a box with >= 2 GPUs): torchrun with 2 ranks, parity vs the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tests", "mgpu_check.py"), "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MGPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
