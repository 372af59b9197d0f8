"""Launches tests/multi_gpu_check.py under torchrun when the box has at least two GPUs (skipped otherwise)."""
import os
import subprocess
import sys

import pytest

from tests.util import ROOT


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_multi_gpu_parity():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-2000:]
