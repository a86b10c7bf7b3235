"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): runs tools/mgpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_and_sharded_codebook_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tools", "mgpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "MGPU OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
