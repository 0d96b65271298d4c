"""GPU (-m gpu), needs >= 2 GPUs: real NCCL run of the sharded path, bit-exact vs one GPU."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_sharded_bank_nccl_bit_exact():
    world = 2 if torch.cuda.device_count() < 4 else (4 if torch.cuda.device_count() < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(ROOT, "tools", "dist_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("bit-exact on all ranks: True") == 3
