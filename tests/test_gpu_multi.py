"""GPU (-m gpu), needs >= 2 devices: N-rank separation under torchrun (NCCL) is bit-identical to one rank."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("prec,batch", [("fp32", 4), ("bf16", 4), ("bf16", 2)])
def test_two_rank_separation_bit_exact(prec, batch):
    n = min(torch.cuda.device_count(), 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tools", "multigpu_check.py"), prec, str(batch)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit_exact=True" in r.stdout
