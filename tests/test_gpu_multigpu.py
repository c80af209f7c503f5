"""GPU, NCCL: runs tests/gpu_multirank.py under torchrun on every visible GPU (skipped with < 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_nccl_global_loss_and_sharded_retrieval():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "gpu_multirank.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
