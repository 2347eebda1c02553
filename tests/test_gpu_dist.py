"""Multi-GPU: the sharded path over NCCL (needs >= 2 GPUs; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharded_trace_and_hpd_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_dist_nccl.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(p.stdout[-3000:])
    assert p.returncode == 0, p.stderr[-3000:]
