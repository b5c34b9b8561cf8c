"""Multi-GPU parity on real NCCL (needs >= 2 visible GPUs; the world-2 choreography is also covered on CPU
under gloo by tests/test_host_logic.py).  Runs tools/multigpu_check.py under torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_row_sharded_loss_and_key_sharded_knn_match_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTIGPU_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
