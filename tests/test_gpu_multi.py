"""Multi-GPU equivalence: N ranks x M envs == 1 rank x N*M envs over NCCL.  Needs
two GPUs; on a single-GPU box the test is skipped (the gloo 2-rank CPU test covers the exchange logic)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_ranks_equal_one_rank():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_equivalence.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTI_GPU_EQUIVALENCE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
