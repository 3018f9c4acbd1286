"""GPU, >= 2 devices: the sharded fit (fused NVLink exchange and the NCCL exchange) gives the same
centroids (bitwise), n_iter and labels as one GPU -- runs tools/multi_gpu_parity.py under torchrun.
Skipped on single-GPU boxes."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_invariance():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "multi_gpu_parity.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    tail = "\n".join((res.stdout + res.stderr).splitlines()[-30:])
    assert res.returncode == 0, tail
    assert "multi-GPU parity: all checks passed" in res.stdout, tail
