"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): runs tools/multi_gpu_check.py under torchrun —
row-sharded ALS and CCD++ over peer memory against one unsharded engine, DSGD with item blocks pushed between ranks
against the same schedule on one engine.  The routing logic itself is covered on CPU by test_dsgd_routing.py."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
def test_two_rank_parity_over_peer_memory():
    if _n_gpus() < 2:
        pytest.skip("needs two GPUs on one node")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)
    out = p.stdout.decode(errors="replace")
    assert p.returncode == 0 and "MULTI_GPU_CHECK PASS" in out, out[-3000:]
