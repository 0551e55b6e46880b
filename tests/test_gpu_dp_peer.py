"""Peer-memory gradient exchange (csrc/dp_peer.cuh) against NCCL on >= 2 GPUs of one box (skipped on a 1-GPU box):
bitwise equality with the NCCL path at world 2, bit-identical replicas, no wait time-outs."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_exchange_matches_nccl_world2():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", "tests/_dp_peer_check.py"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if "peer==nccl" in l]
    if not lines and "PeerGradExchange unavailable" in out.stdout:
        pytest.skip("symmetric memory unavailable on this box")
    assert len(lines) == 2
    for l in lines:
        assert "peer==nccl bitwise: True" in l and "replicas identical: True" in l and "timeout flag: False" in l, l
