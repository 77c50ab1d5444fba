"""Mode B over NVLink peer memory on real ranks (needs >= 2 GPUs; skipped on a single-GPU box):
torchrun runs scripts/modeb_p2p_check.py, which asserts that the peer-memory step, its graphed
pipeline and the NCCL all-to-all step give bitwise-equal parameters on the same batches."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_memory_mode_b_equals_nccl_mode_b_on_two_ranks():
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "modeb_p2p_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "pipelined graphed loop == serial steps (bitwise)" in res.stdout
    assert "peer-mode parity OK" in res.stdout
