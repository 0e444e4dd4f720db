"""GPU tier: the fused coupling + peer-memory exchange (msgpu_comm_p2p_*) with two ranks.

On a one-GPU box both ranks share cuda:0 (CUDA IPC works between processes on one device; NCCL does not,
so the ranks run without an NCCL communicator): every rank's coupling kernel stores its integrals into
both exchange buffers, and the gathered vectors must equal those of one context holding all systems."""
import os
import subprocess
import sys

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_two_ranks_exchange_through_peer_memory():
    env = dict(os.environ, MSGPU_P2P_SAME_DEVICE="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "scripts", "p2p_check.py"), "6", "8"]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "peer-memory exchange equals reference: True" in res.stdout
