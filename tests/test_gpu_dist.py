"""GPU tier: one batch sharded across two ranks under torchrun, verified with bp.verify_batch on the GPU(s), verdicts gathered
(bulletproofs_b200/dist.py).  With two or more GPUs every rank owns one and the table broadcast / gather run over NCCL; on a
one-GPU box both ranks share cuda:0 and the plumbing runs over gloo."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_one_batch_sharded_over_two_ranks(built):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("verdicts ok") == 2, r.stdout
