"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): see tests/mgpu_worker.py."""
import os
import socket
import subprocess
import sys

import pytest

from tests.util import zipf_corpus

pytestmark = pytest.mark.gpu


def test_replica_average_over_nccl(tmp_path):
    w2b = pytest.importorskip("word2bits_b200")
    n = w2b.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    path = zipf_corpus(str(tmp_path / "c.txt"), 400000, 5000, seed=3)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(root, "tests", "mgpu_worker.py"), path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
