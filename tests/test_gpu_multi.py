"""Real-NCCL parity of the row-sharded federation, as a -m gpu test: spawns torch.distributed.run over every
visible GPU (one rank per GPU) on scripts/multi_gpu_check.py -- sharded == unsharded bit for bit for batches
1 ... 2100 on every scan kernel family, plus a tie-heavy integer federation (cross-rank / cross-source ties).
Skipped on a one-GPU box; the committed log of the last multi-GPU run is profiles/r02_multi_gpu_check_*.log."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_equals_unsharded_over_nccl(cuda_required):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible GPU: the NCCL parity check needs at least two")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(REPO, "scripts", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=360, cwd=REPO)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "multi-GPU parity ok" in res.stdout
