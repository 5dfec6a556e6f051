"""-m gpu, needs >= 2 GPUs on the box (skipped otherwise): CUDA shards -> NCCL placement all-gather ->
fragments sent into ONE assembled body on rank 0 == the single-GPU body == the oracle body, including the
cross-rank 'rest of file raw' fold (adaptive_compressor.py:586-590) with the raw chunk in rank 0, a middle
rank, the last rank, and with more ranks than chunks."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4])
def test_assembled_body_equals_single_gpu(world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_multirank_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIRANK OK" in r.stdout
