"""
Multi-GPU parity ON HARDWARE: torchrun launches one rank per GPU (NCCL) and every sharded path
-- fused peer-memory swap+apply (t <= 8 and t > 8 operators on a sharded axis), NCCL swaps,
sharded density matrices (apply / Kraus / probs / partial trace / measurement) -- is compared
with the NumPy oracle (tests/multi_gpu_worker.py).  Needs >= 2 GPUs: `gpurun --gpus 2`.
"""

import os
import socket
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")

pytestmark = [pytest.mark.gpu, pytest.mark.gpu_multi]
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_paths_match_oracle_on_gpus(nproc):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "MULTI_OK" in proc.stdout, proc.stdout[-2000:]
