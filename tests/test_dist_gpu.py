"""Multi-GPU parity on a box with >= 2 GPUs: launches tests/dist_worker_gpu.py under
torchrun, one rank per GPU, NCCL.  Skipped on a single-GPU box (the CPU-side partition logic is
covered by tests/test_partition.py)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_row_partitioned_parity(b200, nproc):
    if b200.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29530 + nproc), str(ROOT / "tests" / "dist_worker_gpu.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist gpu ok" in r.stdout
