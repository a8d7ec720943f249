"""Multi-GPU paths on real devices (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

import celda_b200 as cb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_sharded_em_and_grid_search_two_ranks():
    if cb.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(ROOT, "tests", "multi_gpu_worker.py")], capture_output=True, text=True, timeout=580)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("MULTI_GPU_OK") == 2
