"""One process per GPU: y-slabs wired through CUDA IPC, checked against one GPU and the oracle.
Needs >= 2 GPUs; skipped on a single-GPU box (the same kernels run there as several slabs in one
process, tests/test_gpu_parity.py::test_slab_count_independence)."""
import os
import subprocess
import sys

import pytest

import simuverse_b200 as sb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("nx,ny,steps", [(520, 384, 120), (131, 77, 60)])
def test_one_process_per_gpu_matches_single_gpu_and_oracle(nx, ny, steps):
    n = sb.lib.lbm_device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    if ny // world < 2:
        world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multirank_check.py"),
           str(nx), str(ny), str(steps)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTIRANK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
