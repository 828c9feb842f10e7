"""Multi-GPU parity (needs >= 2 B200s: run with `gpurun --gpus 2`): the SPMD worker compares the
row-partitioned execution — halo exchange for the stencil, NCCL combines for reductions — with
NumPy on every rank."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    from cunumeric_b200 import _lib

    return _lib.load().cnb_device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_execution(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29700 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-6000:]
    assert res.stdout.count(" ok") == world
