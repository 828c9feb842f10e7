"""CPU coverage of the N > 1 path (world_size 2 and 3, gloo): PartitionedArray + the deferred chain
queue + write-after-read renaming + the halo exchange kept in program order, on a host-memory
stand-in for the CUDA library (tests/sim_backend.py, tests/sim_dist_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode", ["always", "0"])
@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_stencil_over_gloo(world, mode):
    port = 29800 + world + (os.getpid() % 150) + (7 if mode == "0" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sim_dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1", SIM_FUSION=mode)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.count(" ok") == world
