"""N > 1 host path on CPU: world_size-2 and -3 gloo runs of tests/mp_cpu_worker.py (partitioned numbering + pattern per
rank, halo plan, halo exchange + distributed SpMV against the oracle's global matrices)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("nproc,port", [(2, 29641), (3, 29643)])
def test_gloo_ranks_partitioned_pattern_and_halo(nproc, port):
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_cpu_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "MPCPU_OK" in r.stdout, r.stdout[-5000:]
