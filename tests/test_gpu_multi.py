"""Multi-GPU parity (one process per GPU over NCCL / NVLink): runs tests/mp_worker.py under torchrun on 2 GPUs
(and on every visible GPU when there are more).  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(nproc, port, env_extra=None, timeout=900):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    assert r.returncode == 0 and "MP_OK" in r.stdout, r.stdout[-6000:]
    return r.stdout


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_two_ranks_halo_exchange_and_solve():
    """Default: per-step ChebIter kernels chained by dependent launch, boundary rows stored into the peers' flag-in-data
    slots from the epilogue (no kernel, fence or flag between steps); P1 fluid-solid and P2 solid fixtures."""
    out = _run(2, 29611)
    print(out[-1500:])


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("env", [dict(NM_SLAB_PERS="1"), dict(NM_SLAB_PERS="1", NM_SLAB_FLOW="0"), dict(NM_HALO_FUSED="0"),
                                 dict(NM_HALO_FUSED="0", NM_HALO_OVERLAP="0"), dict(NM_P2P="0")])
def test_two_ranks_other_transports(env):
    """The persistent kernel with the in-kernel slot exchange (dataflow flags / grid barrier), the round-1 schemes
    (peer-window stores + arrival flags polled by the step kernel; flags awaited in k_halo_push) and the NCCL send/recv
    fallback."""
    out = _run(2, 29615, env_extra=dict(env, NM_MP_CASES="prem3k_p1_j2"))
    print(out[-800:])


@pytest.mark.skipif(_ngpu() < 4, reason="needs >= 4 GPUs")
def test_all_ranks_halo_exchange_and_solve():
    out = _run(min(_ngpu(), 8), 29613, env_extra={"NM_MP_CASES": "prem3k_p1_j2"})
    print(out[-1500:])
