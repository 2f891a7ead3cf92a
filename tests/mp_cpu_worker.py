"""World-size-2 (or more) CPU worker over gloo: the host side of the multi-GPU path -- per-rank numbering + pattern
(matrixstruct for a partition), the receive side of the halo plan (nm_halo_plan_host, the routine the device plan is
built from) and a halo exchange + distributed SpMV carried out with gloo -- against the oracle's GLOBAL matrices.
Launched by tests/test_multirank_cpu.py.  Prints MPCPU_OK on rank 0."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch.distributed as dist
    from conftest import load_case
    from normalmodes_b200._lib import lib, check, iptr, i32
    from normalmodes_b200.create_matrix import Fem
    from oracle import fem as ofem
    rank = int(os.environ["RANK"]); P = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    for name in ("prem3k_p1_j2", "const3k_p2_j1"):
        c = load_case(name)
        g = c["g"]
        f0 = Fem(c["mesh"], c["model"]["vs"], g["porder"], nproc=1)
        nn = f0.nn
        f0.free()
        # an arbitrary but deterministic partition: round-robin blocks of 37 nodes (many ghosts, every rank talks to every rank)
        part = ((np.arange(nn) // 37) % P).astype(np.int32)
        mats, topo, num, geo = ofem.assemble(c["mesh"], c["model"], g["porder"], g["job"], part=part.astype(np.int64), nproc=P)
        f = Fem(c["mesh"], c["model"]["vs"], g["porder"], nproc=P, part=part, rank=rank)
        names = ("Ad", "B", "E", "ET", "Ap") if f.fluidcase else ("A", "B")
        rng = np.random.default_rng(3)
        for k in names:
            m = f.matrix(k, values=False)
            S = ofem.to_scipy(mats[k if k != "Ad" or "Ad" in mats else "A"])
            # every rank sees the same distribution, and the blocks tile the global matrix
            dists = [None] * P
            dist.all_gather_object(dists, (m.sizdist.tolist(), m.coldist.tolist()))
            assert all(d == dists[0] for d in dists), (name, k)
            assert m.sizdist[-1] == S.shape[0] and m.coldist[-1] == S.shape[1]
            r0, r1 = int(m.sizdist[rank]), int(m.sizdist[rank + 1])
            c0, c1 = int(m.coldist[rank]), int(m.coldist[rank + 1])
            ref = S[r0:r1].tocsr(); ref.sort_indices()
            assert (m.rowdist == ref.indptr).all() and (m.col == ref.indices).all(), (name, k, "pattern of the rank block")
            # receive side of the halo plan from the library (host-only entry point)
            nghost = np.zeros(1, dtype=np.int32); ghosts = np.empty(max(m.NNZ, 1), dtype=np.int32); rc = np.zeros(P, dtype=np.int32)
            check(lib().nm_halo_plan_host(P, rank, iptr(m.coldist), m.NNZ, iptr(m.col), iptr(nghost), iptr(ghosts), iptr(rc)))
            ghosts = ghosts[:nghost[0]]
            assert (np.diff(ghosts) > 0).all() and rc.sum() == nghost[0] and rc[rank] == 0
            owner = np.searchsorted(m.coldist, ghosts, side="right") - 1
            assert (np.bincount(owner, minlength=P) == rc).all() and (np.diff(owner) >= 0).all()
            # send side: every owner learns which of its columns are needed (what the device plan does with NCCL)
            need = [ghosts[owner == r].tolist() for r in range(P)]
            gathered = [None] * P
            dist.all_gather_object(gathered, need)
            asked = [gathered[s][rank] for s in range(P)]           # columns rank s needs from me
            for s in range(P):
                assert all(c0 <= cg < c1 for cg in asked[s]), (name, k, "peer asked for a column this rank does not own")
            # exchange of x values + local product in the local column space (owned first, ghosts after)
            x = rng.uniform(-1, 1, S.shape[1])
            sendvals = [x[np.array(asked[s], dtype=np.int64)].tolist() if asked[s] else [] for s in range(P)]
            got = [None] * P
            dist.all_gather_object(got, sendvals)
            xg = np.concatenate([np.array(got[r][rank], dtype=float) for r in range(P)]) if nghost[0] else np.zeros(0)
            assert xg.size == nghost[0]
            lcol = np.where((m.col >= c0) & (m.col < c1), m.col - c0, (c1 - c0) + np.searchsorted(ghosts, m.col))
            vals = ref.data
            Sl = sp.csr_matrix((vals, lcol, m.rowdist), shape=(r1 - r0, (c1 - c0) + nghost[0]))
            y = Sl @ np.concatenate([x[c0:c1], xg])
            yref = (S @ x)[r0:r1]
            assert np.abs(y - yref).max() <= 1e-12 * max(np.abs(yref).max(), 1e-300), (name, k)
        f.free()
        if rank == 0:
            print("MPCPU case %s ok on %d ranks" % (name, P), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MPCPU_OK", flush=True)


if __name__ == "__main__":
    main()
