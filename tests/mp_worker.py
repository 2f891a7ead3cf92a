"""Multi-rank GPU parity worker (one process per GPU, launched by tests/test_gpu_multi.py through torchrun).

Each rank assembles ONLY its rows of a partitioned fixture (reference numbering for that partition),
builds the operators with halo exchange, and checks against the oracle's GLOBAL assembly of the same partition:
distributed SpMV on every matrix, the Chebyshev B- and Ap-solves, the operators, one filter application and the
full filtered-Lanczos solve (count + 1e-10 against the independent truth eigenvalues).  Prints MP_OK on rank 0.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def one_directional_halo(L, rank, nranks, torch):
    """Block upper-triangular rectangular matrix: rank r gathers columns of the ranks > r only, so the last rank sends
    and never receives (no arrival flag of its own to wait for).  60 products are queued back to back without a host
    sync: without an acknowledgement from the readers the sender would overwrite a parity buffer still being read."""
    import scipy.sparse as sp
    from normalmodes_b200 import _lib, matvec as mv
    nr, nc = 300, 4000                                              # rows / columns per rank
    S = sp.random(nr * nranks, nc * nranks, density=0.02, random_state=5, format="lil")
    for r in range(nranks):
        S[r * nr:(r + 1) * nr, :r * nc] = 0
    S = S.tocsr(); S.sort_indices()
    loc = S[rank * nr:(rank + 1) * nr]
    m = mv.COOmat([nr * r for r in range(nranks + 1)], loc.indptr, loc.indices, loc.data,
                  coldist=[nc * r for r in range(nranks + 1)])
    h = mv.parcsr_create(m)
    reps = 60
    xs = np.random.default_rng(3).uniform(-1, 1, (reps, nc * nranks))
    xd = torch.from_numpy(np.ascontiguousarray(xs[:, rank * nc:(rank + 1) * nc])).cuda()
    yd = torch.zeros(reps, nr, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for k in range(reps):
        _lib.check(L.nm_parcsr_matvec_dev(h, C.c_void_p(xd[k].data_ptr()), C.c_void_p(yd[k].data_ptr())))
    _lib.check(L.nm_sync())
    got = yd.cpu().numpy()
    for k in range(reps):
        ref = (S @ xs[k])[rank * nr:(rank + 1) * nr]
        assert np.abs(got[k] - ref).max() <= 1e-12 * (np.abs(ref).max() + 1.0), ("one-directional halo", rank, k)
    _lib.check(L.nm_parcsr_free(h))
    if rank == 0:
        print("MP one-directional halo ok on %d ranks" % nranks, flush=True)


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_case
    from normalmodes_b200 import _lib, matvec as mv, pevsl
    from normalmodes_b200.create_matrix import cg_create_matrix
    from oracle import fem as ofem, solver
    rank = int(os.environ["RANK"]); nranks = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    L = _lib.lib()
    _lib.check(L.nm_init(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = C.create_string_buffer(128)
        _lib.check(L.nm_comm_unique_id(raw))
        idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(idbuf, 0)
    _lib.check(L.nm_comm_init(rank, nranks, bytes(idbuf.cpu().numpy().tobytes())))

    def allgather(v):
        """concatenate the ranks' local slices (host arrays of different length)"""
        out = [None] * nranks
        dist.all_gather_object(out, np.asarray(v))
        return np.concatenate(out)

    for name in (os.environ.get("NM_MP_CASES", "prem3k_p1_j2,const3k_p2_j1")).split(","):
        c = load_case(name)
        g = c["g"]
        # slabs along z with equal node counts per rank (any partition is legal input; ParMETIS is external)
        from normalmodes_b200.create_matrix import Fem
        f0 = Fem(c["mesh"], c["model"]["vs"], g["porder"], nproc=1)
        nn = f0.nn
        if g["porder"] == 1:
            X = c["mesh"]["node"]
        else:
            from normalmodes_b200 import partition
            X = partition.node_coordinates(c["mesh"], f0)
        f0.free()
        order = np.argsort(X[:, 2], kind="stable")
        part = np.empty(nn, dtype=np.int32)
        for r in range(nranks):
            part[order[(r * nn) // nranks:((r + 1) * nn) // nranks]] = r
        mats, topo, num, geo = ofem.assemble(c["mesh"], c["model"], g["porder"], g["job"], part=part.astype(np.int64), nproc=nranks)
        CGM, f = cg_create_matrix(c["mesh"], c["model"], g["porder"], g["job"], nproc=nranks, part=part, rank=rank)
        rng = np.random.default_rng(7)
        # ---- distributed SpMV of every matrix against the oracle's global CSR product
        for k, m in CGM.items():
            S = ofem.to_scipy(mats[k])
            x = rng.uniform(-1, 1, S.shape[1])
            h = mv.parcsr_create(m)
            c0, c1 = int(m.coldist[rank]), int(m.coldist[rank + 1])
            r0, r1 = int(m.sizdist[rank]), int(m.sizdist[rank + 1])
            y = mv.parcsr_matvec(h, x[c0:c1], r1 - r0)
            ref = (S @ x)[r0:r1]
            # the rank's values come from the device assembly, S from the oracle's (they agree to ~1e-12 relative,
            # tests/test_gpu_assembly.py); a halo bug shows up as an O(1) relative error
            bound = 1e-10 * (abs(S) @ np.abs(x))[r0:r1] + 1e-300
            assert (np.abs(y - ref) <= bound).all(), (name, k, rank, np.abs(y - ref).max())
            info = mv.parcsr_info(h)
            assert nranks == 1 or info["nghost"] > 0 or m.NNZ == 0, (name, k, "no ghosts in a partitioned matrix?")
            _lib.check(L.nm_parcsr_free(h))
        # ---- operators + solves
        m = mv.setupmatvec(CGM, g["porder"], rank=rank, nproc=nranks)
        ops = solver.Operators(mats, g["porder"], bounds=dict(B=m.boundsB, Ap=getattr(m, "boundsAp", None)))
        r0, r1 = int(CGM["B"].sizdist[rank]), int(CGM["B"].sizdist[rank + 1])
        x = rng.uniform(-1, 1, ops.n)
        for nm_, got, ref in (("B~", mv.sparseBV(x[r0:r1], m), ops.bmv(x)), ("B-solve", mv.solveBV(x[r0:r1], m), ops.bsol(x)),
                              ("A op", mv.sparseAV(x[r0:r1], m), ops.amv(x))):
            err = np.abs(got - ref[r0:r1]).max() / np.abs(ref).max()
            assert err < 1e-11, (name, nm_, rank, err)
        # bounds agree with a serial Lanczos on the same matrix
        ob = solver.lanbounds(lambda v: ops.Bt @ v, ops.n, 1000, 2000, 1e-12)
        assert abs(m.boundsB[1] - ob[1]) < 1e-6 * ob[1] and abs(m.boundsB[0] - ob[0]) < 1e-3 * ob[1]
        # ---- full solve: eigenvalues do not depend on the partition
        low, up = g["lowfreq"], g["upfreq"]
        r = pevsl.pnm_apply_pevsl(m, low, up, recheck=False)
        truth = np.array(g["truth_eigs"])
        truth = truth[(truth >= r.xintv[0]) & (truth <= r.xintv[1])]
        assert r.nev == len(truth), (name, r.nev, len(truth))
        assert np.max(np.abs(r.eigval - truth) / truth) < 1e-10
        # eigenvectors: gathered residual in the oracle's operators
        if r.nev:
            i = r.nev // 2
            y = allgather(r.eigvec[i])
            res = np.linalg.norm(ops.amv(y) - r.eigval[i] * ops.bmv(y)) / abs(r.eigval[i])
            assert res < 1e-8, (name, res)
        if rank == 0:
            print("MP case %s ok on %d ranks: %d eigenpairs, %d Lanczos steps" % (name, nranks, r.nev, r.steps), flush=True)
        f.free()
    one_directional_halo(L, rank, nranks, torch)
    dist.barrier()
    _lib.check(L.nm_comm_finalize())
    dist.destroy_process_group()
    if rank == 0:
        print("MP_OK", flush=True)


if __name__ == "__main__":
    main()
