"""Host-side maths and bench plumbing that need no GPU: the dense symmetric-definite eigensolver behind the Rayleigh-Ritz
refinement (nm_sym_geneig_host), the P2 node count bench.py sizes its partition vector with, and the stacking of the
ranks' CSR row blocks into the global matrices the CPU oracle multiplies by (bench.py `check` at N > 1)."""
import ctypes as C

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

from conftest import load_case


def _geneig(H, G):
    from normalmodes_b200._lib import lib, check, dptr
    m = H.shape[0]
    w = np.empty(m); Cm = np.empty((m, m), order="F")
    check(lib().nm_sym_geneig_host(m, dptr(np.ascontiguousarray(H.T)), dptr(np.ascontiguousarray(G.T)), dptr(w),
                                   Cm.ctypes.data_as(C.POINTER(C.c_double))))
    return w, Cm


def test_sym_geneig_nearly_diagonal_and_dense():
    rng = np.random.default_rng(0)
    # the real use: H nearly diagonal (accepted Ritz vectors), G ~ I, eigenvalues over two decades
    m = 300
    w0 = np.sort(10 ** rng.uniform(-6, -4, m))
    E = rng.standard_normal((m, m)) * 1e-11 * w0.mean()
    H = np.diag(w0) + (E + E.T) / 2
    G = np.eye(m) + 1e-15 * rng.standard_normal((m, m)); G = (G + G.T) / 2
    w, Cm = _geneig(H, G)
    wr = sla.eigh(H, G, eigvals_only=True)
    assert np.abs((w - wr) / wr).max() < 1e-13
    assert np.abs(Cm.T @ G @ Cm - np.eye(m)).max() < 1e-12
    assert np.abs(H @ Cm - (G @ Cm) * w).max() <= 1e-13 * np.abs(w).max()
    # a dense pencil, clustered and degenerate eigenvalues included
    for m in (1, 2, 7, 60):
        Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
        w0 = np.sort(10 ** rng.uniform(-6, -4, m))
        if m > 4:
            w0[2] = w0[1]                                       # exact multiplicity
        H = (Q * w0) @ Q.T; H = (H + H.T) / 2
        X = rng.standard_normal((m, m)) * 1e-3
        G = np.eye(m) + X @ X.T
        w, Cm = _geneig(H, G)
        wr = sla.eigh(H, G, eigvals_only=True)
        assert np.abs(w - wr).max() <= 1e-12 * np.abs(wr).max()
        assert np.abs(Cm.T @ G @ Cm - np.eye(m)).max() < 1e-12


def test_sym_geneig_rejects_indefinite_G():
    import pytest
    from normalmodes_b200._lib import NmError
    with pytest.raises(NmError):
        _geneig(np.eye(3), np.diag([1.0, -1.0, 1.0]))


def test_p2_node_count_matches_topology():
    import bench
    from normalmodes_b200 import meshgen
    from normalmodes_b200.create_matrix import Fem
    mesh = meshgen.build_mesh(2500, seed=1)
    model = meshgen.build_model(mesh, 2)
    f = Fem(mesh, model["vs"], 2, nproc=1)
    assert bench.p2_node_count(mesh) == f.nn
    f.free()
    c = load_case("const3k_p2_j1")
    f2 = Fem(c["mesh"], c["model"]["vs"], 2, nproc=1)
    assert bench.p2_node_count(dict(ele=c["mesh"]["ele"], nvert=c["mesh"]["nvert"])) == f2.nn
    f2.free()


def test_bench_stacks_rank_blocks_into_global_csr():
    """bench.py gathers the ranks' row blocks (global column ids) for the CPU oracle: stacking them in rank order must give
    the global matrix."""
    import bench
    rng = np.random.default_rng(1)
    n = 90
    S = (sp.random(n, n, density=0.1, random_state=2, format="csr") + sp.identity(n)).tocsr()
    S.sort_indices()
    cuts = [0, 25, 60, n]
    parts = []
    for r in range(3):
        blk = S[cuts[r]:cuts[r + 1]]
        d = dict(d=np.ones(cuts[r + 1] - cuts[r]))
        for name in ("B", "A"):
            d[name + "_ia"] = blk.indptr.astype(np.int32); d[name + "_ja"] = blk.indices.astype(np.int32); d[name + "_a"] = blk.data.copy()
        parts.append(d)

    class MV:
        boundsB = (0.5, 3.0); degB = 5
    ops = bench.cpu_ops_from_arrays(parts, MV, False)
    x = rng.standard_normal(n)
    y = ops.apply_A(x)
    assert np.abs(y - S @ x).max() <= 1e-13 * np.abs(y).max()


def test_bench_workload_partition_host_side():
    """bench.py's multi-rank workload construction without a GPU: P2 node count without a topology probe, RCB parts of
    equal size, and per-rank row blocks that tile the global problem for 2, 3 and 8 ranks."""
    import bench
    from normalmodes_b200 import meshgen, partition
    from normalmodes_b200.create_matrix import Fem
    mesh = meshgen.build_mesh(2500, seed=1)
    model = meshgen.build_model(mesh, 2)
    f1 = Fem(mesh, model["vs"], 2, nproc=1)
    N1, Np1 = f1.N, f1.Np
    f1.free()
    for nranks in (2, 3, 8):
        nn = bench.p2_node_count(mesh)
        f0 = Fem(mesh, model["vs"], 2, nproc=nranks, part=np.zeros(nn, dtype=np.int32), rank=0)
        assert f0.nn == nn
        X = partition.node_coordinates(mesh, f0)
        f0.free()
        part = partition.rcb(X, nranks)
        cnt = np.bincount(part, minlength=nranks)
        assert cnt.max() - cnt.min() <= nranks                  # recursive bisection: one node of slack per level
        rows = 0; prows = 0
        for r in range(nranks):
            f = Fem(mesh, model["vs"], 2, nproc=nranks, part=part, rank=r)
            assert (f.N, f.Np) == (N1, Np1)                    # the global sizes do not depend on the partition
            B = f.matrix("B", values=False)
            Ap = f.matrix("Ap", values=False)
            rows += B.siz(r); prows += Ap.siz(r)
            assert B.col.min() >= 0 and B.col.max() < N1
            f.free()
        assert rows == N1 and prows == Np1
