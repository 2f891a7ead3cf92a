"""Host-side checks of the warp-sliced ELL slab packer behind k_slab (normalmodes_b200/csrc/nm_slab.cu): the
library packs the pattern and walks its own blobs on the CPU exactly as the kernel does (nm_slab_host_selftest);
the result must equal the CSR product.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_case


def _selftest(S, R, x, ncolb=None):
    from normalmodes_b200._lib import lib, check, dptr, iptr, i32, f64
    S = S.tocsr(); S.sort_indices()
    n = S.shape[0]
    ncolb = ncolb or S.shape[1]
    rp, idx, vals = i32(S.indptr), i32(S.indices), f64(S.data)
    y = np.full(R * n, np.nan); order = np.empty(n, dtype=np.int32); info = np.zeros(10, dtype=np.int32)
    check(lib().nm_slab_host_selftest(n, ncolb, R, iptr(rp), iptr(idx), dptr(vals), dptr(f64(x)), dptr(y), iptr(order), iptr(info)))
    return y, order, dict(zip(("nchunk", "grid", "threads", "smem", "nstage", "maxper", "padded", "sum_nd", "x_wavefronts", "x_ideal"), info.tolist()))


CONFIGS = [dict(), dict(NM_SLAB_THREADS="64"), dict(NM_SLAB_THREADS="512", NM_SLAB_SPLIT="8"),
           dict(NM_SLAB_THREADS="512", NM_SLAB_SPLIT="4", NM_SLAB_MAXGRID="2"), dict(NM_SLAB_THREADS="128", NM_SLAB_SPLIT="32"), dict(NM_SLAB_THREADS="256", NM_SLAB_STAGES="3"),
           dict(NM_SLAB_ENTRIES="200", NM_SLAB_DISTINCT="90", NM_SLAB_MAXGRID="2"),
           dict(NM_SLAB_THREADS="64", NM_SLAB_ENTRIES="96", NM_SLAB_DISTINCT="40", NM_SLAB_MAXGRID="1", NM_SLAB_STAGES="2"),
           dict(NM_PACK_ORDER="0", NM_PACK_BANK_AWARE="0"), dict(NM_SLAB_SPLIT="8"),
           dict(NM_SLAB_SPLIT="4", NM_SLAB_THREADS="64", NM_SLAB_MAXGRID="3")]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_slab_pack_walk_equals_csr_product(monkeypatch, cfg):
    from oracle import fem
    for k, v in cfg.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(5)
    # scalar mass pattern M of B = M (x) I3 (R = 3) and the scalar Ap (R = 1) of the fluid-solid fixture
    c = load_case("prem3k_p1_j2")
    B = fem.to_scipy(c["mats"]["B"]).tocsr()
    Ms = B[::3, ::3].tocsr()                                  # one entry per 3x3 block
    assert abs(sp.kron(Ms, sp.identity(3)) - B).max() == 0.0
    for S, R in ((Ms, 3), (fem.to_scipy(c["mats"]["Ap"]).tocsr(), 1)):
        n = S.shape[0]
        x = rng.uniform(-1, 1, R * n)
        y, order, info = _selftest(S, R, x)
        assert sorted(order.tolist()) == list(range(n))
        A = sp.kron(S, sp.identity(R)).tocsr() if R > 1 else S
        ref = (A @ x).reshape(n, R)[order].ravel()            # pack order
        tol = 64 * np.finfo(float).eps * (abs(A) @ np.abs(x)).reshape(n, R)[order].ravel() + 1e-300
        assert (np.abs(y - ref) <= tol).all(), (cfg, R, np.abs(y - ref).max())
        assert info["smem"] <= 226 * 1024 and info["maxper"] <= 256
        if "NM_SLAB_MAXGRID" in cfg:
            assert info["maxper"] > 1                          # several chunks per CTA: the ring wraps
        nnz = S.nnz
        assert info["padded"] >= nnz
        if not cfg:
            assert info["padded"] <= 1.5 * nnz, info          # padding stays bounded (warps hold lanes of similar length)
            # modelled shared-memory wavefronts of the x reads against the conflict-free count (share-aware schedule)
            assert info["x_wavefronts"] <= 1.45 * info["x_ideal"], info


def test_slab_pack_with_ghost_columns(monkeypatch):
    """Rows of one rank of a partitioned matrix: ghost columns (ids >= n) keep their id, owned ones are renumbered."""
    rng = np.random.default_rng(6)
    n, ng = 500, 60
    S = sp.random(n, n + ng, density=0.03, random_state=3, format="csr") + sp.hstack([sp.identity(n), sp.csr_matrix((n, ng))])
    S = S.tocsr()
    for R in (1, 3):
        x = rng.uniform(-1, 1, R * (n + ng))
        y, order, info = _selftest(S, R, x, ncolb=n + ng)
        A = sp.kron(S, sp.identity(R)).tocsr() if R > 1 else S
        ref = (A @ x).reshape(n, R)[order].ravel()
        assert np.abs(y - ref).max() <= 1e-13 * np.abs(ref).max()


def test_slab_pack_ragged_rows(monkeypatch):
    """Empty rows, one very long row, single-row matrix."""
    rng = np.random.default_rng(8)
    n = 300
    S = sp.random(n, n, density=0.02, random_state=1, format="lil")
    S[7, :] = 0.0                                # empty row
    S[11, :200] = rng.uniform(1, 2, 200)         # long row
    S = S.tocsr(); S.eliminate_zeros()
    x = rng.uniform(-1, 1, n)
    y, order, info = _selftest(S, 1, x)
    assert np.abs(y - (S @ x)[order]).max() <= 1e-13
    S1 = sp.csr_matrix(np.array([[2.5]]))
    y, order, info = _selftest(S1, 3, np.array([1.0, 2.0, 3.0]))
    assert np.allclose(y, [2.5, 5.0, 7.5])
