"""CPU suite: the oracle against the committed golden vectors, host maths of the library, and the
C-ABI surface (library loads and exports every declared symbol; no compute calls without a GPU)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import load_case


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES_FAST = ["const3k_p1_j1", "const3k_p1_j2", "prem3k_p1_j2", "rtmdwak8k_p1_j2", "const3k_p2_j1", "prem3k_p2_j2"]


@pytest.mark.parametrize("name", CASES_FAST)
def test_oracle_assembly_matches_golden(name):
    """CSR pattern + DOF numbering integer-exact, values to rounding (same machine arithmetic)."""
    c = load_case(name)
    g = c["g"]
    assert c["num"]["N"] == g["N"] and c["num"]["Np"] == g["Np"]
    for k, gm in g["matrices"].items():
        m = c["mats"][k]
        assert list(m["shape"]) == gm["shape"] and m["ja"].size == gm["nnz"]
        assert _sha(m["ia"].astype(np.int32)) == gm["ia_sha256"]
        assert _sha(m["ja"].astype(np.int32)) == gm["ja_sha256"]
        assert np.isclose(m["a"].sum(), gm["sum"], rtol=1e-9, atol=1e-9 * gm["abssum"])
        assert np.isclose(np.abs(m["a"]).sum(), gm["abssum"], rtol=1e-12)


@pytest.mark.parametrize("name", ["const3k_p1_j1", "prem3k_p1_j2", "const3k_p2_j1"])
def test_oracle_invariants(name):
    """symmetry, B SPD (positive diagonal + diagonally scaled spectrum in (0, 5)), ET = E^T,
    rigid-body null space of A for the solid JOB 1 model."""
    from oracle import fem
    c = load_case(name)
    mats = c["mats"]
    B = fem.to_scipy(mats["B"])
    assert abs(B - B.T).max() <= 1e-12 * abs(B).max()
    assert (B.diagonal() > 0).all()
    if "A" in mats:
        A = fem.to_scipy(mats["A"])
        assert abs(A - A.T).max() <= 1e-9 * abs(A).max()
        if c["g"]["job"] == 1 and c["g"]["porder"] == 1:
            X = c["mesh"]["node"]
            n = X.shape[0]
            for comp in range(3):                      # translations
                t = np.zeros((n, 3)); t[:, comp] = 1.0
                assert np.abs(A @ t.ravel()).max() <= 1e-7 * abs(A).max()
            for ax in range(3):                        # rotations
                w = np.zeros(3); w[ax] = 1.0
                r = np.cross(np.broadcast_to(w, X.shape), X)
                assert np.abs(A @ r.ravel()).max() <= 1e-7 * abs(A).max() * np.abs(X).max()
    else:
        E = fem.to_scipy(mats["E"]); ET = fem.to_scipy(mats["ET"])
        assert abs(E - ET.T).max() == 0.0
        Ap = fem.to_scipy(mats["Ap"])
        assert (Ap.diagonal() < 0).all()


def test_oracle_filtered_lanczos_vs_truth(golden):
    """The oracle's restatement of the pEVSL path finds every eigenvalue of the independent dense solve
    (count exact, 1e-10 relative) on the reference's own demo configuration (demos/global_conf)."""
    from oracle import solver
    c = load_case("const3k_p1_j1")
    g = c["g"]
    ops, lam, Y, res, info = solver.solve(c["mats"], 1, g["lowfreq"], g["upfreq"])
    truth = np.array(g["truth_eigs"])
    assert len(lam) == len(truth) == 271
    assert np.max(np.abs(lam - truth) / truth) < 1e-10
    # B-orthonormality and the reference's RMS residual (README: "typically around 1e-13")
    G = Y.T @ (ops.Bt @ Y)
    assert np.abs(G - np.eye(len(lam))).max() < 1e-8
    rms = max(solver.residual_rms(ops, lam[i], Y[:, i]) for i in range(0, len(lam), 10))
    assert rms < 1e-12


def test_oracle_chebiter_accuracy():
    """App. D table: degree-25 Chebyshev solve of the P1 Jacobi-scaled mass matrix is ~1e-11 accurate."""
    from oracle import solver, fem
    c = load_case("const3k_p1_j1")
    Bs, d = fem.jacobi_scale(c["mats"]["B"])
    Bt = fem.to_scipy(Bs)
    lb, ub = solver.lanbounds(lambda v: Bt @ v, Bt.shape[0], 1000, 2000, 1e-12)
    assert 0.55 < lb < 0.57 and abs(ub - 2.5) < 1e-6
    b = np.random.default_rng(0).standard_normal(Bt.shape[0])
    x = solver.chebiter(Bt, lb, ub, 25, b)
    r = np.linalg.norm(b - Bt @ x) / np.linalg.norm(b)
    assert r < 2e-11


def test_freq_interval_float32_semantics():
    from oracle import solver
    from normalmodes_b200 import pevsl
    a, b = solver.freq_interval(0.2, 2.0, -1.0)
    assert abs(a - 1.579136839123207e-06) < 1e-20 and abs(b - 0.0001579136792061263) < 1e-18
    assert pevsl.freq_interval(0.2, 2.0, -1.0) == (a, b)
    assert solver.freq_interval(0.0, 1.0, -3.0)[0] == -3.0


# ------------------------------------------------------------------ library surface (no GPU needed)
def test_library_builds_and_exports_declared_symbols():
    from normalmodes_b200 import _lib
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    for n in ["pevsl_start_f90_", "pevsl_parcsrcreate_f90_", "pevsl_parcsrmatvec_f90_", "pevsl_cheblannr_f90_",
              "pevsl_copy_result_f90_", "pevsl_setup_chebiter_f90_", "pevsl_chebiter_f90_", "pevsl_lanbounds_f90_",
              "pevsl_findpol_f90_", "pevsl_freepol_f90_", "pevsl_finish_f90_", "pevsl_setamv_f90_", "pevsl_setbmv_f90_",
              "pevsl_setbsol_chebiter_f90_", "pevsl_set_geneig_f90_", "pevsl_get_nev_f90_", "pevsl_setprobsizes_f90_",
              "pevsl_chebiterstatsprint_f90_"]:
        assert n in names


def test_findpol_host_matches_oracle():
    """nm_findpol (host C++) against the oracle's find_pol: same degree, same coefficients."""
    from oracle import solver
    from normalmodes_b200 import pevsl
    for xintv in ([1.579136839123207e-06, 0.0001579136792061263, -2.77e-08, 0.0028953701550054407],
                  [3.9478420978080176e-07, 3.9478419801531574e-05, -4.677924941383012e-07, 0.02812179603432921],
                  [-1.0e-8, 2.0e-5, -1.0e-8, 3.0e-3],          # touches the left end
                  [2.0e-3, 3.0e-3, -1.0e-8, 3.0e-3]):           # touches the right end
        ref = solver.findpol(xintv, 0.8, 0.7)
        pol = pevsl.Pol(xintv, 0.8, 0.7)
        assert pol.deg == ref["deg"]
        assert np.allclose(pol.mu, ref["mu"], rtol=1e-11, atol=1e-13)
        assert abs(pol.bar - ref["bar"]) < 1e-11 and abs(pol.gam - ref["gam"]) < 1e-11
        assert abs(pol.cc - ref["cc"]) <= 1e-18 and abs(pol.dd - ref["dd"]) <= 1e-18
        pol.free()


def test_tridiag_host_matches_lapack():
    import scipy.linalg as sla
    from normalmodes_b200._lib import lib, check, dptr
    rng = np.random.default_rng(3)
    for k in (1, 2, 7, 150):
        d = rng.standard_normal(k); e = rng.standard_normal(max(k - 1, 1))
        w = np.empty(k); Z = np.empty((k, k)); lr = np.empty(k)
        check(lib().nm_tridiag_eig_host(k, dptr(d), dptr(e), dptr(w), dptr(Z), dptr(lr)))
        if k == 1:
            assert w[0] == d[0]; continue
        wr, Vr = sla.eigh_tridiagonal(d, e[:k - 1])
        assert np.allclose(w, wr, atol=1e-12)
        V = Z.T                                                 # column-major k*k -> V[:, j]
        T = np.diag(d) + np.diag(e[:k - 1], 1) + np.diag(e[:k - 1], -1)
        assert np.abs(T @ V - V * w).max() < 1e-11
        assert np.allclose(np.abs(lr), np.abs(V[-1, :]), atol=1e-11)


def test_error_reporting_without_gpu():
    """A compute call on a box without a CUDA device must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from normalmodes_b200._lib import lib
    L = lib()
    h = C.c_void_p()
    rc = L.nm_pevsl_create(C.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in L.nm_last_error_message()


@pytest.mark.parametrize("name", ["const3k_p1_j1", "prem3k_p1_j2"])
def test_c_port_matches_numpy_oracle(name):
    """oracle/c (the CPU baseline bench.py times) against oracle/solver.py: operator, B-solve, ChebAv."""
    from oracle import cpu as ocpu, fem, solver
    c = load_case(name)
    mats = c["mats"]
    ops = solver.Operators(mats, 1)
    Bs, d = fem.jacobi_scale(mats["B"])
    t = lambda m: (m["ia"], m["ja"], m["a"])
    if ops.fluid:
        Aps, dp = fem.jacobi_scale(mats["Ap"], -1.0)
        co = ocpu.CpuOps(t(Bs), t(mats["Ad"]), d, ops.boundsB, ops.degB, E=t(mats["E"]), ET=t(mats["ET"]), Ap=t(Aps), dp=dp,
                         boundsAp=ops.boundsAp, degAp=ops.degAp)
    else:
        co = ocpu.CpuOps(t(Bs), t(mats["A"]), d, ops.boundsB, ops.degB)
    v = np.random.default_rng(0).standard_normal(ops.n)
    ref = ops.amv(v)
    assert np.abs(co.apply_A(v) - ref).max() <= 1e-13 * np.abs(ref).max()
    ref = ops.bsol(v)
    assert np.abs(co.bsol(v) - ref).max() <= 1e-13 * np.abs(ref).max()
    pol = solver.findpol([4e-7, 4e-5, -1e-8, 3e-3], 0.8, 0.7)
    pol["deg"] = min(pol["deg"], 6); pol["mu"] = pol["mu"][:pol["deg"] + 1]
    ref = solver.chebav(pol, v, ops)
    got = co.chebav(pol["deg"], pol["mu"], pol["cc"], pol["dd"], v)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()


def test_truth_eigenvalues_against_survey_probe_and_physics(golden):
    """Pins that do not come from the oracle's own code:
    (a) the survey session's throw-away probe of CONST3k -- standard P1 elasticity + consistent mass with dense LAPACK,
        written independently of oracle/fem.py -- found 271 eigenvalues in the demo band, the lowest at 0.37418, 0.37422,
        0.37459, 0.37515, 0.37527 mHz (0T2), then 0.39331 ... 0.39400 (0S2), 0.51220, and lambda_max(B^-1 A) = 2.89537e-3;
    (b) the analytic 0T2 of a homogeneous ball, (l-1) j_l(x) = x j_{l+1}(x) => x = 2.501, f = x vs / (2 pi a) = 0.361 mHz,
        multiplicity 2l+1 = 5 (3 k elements: percent-level discretisation error);
    (c) multiplet structure 2l+1 of the spherical demo models (splitting by the unstructured mesh stays at the 1e-3 level)."""
    g = golden["const3k_p1_j1"]
    f = np.sqrt(np.array(g["truth_eigs"])) / (2 * np.pi) * 1e3
    assert len(f) == 271
    assert np.allclose(f[:5], [0.37418, 0.37422, 0.37459, 0.37515, 0.37527], atol=6e-6)
    assert np.allclose(f[5:11], [0.39331, 0.39344, 0.39357, 0.39373, 0.39400, 0.51220], atol=6e-6)
    assert abs(g["oracle_lanczos"]["bounds"][1] - 2.89537e-3) < 1e-8
    f0t2 = 2.501 * 5.7735 / (2 * np.pi * 6371.0) * 1e3
    assert abs(f[:5].mean() / f0t2 - 1.0) < 0.05 and f[5] - f[4] > 10 * (f[4] - f[0])

    def multiplets(fr, rel=4e-3):
        out = [1]
        for a, b in zip(fr[:-1], fr[1:]):
            if (b - a) < rel * b:
                out[-1] += 1
            else:
                out.append(1)
        return out
    fm = np.sqrt(np.array(golden["mtopo100k_p1_j1"]["truth_eigs"])) / (2 * np.pi) * 1e3
    assert multiplets(fm) == [5, 5, 3, 7]                      # Moon model: l = 2, 2, 1, 3
    fp = np.sqrt(np.array(golden["prem3k_p2_j2"]["truth_eigs"])) / (2 * np.pi) * 1e3
    assert multiplets(fp) == [3, 5, 5, 3]                      # PREM, P2, gravity: a triplet (l = 1), two quintuplets, a triplet
    fc = np.sqrt(np.array(golden["const3k_p2_j1"]["truth_eigs"])) / (2 * np.pi) * 1e3
    assert multiplets(fc)[:3] == [5, 5, 3]
    assert abs(fc[:5].mean() / f0t2 - 1.0) < 0.012            # P2 on the same mesh: 0T2 within 1.2 % of the analytic value
