"""GPU parity of the FE assembly (K9): element integration + scatter on the device against the oracle's
restatement of CGE3D_ISO / CGFSE3D_ISO on the reference's demo meshes, through the C ABI.

Tolerance: the device sums an entry's element contributions with fp64 atomics (order not fixed), the
reference in ascending element order; each contribution also differs by reassociation inside the dense
element products.  Bound used: |dv| <= 2e-12 * (largest |entry| of that matrix row scale)."""
import numpy as np
import pytest

from conftest import load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nm():
    import torch
    assert torch.cuda.is_available()
    from normalmodes_b200 import _lib
    L = _lib.lib()
    _lib.check(L.nm_init(0))
    return L


@pytest.mark.parametrize("name", ["const3k_p1_j1", "const3k_p1_j2", "prem3k_p1_j2", "const3k_p2_j1", "prem3k_p2_j2",
                                  "rtmdwak8k_p1_j2"])
def test_assembly_values_match_oracle(nm, name):
    from normalmodes_b200.create_matrix import cg_create_matrix
    c = load_case(name)
    g = c["g"]
    CGM, fem = cg_create_matrix(c["mesh"], c["model"], g["porder"], g["job"])
    assert set(CGM) == set(c["mats"])
    for k, m in CGM.items():
        ref = c["mats"][k]
        assert (m.rowdist == ref["ia"]).all() and (m.col == ref["ja"]).all(), k     # pattern: bit-exact
        scale = np.abs(ref["a"]).max()
        err = np.abs(m.val - ref["a"]).max()
        assert err <= 2e-12 * scale, (k, err, scale)
    fem.free()


def test_assembly_partitioned_rows(nm):
    """Each rank assembles only the rows it owns (no value communication): rank blocks == oracle row slices."""
    from oracle import fem as ofem
    from normalmodes_b200.create_matrix import cg_create_matrix
    c = load_case("prem3k_p1_j2")
    g = c["g"]
    nproc = 2
    part = (c["mesh"]["node"][:, 2] > 0).astype(np.int32)             # two half-balls
    mats, topo, num, geo = ofem.assemble(c["mesh"], c["model"], g["porder"], g["job"], part=part.astype(np.int64), nproc=nproc)
    for rank in range(nproc):
        CGM, f = cg_create_matrix(c["mesh"], c["model"], g["porder"], g["job"], nproc=nproc, part=part, rank=rank)
        for k, m in CGM.items():
            ref = mats[k]
            dist = num["psizdist"] if k in ("ET", "Ap") else num["sizdist"]
            r0, r1 = int(dist[rank]), int(dist[rank + 1])
            sl = slice(ref["ia"][r0], ref["ia"][r1])
            assert (m.col == ref["ja"][sl]).all()
            assert np.abs(m.val - ref["a"][sl]).max() <= 2e-12 * np.abs(ref["a"]).max()
        f.free()


def test_end_to_end_from_mesh_files_prem_like(nm):
    """Builder-generated PREM-like mesh -> device assembly -> setupmatvec -> filtered Lanczos, checked against the
    oracle's independent shift-invert eigenvalues of the oracle-assembled pencil (count + 1e-10)."""
    from oracle import fem as ofem, solver
    from normalmodes_b200 import meshgen, matvec as mv, pevsl
    from normalmodes_b200.create_matrix import cg_create_matrix
    mesh = meshgen.build_mesh(2500, seed=1)
    model = meshgen.build_model(mesh, 1)
    CGM, f = cg_create_matrix(mesh, model, 1, 2)
    m = mv.setupmatvec(CGM, 1)
    r = pevsl.pnm_apply_pevsl(m, 0.3, 1.2)
    omesh = dict(ntet=mesh["ntet"], nvert=mesh["nvert"], ele=mesh["ele"], neigh=mesh["neigh"], node=mesh["node"])
    mats, topo, num, geo = ofem.assemble(omesh, model, 1, 2)
    truth = solver.truth_eigs(mats, r.xintv[0], r.xintv[1])
    assert r.nev == len(truth) and r.nev > 5
    assert np.max(np.abs(r.eigval - truth) / truth) < 1e-10
    # the reference's result files (src/mod_pevsl.f90:188-201,225-243) written from the GPU solve and read back: eigenvector i
    # in physical coordinates x = d * y is an eigenvector of the UNSCALED oracle pencil
    import tempfile
    from normalmodes_b200 import io
    with tempfile.TemporaryDirectory() as tmp:
        names = io.output_names(tmp + "/", "prem_like", 2, 1, 1, 0.3, 1.2)
        files = pevsl.pnm_save_eigenvectors(m, r, names["fvdata"])
        assert len(files) == r.nev and files[0].endswith("_0.300000012_1.20000005_1.dat")
        A, B = solver.effective_pencil(mats)
        for i in (0, r.nev // 2, r.nev - 1):
            x = np.fromfile(files[i], dtype="<f8")
            assert x.size == m.Gpbsiz
            res = np.linalg.norm(A @ x - r.eigval[i] * (B @ x)) / (abs(r.eigval[i]) * np.linalg.norm(B @ x))
            assert res < 1e-7, (i, res)
    f.free()
