"""Readers / result writers in the reference's raw-binary formats: round trips on CPU."""
import numpy as np

from conftest import load_case


def test_mesh_and_model_files_round_trip(tmp_path):
    from normalmodes_b200 import meshgen, io
    mesh = meshgen.build_mesh(600, seed=3)
    model = meshgen.build_model(mesh, 2, gravity=True)
    meshgen.write_files(str(tmp_path) + "/", "toy.1", mesh, model, 2)
    m2 = io.read_mesh(str(tmp_path / "toy.1"))
    assert m2["ntet"] == mesh["ntet"] and m2["nvert"] == mesh["nvert"]
    assert (m2["ele"] == mesh["ele"]).all() and (m2["neigh"] == mesh["neigh"]).all() and (m2["node"] == mesh["node"]).all()
    mod2 = io.read_model(str(tmp_path / "toy.1"), mesh["ntet"], 2, 2)
    for k in ("vp", "vs", "rho", "g0"):
        assert (mod2[k] == model[k]).all(), k


def test_output_names_follow_mod_para():
    from normalmodes_b200 import io
    n = io.output_names("out/", "CONST_1L_3k.1", 1, 1, 2, 0.2, 2.0)
    assert n["fvlist"] == "out/CONST_1L_3k.1_pod1_np2_vlist.dat"
    assert n["fvstat"] == "out/CONST_1L_3k.1_pod1_np2_vstat.dat"
    assert n["fvdata"] == "out/CONST_1L_3k.1_JOB1_pod1_np2_0.200000003_2.00000000"
    # gfortran's list-directed REAL(4): 9 significant digits, F form on [0.1, 1e9), ES form outside
    for v, txt in ((0.05, "5.00000007E-02"), (10.0, "10.0000000"), (12.5, "12.5000000"), (0.1, "0.100000001"),
                   (1.0, "1.00000000"), (123.456, "123.456001")):
        assert io._list_directed_real4(v) == txt, (v, io._list_directed_real4(v))


def test_rank_slices_assemble_the_reference_files(tmp_path):
    """Two 'ranks' write their slices at their byte offsets (as mpi_file_set_view + mpi_file_write do), in any order:
    eigenvectors in physical coordinates d*y in global DOF order, vlist = original node ids rank after rank."""
    from normalmodes_b200 import io
    from normalmodes_b200.create_matrix import Fem
    c = load_case("prem3k_p1_j2")
    g = c["g"]
    nproc = 2
    part = (c["mesh"]["node"][:, 2] > 0).astype(np.int32)
    names = io.output_names(str(tmp_path) + "/", g["basename"], g["job"], g["porder"], nproc, g["lowfreq"], g["upfreq"])
    rng = np.random.default_rng(0)
    nev = 3
    pieces, dpieces, vl_all, vs_all, N = {}, {}, [], [], None
    for rank in (1, 0):                                           # out of order on purpose
        f = Fem(c["mesh"], c["model"]["vs"], g["porder"], nproc=nproc, part=part, rank=rank)
        B = f.matrix("B", values=False)
        num = f.numbering()
        n_loc = B.siz(rank); N = B.Gsiz
        y = rng.standard_normal((nev, n_loc)); d = rng.uniform(0.5, 2.0, n_loc)
        pieces[rank], dpieces[rank] = y, d
        io.save_eigenvectors(names["fvdata"], y, d, B.sizdist, rank)
        order = num["order"]                                      # position -> original node id (0-based), ranks concatenated
        counts = np.bincount(part, minlength=nproc)
        vtxdist = np.concatenate([[0], np.cumsum(counts)])
        sl = slice(vtxdist[rank], vtxdist[rank + 1])
        io.save_vlist_vstat(names, order[sl] + 1, vtxdist, rank, vstat_local=num["vstat"][order[sl]])
        f.free()
    for i in range(nev):
        x = io.read_eigenvector(names["fvdata"], i + 1, N)
        ref = np.concatenate([pieces[r][i] * dpieces[r] for r in range(nproc)])
        assert (x == ref).all()
    vl = np.fromfile(names["fvlist"], dtype="<i4")
    assert sorted(vl.tolist()) == list(range(1, len(part) + 1))
    assert (part[vl - 1] == np.repeat(np.arange(nproc), np.bincount(part, minlength=nproc))).all()
    for r in range(nproc):                                        # ascending original id inside a rank (mod_geometry.f90:953-965)
        blk = vl[part[vl - 1] == r]
        assert (np.diff(blk) > 0).all()
    # the dump is the partition fixture format partition.read_vlist understands
    from normalmodes_b200 import partition
    vtxdist = np.concatenate([[0], np.cumsum(np.bincount(part, minlength=nproc))])
    assert (partition.read_vlist(names["fvlist"], len(part), vtxdist) == part).all()
    vs = np.fromfile(names["fvstat"], dtype="<i4")
    assert vs.size == len(part) and set(np.unique(vs)) <= {0, 1, 2}
