"""Host C++ numbering + CSR pattern (nm_fem_create) against the oracle: integer-exact,
for nproc = 1 and for a supplied partition.  Needs no GPU."""
import numpy as np
import pytest

from conftest import load_case


def _oracle_rows(c, name, part, nproc, rank):
    """Row slice [sizdist[r], sizdist[r+1]) of the oracle's global pattern for matrix `name`."""
    from oracle import fem
    if part is None:
        mats, num = c["mats"], c["num"]
    else:
        key = ("part", nproc, hash(part.tobytes()))
        if key not in c:
            topo = fem.build_topology(c["mesh"], c["model"]["vs"], c["g"]["porder"], nproc)
            num_ = fem.numbering(topo, part.astype(np.int64), nproc)
            c[key] = (fem.pattern(topo, num_), num_)
        mats, num = c[key]
    m = mats[name]
    dist = num["psizdist"] if name in ("ET", "Ap") else num["sizdist"]
    r0, r1 = int(dist[rank]), int(dist[rank + 1])
    ia = m["ia"][r0:r1 + 1] - m["ia"][r0]
    ja = m["ja"][m["ia"][r0]:m["ia"][r1]]
    return ia, ja, num


@pytest.mark.parametrize("name", ["const3k_p1_j1", "prem3k_p1_j2", "const3k_p2_j1", "prem3k_p2_j2", "rtmdwak8k_p1_j2"])
def test_pattern_bit_exact_single_rank(name):
    from normalmodes_b200.create_matrix import Fem
    c = load_case(name)
    f = Fem(c["mesh"], c["model"]["vs"], c["g"]["porder"])
    num = c["num"]
    assert (f.N, f.Np, f.nn) == (num["N"], num["Np"], c["topo"]["nn"])
    nb = f.numbering()
    assert (nb["vstat"] == c["topo"]["vstat"]).all()
    assert (nb["vstt"] == num["vstt"]).all() and (nb["pstt"] == num["pstt"]).all()
    assert (nb["vnum"] == num["vnum"]).all() and (nb["pnum"] == num["pnum"]).all()
    assert (f.t2n() == c["topo"]["t2n"]).all()
    for k in c["mats"]:
        m = f.matrix(k, values=False)
        ia, ja, _ = _oracle_rows(c, k, None, 1, 0)
        assert (m.rowdist == ia).all(), k
        assert (m.col == ja).all(), k
    f.free()


@pytest.mark.parametrize("name,nproc", [("const3k_p1_j1", 2), ("prem3k_p1_j2", 3), ("prem3k_p2_j2", 2)])
def test_pattern_bit_exact_with_partition(name, nproc):
    """Rank-local rows for a supplied part[] (what a ParMETIS run would hand over): each rank's block equals the
    corresponding row slice of the oracle's global pattern, and the per-rank offsets agree."""
    from normalmodes_b200.create_matrix import Fem
    c = load_case(name)
    nn = c["topo"]["nn"]
    rng = np.random.default_rng(nproc)
    # P2 edge-node ids depend on nproc (App. B item 5): the oracle topology for this nproc defines nn ordering
    part = rng.integers(0, nproc, nn).astype(np.int32)
    for rank in range(nproc):
        f = Fem(c["mesh"], c["model"]["vs"], c["g"]["porder"], nproc=nproc, part=part, rank=rank)
        for k in c["mats"]:
            m = f.matrix(k, values=False)
            ia, ja, num = _oracle_rows(c, k, part, nproc, rank)
            dist = num["psizdist"] if k in ("ET", "Ap") else num["sizdist"]
            cdist = num["psizdist"] if k in ("E", "Ap") else num["sizdist"]
            assert (m.sizdist == dist).all() and (m.coldist == cdist).all(), k
            assert (m.rowdist == ia).all() and (m.col == ja).all(), k
        f.free()


def test_mesh_generator_is_valid_input():
    """Builder-generated PREM-like mesh: positive orientation, consistent neighbours, fluid outer core."""
    from normalmodes_b200 import meshgen
    from normalmodes_b200.create_matrix import Fem
    m = meshgen.build_mesh(14000, seed=0)
    X = m["node"][m["ele"]]
    det = np.linalg.det(X[:, 1:4] - X[:, 0:1])
    assert (det > 0).all()
    vol = det.sum() / 6.0
    assert abs(vol / (4.0 / 3.0 * np.pi * 6371.0 ** 3) - 1.0) < 0.06      # polyhedral approximation of the ball
    assert (m["neigh"] < 0).sum() == 12 * m["nx"] ** 2                    # conforming: only the free surface is open
    for j in range(4):
        nb = m["neigh"][:, j]; msk = nb >= 0
        oth = np.delete(m["ele"], j, axis=1)[msk]
        for cc in range(3):
            assert (oth[:, cc][:, None] == m["ele"][nb[msk]]).any(axis=1).all()
    model = meshgen.build_model(m, 1)
    assert 0.05 < (model["layer"] == 1).mean() < 0.6
    assert (model["vs"][model["layer"] == 1] == 0).all() and (model["vs"][model["layer"] != 1] > 1).all()
    assert 9.0 < np.abs(model["g0"]).max() < 11.0
    f = Fem(m, model["vs"], 1)
    assert f.fluidcase == 1 and f.Np > 0 and f.N > 3 * m["nvert"]
    f.free()
