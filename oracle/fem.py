"""ORACLE (test infrastructure, NOT product code) -- mesh, numbering, CSR pattern
and element integration of NormalModes, restated on the CPU with numpy.

Parity status: UNPINNED (no reference golden vectors exist; the reference cannot
be built in this image -- SURVEY.md section 8c).  What is pinned instead: the four
shipped demo inputs run through this restatement, committed as fixtures under
tests/golden/ by tests/golden/make_golden.py.

Conventions: everything is 0-based here.  `order` is the rank-major ordering of
nodes (rank after rank, ascending original node id inside a rank), which is what
the reference's redistribution produces (src/mod_geometry.f90:953-965, 1129).
A rank-local matrix of the reference is the row slice [sizdist[r], sizdist[r+1])
of the global matrix built here, because every rank visits all elements touching
its owned nodes in ascending element id (Clelist, :1016-1026) and writes owned
rows only (src/mod_cg_create_matrix.f90:414,711,1226).

Follows (paths relative to /root/reference):
  file formats          src/mod_geometry.f90:83-135,179-220,1268-1349; src/mod_cg_models.f90:301-429
  vstat / v2v           src/mod_geometry.f90:284-421
  P2 edge nodes         src/mod_geometry.f90:428-689, 1566-1717
  Jacobians / normals   src/mod_geometry.f90:2216-2304, 1392-1428
  numbering + pattern   src/mod_cg_create_matrix.f90:1269-1455 (solid), 1458-2033 (fluid / fluid-solid)
  element values        src/mod_cg_create_matrix.f90:980-1266 (CGE3D_ISO), 103-977 (CGFSE3D_ISO)
  Jacobi scaling        src/mod_matvec.f90:252-441
"""
import os
import numpy as np
import scipy.sparse as sp

from .refelem import RefTet, TOL

EPS0 = 1.0e-15        # src/mod_cg_create_matrix.f90:30


# --------------------------------------------------------------------------- files
def read_mesh(inputdir, basename):
    pre = os.path.join(inputdir, basename)
    ntet, nvert = [int(x) for x in open(pre + "_mesh.header").read().split()[:2]]
    ele = np.fromfile(pre + "_ele.dat", dtype="<i4").reshape(ntet, 4).astype(np.int64) - 1
    neigh = np.fromfile(pre + "_neigh.dat", dtype="<i4").reshape(ntet, 4).astype(np.int64)
    neigh = np.where(neigh > 0, neigh - 1, -1)
    node = np.fromfile(pre + "_node.dat", dtype="<f8").reshape(nvert, 3)
    return dict(ntet=ntet, nvert=nvert, ele=ele, neigh=neigh, node=node)


def read_model(inputdir, basename, porder, job, ntet):
    pre = os.path.join(inputdir, basename)
    pNp = (porder + 1) * (porder + 2) * (porder + 3) // 6
    out = {}
    for k in ("vp", "vs", "rho"):
        out[k] = np.fromfile("%s_%s_pod_%d_true.dat" % (pre, k, porder), dtype="<f8").reshape(ntet, pNp)
    if job >= 2:
        out["g0"] = np.fromfile("%s_pod_%d_potential_acceleration_true.dat" % (pre, porder),
                                dtype="<f8").reshape(ntet, pNp, 3)
    else:
        out["g0"] = None
    return out


def write_mesh(outdir, basename, ele, neigh, node):
    """Write the App. A input formats (used by the synthetic mesh builder tests)."""
    os.makedirs(outdir, exist_ok=True)
    pre = os.path.join(outdir, basename)
    open(pre + "_mesh.header", "w").write("%d %d\n" % (ele.shape[0], node.shape[0]))
    (ele + 1).astype("<i4").tofile(pre + "_ele.dat")
    np.where(neigh >= 0, neigh + 1, -1).astype("<i4").tofile(pre + "_neigh.dat")
    node.astype("<f8").tofile(pre + "_node.dat")


# --------------------------------------------------------------------------- topology
def block_dist(n, nproc):
    """Initial block distribution, src/mod_geometry.f90:110-120."""
    tmp1 = n % nproc; tmp3 = (n - tmp1) // nproc
    d = np.zeros(nproc + 1, dtype=np.int64)
    for i in range(nproc):
        d[i + 1] = d[i] + tmp3 + (1 if i < tmp1 else 0)
    d[nproc] = n
    return d


def vertex_status(ele, vs, nvert):
    """vstat: 0 solid, 1 fluid, 2 fluid-solid (src/mod_geometry.f90:284-312)."""
    efl = (vs.max(axis=1) < 1.0e-6)
    nfl = np.zeros(nvert, dtype=np.int64); ntot = np.zeros(nvert, dtype=np.int64)
    np.add.at(nfl, ele.ravel(), np.repeat(efl.astype(np.int64), 4))
    np.add.at(ntot, ele.ravel(), 1)
    vstat = np.where(nfl == 0, 0, np.where(nfl == ntot, 1, 2))
    return vstat, efl


def _adjacency(cells, nn):
    """Sorted unique node->node adjacency (incl. self) of nodes sharing a cell."""
    k = cells.shape[1]
    ii = np.repeat(cells, k, axis=1).ravel()
    jj = np.tile(cells, (1, k)).ravel()
    A = sp.csr_matrix((np.ones(ii.size, dtype=np.int8), (ii, jj)), shape=(nn, nn))
    A.sum_duplicates(); A.sort_indices()
    return A.indptr.astype(np.int64), A.indices.astype(np.int64)


def p2_edges(ele, nvert, nproc):
    """Edge-node numbering for P2 (src/mod_geometry.f90:452-580, edge0%list :509-514).

    An edge is owned by the higher-ranked endpoint owner under the initial block
    distribution; inside a rank edges are met scanning the rank's vertices in
    ascending id and each vertex's neighbours in ascending id (when both ends are
    on the rank the edge is counted at its smaller endpoint)."""
    pa = np.array([[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]])
    e = np.sort(ele[:, pa].reshape(-1, 2), axis=1)
    e = np.unique(e, axis=0)                                     # (min,max), lexicographic
    vtxdist = block_dist(nvert, nproc)
    pid = np.searchsorted(vtxdist, np.arange(nvert), side="right") - 1
    pa_, pb_ = pid[e[:, 0]], pid[e[:, 1]]
    owner = np.maximum(pa_, pb_)
    # scanning vertex i (owned by `owner`), neighbour nb
    i = np.where(pa_ == pb_, e[:, 0], np.where(pa_ > pb_, e[:, 0], e[:, 1]))
    nb = np.where(i == e[:, 0], e[:, 1], e[:, 0])
    o = np.lexsort((nb, i, owner))
    return e[o]


def build_topology(mesh, vs, porder, nproc=1):
    """Global node set, status, per-element node table, node adjacency."""
    ele = mesh["ele"]; nvert = mesh["nvert"]; ntet = mesh["ntet"]
    vstat_v, efl = vertex_status(ele, vs, nvert)
    if porder == 1:
        t2n = ele.copy(); nn = nvert; vstat = vstat_v; edges = None
    else:
        edges = p2_edges(ele, nvert, nproc)
        ne = edges.shape[0]
        key = edges[:, 0] * nvert + edges[:, 1]
        ks = np.argsort(key); key_s = key[ks]
        pa = np.array([[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]])   # e12,e13,e23,e14,e24,e34 (:1575-1586)
        ee = np.sort(ele[:, pa], axis=2)
        eid = ks[np.searchsorted(key_s, ee[:, :, 0] * nvert + ee[:, :, 1])] + nvert
        t2n = np.zeros((ntet, 10), dtype=np.int64)
        ordv = np.array([0, 2, 5, 9, 1, 3, 4, 6, 7, 8])                   # ord (:1707)
        t2n[:, ordv[:4]] = ele
        t2n[:, ordv[4:]] = eid
        nn = nvert + ne
        est = np.minimum(vstat_v[edges[:, 0]], vstat_v[edges[:, 1]]) + 3   # :672-680
        vstat = np.concatenate([vstat_v, est])
    ip, ix = _adjacency(t2n, nn)
    mx = vstat_v.max()
    fsexist = bool(mx == 2); purefluid = bool(mx == 1)                    # :329-341
    return dict(porder=porder, nn=nn, nvert=nvert, ntet=ntet, t2n=t2n, vstat=vstat, efl=efl,
                v2v_ptr=ip, v2v=ix, fsexist=fsexist, purefluid=purefluid, edges=edges)


def element_geometry(mesh, ref):
    """invJ, detJ, outward normals, surface Jacobians, node coordinates.
    src/mod_geometry.f90:2216-2304, 1401-1428, 2243-2262."""
    X = mesh["node"][mesh["ele"]]                                         # (Ne,4,3)
    B = X[:, 1:4, :] - X[:, 0:1, :]                                       # rows v2-v1, v3-v1, v4-v1
    detB = (B[:, 0, 0] * (B[:, 1, 1] * B[:, 2, 2] - B[:, 2, 1] * B[:, 1, 2])
            + B[:, 0, 1] * (B[:, 2, 0] * B[:, 1, 2] - B[:, 1, 0] * B[:, 2, 2])
            + B[:, 0, 2] * (B[:, 1, 0] * B[:, 2, 1] - B[:, 2, 0] * B[:, 1, 1]))
    invJ = np.linalg.inv(B) * 2.0
    detJ = detB / 8.0
    n = np.zeros((X.shape[0], 4, 3)); sJ = np.zeros((X.shape[0], 4))
    for f in range(4):
        oth = [j for j in range(4) if j != f]
        v12 = X[:, oth[1]] - X[:, oth[0]]; v13 = X[:, oth[2]] - X[:, oth[0]]
        c = np.cross(v12, v13)
        a = np.sqrt((c ** 2).sum(axis=1))
        c = c / a[:, None]
        flip = ((X[:, f] - X[:, oth[0]]) * c).sum(axis=1) > 0
        c[flip] *= -1.0
        n[:, f] = c; sJ[:, f] = a / 4.0
    # node coordinates: X = a + (B/2)^T (ref+1)
    nods = X[:, 0:1, :] + np.einsum("eai,an->eni", B / 2.0, ref.nodes + 1.0)
    return dict(invJ=invJ, detJ=detJ, n=n, sJac=sJ, nods=nods)


# --------------------------------------------------------------------------- numbering
def numbering(topo, part=None, nproc=1):
    """DOF numbering (src/mod_cg_create_matrix.f90:1278-1303, 1467-1499, 1581-1613)."""
    nn = topo["nn"]; vstat = topo["vstat"]
    if part is None:
        part = np.zeros(nn, dtype=np.int64)
    order = np.lexsort((np.arange(nn), part))          # rank-major, ascending id inside rank
    pos = np.empty(nn, dtype=np.int64); pos[order] = np.arange(nn)
    fluidcase = topo["fsexist"] or topo["purefluid"]
    if fluidcase:
        vnum = np.where((vstat == 2) | (vstat == 5), 6, 3)
        pnum = np.where((vstat != 0) & (vstat != 3), 1, 0)
    else:
        vnum = np.full(nn, 3, dtype=np.int64); pnum = np.zeros(nn, dtype=np.int64)
    vstt = np.empty(nn, dtype=np.int64); pstt = np.full(nn, -1, dtype=np.int64)
    cs = np.concatenate([[0], np.cumsum(vnum[order])]); vstt[order] = cs[:-1]
    cp = np.concatenate([[0], np.cumsum(pnum[order])]); pstt[order] = np.where(pnum[order] == 1, cp[:-1], -1)
    nown = np.bincount(part, minlength=nproc)
    vtxdist = np.concatenate([[0], np.cumsum(nown)])
    sizdist = cs[vtxdist]; psizdist = cp[vtxdist]
    return dict(order=order, pos=pos, vnum=vnum, pnum=pnum, vstt=vstt, pstt=pstt, vtxdist=vtxdist,
                sizdist=sizdist.astype(np.int64), psizdist=psizdist.astype(np.int64),
                N=int(cs[-1]), Np=int(cp[-1]), part=part, fluidcase=fluidcase)


def _csr_from_pairs(rows, cols, nrow, ncol):
    o = np.lexsort((cols, rows))
    rows = rows[o]; cols = cols[o]
    if rows.size:
        dup = (rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])
        assert not dup.any(), "duplicate entries in pattern"
    ia = np.zeros(nrow + 1, dtype=np.int64)
    np.add.at(ia, rows + 1, 1)
    ia = np.cumsum(ia)
    return dict(ia=ia, ja=cols.astype(np.int64), a=np.zeros(cols.size), shape=(nrow, ncol))


def pattern(topo, num):
    """CSR patterns with explicit zeros.  Solid: matrixstruct (:1358-1435);
    fluid / fluid-solid: matrixstruct_general (:1682-2028)."""
    ip, ix = topo["v2v_ptr"], topo["v2v"]
    nn = topo["nn"]
    i = np.repeat(np.arange(nn), np.diff(ip)); j = ix
    vstt, vnum, pstt, pnum = num["vstt"], num["vnum"], num["pstt"], num["pnum"]
    N, Np = num["N"], num["Np"]
    three = np.arange(3)

    def blk(ri, cj, roff, coff, both=True):
        """rows vstt[ri]+roff+p ; cols vstt[cj]+coff+q (all p,q) or q==p only."""
        r0 = vstt[ri] + roff; c0 = vstt[cj] + coff
        if both:
            r = (r0[:, None, None] + three[None, :, None] + 0 * three[None, None, :]).ravel()
            c = (c0[:, None, None] + 0 * three[None, :, None] + three[None, None, :]).ravel()
        else:
            r = (r0[:, None] + three[None, :]).ravel(); c = (c0[:, None] + three[None, :]).ravel()
        return r, c

    if not num["fluidcase"]:
        z = np.zeros(i.size, dtype=np.int64)
        rA, cA = blk(i, j, z, z, True)
        rB, cB = blk(i, j, z, z, False)
        return dict(A=_csr_from_pairs(rA, cA, N, N), B=_csr_from_pairs(rB, cB, N, N))

    pi, pj, vi, vj = pnum[i], pnum[j], vnum[i], vnum[j]
    fo_j = vnum[j] - 3                                  # offset of the fluid-side triple of j
    R = {k: ([], []) for k in ("Ad", "B", "E", "ET", "Ap")}

    def add(name, r, c):
        R[name][0].append(r); R[name][1].append(c)

    # pure solid rows (:1987-2022): all neighbours, solid-side triple
    m = (pi == 0) & (vi == 3)
    z = np.zeros(m.sum(), dtype=np.int64)
    r, c = blk(i[m], j[m], z, z, True); add("Ad", r, c)
    r, c = blk(i[m], j[m], z, z, False); add("B", r, c)
    # pure fluid rows (:1820-1859): every neighbour carries pressure
    m = (pi == 1) & (vi == 3)
    assert (pj[m] == 1).all(), "Error: pure fluid"      # :1694-1697
    z = np.zeros(m.sum(), dtype=np.int64)
    r, c = blk(i[m], j[m], z, fo_j[m], True); add("Ad", r, c)
    r, c = blk(i[m], j[m], z, fo_j[m], False); add("B", r, c)
    add("E", (vstt[i[m]][:, None] + three[None, :]).ravel(), np.repeat(pstt[j[m]], 3))
    add("Ap", pstt[i[m]], pstt[j[m]])
    add("ET", np.repeat(pstt[i[m]], 3), (vstt[j[m]] + fo_j[m])[:, None].repeat(3, 1).ravel() + np.tile(three, m.sum()))
    # interface nodes (:1861-1985)
    mi = (vi == 6)
    ms = mi & ((pj == 0) | (vj == 6))                   # solid rows: solid + interface neighbours
    z = np.zeros(ms.sum(), dtype=np.int64)
    r, c = blk(i[ms], j[ms], z, z, True); add("Ad", r, c)
    r, c = blk(i[ms], j[ms], z, z, False); add("B", r, c)
    m6 = mi & (vj == 6)
    add("E", (vstt[i[m6]][:, None] + three[None, :]).ravel(), np.repeat(pstt[j[m6]], 3))
    mf = mi & (pj == 1)                                 # pressure row + fluid rows: pressure-carrying neighbours
    add("Ap", pstt[i[mf]], pstt[j[mf]])
    # ET row: every DOF (3 or 6) of pressure-carrying neighbours
    ii_, jj_ = i[mf], j[mf]
    cnt = vnum[jj_]
    rr = np.repeat(pstt[ii_], cnt)
    base = np.repeat(vstt[jj_], cnt)
    off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    add("ET", rr, base + off)
    t3 = np.full(mf.sum(), 3, dtype=np.int64)
    r, c = blk(ii_, jj_, t3, fo_j[mf], True); add("Ad", r, c)
    r, c = blk(ii_, jj_, t3, fo_j[mf], False); add("B", r, c)
    add("E", (vstt[ii_] + 3)[:, None].repeat(3, 1).ravel() + np.tile(three, mf.sum()), np.repeat(pstt[jj_], 3))
    shp = dict(Ad=(N, N), B=(N, N), E=(N, Np), ET=(Np, N), Ap=(Np, Np))
    out = {}
    for k, (rl, cl) in R.items():
        out[k] = _csr_from_pairs(np.concatenate(rl), np.concatenate(cl), *shp[k])
    return out


# --------------------------------------------------------------------------- values
def _T(x):
    return np.swapaxes(x, -1, -2)


def _locate(csr, rows, cols):
    """Position of (row,col) inside the CSR (the reference's findorder on the row)."""
    ncol = csr["shape"][1]
    ia = csr["ia"]
    key = np.repeat(np.arange(ia.size - 1), np.diff(ia)) * ncol + csr["ja"]
    k = rows * ncol + cols
    p = np.searchsorted(key, k)
    assert (key[p] == k).all(), "error: can not find the id"
    return p


def _scatter(csr, rows, cols, vals):
    """Sequential add in the given order == reference summation order (elements ascending)."""
    np.add.at(csr["a"], _locate(csr, rows.ravel(), cols.ravel()), vals.ravel())


def _solid_blocks(ref, invJ, detJ, lam, mu, rho, nods, g0, mass_mode):
    """TM (Ne,3,3,pNp,pNp) [block (i,j)] and Mrho*detJ (Ne,pNp,pNp).
    CGE3D_ISO :1043-1219 (mass_mode 'sym') / CGFSE3D_ISO solid branch :281-409 (mass_mode 'avg')."""
    M = ref.MassM; pNp = ref.pNp; Ne = invJ.shape[0]
    D = np.einsum("eia,amn->eimn", invJ, ref.Drst)                       # Derv(:,:,i)
    DT = _T(D)
    if mass_mode == "sym":
        Mrho = (rho[:, :, None] * M[None] + M[None] * rho[:, None, :]) / 2.0    # :1089-1091
    else:
        Mrho = M[None] * (rho.sum(axis=1) / float(pNp))[:, None, None]          # :293
    Ll = M[None] * lam[:, None, :]; Ll = (Ll + _T(Ll)) / 2.0                    # :1112-1114
    Lm = M[None] * mu[:, None, :];  Lm = (Lm + _T(Lm)) / 2.0
    OP1 = np.einsum("emk,eikn->eimn", Ll, D)
    OP2 = np.einsum("emk,eikn->eimn", Lm, D)
    OPt = np.einsum("eimk,eikn->eimn", DT, OP2)
    OPs = OPt.sum(axis=1)
    TM = np.zeros((Ne, 3, 3, pNp, pNp))
    OPrho = _gravity_solid(ref, D, rho, nods, g0) if g0 is not None else None
    for i in range(3):
        for j in range(3):
            if i == j:
                X = OPs + OPt[:, i] + DT[:, i] @ OP1[:, i]
                X = (X + _T(X)) / 2.0
            else:
                a = DT[:, i] @ OP1[:, j] + DT[:, j] @ OP2[:, i]
                b = DT[:, j] @ OP1[:, i] + DT[:, i] @ OP2[:, j]
                X = (_T(b) + a) / 2.0
            if OPrho is not None:
                X = X + OPrho[:, i, j]
            TM[:, i, j] = X * detJ[:, None, None]
    return TM, Mrho * detJ[:, None, None]


def _grad_ls(nods, f):
    """Least-squares gradient used for g and rho (:1063-1078, :197-218).  Returns (Ne,3,ncomp)."""
    pNp = nods.shape[1]
    dism = nods - nods.sum(axis=1, keepdims=True) / float(pNp)
    f0 = f - f.sum(axis=1, keepdims=True) / float(pNp)
    nd = np.einsum("eni,enj->eij", dism, dism)
    return np.linalg.solve(nd, np.einsum("eni,enc->eic", dism, f0)), dism


def _gravity_solid(ref, D, rho, nods, g0):
    """OPrho(i,j) (:1132-1177 / :331-377)."""
    M = ref.MassM
    gk1 = g0 / 1.0e3
    dgk1, _ = _grad_ls(nods, gk1)                                         # (Ne,3,3)
    Ne, pNp = rho.shape
    DT = _T(D)
    out = np.zeros((Ne, 3, 3, pNp, pNp))
    G = [gk1[:, :, c] for c in range(3)]

    def dl(g, X):      # diag(g) X
        return g[:, :, None] * X

    def dr(X, g):      # X diag(g)
        return X * g[:, None, :]
    Mb = np.broadcast_to(M, (Ne, pNp, pNp))
    for i in range(3):
        for j in range(3):
            a1 = DT[:, i] @ dr(Mb, G[j]); a1t = dl(G[j], Mb @ D[:, i])
            a1 = (a1 + _T(a1t)) / 2.0
            a2 = dl(G[i], Mb @ D[:, j]); a2t = DT[:, j] @ dr(Mb, G[i])
            a2 = (a2 + _T(a2t)) / 2.0
            O = (a1 + a2) / 2.0
            O = O - Mb * ((dgk1[:, i, j] + dgk1[:, j, i]) / 2.0)[:, None, None]
            b1 = DT[:, j] @ dl(G[i], Mb); b1t = Mb @ dl(G[i], D[:, j])
            b1 = (b1 + _T(b1t)) / 2.0
            b2 = Mb @ dl(G[j], D[:, i]); b2t = DT[:, i] @ dl(G[j], Mb)
            b2 = (b2 + _T(b2t)) / 2.0
            O = O - (b1 + b2) / 2.0
            O = (dr(O, rho) + dl(rho, O)) / 2.0
            out[:, i, j] = O
    return out


def assemble(mesh, model, porder, job, part=None, nproc=1, ref=None):
    """Full restatement of cg_create_matrix (src/mod_cg_create_matrix.f90:35-61).
    Returns (matrices, topo, num, geo).  Columns 0-based, values unscaled (CGM%*)."""
    ref = ref or RefTet(porder)
    topo = build_topology(mesh, model["vs"], porder, nproc)
    num = numbering(topo, part, nproc)
    geo = element_geometry(mesh, ref)
    mats = pattern(topo, num)
    pNp = ref.pNp
    rho = model["rho"]; mu = rho * model["vs"] ** 2; lam = rho * model["vp"] ** 2 - 2 * mu   # :1019-1020
    g0 = model["g0"] if job >= 2 else None
    t2n = topo["t2n"]; vstt = num["vstt"]; vnum = num["vnum"]; pstt = num["pstt"]
    three = np.arange(3)

    def scatter_u(name, el, TM, roff_n, coff_n):
        """rows vstt(m)+roff(m)+p ; cols vstt(n)+coff(n)+q ; TM (ne,3,3,pNp,pNp) indexed [p,q,m,n]."""
        nd = t2n[el]
        r = (vstt[nd] + roff_n)[:, None, None, :, None] + three[None, :, None, None, None]
        c = (vstt[nd] + coff_n)[:, None, None, None, :] + three[None, None, :, None, None]
        r, c = np.broadcast_arrays(r, c)
        # reference loop order inside an element is irrelevant: every (row,col) is hit once per element
        _scatter(mats[name], r.reshape(len(el), -1), c.reshape(len(el), -1), TM.reshape(len(el), -1))

    def scatter_b(name, el, MM, roff_n, coff_n):
        nd = t2n[el]
        r = (vstt[nd] + roff_n)[:, None, :, None] + three[None, :, None, None]
        c = (vstt[nd] + coff_n)[:, None, None, :] + three[None, :, None, None]
        r, c = np.broadcast_arrays(r, c)
        v = np.broadcast_to(MM[:, None], (len(el), 3, pNp, pNp))
        _scatter(mats[name], r.reshape(len(el), -1), c.reshape(len(el), -1), v.reshape(len(el), -1))

    if not num["fluidcase"]:
        el = np.arange(mesh["ntet"])
        for s in range(0, len(el), 20000):
            e = el[s:s + 20000]
            TM, MM = _solid_blocks(ref, geo["invJ"][e], geo["detJ"][e], lam[e], mu[e], rho[e], geo["nods"][e],
                                   None if g0 is None else g0[e], "sym")
            z = np.zeros((len(e), pNp), dtype=np.int64)
            scatter_u("A", e, TM, z, z); scatter_b("B", e, MM, z, z)
        return mats, topo, num, geo

    # ---- fluid / fluid-solid: CGFSE3D_ISO.  Elements must be visited in ascending id because
    # solid and fluid elements both add into Ad and B; process in id-ordered chunks.
    solid = mu.max(axis=1) >= TOL                                          # :279
    ne = mesh["ntet"]
    CH = 20000
    for s in range(0, ne, CH):
        e_all = np.arange(s, min(ne, s + CH))
        es = e_all[solid[e_all]]; ef = e_all[~solid[e_all]]
        contrib = []      # (matrix, rows(ne,k), cols(ne,k), vals(ne,k), element ids)
        if len(es):
            TM, MM = _solid_blocks(ref, geo["invJ"][es], geo["detJ"][es], lam[es], mu[es], rho[es], geo["nods"][es],
                                   None if g0 is None else g0[es], "avg")
            nd = t2n[es]
            r = vstt[nd][:, None, None, :, None] + three[None, :, None, None, None]
            c = vstt[nd][:, None, None, None, :] + three[None, None, :, None, None]
            r, c = np.broadcast_arrays(r, c)
            contrib.append(("Ad", r.reshape(len(es), -1), c.reshape(len(es), -1), TM.reshape(len(es), -1), es))
            r = vstt[nd][:, None, :, None] + three[None, :, None, None]
            c = vstt[nd][:, None, None, :] + three[None, :, None, None]
            r, c = np.broadcast_arrays(r, c)
            v = np.broadcast_to(MM[:, None], (len(es), 3, pNp, pNp))
            contrib.append(("B", r.reshape(len(es), -1), c.reshape(len(es), -1), v.reshape(len(es), -1), es))
        if len(ef):
            contrib += _fluid_contrib(ref, mesh, topo, num, geo, lam, rho, g0, ef)
        # merge by matrix, ordered by element id (stable)
        for name in ("Ad", "B", "E", "ET", "Ap"):
            items = [c for c in contrib if c[0] == name]
            if not items:
                continue
            rows = np.concatenate([np.asarray(c[1]).ravel() for c in items])
            cols = np.concatenate([np.asarray(c[2]).ravel() for c in items])
            vals = np.concatenate([np.asarray(c[3]).ravel() for c in items])
            eid = np.concatenate([np.repeat(c[4], np.asarray(c[1]).shape[1]) for c in items])
            o = np.argsort(eid, kind="stable")
            _scatter(mats[name], rows[o], cols[o], vals[o])
    return mats, topo, num, geo


def _fluid_contrib(ref, mesh, topo, num, geo, lam_all, rho_all, g0_all, ef):
    """Fluid-element blocks FT/FM and their scatter lists (:454-952)."""
    M = ref.MassM; pNp = ref.pNp; Nfp = ref.Nfp
    t2n = topo["t2n"]; vstt = num["vstt"]; vnum = num["vnum"]; pstt = num["pstt"]
    three = np.arange(3)
    ne = len(ef)
    invJ = geo["invJ"][ef]; detJ = geo["detJ"][ef]; nods = geo["nods"][ef]
    lam = lam_all[ef]; rho = rho_all[ef]
    nrm = geo["n"][ef]; sJ = geo["sJac"][ef]
    selfG = g0_all is not None
    gk1 = g0_all[ef] / 1.0e3 if selfG else np.zeros((ne, pNp, 3))         # JOB 1: gravity terms defined as zero
    D = np.einsum("eia,amn->eimn", invJ, ref.Drst)
    Mb = np.broadcast_to(M, (ne, pNp, pNp))
    lamiv = 1.0 / np.sqrt(lam)
    flam = lamiv[:, :, None] * Mb * lamiv[:, None, :]                     # :470-472
    Mrho = Mb * (rho.sum(axis=1) / float(pNp))[:, None, None]             # :480
    FTuu = np.zeros((ne, 3, 3, pNp, pNp)); FTup = np.zeros((ne, 3, pNp, pNp))
    FMuu = Mrho * detJ[:, None, None]
    for i in range(3):
        fOP = (Mb @ D[:, i] + _T(_T(D[:, i]) @ Mb)) / 2.0                 # :521-525 / :618-622
        FTup[:, i] = fOP
    if selfG:
        rhoavg = rho.sum(axis=1) / float(pNp)
        drho0, _ = _grad_ls(nods, (rho - rhoavg[:, None])[:, :, None])    # :216-218
        drho0 = drho0[:, :, 0]
        normg = np.sqrt((gk1 ** 2).sum(axis=2))
        N2 = (drho0[:, None, :] * gk1).sum(axis=2)
        normalg = gk1 / np.maximum(normg, EPS0)[:, :, None]
        N2 = N2 / rho - normg ** 2 / lam * rho                            # :246
        N2[normg < EPS0] = 0.0
        if topo["purefluid"]:
            N2[:] = 0.0
        for i in range(3):
            for j in range(3):
                Rj = normalg[:, :, j] * N2 * rho
                Ni = normalg[:, :, i]
                if i == j:
                    O = (Ni[:, :, None] * (Mb * Rj[:, None, :]) + Rj[:, :, None] * (Mb * Ni[:, None, :])) / 2.0   # :497-511
                    rinl = rho / lam
                    dg = rinl * gk1[:, :, i]
                    Od = (Mb * dg[:, None, :] + _T(dg[:, :, None] * Mb)) / 2.0                                   # :550-552
                    FTup[:, i] = FTup[:, i] - Od
                else:
                    O = (Ni[:, :, None] * (Mb * Rj[:, None, :]) + _T(Rj[:, :, None] * (Mb * Ni[:, None, :]))) / 2.0
                    Ri = normalg[:, :, i] * N2 * rho
                    Nj = normalg[:, :, j]
                    Ot = (Ri[:, :, None] * (Mb * Nj[:, None, :]) + _T(Nj[:, :, None] * (Mb * Ri[:, None, :]))) / 2.0
                    O = (O + Ot) / 2.0
                FTuu[:, i, j] = O
    FTuu = FTuu * detJ[:, None, None, None, None]
    FTup = FTup * detJ[:, None, None, None]
    FTpp = -flam * detJ[:, None, None]                                    # :638-639
    # boundary / interior-face block, always executed (:643-706)
    neigh = mesh["neigh"][ef]
    nd = t2n[ef]
    v6 = (vnum[nd[:, ref.vord]] == 6)                                     # (ne,4)
    for f in range(4):
        Fm = ref.Fmask[:, f]
        surfrho = rho[:, Fm].sum(axis=1) / float(Nfp)
        gn = (gk1[:, Fm, :] * nrm[:, f, None, :]).sum(axis=2)            # (ne,Nfp)
        bnd = neigh[:, f] == -1
        if bnd.any():
            sgn = np.sqrt((-gn[bnd]) ** 2).sum(axis=1) / float(Nfp)
            if (sgn == 0).any():
                raise ValueError("fluid free-surface face without gravity: undefined in the reference (:656-657)")
            surfp = ref.MassF[f][None] / sgn[:, None, None] / surfrho[bnd][:, None, None]
            idx = np.where(bnd)[0]
            FTpp[np.ix_(idx, Fm, Fm)] += -surfp * sJ[bnd, f][:, None, None]
        ins = ~bnd
        sgn = gn.sum(axis=1) / float(Nfp)
        cout = v6.sum(axis=1) - v6[:, f]
        m = ins & (cout < 3)
        if m.any():
            idx = np.where(m)[0]
            surfp = ref.MassF[f][None] * (sgn[m] * surfrho[m])[:, None, None]
            for j in range(3):
                FTuu[np.ix_(idx, [j], [j], Fm, Fm)] += (surfp * (sJ[m, f] * nrm[m, f, j] ** 2)[:, None, None])[:, None, None]
    out = []
    fo = vnum[nd] - 3                                                     # fluid-side triple offset
    pr = pstt[nd]
    # Ap (:715-728)
    r = np.broadcast_to(pr[:, :, None], (ne, pNp, pNp)); c = np.broadcast_to(pr[:, None, :], (ne, pNp, pNp))
    out.append(("Ap", r.reshape(ne, -1), c.reshape(ne, -1), FTpp.reshape(ne, -1), ef))
    # ET (:730-746): FT(p_m, u_q n) = FTup[q]^T[m,n]
    r = np.broadcast_to(pr[:, None, :, None], (ne, 3, pNp, pNp))
    c = np.broadcast_to((vstt[nd] + fo)[:, None, None, :] + three[None, :, None, None], (ne, 3, pNp, pNp))
    out.append(("ET", r.reshape(ne, -1), c.reshape(ne, -1), _T(FTup).reshape(ne, -1), ef))
    # Ad fluid rows (:749-765)
    r = (vstt[nd] + fo)[:, None, None, :, None] + three[None, :, None, None, None]
    c = (vstt[nd] + fo)[:, None, None, None, :] + three[None, None, :, None, None]
    r, c = np.broadcast_arrays(r, c)
    out.append(("Ad", r.reshape(ne, -1), c.reshape(ne, -1), FTuu.reshape(ne, -1), ef))
    # B fluid rows (:767-779)
    r = (vstt[nd] + fo)[:, None, :, None] + three[None, :, None, None]
    c = (vstt[nd] + fo)[:, None, None, :] + three[None, :, None, None]
    r, c = np.broadcast_arrays(r, c)
    v = np.broadcast_to(FMuu[:, None], (ne, 3, pNp, pNp))
    out.append(("B", r.reshape(ne, -1), c.reshape(ne, -1), v.reshape(ne, -1), ef))
    # E fluid rows (:781-794)
    r = np.broadcast_to((vstt[nd] + fo)[:, None, :, None] + three[None, :, None, None], (ne, 3, pNp, pNp))
    c = np.broadcast_to(pr[:, None, None, :], (ne, 3, pNp, pNp))
    out.append(("E", r.reshape(ne, -1), c.reshape(ne, -1), FTup.reshape(ne, -1), ef))
    # fluid-solid interface faces (:804-952)
    cnt = v6.sum(axis=1)
    for le in np.where(cnt == 3)[0]:
        fc = int(np.where(~v6[le])[0][-1])
        Fm = ref.Fmask[:, fc]
        nf = nd[le, Fm]
        nq = nrm[le, fc]; sj = sJ[le, fc]; MF = ref.MassF[fc]
        eid = np.array([ef[le]])
        if selfG:
            rhof = rho[le, Fm].sum() / float(Nfp)
            Gd = [np.diag(gk1[le, Fm, c_]) for c_ in range(3)]
            SCM = np.zeros((3, 3, Nfp, Nfp))
            for i in range(3):
                for j in range(3):
                    si = sj * nq[i] * (MF @ Gd[j]); sit = sj * nq[i] * (Gd[j] @ MF); si = (si + sit.T) / 2.0
                    s2 = sj * nq[j] * (MF @ Gd[i]); s2t = sj * nq[j] * (Gd[i] @ MF); s2 = (s2 + s2t.T) / 2.0
                    SCM[i, j] = (si + s2.T) / 2.0 * rhof
            r = vstt[nf][None, None, :, None] + three[:, None, None, None]
            c = vstt[nf][None, None, None, :] + three[None, :, None, None]
            r, c = np.broadcast_arrays(r, c)
            out.append(("Ad", r.reshape(1, -1), c.reshape(1, -1), (-SCM).reshape(1, -1), eid))
        # ET(p_m, u_q n) -= n_q sJac MassF(n,m) (:895-924)
        r = np.broadcast_to(pstt[nf][:, None, None], (Nfp, Nfp, 3))
        c = vstt[nf][None, :, None] + three[None, None, :]
        r, c = np.broadcast_arrays(r, c)
        v = -(MF.T[:, :, None] * nq[None, None, :]) * sj
        out.append(("ET", r.reshape(1, -1), c.reshape(1, -1), v.reshape(1, -1), eid))
        # E(u_p m [solid side], p_n) -= n_p sJac MassF(m,n) (:926-948)
        r = vstt[nf][:, None, None] + three[None, :, None]
        c = pstt[nf][None, None, :]
        r, c = np.broadcast_arrays(r, c)
        v = -(MF[:, None, :] * nq[None, :, None]) * sj
        out.append(("E", r.reshape(1, -1), c.reshape(1, -1), v.reshape(1, -1), eid))
    return out


# --------------------------------------------------------------------------- scaling / operators
def csr_diag(m):
    ia, ja = m["ia"], m["ja"]
    rows = np.repeat(np.arange(ia.size - 1), np.diff(ia))
    d = np.zeros(ia.size - 1)
    sel = rows == ja
    d[rows[sel]] = m["a"][sel]
    return d


def jacobi_scale(m, sign=1.0):
    """Bdiagscaling / Apdiagscaling: d=1/sqrt(sign*diag); val=(sign*a*d_j)*d_i (src/mod_matvec.f90:275,336,361,422)."""
    d = 1.0 / np.sqrt(sign * csr_diag(m))
    rows = np.repeat(np.arange(m["ia"].size - 1), np.diff(m["ia"]))
    out = dict(m); out["a"] = (sign * m["a"] * d[m["ja"]]) * d[rows]
    return out, d


def to_scipy(m):
    return sp.csr_matrix((m["a"], m["ja"], m["ia"]), shape=m["shape"])
