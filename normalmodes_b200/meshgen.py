"""Builder-generated PREM-like planet models for the benchmark configs (BASELINE.json configs 2-5).

PlanetaryModels / TetGen are not available offline, so the mesh is built here: a conforming cubed-sphere
grid (central cube + 6 equiangular radial chunks) whose grid surfaces include the free surface and the two
PREM discontinuities kept by the reference's `prem_3L` models (ICB 1221.5 km, CMB 3480 km), every
hexahedron cut into 12 tetrahedra.  Elements are classified by centroid radius into inner core (solid) /
outer core (fluid, vs = 0) / mantle+crust (solid); the fluid-solid interfaces are triangulated spheres, so
every mesh edge joining two interface vertices lies ON the interface -- which the reference's P2 edge-node
status rule (src/mod_geometry.f90:672-680) silently requires.
Material values are PREM's polynomials (Dziewonski & Anderson 1981, isotropic, ocean replaced by upper
crust) evaluated at the element nodes inside the element's own layer; the reference gravity
g0 = -g(r) r^ (m/s^2, as the reference's *_potential_acceleration_true.dat files) comes from the radial
integration of that density.  Output arrays follow the reference's on-disk layout
and `write_files` stores them under the reference's file names, so the same inputs can be fed to the
real reference elsewhere.
"""
import os

import numpy as np

R_EARTH = 6371.0
R_ICB = 1221.5
R_CMB = 3480.0
GRAV = 6.6743e-11

# (r_lo, r_hi, rho[4], vp[4], vs[4]) polynomial coefficients in x = r / 6371
_PREM = [
    (0.0, 1221.5, (13.0885, 0.0, -8.8381, 0.0), (11.2622, 0.0, -6.3640, 0.0), (3.6678, 0.0, -4.4475, 0.0)),
    (1221.5, 3480.0, (12.5815, -1.2638, -3.6426, -5.5281), (11.0487, -4.0362, 4.8023, -13.5732), (0.0, 0.0, 0.0, 0.0)),
    (3480.0, 3630.0, (7.9565, -6.4761, 5.5283, -3.0807), (15.3891, -5.3181, 5.5242, -2.5514), (6.9254, 1.4672, -2.0834, 0.9783)),
    (3630.0, 5600.0, (7.9565, -6.4761, 5.5283, -3.0807), (24.9520, -40.4673, 51.4832, -26.6419), (11.1671, -13.7818, 17.4575, -9.2777)),
    (5600.0, 5701.0, (7.9565, -6.4761, 5.5283, -3.0807), (29.2766, -23.6027, 5.5242, -2.5514), (22.3459, -17.2473, -2.0834, 0.9783)),
    (5701.0, 5771.0, (5.3197, -1.4836, 0.0, 0.0), (19.0957, -9.8672, 0.0, 0.0), (9.9839, -4.9324, 0.0, 0.0)),
    (5771.0, 5971.0, (11.2494, -8.0298, 0.0, 0.0), (39.7027, -32.6166, 0.0, 0.0), (22.3512, -18.5856, 0.0, 0.0)),
    (5971.0, 6151.0, (7.1089, -3.8045, 0.0, 0.0), (20.3926, -12.2569, 0.0, 0.0), (8.9496, -4.4597, 0.0, 0.0)),
    (6151.0, 6346.6, (2.6910, 0.6924, 0.0, 0.0), (4.1875, 3.9382, 0.0, 0.0), (2.1519, 2.3481, 0.0, 0.0)),
    (6346.6, 6356.0, (2.9, 0.0, 0.0, 0.0), (6.8, 0.0, 0.0, 0.0), (3.9, 0.0, 0.0, 0.0)),
    (6356.0, 6371.0, (2.6, 0.0, 0.0, 0.0), (5.8, 0.0, 0.0, 0.0), (3.2, 0.0, 0.0, 0.0)),
]
_LAYERS = [(0.0, R_ICB), (R_ICB, R_CMB), (R_CMB, R_EARTH)]     # inner core, outer core (fluid), mantle + crust


def prem(r, layer):
    """(rho, vp, vs) of the 3-layer PREM at radius r (km), evaluated INSIDE `layer` (0, 1, 2)."""
    r = np.asarray(r, dtype=float)
    L = np.asarray(_LAYERS)[np.asarray(layer)]
    lo, hi = L[..., 0], L[..., 1]
    rc = np.minimum(np.maximum(r, lo + 1e-6), hi - 1e-6)
    x = rc / R_EARTH
    rho = np.zeros_like(x); vp = np.zeros_like(x); vs = np.zeros_like(x)
    for (a, b, cr, cp, cs) in _PREM:
        m = (rc >= a) & (rc < b)
        for out, c in ((rho, cr), (vp, cp), (vs, cs)):
            out[m] = c[0] + x[m] * (c[1] + x[m] * (c[2] + x[m] * c[3]))
    return rho, vp, vs


def gravity_profile(nr=20001):
    """g(r) in m/s^2 on a radial grid from the 3-layer PREM density (g/cm^3 -> kg/m^3, km -> m)."""
    r = np.linspace(0.0, R_EARTH, nr)
    layer = np.where(r < R_ICB, 0, np.where(r < R_CMB, 1, 2))
    rho = prem(r, layer)[0] * 1000.0
    rm = r * 1000.0
    f = rho * rm ** 2
    mass = 4.0 * np.pi * np.concatenate([[0.0], np.cumsum(0.5 * (f[1:] + f[:-1]) * np.diff(rm))])
    g = np.zeros_like(r)
    g[1:] = GRAV * mass[1:] / rm[1:] ** 2
    return r, g


def _face_neighbours(ele):
    """neigh[k, j] = element sharing the face opposite local vertex j (-1: boundary)."""
    ne = ele.shape[0]
    faces = np.stack([np.delete(ele, j, axis=1) for j in range(4)], axis=1).reshape(-1, 3)   # row 4k+j
    fs = np.sort(faces, axis=1)
    o = np.lexsort((fs[:, 2], fs[:, 1], fs[:, 0]))
    f = fs[o]
    same = (f[1:] == f[:-1]).all(axis=1)
    neigh = np.full(4 * ne, -1, dtype=np.int64)
    a = o[:-1][same]; b = o[1:][same]
    neigh[a] = b // 4; neigh[b] = a // 4
    return neigh.reshape(ne, 4)


def _radial_levels(nx):
    """Radii of the spherical grid surfaces from the ICB outwards (ICB, CMB and the free surface are grid
    surfaces), spaced like the lateral cell size at the middle of each shell."""
    levels = [R_ICB]
    for lo, hi in ((R_ICB, R_CMB), (R_CMB, R_EARTH)):
        lateral = 0.5 * np.pi * 0.5 * (lo + hi) / nx
        n = max(1, int(round((hi - lo) / lateral)))
        levels += list(lo + (hi - lo) * np.arange(1, n + 1) / n)
    return np.array(levels)


def _count_tets(nx):
    c = 0.45 * R_ICB
    ncore = max(1, int(round((R_ICB - c) / (0.5 * np.pi * 0.75 * R_ICB / nx))))
    return 12 * (nx ** 3 + 6 * nx * nx * (ncore + len(_radial_levels(nx)) - 1)), ncore


def build_mesh(ntet_target, seed=0):
    """Conforming tetrahedral mesh of the ball with ~ntet_target elements: a cubed sphere (central cube +
    6 radial chunks, equiangular) whose grid surfaces include the ICB, the CMB and the free surface, every
    hexahedron cut into 6 pyramids around its centre and every pyramid into 2 tets along the base diagonal
    through the lowest-numbered vertex (so neighbouring hexahedra agree).  Returns dict(ele, neigh, node),
    0-based, positively oriented, neigh[k, j] = element across the face opposite local vertex j (-1 boundary)."""
    nx = 2
    while _count_tets(nx + 1)[0] <= ntet_target * 1.05:
        nx += 1
    _, ncore = _count_tets(nx)
    c = 0.45 * R_ICB
    ang = np.tan(np.linspace(-1.0, 1.0, nx + 1) * np.pi / 4.0)
    pts = []; hexes = []; off = 0

    def add_block(P):                      # P: (n0+1, n1+1, n2+1, 3) structured points -> hexes
        nonlocal off
        n0, n1, n2 = P.shape[0] - 1, P.shape[1] - 1, P.shape[2] - 1
        idx = (np.arange(P.shape[0] * P.shape[1] * P.shape[2]) + off).reshape(P.shape[:3])
        off += idx.size
        pts.append(P.reshape(-1, 3))
        h = np.stack([idx[:-1, :-1, :-1], idx[1:, :-1, :-1], idx[1:, 1:, :-1], idx[:-1, 1:, :-1],
                      idx[:-1, :-1, 1:], idx[1:, :-1, 1:], idx[1:, 1:, 1:], idx[:-1, 1:, 1:]], axis=-1)
        hexes.append(h.reshape(-1, 8))
    # central cube
    X, Y, Z = np.meshgrid(c * ang, c * ang, c * ang, indexing="ij")
    add_block(np.stack([X, Y, Z], -1))
    # six chunks: (u, v) on the cube face, w radial level
    U, V = np.meshgrid(ang, ang, indexing="ij")
    one = np.ones_like(U)
    faces = [np.stack([U, V, one], -1), np.stack([U, V, -one], -1), np.stack([U, one, V], -1),
             np.stack([U, -one, V], -1), np.stack([one, U, V], -1), np.stack([-one, U, V], -1)]
    shells = _radial_levels(nx)
    for Fc in faces:
        d = Fc / np.linalg.norm(Fc, axis=-1, keepdims=True)
        lev = [(1.0 - t) * c * Fc + t * R_ICB * d for t in np.arange(0, ncore) / ncore]
        lev += [r * d for r in shells]
        add_block(np.stack(lev, axis=2))
    P = np.concatenate(pts); H = np.concatenate(hexes)
    # merge coincident vertices (chunk edges / cube faces)
    key = np.round(P / 1.0e-5).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    node = P[first]; H = inv.reshape(-1)[H]
    nv = node.shape[0]
    # spatially coherent vertex numbering (coarse grid, z-major), as mesh generators usually emit
    a = (4.0 / 3.0 * np.pi * R_EARTH ** 3 / max(nv, 1)) ** (1.0 / 3.0)
    g = np.floor((node + R_EARTH) / (4.0 * a)).astype(np.int64)
    order = np.lexsort((node[:, 0], g[:, 0], g[:, 1], g[:, 2]))
    rank = np.empty(nv, dtype=np.int64); rank[order] = np.arange(nv)
    node = node[order]; H = rank[H]
    # hexahedron -> 6 pyramids around the centre -> 12 tets
    cen = node[H].mean(axis=1)
    cid = nv + np.arange(H.shape[0])
    node = np.concatenate([node, cen])
    quads = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [3, 2, 6, 7], [0, 3, 7, 4], [1, 2, 6, 5]])
    tets = []
    for q in quads:
        Q = H[:, q]                                                   # cyclic corner order
        k = np.argmin(Q, axis=1)
        even = (k % 2 == 0)[:, None]
        t1 = np.where(even, np.stack([Q[:, 0], Q[:, 1], Q[:, 2], cid], 1), np.stack([Q[:, 1], Q[:, 2], Q[:, 3], cid], 1))
        t2 = np.where(even, np.stack([Q[:, 0], Q[:, 2], Q[:, 3], cid], 1), np.stack([Q[:, 1], Q[:, 3], Q[:, 0], cid], 1))
        tets += [t1, t2]
    ele = np.concatenate(tets).astype(np.int64)
    # interleave so that the 12 tets of a hexahedron are consecutive (element locality)
    nh = H.shape[0]
    ele = ele.reshape(12, nh, 4).transpose(1, 0, 2).reshape(-1, 4)
    Xe = node[ele]
    det = np.linalg.det(Xe[:, 1:4, :] - Xe[:, 0:1, :])
    flip = det < 0
    ele[flip, 0], ele[flip, 1] = ele[flip, 1].copy(), ele[flip, 0].copy()
    # elements ordered like their hexahedra's centres (z-major coarse grid)
    hc = np.floor((cen + R_EARTH) / (4.0 * a)).astype(np.int64)
    ho = np.lexsort((cen[:, 0], hc[:, 0], hc[:, 1], hc[:, 2]))
    ele = ele.reshape(nh, 12, 4)[ho].reshape(-1, 4)
    neigh = _face_neighbours(ele)
    return dict(ntet=int(ele.shape[0]), nvert=int(node.shape[0]), ele=ele, neigh=neigh, node=node, nx=nx)


_P2_PAIRS = np.array([[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]])       # e12,e13,e23,e14,e24,e34
_P2_ORD = np.array([0, 2, 5, 9, 1, 3, 4, 6, 7, 8])                            # src/mod_geometry.f90:1707


def element_nodes(mesh, porder):
    """Coordinates of the pNp nodes of every element in the reference's local order (App. C)."""
    X = mesh["node"][mesh["ele"]]
    if porder == 1:
        return X
    out = np.empty((X.shape[0], 10, 3))
    out[:, _P2_ORD[:4]] = X
    out[:, _P2_ORD[4:]] = 0.5 * (X[:, _P2_PAIRS[:, 0]] + X[:, _P2_PAIRS[:, 1]])
    return out


def build_model(mesh, porder, gravity=True):
    """vp, vs, rho [Ntet][pNp] and g0 [Ntet][pNp][3] (m/s^2) of the PREM-like model."""
    cent = mesh["node"][mesh["ele"]].mean(axis=1)
    rc = np.linalg.norm(cent, axis=1)
    layer = np.where(rc < R_ICB, 0, np.where(rc < R_CMB, 1, 2))
    Xn = element_nodes(mesh, porder)
    rn = np.linalg.norm(Xn, axis=2)
    rho, vp, vs = prem(rn, layer[:, None] * np.ones_like(rn, dtype=np.int64))
    vs[layer == 1] = 0.0
    g0 = None
    if gravity:
        rr, gg = gravity_profile()
        gmag = np.interp(rn, rr, gg)
        g0 = -gmag[:, :, None] * Xn / np.maximum(rn, 1e-9)[:, :, None]
    return dict(vp=vp, vs=vs, rho=rho, g0=g0, layer=layer)


def write_files(outdir, basename, mesh, model, porder):
    """Write the reference's input files for this mesh/model."""
    os.makedirs(outdir, exist_ok=True)
    pre = os.path.join(outdir, basename)
    open(pre + "_mesh.header", "w").write("%d %d\n" % (mesh["ntet"], mesh["nvert"]))
    (mesh["ele"] + 1).astype("<i4").tofile(pre + "_ele.dat")
    np.where(mesh["neigh"] >= 0, mesh["neigh"] + 1, -1).astype("<i4").tofile(pre + "_neigh.dat")
    mesh["node"].astype("<f8").tofile(pre + "_node.dat")
    for k in ("vp", "vs", "rho"):
        model[k].astype("<f8").tofile("%s_%s_pod_%d_true.dat" % (pre, k, porder))
    if model["g0"] is not None:
        model["g0"].astype("<f8").tofile("%s_pod_%d_potential_acceleration_true.dat" % (pre, porder))
