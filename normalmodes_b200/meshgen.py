"""Builder-generated PREM-like planet models for the benchmark configs (BASELINE.json configs 2-5).

PlanetaryModels / TetGen are not available offline, so the mesh is built here: a jittered BCC point
lattice inside the ball plus Fibonacci point shells on the free surface and on the two PREM
discontinuities kept by the reference's `prem_3L` models (ICB 1221.5 km, CMB 3480 km), tetrahedralised
with Qhull (scipy.spatial.Delaunay).  Elements are classified by centroid radius into inner core
(solid) / outer core (fluid, vs = 0) / mantle+crust (solid), so fluid-solid interfaces are mesh faces.
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


def _fibonacci_sphere(n, radius):
    i = np.arange(n) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    th = np.pi * (1.0 + 5.0 ** 0.5) * i
    return radius * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], axis=1)


def build_mesh(ntet_target, seed=0):
    """Tetrahedral mesh of the ball with ~ntet_target elements.  Returns dict(ele, neigh, node), 0-based,
    positively oriented, neigh[k, j] = element across the face opposite local vertex j (-1 boundary)."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    npts = max(ntet_target / 6.2, 60.0)
    a = (2.0 * (4.0 / 3.0) * np.pi * R_EARTH ** 3 / npts) ** (1.0 / 3.0)      # BCC cell: 2 points per a^3
    m = int(np.ceil(R_EARTH / a)) + 1
    g = np.arange(-m, m + 1) * a
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    P = np.concatenate([np.stack([X, Y, Z], -1).reshape(-1, 3), np.stack([X, Y, Z], -1).reshape(-1, 3) + a / 2.0])
    P = P + rng.uniform(-0.02 * a, 0.02 * a, P.shape)
    r = np.linalg.norm(P, axis=1)
    s = 0.87 * a                                                               # nearest-neighbour distance of BCC
    keep = r < R_EARTH - 0.45 * s
    for rs in (R_ICB, R_CMB):
        keep &= np.abs(r - rs) > 0.4 * s
    shells = [P[keep]]
    for rs in (R_ICB, R_CMB, R_EARTH):
        n = max(int(4.0 * np.pi * rs ** 2 / (0.866 * s * s)), 12)
        shells.append(_fibonacci_sphere(n, rs))
    node = np.concatenate(shells)
    node = node[rng.permutation(node.shape[0])]                               # unstructured numbering, like TetGen's
    tri = Delaunay(node)
    ele = tri.simplices.astype(np.int64)
    neigh = tri.neighbors.astype(np.int64)
    # positive orientation (reference meshes: all detJ > 0)
    Xe = node[ele]
    B = Xe[:, 1:4, :] - Xe[:, 0:1, :]
    det = np.linalg.det(B)
    flip = det < 0
    ele[flip, 0], ele[flip, 1] = ele[flip, 1].copy(), ele[flip, 0].copy()
    neigh[flip, 0], neigh[flip, 1] = neigh[flip, 1].copy(), neigh[flip, 0].copy()
    # drop zero-volume hull slivers (4 nearly coplanar surface points); only boundary elements qualify
    vol = np.abs(det) / 6.0
    bad = (vol < 1e-6 * a ** 3) & (neigh < 0).any(axis=1)
    if bad.any():
        newid = np.cumsum(~bad) - 1
        ele = ele[~bad]
        neigh = neigh[~bad]
        neigh = np.where(neigh >= 0, np.where(bad[np.maximum(neigh, 0)], -1, newid[np.maximum(neigh, 0)]), -1)
    used = np.zeros(node.shape[0], dtype=bool); used[ele.ravel()] = True
    if not used.all():
        newv = np.cumsum(used) - 1
        node = node[used]; ele = newv[ele]
    # cache-friendly numbering: order vertices along a coarse spatial grid (Morton-like), as mesh
    # generators' output usually is; the DOF numbering contract is "whatever ids the input files carry"
    key = np.floor((node + R_EARTH) / (4.0 * a)).astype(np.int64)
    order = np.lexsort((key[:, 0], key[:, 1], key[:, 2]))
    inv = np.empty_like(order); inv[order] = np.arange(order.size)
    node = node[order]; ele = inv[ele]
    cent = node[ele].mean(axis=1)
    eo = np.lexsort((cent[:, 0], cent[:, 1], cent[:, 2]))
    einv = np.empty_like(eo); einv[eo] = np.arange(eo.size)
    ele = ele[eo]; neigh = neigh[eo]
    neigh = np.where(neigh >= 0, einv[np.maximum(neigh, 0)], -1)
    return dict(ntet=int(ele.shape[0]), nvert=int(node.shape[0]), ele=ele, neigh=neigh, node=node, spacing=a)


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
