"""ORACLE (test infrastructure, NOT product code) -- reference tetrahedron.

CPU restatement (numpy) of the nodal-basis machinery NormalModes uses to build
its reference element.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this package.

Parity status: UNPINNED -- the reference ships no golden vectors and cannot be
compiled here (no Fortran/MPI/pEVSL/ParMETIS in the image).

Follows (all paths relative to /root/reference):
  src/mod_utility.f90:160-242   matdet / matinv (Gauss-Jordan)
  src/mod_utility.f90:299-486   JacobiP, Basis3D, Vandermonde3D, GradBasis3D,
                                GradVandermonde3D, Basis2D, Vandermonde2D
  src/mod_utility.f90:494-635   blend_nodes (equispaced for pOrder<=2: warp==0)
  src/mod_geometry.f90:2307-2523 Build_reference
"""
import math
import numpy as np

TOL = 1.0e-8          # pin%TOL for rkind=8, src/mod_para.f90:181-185


def matinv(a):
    """Gauss-Jordan inverse with partial pivoting, src/mod_utility.f90:174-242."""
    a = np.array(a, dtype=np.float64, copy=True)
    n = a.shape[0]
    b = np.eye(n)
    for i in range(n):
        big = abs(a[i, i]); irow = i
        for j in range(i, n):
            if abs(a[j, i]) > big:
                big = abs(a[j, i]); irow = j
        if big > abs(a[i, i]):
            a[[i, irow], :] = a[[irow, i], :]
            b[[i, irow], :] = b[[irow, i], :]
        dum = a[i, i]
        a[i, :] = a[i, :] / dum
        b[i, :] = b[i, :] / dum
        for j in range(i + 1, n):
            dum = a[j, i]
            a[j, :] = a[j, :] - dum * a[i, :]
            b[j, :] = b[j, :] - dum * b[i, :]
    for i in range(n - 1):
        for j in range(i + 1, n):
            dum = a[i, j]
            a[i, :] = a[i, :] - dum * a[j, :]
            b[i, :] = b[i, :] - dum * b[j, :]
    return b


def _combination(n, alpha):
    beta = n - alpha
    tmp = 1.0
    for i in range(1, beta + 1):
        tmp = tmp * float(alpha + i) / float(i)
    return tmp


def jacobi_p(x, alpha, beta, N):
    """Orthonormal Jacobi polynomial, src/mod_utility.f90:299-329 (scalar x)."""
    gamma0 = 2.0 ** (alpha + beta + 1) / float(alpha + beta + 1) / _combination(alpha + beta, alpha)
    gamma1 = float(alpha + 1) * float(beta + 1) / float(alpha + beta + 3) * gamma0
    if N == 0:
        return 1.0 / math.sqrt(gamma0)
    p2 = (float(alpha + beta + 2) * x / 2.0 + float(alpha - beta) / 2.0) / math.sqrt(gamma1)
    if N == 1:
        return p2
    p1 = 1.0 / math.sqrt(gamma0)
    a1 = 2.0 / float(2 + alpha + beta) * math.sqrt(float(alpha + 1) * float(beta + 1) / float(alpha + beta + 3))
    res = p2
    for i in range(1, N):
        hh8 = float(2 * i + alpha + beta); ir8 = float(i)
        a2 = 2.0 / (hh8 + 2.0) * math.sqrt((ir8 + 1.0) * (ir8 + 1.0 + float(alpha + beta))
                                           * float(i + 1 + alpha) * float(i + 1 + beta) / (hh8 + 1.0) / (hh8 + 3.0))
        a3 = -float(alpha ** 2 - beta ** 2) / hh8 / (hh8 + 2.0)
        res = 1.0 / a2 * (-a1 * p1 + (x - a3) * p2)
        a1 = a2; p1 = p2; p2 = res
    return res


def djacobi_p(x, alpha, beta, N):
    if N == 0:
        return 0.0
    return math.sqrt(float(N) * float(N + alpha + beta + 1)) * jacobi_p(x, alpha + 1, beta + 1, N - 1)


def _abc(r, s, t):
    a = -1.0 if abs(s + t) <= TOL else 2.0 * (1.0 + r) / (-s - t) - 1.0
    b = -1.0 if abs(t - 1) <= TOL else 2.0 * (1.0 + s) / (1.0 - t) - 1.0
    return a, b, t


def basis3d(r, s, t, i, j, k):
    p = np.zeros(len(r))
    for n in range(len(r)):
        a, b, c = _abc(r[n], s[n], t[n])
        h1 = jacobi_p(a, 0, 0, i); h2 = jacobi_p(b, 2 * i + 1, 0, j); h3 = jacobi_p(c, 2 * (i + j) + 2, 0, k)
        p[n] = 2.0 * math.sqrt(2.0) * h1 * ((1.0 - b) ** i) * h2 * ((1.0 - c) ** (i + j)) * h3
    return p


def vandermonde3d(N, r, s, t):
    cols = []
    for i in range(N + 1):
        for j in range(N - i + 1):
            for k in range(N - i - j + 1):
                cols.append(basis3d(r, s, t, i, j, k))
    return np.array(cols).T


def grad_basis3d(r, s, t, i, j, k):
    """src/mod_utility.f90:388-436."""
    n_ = len(r)
    pr = np.zeros(n_); ps = np.zeros(n_); pt = np.zeros(n_)
    for n in range(n_):
        a, b, c = _abc(r[n], s[n], t[n])
        h1 = jacobi_p(a, 0, 0, i); dh1 = djacobi_p(a, 0, 0, i)
        h2 = jacobi_p(b, 2 * i + 1, 0, j); dh2 = djacobi_p(b, 2 * i + 1, 0, j)
        h3 = jacobi_p(c, 2 * (i + j) + 2, 0, k); dh3 = djacobi_p(c, 2 * (i + j) + 2, 0, k)
        v = dh1 * h2 * h3
        if i > 1:
            v = v * ((0.5 * (1.0 - b)) ** (i - 1))
        if (i + j) > 1:
            v = v * ((0.5 * (1.0 - c)) ** (i + j - 1))
        pr[n] = v
        ps[n] = 0.5 * (1.0 + a) * v
        tmp = dh2 * ((0.5 * (1.0 - b)) ** i)
        if i > 0:
            tmp = tmp + (-0.5 * float(i)) * (h2 * (0.5 * (1.0 - b)) ** (i - 1))
        if (i + j) > 1:
            tmp = tmp * ((0.5 * (1.0 - c)) ** (i + j - 1))
        tmp = tmp * h1 * h3
        ps[n] = ps[n] + tmp
        pt[n] = 0.5 * (1.0 + a) * v + 0.5 * (1.0 + b) * tmp
        tmp = dh3 * ((0.5 * (1.0 - c)) ** (i + j))
        if (i + j) > 0:
            tmp = tmp - 0.5 * float(i + j) * (h3 * ((0.5 * (1.0 - c)) ** (i + j - 1)))
        tmp = h1 * h2 * tmp * ((0.5 * (1.0 - b)) ** i)
        pt[n] = pt[n] + tmp
    sc = 2.0 ** (float(2 * i + j) + 1.5)
    return pr * sc, ps * sc, pt * sc


def grad_vandermonde3d(N, r, s, t):
    Vr, Vs, Vt = [], [], []
    for i in range(N + 1):
        for j in range(N - i + 1):
            for k in range(N - i - j + 1):
                a, b, c = grad_basis3d(r, s, t, i, j, k)
                Vr.append(a); Vs.append(b); Vt.append(c)
    return np.array(Vr).T, np.array(Vs).T, np.array(Vt).T


def basis2d(r, s, i, j):
    p = np.zeros(len(r))
    for n in range(len(r)):
        a = -1.0 if abs(s[n] - 1.0) <= TOL else 2.0 * (1.0 + r[n]) / (1.0 - s[n]) - 1.0
        b = s[n]
        p[n] = math.sqrt(2.0) * jacobi_p(a, 0, 0, i) * jacobi_p(b, 2 * i + 1, 0, j) * (1.0 - b) ** i
    return p


def vandermonde2d(N, r, s):
    cols = []
    for i in range(N + 1):
        for j in range(N - i + 1):
            cols.append(basis2d(r, s, i, j))
    return np.array(cols).T


def blend_nodes(porder):
    """Reference nodes, src/mod_utility.f90:494-635.  For pOrder<=2 the
    Gauss-Lobatto points coincide with the equispaced ones so the warp/shift is
    identically zero; only the affine round trip through the equilateral tet
    remains (kept so rounding follows the reference)."""
    if porder > 2:
        raise NotImplementedError("reference is used with pOrder 1 or 2 (src/mod_para.f90:111)")
    r, s, t = [], [], []
    for i in range(1, porder + 2):
        for j in range(1, porder + 3 - i):
            for k in range(1, porder + 4 - i - j):
                r.append(-1.0 + float(k - 1) * 2.0 / float(porder))
                s.append(-1.0 + float(j - 1) * 2.0 / float(porder))
                t.append(-1.0 + float(i - 1) * 2.0 / float(porder))
    r = np.array(r); s = np.array(s); t = np.array(t)
    sq3 = math.sqrt(3.0); sq6 = math.sqrt(6.0)
    v1 = np.array([-1.0, -1.0 / sq3, -1.0 / sq6]); v2 = np.array([1.0, -1.0 / sq3, -1.0 / sq6])
    v3 = np.array([0.0, 2.0 / sq3, -1.0 / sq6]);   v4 = np.array([0.0, 0.0, 3.0 / sq6])
    L1 = (1.0 + t) / 2.0; L2 = (1.0 + s) / 2.0; L3 = -(1.0 + r + s + t) / 2.0; L4 = (1.0 + r) / 2.0
    XYZ = np.array([L3 * v1[i] + L4 * v2[i] + L2 * v3[i] + L1 * v4[i] for i in range(3)])
    r, s, t = XYZ[0], XYZ[1], XYZ[2]
    den = 4.0 * math.sqrt(2.0)
    x = (v1[0] * (v4[1] * v3[2] - v3[1] * v4[2]) + v3[0] * (v1[1] * v4[2] - v4[1] * v1[2])
         + v4[0] * (v3[1] * v1[2] - v1[1] * v3[2])
         + (v1[1] * (v3[2] - v4[2]) + v3[1] * (v4[2] - v1[2]) + v4[1] * (v1[2] - v3[2])) * r
         + (v1[0] * (v4[2] - v3[2]) + v3[0] * (v1[2] - v4[2]) + v4[0] * (v3[2] - v1[2])) * s
         + (v1[0] * (v3[1] - v4[1]) + v3[0] * (v4[1] - v1[1]) + v4[0] * (v1[1] - v3[1])) * t) / 4.0 / math.sqrt(2.0)
    y = (v1[1] * (v4[0] * v2[2] - v2[0] * v4[2]) + v2[1] * (v1[0] * v4[2] - v4[0] * v1[2])
         + v4[1] * (v2[0] * v1[2] - v1[0] * v2[2])
         + (v1[1] * (v4[2] - v2[2]) + v2[1] * (v1[2] - v4[2]) + v4[1] * (v2[2] - v1[2])) * r
         + (v1[0] * (v2[2] - v4[2]) + v2[0] * (v4[2] - v1[2]) + v4[0] * (v1[2] - v2[2])) * s
         + (v1[0] * (v4[1] - v2[1]) + v2[0] * (v1[1] - v4[1]) + v4[0] * (v2[1] - v1[1])) * t) / 4.0 / math.sqrt(2.0)
    z = (v1[2] * (v3[0] * v2[1] - v2[0] * v3[1]) + v2[2] * (v1[0] * v3[1] - v3[0] * v1[1])
         + v3[2] * (v2[0] * v1[1] - v1[0] * v2[1])
         + (v1[1] * (v2[2] - v3[2]) + v2[1] * (v3[2] - v1[2]) + v3[1] * (v1[2] - v2[2])) * r
         + (v1[0] * (v3[2] - v2[2]) + v2[0] * (v1[2] - v3[2]) + v3[0] * (v2[2] - v1[2])) * s
         + (v1[0] * (v2[1] - v3[1]) + v2[0] * (v3[1] - v1[1]) + v3[0] * (v1[1] - v2[1])) * t) / 4.0 / math.sqrt(2.0)
    del den
    return x * 2.0 - 1.0, y * 2.0 - 1.0, z * 2.0 - 1.0


class RefTet:
    """Build_reference, src/mod_geometry.f90:2307-2523.  All index arrays 0-based."""

    def __init__(self, porder):
        self.porder = porder
        self.pNp = (porder + 1) * (porder + 2) * (porder + 3) // 6
        self.Nfp = (porder + 1) * (porder + 2) // 2
        r, s, t = blend_nodes(porder)
        self.nodes = np.array([r, s, t])                      # (3,pNp)
        # FtoV: face f lacks vertex f (src/mod_geometry.f90:2341-2347)
        self.FtoV = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
        fm = [[], [], [], []]
        for n in range(self.pNp):                              # :2392-2405
            if abs(1.0 + r[n]) <= TOL: fm[1].append(n)
            if abs(1.0 + s[n]) <= TOL: fm[2].append(n)
            if abs(1.0 + t[n]) <= TOL: fm[3].append(n)
            if abs(1.0 + r[n] + s[n] + t[n]) <= TOL: fm[0].append(n)
        assert all(len(f) == self.Nfp for f in fm)
        self.Fmask = np.array(fm).T                            # (Nfp,4)
        self.V3D = vandermonde3d(porder, r, s, t)
        self.invV = matinv(self.V3D)
        self.MassM = self.invV.T @ self.invV                   # :2473
        D1, D2, D3 = grad_vandermonde3d(porder, r, s, t)
        self.Drst = np.array([D1 @ self.invV, D2 @ self.invV, D3 @ self.invV])   # :2485-2487
        self.MassF = np.zeros((4, self.Nfp, self.Nfp))         # :2491-2510
        pairs = [(s, t), (s, t), (r, t), (r, s)]
        for f in range(4):
            a, b = pairs[f]
            V2D = vandermonde2d(porder, a[self.Fmask[:, f]], b[self.Fmask[:, f]])
            self.MassF[f] = matinv(V2D @ V2D.T)
        self.vord = np.array([0, 1, 2, 3]) if porder == 1 else np.array([0, 2, 5, 9])   # :2514-2518
