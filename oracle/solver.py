"""ORACLE (test infrastructure, NOT product code) -- the operator layer and the
Chebyshev-filtered Lanczos solve of NormalModes, restated on the CPU with numpy/scipy.

Parity status: UNPINNED.  The arithmetic of this path lives in the third-party C
library pEVSL (fork js1019/pEVSL, no pinned version; call sites src/mod_matvec.f90:69-95,
146-174,453-515 and src/mod_pevsl.f90:54-134), whose source is absent from
/root/reference.  The routines below restate pEVSL's *published* algorithms
(ChebIter = Saad, Iterative Methods, Alg. 12.1; find_pol / ChebAv / ChebLanNr /
CGS_DGKS2 / LanTrbounds as described in Li, Xi, Erlandson, Saad, "The Eigenvalues
Slicing Library (EVSL)", SISC 2019) and are anchored on the reference's own call
sites, constants and acceptance rules.  What pins results
instead: `truth_eigs` below -- an INDEPENDENT dense / shift-invert solve of the
same assembled pencil -- and the committed fixtures under tests/golden/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla

from . import fem

DBL_EPS_MULT = 10.0
DBL_EPSILON = 2.220446049250313e-16


# --------------------------------------------------------------------------- operator layer (mod_matvec.f90)
class Operators:
    """setupmatvec (src/mod_matvec.f90:24-250) for nproc = 1.

    B~ = D B D (Bdiagscaling :252-342), A applied as D A D (sparseAV :445-458) or
    D [Ad + E Dp Ap~^-1 Dp ET] D (sparsefsAV :498-520) with Ap~ = Dp (-Ap) Dp."""

    def __init__(self, mats, porder, degB=None, degAp=None, bounds=None):
        self.fluid = "Ad" in mats
        Bs, self.d = fem.jacobi_scale(mats["B"])
        self.Bt = fem.to_scipy(Bs)
        self.n = self.Bt.shape[0]
        self.degB = degB if degB is not None else (25 if porder == 1 else 45)       # :88-92
        self.degAp = degAp if degAp is not None else (25 if porder == 1 else 100)    # :167-171
        self.nmatvec = 0
        if self.fluid:
            self.Ad = fem.to_scipy(mats["Ad"]); self.E = fem.to_scipy(mats["E"]); self.ET = fem.to_scipy(mats["ET"])
            Aps, self.dp = fem.jacobi_scale(mats["Ap"], sign=-1.0)                   # :137, 345-441
            self.Apt = fem.to_scipy(Aps)
        else:
            self.A = fem.to_scipy(mats["A"])
        bounds = bounds or {}
        self.boundsB = bounds.get("B") or lanbounds(lambda v: self.Bt @ v, self.n, 1000, 2000, 1e-12)   # :83-86
        if self.fluid:
            self.boundsAp = bounds.get("Ap") or lanbounds(lambda v: self.Apt @ v, self.Apt.shape[0], 2000, 3000, 1e-12)

    def bmv(self, v):                       # sparseBV :461-472
        return self.Bt @ v

    def bsol(self, b):                      # pevsl_setbsol_chebiter (src/mod_pevsl.f90:72-73)
        return chebiter(self.Bt, self.boundsB[0], self.boundsB[1], self.degB, b)

    def amv(self, v):                       # sparseAV / sparsefsAV
        self.nmatvec += 1
        v0 = v * self.d
        if not self.fluid:
            return (self.A @ v0) * self.d
        w0 = self.Ad @ v0
        x1 = (self.ET @ v0) * self.dp
        y0 = chebiter(self.Apt, self.boundsAp[0], self.boundsAp[1], self.degAp, x1)
        w1 = self.E @ (y0 * self.dp)
        return (w0 + w1) * self.d


# --------------------------------------------------------------------------- ChebIter
def chebiter_coeffs(lb, ub, deg):
    """Scalars of the Chebyshev iteration (Saad Alg. 12.1) with zero initial guess:
    returns theta and per-step (a_k, b_k) with d_{k+1} = a_k d_k + b_k r_{k+1}."""
    theta = (ub + lb) / 2.0; delta = (ub - lb) / 2.0
    sigma1 = theta / delta; rho = 1.0 / sigma1
    ab = []
    for _ in range(deg):
        rho1 = 1.0 / (2.0 * sigma1 - rho)
        ab.append((rho1 * rho, 2.0 * rho1 / delta))
        rho = rho1
    return theta, ab


def chebiter(M, lb, ub, deg, b):
    """x = q_deg(M) b ~ M^-1 b: `deg` products with M, residual polynomial of degree deg+1
    (the scaled Chebyshev polynomial on [lb,ub]).  No inner products."""
    theta, ab = chebiter_coeffs(lb, ub, deg)
    r = b.copy(); d = r / theta; x = np.zeros_like(b)
    for (a_k, b_k) in ab:
        x += d
        r -= M @ d
        d = a_k * d + b_k * r
    x += d
    return x


# --------------------------------------------------------------------------- tridiagonal helpers
def _tridiag_eig(dT, eT, vectors=False):
    k = len(dT)
    if k == 1:
        return (np.array(dT[:1]), np.ones((1, 1))) if vectors else np.array(dT[:1])
    if vectors:
        return sla.eigh_tridiagonal(np.asarray(dT), np.asarray(eT[:k - 1]))
    return sla.eigh_tridiagonal(np.asarray(dT), np.asarray(eT[:k - 1]), eigvals_only=True)


# --------------------------------------------------------------------------- LanTrbounds
def lanbounds(amv, n, mlan, lanstep, tol, bsol=None, bmv=None, seed=1234, check_every=10):
    """Outer bounds [lmin, lmax] of the spectrum of the operator (pEVSL LanTrbounds, bndtype 1:
    theta_min - |beta s_min|, theta_max + |beta s_max|; stop when the two residuals sum to less
    than tol*(|lmin|+|lmax|)).  Generalised case: Lanczos on B^-1 A in the B inner product.
    Full reorthogonalisation; a simple thick restart keeps the two extreme Ritz vectors."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal(n)
    gen = bsol is not None
    m = min(mlan, n)
    V = np.zeros((n, m + 1)); Z = np.zeros((n, m + 1)) if gen else V
    if gen:
        z = bmv(v); t = 1.0 / np.sqrt(v @ z); V[:, 0] = v * t; Z[:, 0] = z * t
    else:
        V[:, 0] = v / np.linalg.norm(v)
    dT = []; eT = []; steps = 0; lmin = lmax = 0.0
    T_extra = None                                   # arrow part after a thick restart
    k0 = 0
    while True:
        for k in range(k0, m):
            steps += 1
            if gen:
                zn = amv(V[:, k])
                if k > 0 and T_extra is None:
                    zn -= eT[k - 1] * Z[:, k - 1]
                a = V[:, k] @ zn; dT.append(a)
                for _ in range(2):
                    c = V[:, :k + 1].T @ zn; zn -= Z[:, :k + 1] @ c
                vn = bsol(zn); beta = np.sqrt(abs(vn @ zn))
                V[:, k + 1] = vn / beta; Z[:, k + 1] = zn / beta
            else:
                w = amv(V[:, k])
                a = V[:, k] @ w; dT.append(a)
                for _ in range(2):
                    c = V[:, :k + 1].T @ w; w -= V[:, :k + 1] @ c
                beta = np.linalg.norm(w)
                V[:, k + 1] = w / beta
            eT.append(beta)
            kk = k + 1
            if kk % check_every and kk != m and steps < lanstep:
                continue
            # projected matrix (tridiagonal, or arrowhead + tridiagonal after a restart)
            if T_extra is None:
                th, S = _tridiag_eig(dT, eT, True)
            else:
                T = np.diag(np.array(dT))
                nk, s = T_extra
                for i in range(nk):
                    T[i, nk] = T[nk, i] = s[i]
                for i in range(nk, kk - 1):
                    T[i, i + 1] = T[i + 1, i] = eT[i]
                th, S = np.linalg.eigh(T)
            r1 = abs(beta * S[-1, 0]); r2 = abs(beta * S[-1, -1])
            lmin = th[0] - r1; lmax = th[-1] + r2
            if r1 + r2 < tol * (abs(lmin) + abs(lmax)) or steps >= lanstep:
                return lmin, lmax
        # thick restart with the two extreme Ritz pairs
        Y = S[:, [0, -1]]
        Vn = V[:, :m] @ Y
        s = beta * Y[-1, :]
        V[:, 0:2] = Vn; V[:, 2] = V[:, m]
        if gen:
            Zn = Z[:, :m] @ Y; Z[:, 0:2] = Zn; Z[:, 2] = Z[:, m]
        dT = [th[0], th[-1]]; eT = [0.0, 0.0]
        T_extra = (2, s); k0 = 2


# --------------------------------------------------------------------------- find_pol
def dampcf(m, damping):
    """Damping coefficients (0 none, 1 Jackson, 2 Lanczos sigma); jac[0] = 1/2 carries the
    half weight of the zeroth Chebyshev term."""
    jac = np.ones(m + 1); jac[0] = 0.5
    k = np.arange(1, m + 1, dtype=float)
    if damping == 1:
        thJ = np.pi / (m + 2.0); a1 = 1.0 / (m + 2.0); a2 = np.sin(thJ)
        jac[1:] = a1 * np.sin((k + 1) * thJ) / a2 + (1.0 - (k + 1) * a1) * np.cos(k * thJ)
    elif damping == 2:
        thL = np.pi / (m + 1.0)
        jac[1:] = np.sin(k * thL) / (k * thL)
    return jac


def chebxpltd(mu, x):
    """p(x) = sum_k mu_k T_k(x) on [-1,1] by the three-term recurrence."""
    x = np.atleast_1d(np.asarray(x, dtype=float))
    vkm1 = np.zeros_like(x); vk = np.ones_like(x)
    y = mu[0] * vk
    for k in range(1, len(mu)):
        s = 1.0 if k == 1 else 2.0
        vkp1 = s * x * vk - vkm1
        y = y + mu[k] * vkp1
        vkm1, vk = vk, vkp1
    return y


def rootchb(m, jac, tha, thb):
    """Centre thc of the damped delta expansion mu_j = cos(j thc) jac_j such that
    p(cos tha) = p(cos thb): safeguarded Newton on [thb, tha]."""
    j = np.arange(m + 1, dtype=float)
    w = jac * (np.cos(j * tha) - np.cos(j * thb))

    def f(t):
        return float(np.sum(w * np.cos(j * t)))

    def df(t):
        return float(-np.sum(w * j * np.sin(j * t)))
    lo, hi = thb, tha
    flo, fhi = f(lo), f(hi)
    thc = 0.5 * (tha + thb)
    tol = abs(tha - thb) * 1.0e-13
    if flo * fhi > 0:
        return thc
    for _ in range(200):
        fv = f(thc)
        if fv == 0.0:
            break
        if (fv > 0) == (flo > 0):
            lo, flo = thc, fv
        else:
            hi, fhi = thc, fv
        dv = df(thc)
        tn = thc - fv / dv if dv != 0.0 else 0.5 * (lo + hi)
        if not (lo < tn < hi):
            tn = 0.5 * (lo + hi)
        if abs(tn - thc) < tol:
            thc = tn
            break
        thc = tn
    return thc


def findpol(xintv, thresh_int=0.8, thresh_ext=0.7, max_deg=10000, min_deg=2, damping=2, intvtol=1e-9):
    """pEVSL find_pol as called from src/mod_pevsl.f90:108-115.  xintv = [a, b, lmin, lmax]."""
    a, b, lmin, lmax = [float(t) for t in xintv]
    cc = (lmax + lmin) / 2.0; dd = (lmax - lmin) / 2.0
    aa = max(a, lmin); bb = min(b, lmax)
    itv = [max(-1.0, (aa - cc) / dd), min(1.0, (bb - cc) / dd)]
    tha = np.arccos(itv[0]); thb = np.arccos(itv[1])
    pol = dict(cc=cc, dd=dd, intv=[a, b, lmin, lmax], damping=damping)
    if aa - intvtol <= lmin or bb + intvtol >= lmax:
        left = aa - intvtol <= lmin
        thc = tha if left else thb
        xin = itv[0] if left else itv[1]           # where p peaks (spectrum end)
        xout = itv[1] if left else itv[0]          # inner edge of the wanted interval
        for m in range(min_deg, max_deg):
            jac = dampcf(m, damping)
            mu = np.cos(np.arange(m + 1) * thc) * jac
            t = chebxpltd(mu, xin)[0]; v = chebxpltd(mu, xout)[0]
            if v <= t * thresh_ext:
                break
        mu = mu / t
        pol.update(deg=m, mu=mu, gam=xin, bar=v / t, type=1 if left else 3)
        return pol
    for m in range(min_deg, max_deg):
        jac = dampcf(m, damping)
        thc = rootchb(m, jac, tha, thb)
        mu = np.cos(np.arange(m + 1) * thc) * jac
        gam = np.cos(thc)
        t = chebxpltd(mu, gam)[0]
        vals = chebxpltd(mu, itv)
        if vals[0] <= t * thresh_int and vals[1] <= t * thresh_int:
            break
    mu = mu / t
    pol.update(deg=m, mu=mu, gam=gam, bar=min(vals[0], vals[1]) / t, type=2)
    return pol


# --------------------------------------------------------------------------- ChebAv
def chebav(pol, z, ops):
    """y = p(A B^-1) z,  p = sum mu_k T_k((. - cc)/dd)  (pEVSL ChebAv, generalised branch)."""
    mu, cc, dd, m = pol["mu"], pol["cc"], pol["dd"], pol["deg"]
    vk = z.copy(); vkm1 = np.zeros_like(z)
    y = mu[0] * vk
    for k in range(1, m + 1):
        t = (1.0 if k == 1 else 2.0) / dd
        w = ops.amv(ops.bsol(vk))
        vkp1 = t * (w - cc * vk) - vkm1
        y += mu[k] * vkp1
        vkm1, vk = vk, vkp1
    return y


# --------------------------------------------------------------------------- ChebLanNr
def cheblannr(ops, xintv, maxit, tol, pol, seed=4321, ntest=30, cycle=20, ngs=2, log=None):
    """Non-restarted polynomial-filtered Lanczos with full reorthogonalisation in the
    B inner product (pEVSL ChebLanNr).  Returns (lam, Y, res, info):
    eigenvalues in [a,b], B-orthonormal eigenvectors (scaled coordinates, columns), and
    ||A y - lam B y||_2."""
    n = ops.n
    aa, bb = xintv[0], xintv[1]
    bar = pol["bar"]
    maxit = min(n, maxit)
    rng = np.random.default_rng(seed)
    V = np.zeros((n, maxit + 1)); Z = np.zeros((n, maxit + 1))
    v = rng.standard_normal(n)
    z = ops.bmv(v); t = 1.0 / np.sqrt(v @ z)
    V[:, 0] = v * t; Z[:, 0] = z * t
    dT = []; eT = []
    beta = 0.0; wn = 0.0; nwn = 0; tr0 = 0.0
    orthtol = 1e-14
    kdim = 0
    for k in range(maxit):
        znew = chebav(pol, Z[:, k], ops)
        if k > 0:
            znew -= beta * Z[:, k - 1]
        alpha = V[:, k] @ znew
        dT.append(alpha); wn += abs(alpha)
        znew -= alpha * Z[:, k]
        for _ in range(ngs):                               # CGS_DGKS2(n, k+1, NGS_MAX, Z, V, znew)
            c = V[:, :k + 1].T @ znew
            znew -= Z[:, :k + 1] @ c
        vnew = ops.bsol(znew)
        beta = np.sqrt(vnew @ znew)
        wn += 2.0 * beta; nwn += 3
        if beta * nwn < orthtol * wn:                      # lucky breakdown: restart with a random vector
            vnew = rng.standard_normal(n)
            for _ in range(ngs):
                c = Z[:, :k + 1].T @ vnew
                vnew -= V[:, :k + 1] @ c
            znew = ops.bmv(vnew)
            beta = np.sqrt(vnew @ znew)
            V[:, k + 1] = vnew / beta; Z[:, k + 1] = znew / beta
            beta = 0.0
        else:
            V[:, k + 1] = vnew / beta; Z[:, k + 1] = znew / beta
        eT.append(beta)
        kdim = k + 1
        if (k < ntest or (k - ntest) % cycle != 0) and k != maxit - 1:
            continue
        th = _tridiag_eig(dT, eT)
        sel = th + DBL_EPS_MULT * DBL_EPSILON >= bar
        tr1 = th[sel].sum()
        if log is not None:
            log.append((k, tr1, int(sel.sum())))
        if abs(tr1 - tr0) < tol * abs(tr1):
            break
        tr0 = tr1
    th, S = _tridiag_eig(dT, eT, True)
    lam = []; Y = []; res = []
    for i in range(kdim):
        if th[i] < bar:
            continue
        u = V[:, :kdim] @ S[:, i]
        w2 = ops.bmv(u); t = 1.0 / np.sqrt(u @ w2)
        u = u * t; w2 = w2 * t
        wk = ops.amv(u)
        t = wk @ u
        if t < aa - DBL_EPS_MULT * DBL_EPSILON or t > bb + DBL_EPS_MULT * DBL_EPSILON:
            continue
        lam.append(t); Y.append(u); res.append(np.linalg.norm(wk - t * w2))
    lam = np.array(lam); o = np.argsort(lam)
    Y = np.array(Y).T if len(Y) else np.zeros((n, 0))
    return lam[o], Y[:, o], np.array(res)[o], dict(steps=kdim, deg=pol["deg"])


# --------------------------------------------------------------------------- mod_pevsl.f90 driver pieces
def freq_interval(lowfreq, upfreq, lmin):
    """XINTV(1:2) exactly as src/mod_pevsl.f90:43,93-103: PI and the frequencies are float32."""
    pi32 = float(np.float32(3.14159265359))
    lo = float(np.float32(lowfreq)); up = float(np.float32(upfreq))
    a = (2.0 * pi32 * lo) ** 2 * 1.0e-6
    b = (2.0 * pi32 * up) ** 2 * 1.0e-6
    if a < 1.0e-10:
        a = lmin
    return a, b


def residual_rms(ops, lam, y):
    """'relative err.' of src/mod_pevsl.f90:144-162: sqrt(sum((A~y - lam B~y)^2)/N)/|lam|."""
    r = ops.amv(y) - lam * ops.bmv(y)
    return np.sqrt((r @ r) / ops.n) / abs(lam)


def solve(mats, porder, lowfreq, upfreq, degB=None, degAp=None, maxit=9624, tol=1e-5, log=None):
    """pnm_apply_pevsl (src/mod_pevsl.f90:16-222) for nproc = 1."""
    ops = Operators(mats, porder, degB, degAp)
    lmin, lmax = lanbounds(ops.amv, ops.n, 3000, 5000, 1e-5, bsol=ops.bsol, bmv=ops.bmv)    # :84
    a, b = freq_interval(lowfreq, upfreq, lmin)
    xintv = [a, b, lmin, lmax]
    pol = findpol(xintv, 0.8, 0.7)                                                            # :108-115
    lam, Y, res, info = cheblannr(ops, xintv, maxit, tol, pol, log=log)
    info.update(xintv=xintv, pol=pol)
    return ops, lam, Y, res, info


# --------------------------------------------------------------------------- independent truth
def effective_pencil(mats):
    """(A_eff, B) in physical (unscaled) coordinates; fluid case: dense-free Schur complement
    A_eff = Ad + E (-Ap)^-1 ET with a sparse LU of -Ap (exact, unlike the reference's ChebIter)."""
    B = fem.to_scipy(mats["B"]).tocsc()
    if "A" in mats:
        return fem.to_scipy(mats["A"]).tocsc(), B
    Ad = fem.to_scipy(mats["Ad"]); E = fem.to_scipy(mats["E"]); ET = fem.to_scipy(mats["ET"])
    mAp = (-fem.to_scipy(mats["Ap"])).tocsc()
    X = spla.splu(mAp).solve(ET.toarray())
    return sp.csc_matrix(Ad + sp.csr_matrix(E @ X)), B


def truth_eigs(mats, a, b, dense_limit=6000):
    """All eigenvalues of A_eff x = lam B x in [a,b] by a method independent of the filtered
    Lanczos: dense LAPACK for small N, otherwise shift-invert Lanczos around the band centre."""
    n = mats["B"]["shape"][0]
    if n > dense_limit and "Ad" in mats:
        # fluid case at scale: shift-invert on the sparse augmented (u,p) pencil
        # [Ad E; ET Ap][u;p] = lam [B 0; 0 0][u;p], whose finite eigenvalues are those of the Schur complement
        Ad = fem.to_scipy(mats["Ad"]); E = fem.to_scipy(mats["E"]); ET = fem.to_scipy(mats["ET"])
        Ap = fem.to_scipy(mats["Ap"]); Bm = fem.to_scipy(mats["B"])
        sigma = (a + b) / 2.0
        K = sp.bmat([[Ad - sigma * Bm, E], [ET, Ap]], format="csc")
        lu = spla.splu(K)
        npr = Ap.shape[0]

        def opinv(rhs):                      # (A_eff - sigma B)^-1 rhs through the augmented LU
            return lu.solve(np.concatenate([rhs, np.zeros(npr)]))[:n]
        OPinv = spla.LinearOperator((n, n), matvec=opinv, dtype=float)
        Aop = spla.LinearOperator((n, n), matvec=lambda v: v, dtype=float)   # unused in shift-invert mode
        k = 64
        while True:
            w = spla.eigsh(Aop, k=min(k, n - 2), M=Bm.tocsc(), sigma=sigma, which="LM", OPinv=OPinv,
                           return_eigenvectors=False)
            w = np.sort(w)
            # shift-invert returns the k eigenvalues NEAREST sigma: everything within rho = max|w - sigma| of it is in w.
            # (Fluid models have a dense cluster of small eigenvalues below the band: waiting for w.max() > b as well
            # would drag the whole cluster in.)
            rho = np.abs(w - sigma).max()
            if (sigma - rho <= a and sigma + rho >= b) or k >= n - 2:
                return w[(w >= a) & (w <= b)]
            k *= 2
    A, B = effective_pencil(mats)
    if n <= dense_limit:
        Ad_ = A.toarray(); Ad_ = (Ad_ + Ad_.T) / 2.0
        w = sla.eigh(Ad_, B.toarray(), eigvals_only=True)
        return w[(w >= a) & (w <= b)]
    sigma = (a + b) / 2.0
    k = 64
    while True:
        w = spla.eigsh(A, k=min(k, n - 2), M=B, sigma=sigma, which="LM", return_eigenvectors=False)
        w = np.sort(w)
        rho = np.abs(w - sigma).max()
        if (sigma - rho <= a and sigma + rho >= b) or k >= n - 2:
            return w[(w >= a) & (w <= b)]
        k *= 2
