"""Host-side mirror of the reference's solve driver `pevsl_mod` (src/mod_pevsl.f90): same call
order, same hard-coded constants; every pEVSL call lands in libnm_b200.so.

    pnm_apply_pevsl   src/mod_pevsl.f90:16-222
"""
import ctypes as C
import time

import numpy as np

from ._lib import lib, check, dptr, f64
from . import matvec as mvmod


def freq_interval(lowfreq, upfreq, lmin):
    """XINTV(1:2), src/mod_pevsl.f90:43,93-103.  PI is assigned from a default-real literal and the
    frequencies are default reals (src/mod_para.f90:51-52) => float32 values promoted to double."""
    pi32 = float(np.float32(3.14159265359))
    lo = float(np.float32(lowfreq)); up = float(np.float32(upfreq))
    a = (2.0 * pi32 * lo) ** 2 * 1.0e-6
    b = (2.0 * pi32 * up) ** 2 * 1.0e-6
    if a < 1.0e-10:
        a = lmin
    return a, b


class Pol:
    def __init__(self, xintv, thre_int, thre_ext):
        self.h = C.c_void_p()
        x = f64(xintv)
        check(lib().nm_findpol_create(dptr(x), C.c_double(thre_int), C.c_double(thre_ext), C.byref(self.h)))
        deg = C.c_int(); cc = C.c_double(); dd = C.c_double(); gam = C.c_double(); bar = C.c_double(); typ = C.c_int()
        check(lib().nm_pol_info(self.h, C.byref(deg), C.byref(cc), C.byref(dd), C.byref(gam), C.byref(bar), C.byref(typ)))
        self.deg, self.cc, self.dd, self.gam, self.bar, self.type = deg.value, cc.value, dd.value, gam.value, bar.value, typ.value
        self.mu = np.empty(self.deg + 1)
        check(lib().nm_pol_coeffs(self.h, dptr(self.mu)))

    def free(self):
        if self.h:
            check(lib().nm_pol_free(self.h)); self.h = None


class Result:
    pass


def pnm_apply_pevsl(mv, lowfreq, upfreq, log=None, maxit=None, tol=1.0e-5, seed=None, recheck=True, ritz_tol=None):
    """src/mod_pevsl.f90:16-222 (without the MPI-IO writers).  mv: matvec.MatVec from setupmatvec."""
    L = lib()
    t0 = time.time()
    MLAN, LANSTEP, TOL = 3000, 5000, tol                                       # :49-52
    pevslAB = mvmod.Pevsl(mv.Gpbsiz, mv.pbsiz, mv.nfirst)                      # :54-57
    if seed is not None:
        check(L.nm_pevsl_set_seed(pevslAB.h, C.c_ulonglong(seed)))
    if ritz_tol is not None:
        check(L.nm_pevsl_set_ritz_tol(pevslAB.h, C.c_double(ritz_tol)))
    pevslAB.setbmv_op(mv.opB)                                                   # PEVSL_SETBMV_F90(sparseBV) :69
    pevslAB.setbsol_chebiter(mv.chebB)                                          # CHEBTYPE = 2 :72-73
    pevslAB.setamv_op(mv.opA)                                                   # sparseAV | sparsefsAV :76-80
    pevslAB.set_geneig()                                                        # :82
    LMIN, LMAX = pevslAB.lanbounds(MLAN, LANSTEP, TOL)                          # :84
    t_bounds = time.time() - t0
    if log:
        log("step 0: eigenvalue bounds for B^{-1}A: lmin %.15g lmax %.15g" % (LMIN, LMAX))
    a, b = freq_interval(lowfreq, upfreq, LMIN)                                 # :93-100
    XINTV = np.array([a, b, LMIN, LMAX])                                        # :102-103
    THRE_INT, THRE_EXT = 0.8, 0.7                                               # :108-109
    EVINT = 800
    pol = Pol(XINTV, THRE_INT, THRE_EXT)                                        # :115
    NEV = EVINT + 2
    MLAN = max(4 * NEV, 2000)
    MAXIT = 3 * MLAN if maxit is None else maxit                                # :116-119
    t1 = time.time()
    check(L.nm_pevsl_cheblannr(pevslAB.h, dptr(XINTV), int(MAXIT), C.c_double(TOL), pol.h))     # :122
    t_lan = time.time() - t1
    nev = C.c_int()
    check(L.nm_pevsl_get_nev(pevslAB.h, C.byref(nev)))                          # :124
    NEVOUT = nev.value
    r = Result()
    r.xintv = XINTV; r.pol = pol; r.nev = NEVOUT
    r.eigval = np.empty(NEVOUT); r.res2 = np.empty(NEVOUT)
    r.eigvec = np.empty((NEVOUT, mv.pbsiz))                                     # row i = eigenvector i (ld = pbsiz)
    if NEVOUT > 0:
        check(L.nm_pevsl_copy_result(pevslAB.h, dptr(r.eigval), dptr(r.eigvec), mv.pbsiz, dptr(r.res2)))   # :130
    steps = C.c_int(); deg = C.c_int(); tt = C.c_double(); tf = C.c_double(); tr = C.c_double(); tz = C.c_double()
    nf = C.c_longlong()
    check(L.nm_pevsl_stats(pevslAB.h, C.byref(steps), C.byref(deg), C.byref(tt), C.byref(tf), C.byref(tr), C.byref(tz),
                           C.byref(nf)))
    r.steps, r.deg = steps.value, deg.value
    r.t_bounds, r.t_cheblannr, r.t_filter, r.t_reorth, r.t_ritz = t_bounds, t_lan, tf.value, tr.value, tz.value
    # residual recheck (:144-162): sqrt(sum((A~y - lam B~y)^2)/N)/|lam|, N = global size (single rank here;
    # multi-rank callers reduce the partial sums themselves)
    r.eigerr = np.zeros(NEVOUT)
    if recheck:
        for i in range(NEVOUT):
            y = r.eigvec[i]
            pv = mvmod.sparseAV(y, mv) - r.eigval[i] * mvmod.sparseBV(y, mv)
            r.eigerr[i] = float(pv @ pv)
    o = np.argsort(r.eigval, kind="stable")                                     # ssort_real :165
    r.eigval, r.res2, r.eigvec, r.eigerr = r.eigval[o], r.res2[o], r.eigvec[o], r.eigerr[o]
    pi32 = float(np.float32(3.14159265359))
    r.freq_mhz = np.sqrt(np.abs(r.eigval)) / (2.0 * pi32) * 1.0e3               # :179
    r.t_total = time.time() - t0
    pevslAB.finish()                                                            # :220
    return r


def finalize_eigerr(r, N):
    """Turn the (already globally summed) squared residuals into the reference's 'relative err.'."""
    return np.sqrt(r.eigerr / N) / np.abs(r.eigval)


def pnm_save_eigenvectors(mv, r, fvdata):
    """src/mod_pevsl.f90:188-201: eigenvector i (ascending eigenvalue) of this rank's rows, in physical coordinates
    `EIGVEC * B%diag`, at byte offset `B%sizdist(rank)*8` of `<fvdata>_<i>.dat` (normalmodes_b200.io)."""
    from . import io
    return io.save_eigenvectors(fvdata, r.eigvec, mv.B.diag, mv.B.sizdist, mv.rank)


def pnm_save_results(names, vlist_local, vtxdist, rank, vstat_local=None):
    """src/mod_pevsl.f90:225-243: `unstrM%new%vlist` (and `vstat` when the model has fluid) of this rank."""
    from . import io
    io.save_vlist_vstat(names, vlist_local, vtxdist, rank, vstat_local)
