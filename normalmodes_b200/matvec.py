"""Host-side mirror of the reference's operator layer `cg_matvec_mod` (src/mod_matvec.f90): same
names, same order of operations, same constants -- the arithmetic runs in libnm_b200.so.

    setupmatvec   src/mod_matvec.f90:24-250
    sparseAV      :445-458      sparseBV  :461-472
    sparseApV     :485-496      sparsefsAV :498-520
"""
import ctypes as C

import numpy as np

from ._lib import lib, check, dptr, iptr, f64, i32


class COOmat:
    """`type COOmat` (src/mod_cg_datatype.f90:35-49): one rank's block of a distributed CSR matrix.
    sizdist: global row offsets per rank; coldist: global column offsets per rank (square: same);
    rowdist: 0-based local row pointers; col: 0-based GLOBAL column ids; val: fp64."""

    def __init__(self, sizdist, rowdist, col, val, coldist=None):
        self.sizdist = i32(sizdist)
        self.coldist = i32(coldist if coldist is not None else sizdist)
        self.rowdist = i32(rowdist)
        self.col = i32(col)
        self.val = f64(val)
        self.Gsiz = int(self.sizdist[-1])
        self.Gcol = int(self.coldist[-1])
        self.NNZ = int(self.rowdist[-1])
        self.diag = None
        self.handle = None

    def siz(self, rank):
        return int(self.sizdist[rank + 1] - self.sizdist[rank])


def parcsr_create(m):
    """PEVSL_PARCSRCREATE_F90 (src/mod_matvec.f90:69-71)."""
    h = C.c_void_p()
    check(lib().nm_parcsr_create(m.Gsiz, m.Gcol, iptr(m.sizdist), iptr(m.coldist), iptr(m.rowdist), iptr(m.col),
                                 dptr(m.val), C.byref(h)))
    m.handle = h
    return h


def parcsr_matvec(h, x, nrow):
    y = np.empty(nrow)
    check(lib().nm_parcsr_matvec(h, dptr(f64(x)), dptr(y)))
    return y


def parcsr_info(h):
    nrow = C.c_int(); ncol = C.c_int(); nnz = C.c_longlong(); fmt = C.c_int(); ng = C.c_int(); fb = C.c_longlong()
    check(lib().nm_parcsr_info(h, C.byref(nrow), C.byref(ncol), C.byref(nnz), C.byref(fmt), C.byref(ng), C.byref(fb)))
    return dict(nrow=nrow.value, ncol=ncol.value, nnz=nnz.value, format=("CSR", "ROW3", "KRON3")[fmt.value],
                nghost=ng.value, fmt_bytes=fb.value)


class Pevsl:
    """A pEVSL context handle (pEVSL_Start_F90 ... PEVSL_FINISH_F90)."""

    def __init__(self, N, n, nfirst=-1):
        self.h = C.c_void_p()
        check(lib().nm_pevsl_create(C.byref(self.h)))
        check(lib().nm_pevsl_setprobsizes(self.h, int(N), int(n), int(nfirst)))
        self._keep = []

    def setamv_op(self, op):
        check(lib().nm_pevsl_setamv_op(self.h, op))

    def setbmv_op(self, op):
        check(lib().nm_pevsl_setbmv_op(self.h, op))

    def setamv_callback(self, fn):
        from ._lib import MATVEC_FN
        cb = MATVEC_FN(fn); self._keep.append(cb)
        check(lib().nm_pevsl_setamv_callback(self.h, cb, None))

    def setbmv_callback(self, fn):
        from ._lib import MATVEC_FN
        cb = MATVEC_FN(fn); self._keep.append(cb)
        check(lib().nm_pevsl_setbmv_callback(self.h, cb, None))

    def setbsol_chebiter(self, cheb):
        check(lib().nm_pevsl_setbsol_chebiter(self.h, cheb))

    def set_geneig(self):
        check(lib().nm_pevsl_set_geneig(self.h))

    def lanbounds(self, mlan, lanstep, tol):
        lmin = C.c_double(); lmax = C.c_double()
        check(lib().nm_pevsl_lanbounds(self.h, int(mlan), int(lanstep), C.c_double(tol), C.byref(lmin), C.byref(lmax)))
        return lmin.value, lmax.value

    def finish(self):
        if self.h:
            check(lib().nm_pevsl_free(self.h))
            self.h = None


def op_csr(h):
    op = C.c_void_p()
    check(lib().nm_op_create_csr(h, C.byref(op)))
    return op


def chebiter_setup(lmin, lmax, deg, h):
    """pEVSL_SETUP_CHEBITER_F90 (src/mod_matvec.f90:93,174)."""
    cheb = C.c_void_p()
    check(lib().nm_chebiter_create(C.c_double(lmin), C.c_double(lmax), int(deg), h, C.byref(cheb)))
    return cheb


def chebiter_solve(cheb, b):
    x = np.empty_like(b)
    check(lib().nm_chebiter_solve_host(cheb, dptr(f64(b)), dptr(x)))
    return x


class MatVec:
    """`type mvparameters` (src/mod_cg_datatype.f90:88-115): the saved state `mymatvec`."""
    pass


def setupmatvec(CGM, porder, rank=0, nproc=1, degB=None, degAp=None, log=None):
    """src/mod_matvec.f90:24-250.  CGM: dict of COOmat -- {'A','B'} (solid) or {'Ad','B','E','ET','Ap'}
    (fluid / fluid-solid), unscaled, 0-based columns.  Returns the MatVec with device handles."""
    L = lib()
    mv = MatVec()
    mv.rank = rank; mv.nproc = nproc
    mv.fluid = "Ad" in CGM
    B = CGM["B"]
    mv.B = B
    mv.Gpbsiz = B.Gsiz
    mv.pbsiz = B.siz(rank)
    mv.nfirst = int(B.sizdist[rank])
    # -- B: handle, Jacobi scaling on the device (Bdiagscaling :65, 252-342)
    mv.sBV = parcsr_create(B)
    B.diag = np.empty(mv.pbsiz)
    check(L.nm_parcsr_jacobi_scale(mv.sBV, C.c_double(1.0), dptr(B.diag)))
    # -- bounds of B~ (:75-86) and the fixed-degree Chebyshev B-solve (:88-93)
    pevslB = Pevsl(mv.Gpbsiz, mv.pbsiz, mv.nfirst)
    mv.opB = op_csr(mv.sBV)
    pevslB.setamv_op(mv.opB)
    MLAN, LANSTEP, TOL = 1000, 2000, 1.0e-12
    mv.boundsB = pevslB.lanbounds(MLAN, LANSTEP, TOL)
    if log is not None:
        log("bounds of B~: %.15g %.15g" % mv.boundsB)
    mv.degB = degB if degB is not None else (25 if porder == 1 else 45)
    mv.chebB = chebiter_setup(mv.boundsB[0], mv.boundsB[1], mv.degB, mv.sBV)
    pevslB.finish()
    if mv.fluid:
        Ad, Ap, E, ET = CGM["Ad"], CGM["Ap"], CGM["E"], CGM["ET"]
        mv.Ad, mv.Ap, mv.E, mv.ET = Ad, Ap, E, ET
        mv.sAdV = parcsr_create(Ad)                                   # :117-119
        # Ap := -CGM%Ap, Jacobi-scaled (:137, Apdiagscaling :345-441)
        mv.sApV = parcsr_create(Ap)                                   # :146-148
        Ap.diag = np.empty(Ap.siz(rank))
        check(L.nm_parcsr_jacobi_scale(mv.sApV, C.c_double(-1.0), dptr(Ap.diag)))
        pevslAp = Pevsl(Ap.Gsiz, Ap.siz(rank), int(Ap.sizdist[rank]))
        mv.opAp = op_csr(mv.sApV)
        pevslAp.setamv_op(mv.opAp)
        MLAN, LANSTEP = 2000, 3000                                    # :160-161
        mv.boundsAp = pevslAp.lanbounds(MLAN, LANSTEP, TOL)
        if log is not None:
            log("bounds of Ap~: %.15g %.15g" % mv.boundsAp)
        mv.degAp = degAp if degAp is not None else (25 if porder == 1 else 100)   # :167-171
        mv.chebAp = chebiter_setup(mv.boundsAp[0], mv.boundsAp[1], mv.degAp, mv.sApV)
        pevslAp.finish()
        mv.sEV = parcsr_create(E)                                     # :196-198
        mv.sETV = parcsr_create(ET)                                   # :218-220
        # device-resident sparsefsAV
        mv.opA = C.c_void_p()
        check(L.nm_op_create_fluidsolid(mv.sAdV, mv.sEV, mv.sETV, mv.chebAp, dptr(B.diag), dptr(Ap.diag),
                                        C.byref(mv.opA)))
    else:
        A = CGM["A"]
        mv.A = A
        mv.sAV = parcsr_create(A)                                     # :242-244
        mv.opA = C.c_void_p()
        check(L.nm_op_create_solid(mv.sAV, dptr(B.diag), C.byref(mv.opA)))
    return mv


# ---- the reference's callbacks, host-vector form (used for the residual recheck, src/mod_pevsl.f90:144-162)
def op_apply(op, v):
    w = np.empty_like(v)
    check(lib().nm_op_apply_host(op, dptr(f64(v)), dptr(w)))
    return w


def sparseAV(v, mv):
    """w = D A D v (:445-458) -- or sparsefsAV (:498-520) when the model has fluid."""
    return op_apply(mv.opA, v)


sparsefsAV = sparseAV


def sparseBV(v, mv):
    """w = B~ v (:461-472)."""
    return parcsr_matvec(mv.sBV, v, mv.pbsiz)


def solveBV(v, mv):
    """w = q(B~) v (:475-482)."""
    return chebiter_solve(mv.chebB, v)


def sparseApV(v, mv):
    """w = Ap~ v (:485-496)."""
    return parcsr_matvec(mv.sApV, v, mv.Ap.siz(mv.rank))
