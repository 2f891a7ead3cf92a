"""ORACLE (test infrastructure, NOT product code): ctypes wrapper of oracle/c/nm_cpu.c, the C + OpenMP
restatement of the reference stack's CPU hot loops.  Used by tests (second checker) and by bench.py's
cpu_baseline / --impl reference legs only.  Parity status: UNPINNED (see the C file's header)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "nm_cpu.c")
LIB = os.path.join(HERE, "c", "libnm_cpu.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Csr(C.Structure):
    _fields_ = [("nrow", C.c_int), ("ia", _ip), ("ja", _ip), ("a", _dp)]


class _Ops(C.Structure):
    _fields_ = [("n", C.c_int), ("np", C.c_int), ("fluid", C.c_int),
                ("B", _Csr), ("A", _Csr), ("E", _Csr), ("ET", _Csr), ("Ap", _Csr),
                ("d", _dp), ("dp", _dp),
                ("lbB", C.c_double), ("ubB", C.c_double), ("lbAp", C.c_double), ("ubAp", C.c_double),
                ("degB", C.c_int), ("degAp", C.c_int)]


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        # all the host cores this process may run on, whatever OMP_NUM_THREADS says: torchrun exports
        # OMP_NUM_THREADS=1 to every rank, which would time the CPU arm on one core (NM_CPU_THREADS overrides)
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
        n = int(os.environ.get("NM_CPU_THREADS", n))
        _lib.nmcpu_set_threads(int(n))
    return _lib


def threads():
    return int(lib().nmcpu_threads())


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class CpuOps:
    """Holds the CSR arrays (int32 / float64, C-contiguous) of one setupmatvec and the ops_t view of them.
    B: Jacobi-scaled mass; A (or Ad), E, ET: unscaled; Ap: Jacobi-scaled -Ap; d, dp: the scalings."""

    def __init__(self, B, A, d, boundsB, degB, E=None, ET=None, Ap=None, dp=None, boundsAp=None, degAp=0):
        self.keep = []

        def csr(m):
            ia = np.ascontiguousarray(m[0], dtype=np.int32); ja = np.ascontiguousarray(m[1], dtype=np.int32)
            a = np.ascontiguousarray(m[2], dtype=np.float64)
            self.keep += [ia, ja, a]
            return _Csr(len(ia) - 1, _i(ia), _i(ja), _d(a))
        o = _Ops()
        o.n = len(B[0]) - 1
        o.fluid = 1 if E is not None else 0
        o.np = (len(Ap[0]) - 1) if o.fluid else 0
        o.B = csr(B); o.A = csr(A)
        self.d = np.ascontiguousarray(d, dtype=np.float64); o.d = _d(self.d)
        o.lbB, o.ubB, o.degB = boundsB[0], boundsB[1], degB
        if o.fluid:
            o.E = csr(E); o.ET = csr(ET); o.Ap = csr(Ap)
            self.dp = np.ascontiguousarray(dp, dtype=np.float64); o.dp = _d(self.dp)
            o.lbAp, o.ubAp, o.degAp = boundsAp[0], boundsAp[1], degAp
        self.o = o
        self.n = o.n

    def apply_A(self, v):
        w = np.empty(self.n)
        lib().nmcpu_apply_A(C.byref(self.o), _d(np.ascontiguousarray(v)), _d(w))
        return w

    def bsol(self, b):
        x = np.empty(self.n); work = np.empty(3 * self.n)
        B = self.o.B
        lib().nmcpu_chebiter(self.n, B.ia, B.ja, B.a, C.c_double(self.o.lbB), C.c_double(self.o.ubB), self.o.degB,
                             _d(np.ascontiguousarray(b)), _d(x), _d(work))
        return x

    def chebav(self, deg, mu, cc, dd, z, kmax=None):
        y = np.empty(self.n)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        lib().nmcpu_chebav(C.byref(self.o), int(deg), _d(mu), C.c_double(cc), C.c_double(dd),
                           int(deg if kmax is None else kmax), _d(np.ascontiguousarray(z)), _d(y))
        return y


def spmv(ia, ja, a, x):
    ia = np.ascontiguousarray(ia, dtype=np.int32); ja = np.ascontiguousarray(ja, dtype=np.int32)
    a = np.ascontiguousarray(a, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(len(ia) - 1)
    lib().nmcpu_spmv(len(ia) - 1, _i(ia), _i(ja), _d(a), _d(x), _d(y))
    return y
