"""Host-side mirror of `cg_create_matrix_mod` (src/mod_cg_create_matrix.f90:35-61): pattern first
(matrixstruct / matrixstruct_general, host C++), then values (CGE3D_ISO / CGFSE3D_ISO, CUDA)."""
import ctypes as C

import numpy as np

from ._lib import lib, check, dptr, iptr, f64, i32
from .matvec import COOmat

MAT_IDS = {"A": 0, "B": 1, "E": 2, "ET": 3, "Ap": 4}


class Fem:
    """Topology + numbering + patterns of one rank (the reference's unstrM / CGM globals)."""

    def __init__(self, mesh, vs, porder, nproc=1, part=None, rank=0):
        self.h = C.c_void_p()
        self.porder = porder
        self.ntet = int(mesh["ele"].shape[0]); self.nvert = int(mesh["node"].shape[0])
        self.pNp = 4 if porder == 1 else 10
        ele = i32(mesh["ele"]); neigh = i32(mesh["neigh"]); node = f64(mesh["node"]); vs = f64(vs)
        assert vs.shape == (self.ntet, self.pNp)
        p = iptr(i32(part)) if part is not None else None
        check(lib().nm_fem_create(self.ntet, self.nvert, iptr(ele), iptr(neigh), dptr(node), int(porder), dptr(vs),
                                  int(nproc), p, int(rank), C.byref(self.h)))
        nn = C.c_int(); N = C.c_int(); Np = C.c_int(); fc = C.c_int(); nle = C.c_int()
        check(lib().nm_fem_info(self.h, C.byref(nn), C.byref(N), C.byref(Np), C.byref(fc), C.byref(nle)))
        self.nn, self.N, self.Np, self.fluidcase, self.nle = nn.value, N.value, Np.value, fc.value, nle.value
        self.nproc, self.rank = nproc, rank

    def numbering(self):
        out = {k: np.empty(self.nn, dtype=np.int32) for k in ("vstat", "vnum", "pnum", "vstt", "pstt", "order")}
        check(lib().nm_fem_numbering(self.h, *[iptr(out[k]) for k in ("vstat", "vnum", "pnum", "vstt", "pstt", "order")]))
        return out

    def t2n(self):
        t = np.empty((self.ntet, self.pNp), dtype=np.int32)
        check(lib().nm_fem_t2n(self.h, iptr(t)))
        return t

    def matrix(self, name, values=True):
        """COOmat of this rank's rows (0-based global columns); name 'A' is Ad in the fluid case."""
        which = MAT_IDS[name if name != "Ad" else "A"]
        present = C.c_int(); nrow = C.c_int(); nnz = C.c_longlong()
        check(lib().nm_fem_matrix_sizes(self.h, which, C.byref(present), C.byref(nrow), C.byref(nnz)))
        if not present.value:
            return None
        rd = np.empty(self.nproc + 1, dtype=np.int32); cd = np.empty(self.nproc + 1, dtype=np.int32)
        ia = np.empty(nrow.value + 1, dtype=np.int32); ja = np.empty(nnz.value, dtype=np.int32)
        val = np.empty(nnz.value) if values else None
        check(lib().nm_fem_matrix_get(self.h, which, iptr(rd), iptr(cd), iptr(ia), iptr(ja),
                                      dptr(val) if values else None))
        return COOmat(rd, ia, ja, val if values else np.zeros(nnz.value), coldist=cd)

    def assemble(self, job, model):
        g0 = model.get("g0")
        vp, vs, rho = f64(model["vp"]), f64(model["vs"]), f64(model["rho"])
        g0 = f64(g0) if (job >= 2 and g0 is not None) else None
        check(lib().nm_fem_assemble_values(self.h, int(job), dptr(vp), dptr(vs), dptr(rho),
                                           dptr(g0) if g0 is not None else None))

    def free(self):
        if self.h:
            check(lib().nm_fem_free(self.h)); self.h = None


def cg_create_matrix(mesh, model, porder, job, nproc=1, part=None, rank=0):
    """Returns (CGM, fem): CGM = {'A','B'} or {'Ad','B','E','ET','Ap'} as COOmat, unscaled."""
    fem = Fem(mesh, model["vs"], porder, nproc, part, rank)
    fem.assemble(job, model)
    names = ("Ad", "B", "E", "ET", "Ap") if fem.fluidcase else ("A", "B")
    CGM = {k: fem.matrix(k) for k in names}
    return CGM, fem
