"""Per-phase clock64 breakdown of one k_slab launch on the bench workload (NM_SLAB_TRACE=1).  Diagnostic only."""
import ctypes as C
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NM_SLAB_TRACE"] = "1"
import bench  # noqa: E402


def main():
    import argparse
    p = argparse.ArgumentParser(); p.add_argument("--ntet", type=int, default=200000); p.add_argument("--porder", type=int, default=2)
    p.add_argument("--job", type=int, default=2); p.add_argument("--which", default="B,Ap")
    a = p.parse_args()
    import torch
    torch.cuda.set_device(0)
    from normalmodes_b200 import _lib, matvec as mvmod
    from normalmodes_b200._lib import check
    L = _lib.lib(); check(L.nm_init(0))
    mesh, model, fem = bench.build_workload(a, 0, 1)
    fem.assemble(a.job, model)
    for which in a.which.split(","):
        m = fem.matrix(which); n = m.siz(0)
        h = mvmod.parcsr_create(m)
        check(L.nm_parcsr_jacobi_scale(h, C.c_double(1.0 if which == "B" else -1.0), None))
        cheb = mvmod.chebiter_setup(0.25, 4.35, 6, h)
        z = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(-1, 1); y = torch.empty_like(z)
        for _ in range(3):
            check(L.nm_chebiter_solve_dev(cheb, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
        torch.cuda.synchronize()
        cap = 4096 * 256 * 8
        buf = np.zeros(cap, dtype=np.int64); grid = C.c_int(); cf = np.zeros(4097, dtype=np.int32)
        check(L.nm_chebiter_trace_dump(cheb, buf.ctypes.data_as(C.POINTER(C.c_longlong)), cap, C.byref(grid), cf.ctypes.data_as(C.POINTER(C.c_int))))
        g = grid.value
        T = buf[:g * 256 * 8].reshape(g, 256, 8)      # NM_SLAB_MAXDESC chunks per CTA
        nm = np.diff(cf[:g + 1])
        names = ["0 start->blob(it+1) arrived", "1->2 gather issue", "2->3 epi.load issue", "3->4 walk", "4->5 shuffles+epilogue",
                 "5->6 cp.async wait", "6->7 barrier", "7->next start"]
        rows = []
        for b in range(g):
            for it in range(nm[b]):
                t = T[b, it]
                nxt = T[b, it + 1, 0] if it + 1 < nm[b] else 0
                last = it + 1 >= nm[b]
                rows.append([0 if last else t[1] - t[0], 0 if last else t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6],
                             (nxt - t[7]) if nxt else 0, t[7] - t[0], it, int(last)])
        R = np.array(rows, dtype=float)
        mid = R[(R[:, 9] > 0) & (R[:, 10] == 0)]
        print("== %s: grid %d, chunks/CTA %.1f, kernel span %.0f cycles (min start -> max end)" % (
            which, g, nm.mean(), T[:, :, 7].max() - T[:, 0, 0][T[:, 0, 0] > 0].min()))
        for i, nme in enumerate(names):
            print("   %-28s mean %7.0f  median %7.0f  p90 %7.0f" % (nme, mid[:, i].mean(), np.median(mid[:, i]), np.percentile(mid[:, i], 90)))
        print("   %-28s mean %7.0f  median %7.0f" % ("iteration total", mid[:, 8].mean(), np.median(mid[:, 8])))
        first = R[R[:, 9] == 0]
        print("   first iteration total mean %.0f ; CTA busy span mean %.0f" % (first[:, 8].mean(), np.mean([T[b, nm[b] - 1, 7] - T[b, 0, 0] for b in range(g)])))
        L.nm_chebiter_free(cheb); L.nm_parcsr_free(h)


if __name__ == "__main__":
    main()
