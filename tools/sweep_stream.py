"""Tuning sweep of the streaming SpMV (stage bytes x stages x CTAs/SM) on the bench workload's B~ (KRON3, fused
ChebIter step), Ad (ROW3) and Ap~ (CSR).  Diagnostic only.   python tools/sweep_stream.py [--ntet N] [--porder P]"""
import argparse
import ctypes as C
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--ntet", type=int, default=200000)
    p.add_argument("--porder", type=int, default=2)
    p.add_argument("--job", type=int, default=2)
    p.add_argument("--which", default="B,Ad,Ap")
    p.add_argument("--entries", default="1024,1536,2048")
    p.add_argument("--distinct", default="768")
    p.add_argument("--stages", default="2,3")
    p.add_argument("--ctas", default="3,4,5")
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_stream.json"))
    a = p.parse_args()
    import torch
    torch.cuda.set_device(0)
    from normalmodes_b200 import _lib, matvec as mvmod
    from normalmodes_b200._lib import check, dptr
    L = _lib.lib()
    check(L.nm_init(0))
    mesh, model, fem = bench.build_workload(a, 0, 1)
    fem.assemble(a.job, model)
    stream = torch.cuda.ExternalStream(L.nm_stream())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    res = []
    for which in a.which.split(","):
        if which not in ("B", "Ad", "A", "Ap", "E", "ET"):
            continue
        m = fem.matrix(which)
        nrow = m.siz(0); ncol = int(m.coldist[1] - m.coldist[0])
        z = torch.empty(max(nrow, ncol), dtype=torch.float64, device="cuda").uniform_(-1, 1)
        y = torch.empty_like(z)
        combos = [(None, None, None, None)] + list(itertools.product([int(v) for v in a.entries.split(",")],
                                                                     [int(v) for v in a.distinct.split(",")],
                                                                     [int(v) for v in a.stages.split(",")],
                                                                     [int(v) for v in a.ctas.split(",")]))
        for vb, dc, st, ct in combos:
            if vb is None:
                os.environ["NM_NO_PACK"] = "1"; os.environ["NM_NO_SELL"] = "1"
            else:
                os.environ["NM_NO_PACK"] = "0"; os.environ["NM_NO_SELL"] = "0"
                os.environ["NM_SELL_TARGET"] = str(dc)
                os.environ["NM_PACK_ENTRIES"] = str(vb // 6 if which in ("Ad", "A") else vb)
                os.environ["NM_PACK_DISTINCT"] = str(dc // 3 if which in ("Ad", "A") else dc)
                os.environ["NM_PACK_STAGES"] = str(st)
                os.environ["NM_PACK_CTAS_PER_SM"] = str(ct)
            h = mvmod.parcsr_create(m)
            info = mvmod.parcsr_info(h)
            cheb = None
            if which in ("B", "Ap"):
                check(L.nm_parcsr_jacobi_scale(h, C.c_double(1.0 if which == "B" else -1.0), None))
                deg = 20
                cheb = mvmod.chebiter_setup(0.25, 4.35, deg, h)
                fn = lambda: check(L.nm_chebiter_solve_dev(cheb, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
                per = deg; reps = 10
                nbytes = info["fmt_bytes"] + 48 * nrow
            else:
                fn = lambda: check(L.nm_parcsr_matvec_dev(h, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
                per = 1; reps = 50
                nbytes = info["fmt_bytes"] + 8 * nrow + 8 * ncol
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * per)
            r = dict(which=which, format=info["format"], entries=vb, distinct=dc, stages=st, ctas=ct, us=us, fmt_bytes=info["fmt_bytes"], format_gbs=nbytes / us / 1e3)
            res.append(r)
            bench.log(json.dumps(r))
            if cheb is not None:
                L.nm_chebiter_free(cheb)
            L.nm_parcsr_free(h)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
