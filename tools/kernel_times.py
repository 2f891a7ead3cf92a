"""Per-kernel device timings of the filter application on the bench workload (diagnostic, not a bench value).

    python tools/kernel_times.py [--ntet 200000] [--porder 2] [--out gpurun_out/kernel_times.json]

For each kernel class of one filter-degree step (fused ChebIter step on B~, on Ap~, ET / E products, the fused
A product) it reports microseconds per launch, algorithmic (CSR) GB/s and format GB/s, then the time of a
whole degree step and what the sum of the parts predicts.
"""
import argparse
import ctypes as C
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
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_times.json"))
    p.add_argument("--reps", type=int, default=20)
    p.add_argument("--ncu-range", action="store_true",
                   help="only run [cudaProfilerStart; one Ap~ solve; one B~ solve; cudaProfilerStop] after warm-up (for "
                        "ncu --profile-from-start off --kernel-name-base demangled -k regex:k_slabws --launch-skip/-c)")
    a = p.parse_args()
    import torch
    torch.cuda.set_device(0)
    from normalmodes_b200 import _lib, matvec as mvmod
    L = _lib.lib()
    _lib.check(L.nm_init(0))
    mesh, model, fem = bench.build_workload(a, 0, 1)
    fem.assemble(a.job, model)
    names = ("Ad", "B", "E", "ET", "Ap") if fem.fluidcase else ("A", "B")
    CGM = {k: fem.matrix(k) for k in names}
    mv = mvmod.setupmatvec(CGM, a.porder, log=bench.log)
    stream = torch.cuda.ExternalStream(L.nm_stream())
    n = mv.pbsiz
    g = torch.Generator("cuda").manual_seed(1)
    z = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(-1, 1, generator=g)
    y = torch.empty_like(z)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps, per=1):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * per)

    if a.ncu_range:
        solveB = lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebB, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
        zp = torch.empty(max(mv.Ap.siz(0), 1) if fem.fluidcase else 1, dtype=torch.float64, device="cuda").uniform_(-1, 1, generator=g)
        yp = torch.empty_like(zp)
        solveAp = lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebAp, C.c_void_p(zp.data_ptr()), C.c_void_p(yp.data_ptr())))
        for _ in range(2):
            if fem.fluidcase:
                solveAp()
            solveB()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        if fem.fluidcase:
            solveAp()
        solveB()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return

    res = {}

    def rec(name, us, info, algo_bytes, fmt_bytes):
        res[name] = dict(us_per_launch=us, algorithmic_gbs=algo_bytes / us / 1e3, format_gbs=fmt_bytes / us / 1e3,
                         algorithmic_bytes=algo_bytes, format_bytes=fmt_bytes, **info)
        bench.log("%-28s %8.2f us  algo %7.0f GB/s  format %7.0f GB/s  %s" % (
            name, us, algo_bytes / us / 1e3, fmt_bytes / us / 1e3, info))

    # fused ChebIter step on B~
    iB = mvmod.parcsr_info(mv.sBV)
    us = timeit(lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebB, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))),
                a.reps, mv.degB)
    rec("chebiter_step_B", us, iB, bench.cheb_step_bytes(iB["nnz"], iB["nrow"]), iB["fmt_bytes"] + 48 * iB["nrow"])
    us = timeit(lambda: _lib.check(L.nm_parcsr_matvec_dev(mv.sBV, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 100)
    rec("spmv_B", us, iB, bench.spmv_bytes(iB["nnz"], iB["nrow"], iB["ncol"]), iB["fmt_bytes"] + 16 * iB["nrow"])
    if fem.fluidcase:
        npz = mv.Ap.siz(0)
        zp = torch.empty(npz, dtype=torch.float64, device="cuda").uniform_(-1, 1, generator=g)
        yp = torch.empty_like(zp)
        iAp = mvmod.parcsr_info(mv.sApV)
        us = timeit(lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebAp, C.c_void_p(zp.data_ptr()), C.c_void_p(yp.data_ptr()))),
                    a.reps, mv.degAp)
        rec("chebiter_step_Ap", us, iAp, bench.cheb_step_bytes(iAp["nnz"], iAp["nrow"]), iAp["fmt_bytes"] + 48 * iAp["nrow"])
        iE = mvmod.parcsr_info(mv.sEV); iET = mvmod.parcsr_info(mv.sETV)
        us = timeit(lambda: _lib.check(L.nm_parcsr_matvec_dev(mv.sETV, C.c_void_p(z.data_ptr()), C.c_void_p(yp.data_ptr()))), 100)
        rec("spmv_ET", us, iET, bench.spmv_bytes(iET["nnz"], iET["nrow"], iET["ncol"]), iET["fmt_bytes"] + 8 * iET["nrow"] + 8 * iET["ncol"])
        us = timeit(lambda: _lib.check(L.nm_parcsr_matvec_dev(mv.sEV, C.c_void_p(zp.data_ptr()), C.c_void_p(y.data_ptr()))), 100)
        rec("spmv_E", us, iE, bench.spmv_bytes(iE["nnz"], iE["nrow"], iE["ncol"]), iE["fmt_bytes"] + 8 * iE["nrow"] + 8 * iE["ncol"])
        iA = mvmod.parcsr_info(mv.sAdV)
        us = timeit(lambda: _lib.check(L.nm_parcsr_matvec_dev(mv.sAdV, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 50)
        rec("spmv_Ad", us, iA, bench.spmv_bytes(iA["nnz"], iA["nrow"], iA["ncol"]), iA["fmt_bytes"] + 16 * iA["nrow"])
    else:
        iA = mvmod.parcsr_info(mv.sAV)
        us = timeit(lambda: _lib.check(L.nm_parcsr_matvec_dev(mv.sAV, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 50)
        rec("spmv_A", us, iA, bench.spmv_bytes(iA["nnz"], iA["nrow"], iA["ncol"]), iA["fmt_bytes"] + 16 * iA["nrow"])
    us = timeit(lambda: _lib.check(L.nm_op_apply_dev(mv.opA, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 10)
    res["op_A_apply_us"] = us
    bench.log("operator A apply (incl. fluid Schur term): %.1f us" % us)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
