"""Live timings of the dense kernels of the Lanczos phase (K6 reorthogonalisation GEMVs, K7 Ritz GEMM and the Gram block
of the Rayleigh-Ritz refinement, both on fp64 tensor cores) at the bench size.  Diagnostic, and the ncu target:

    python tools/lanczos_kernels.py [--n 8364411] [--k 500] [--ns 128] [--out gpurun_out/lanczos_kernels.json]
    ncu --set full -k regex:"k_gemvT|k_gemvN|k_ritz_gemm|k_gram_dmma" -c 8 python tools/lanczos_kernels.py --k 200
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n", type=int, default=814323)
    p.add_argument("--k", type=int, nargs="+", default=[500, 2000])
    p.add_argument("--ns", type=int, default=128)
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "lanczos_kernels.json"))
    a = p.parse_args()
    import torch
    torch.cuda.set_device(0)
    from normalmodes_b200 import _lib
    L = _lib.lib()
    _lib.check(L.nm_init(0))
    peak = 6552.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = []
    for k in a.k:
        us = (C.c_double * 4)()
        _lib.check(L.nm_diag_lanczos_kernels(C.c_longlong(a.n), int(k), int(a.ns), us))
        r = dict(n=a.n, k=k, ns=a.ns, us_gemvT=us[0], us_gemvN=us[1], us_ritz_gemm=us[2], us_gram=us[3],
                 gemvT_gbs=8.0 * a.n * k / us[0] / 1e3, gemvN_gbs=8.0 * a.n * k / us[1] / 1e3,
                 ritz_tflops=2.0 * a.n * k * a.ns / us[2] / 1e6, ritz_gbs=8.0 * a.n * (k + a.ns) / us[2] / 1e3,
                 gram_tflops=2.0 * a.n * a.ns * a.ns / us[3] / 1e6, hbm_peak_gbs=peak)
        r["gemvT_frac_of_hbm"] = r["gemvT_gbs"] / peak; r["gemvN_frac_of_hbm"] = r["gemvN_gbs"] / peak
        res.append(r)
        print(json.dumps(r), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
