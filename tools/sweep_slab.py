"""Tuning sweep of the fused ChebIter step kernels on the bench workload's B~ (KRON3) and Ap~ (CSR): k_slab
configurations (threads, stages, split, caps, persistent variant).  Diagnostic only, not a bench value.
    python tools/sweep_slab.py [--ntet N] [--which B,Ap] [--out gpurun_out/sweep_slab.json]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = ("NM_CHEB_KERNEL", "NM_SLAB_THREADS", "NM_SLAB_STAGES", "NM_SLAB_SPLIT", "NM_SLAB_ENTRIES", "NM_SLAB_DISTINCT",
        "NM_SLAB_CTAS_PER_SM", "NM_PACK_BANK_AWARE", "NM_PACK_ORDER", "NM_SLAB_WS", "NM_SLAB_PRODUCERS", "NM_SLAB_XS", "NM_SLAB_PDL",
        "NM_SLAB_PERS", "NM_SLAB_FLOW")

CONFIGS = [
    dict(),
    dict(NM_SLAB_SPLIT="16"),
    dict(NM_SLAB_PRODUCERS="6"),
    dict(NM_SLAB_XS="3"),
    dict(NM_SLAB_ENTRIES="2560", NM_SLAB_DISTINCT="480", NM_SLAB_STAGES="3", NM_SLAB_XS="3"),
    dict(NM_SLAB_THREADS="512", NM_SLAB_SPLIT="12"),
    dict(NM_PACK_BANK_AWARE="1"),
    dict(NM_SLAB_SPLIT="8"),
    dict(NM_SLAB_PRODUCERS="2"),
    dict(NM_SLAB_THREADS="128", NM_SLAB_SPLIT="12", NM_SLAB_ENTRIES="1792", NM_SLAB_DISTINCT="380"),
    dict(NM_SLAB_PERS="1", NM_SLAB_FLOW="0"),
]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--ntet", type=int, default=200000)
    p.add_argument("--porder", type=int, default=2)
    p.add_argument("--job", type=int, default=2)
    p.add_argument("--which", default="B,Ap")
    p.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_slab.json"))
    a = p.parse_args()
    import torch
    torch.cuda.set_device(0)
    from normalmodes_b200 import _lib, matvec as mvmod
    from normalmodes_b200._lib import check
    L = _lib.lib()
    check(L.nm_init(0))
    mesh, model, fem = bench.build_workload(a, 0, 1)
    fem.assemble(a.job, model)
    stream = torch.cuda.ExternalStream(L.nm_stream())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    res = []
    for which in a.which.split(","):
        m = fem.matrix(which)
        nrow = m.siz(0)
        z = torch.empty(nrow, dtype=torch.float64, device="cuda").uniform_(-1, 1)
        y = torch.empty_like(z)
        h = mvmod.parcsr_create(m)
        check(L.nm_parcsr_jacobi_scale(h, C.c_double(1.0 if which == "B" else -1.0), None))
        yref = None
        for cfg in CONFIGS:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(cfg)
            t0 = time.time()
            deg = 12
            cheb = mvmod.chebiter_setup(0.25, 4.35, deg, h)
            tb = time.time() - t0
            kind = C.c_int(); nb = C.c_longlong()
            check(L.nm_chebiter_pack_info(cheb, C.byref(kind), C.byref(nb)))
            fn = lambda: check(L.nm_chebiter_solve_dev(cheb, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            reps = 4
            e0.record(stream)
            for _ in range(reps):
                fn()
            e1.record(stream)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * deg)
            if yref is None:
                yref = y.clone()
            err = float((y - yref).abs().max() / yref.abs().max())
            r = dict(which=which, cfg=cfg, kind=kind.value, us_per_step=us, matrix_bytes=nb.value,
                     format_gbs=(nb.value + 48 * nrow) / us / 1e3, build_s=tb, rel_diff_vs_first=err)
            res.append(r)
            bench.log(json.dumps(r))
            L.nm_chebiter_free(cheb)
        L.nm_parcsr_free(h)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
