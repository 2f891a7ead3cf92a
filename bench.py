"""Benchmark of the hot path: one application y = p(A B^-1) z of the Chebyshev polynomial filter
(pEVSL ChebAv inside pEVSL_CHEBLANNR_F90, src/mod_pevsl.f90:122) on a builder-generated PREM-like mesh.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--ntet 200000] [--porder 2] [--solve]

A "step" is one filter application (degree m from find_pol for the band): m x { ChebIter B-solve
(degB fused SpMV steps) + fluid Schur term (ET, degAp fused SpMV steps on Ap~, E) + A product fused with
the three-term update }.  `value` = ALGORITHMIC bytes of one application (CSR, 8-byte values, 4-byte
indices) / device time, inputs resident in HBM; `e2e` = the same through the C ABI with HOST
vectors (H2D of z and D2H of y inside the timed region).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--ntet", type=int, default=200000)
    p.add_argument("--porder", type=int, default=2)
    p.add_argument("--job", type=int, default=2)
    p.add_argument("--lowfreq", type=float, default=0.1)
    p.add_argument("--upfreq", type=float, default=1.0)
    p.add_argument("--solve", action="store_true", help="also run the full eigen-solve (time-to-all-eigenpairs)")
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline sample")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=2, help="filter applications timed through the host-vector C ABI")
    return p.parse_args()


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------- algorithmic bytes
def spmv_bytes(nnz, nrow, ncol_touched):
    return 12 * nnz + 4 * (nrow + 1) + 8 * ncol_touched + 8 * nrow


def cheb_step_bytes(nnz, n):           # fused ChebIter step: SpMV on d + r,x read/write + d write
    return 12 * nnz + 4 * (n + 1) + 8 * n + 40 * n


def filter_step_bytes(nnz, n):         # fused filter step: SpMV on w + v, v-, y read + v+, y write
    return 12 * nnz + 4 * (n + 1) + 8 * n + 40 * n


def filter_bytes(sz, deg, degB, degAp):
    """sz: dict of GLOBAL nnz/rows.  Bytes of one y = p(A B^-1) z."""
    per = degB * cheb_step_bytes(sz["nnzB"], sz["N"]) + filter_step_bytes(sz["nnzA"], sz["N"])
    if sz["fluid"]:
        per += degAp * cheb_step_bytes(sz["nnzAp"], sz["Np"])
        per += spmv_bytes(sz["nnzET"], sz["Np"], sz["N"]) + spmv_bytes(sz["nnzE"], sz["N"], sz["Np"])
    return deg * per


# ---------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.rows = []; self.stop = False; self.t = None

    def _run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True); self.t.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ---------------------------------------------------------------------------- workload construction (product path)
def build_workload(a, rank, nranks):
    from normalmodes_b200 import meshgen, partition
    from normalmodes_b200.create_matrix import Fem
    t0 = time.time()
    mesh = meshgen.build_mesh(a.ntet, seed=0)
    model = meshgen.build_model(mesh, a.porder, gravity=a.job >= 2)
    log("mesh: %d tets, %d vertices (%.1fs)" % (mesh["ntet"], mesh["nvert"], time.time() - t0))
    part = None
    if nranks > 1:
        # topology with a trivial partition gives the node ids (P2 edge numbering depends on nproc only)
        nn_probe = Fem(mesh, model["vs"], a.porder, nproc=1)
        nn1 = nn_probe.nn
        nn_probe.free()
        f0 = Fem(mesh, model["vs"], a.porder, nproc=nranks, part=np.zeros(nn1, dtype=np.int32), rank=0)
        X = partition.node_coordinates(mesh, f0)
        f0.free()
        part = partition.rcb(X, nranks)
    t1 = time.time()
    fem = Fem(mesh, model["vs"], a.porder, nproc=nranks, part=part, rank=rank)
    log("topology + numbering + pattern: N=%d Np=%d (%.1fs)" % (fem.N, fem.Np, time.time() - t1))
    return mesh, model, fem


def gather_sizes(CGM, fem, nranks):
    sz = dict(N=fem.N, Np=fem.Np, fluid=bool(fem.fluidcase))
    loc = dict(nnzA=CGM["Ad" if fem.fluidcase else "A"].NNZ, nnzB=CGM["B"].NNZ)
    if fem.fluidcase:
        loc.update(nnzAp=CGM["Ap"].NNZ, nnzE=CGM["E"].NNZ, nnzET=CGM["ET"].NNZ)
    if nranks > 1:
        import torch
        import torch.distributed as dist
        keys = sorted(loc)
        t = torch.tensor([loc[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        loc = {k: int(v) for k, v in zip(keys, t.tolist())}
    sz.update(loc)
    return sz


# ---------------------------------------------------------------------------- CPU arm (oracle port; reference unbuildable here)
def cpu_ops_from(CGM, mv, fluid):
    """CpuOps over the SAME CSR arrays the GPU path uses (B~ / Ap~ values read back from the device)."""
    from oracle import cpu as ocpu
    from normalmodes_b200._lib import lib, check, dptr
    B = CGM["B"]
    Bv = np.empty(B.NNZ); check(lib().nm_parcsr_get_values(mv.sBV, dptr(Bv)))
    A = CGM["Ad" if fluid else "A"]
    kw = {}
    if fluid:
        Ap = CGM["Ap"]; Apv = np.empty(Ap.NNZ); check(lib().nm_parcsr_get_values(mv.sApV, dptr(Apv)))
        kw = dict(E=(CGM["E"].rowdist, CGM["E"].col, CGM["E"].val), ET=(CGM["ET"].rowdist, CGM["ET"].col, CGM["ET"].val),
                  Ap=(Ap.rowdist, Ap.col, Apv), dp=Ap.diag, boundsAp=mv.boundsAp, degAp=mv.degAp)
    return ocpu.CpuOps((B.rowdist, B.col, Bv), (A.rowdist, A.col, A.val), B.diag, mv.boundsB, mv.degB, **kw)


def time_cpu_sample(cops, pol, sz, degB, degAp, target_s, steps=1, warmup=0):
    """Times ChebAv truncated to kmax degree steps (a bounded sample of one filter application)."""
    from oracle import cpu as ocpu
    z = np.random.default_rng(12345).uniform(-1, 1, cops.n)
    t0 = time.time(); cops.chebav(pol.deg, pol.mu, pol.cc, pol.dd, z, kmax=1); t1 = time.time() - t0
    kmax = int(max(1, min(pol.deg, target_s / max(t1, 1e-6))))
    for _ in range(warmup):
        cops.chebav(pol.deg, pol.mu, pol.cc, pol.dd, z, kmax=kmax)
    ts = []
    for _ in range(steps):
        t0 = time.time(); cops.chebav(pol.deg, pol.mu, pol.cc, pol.dd, z, kmax=kmax); ts.append(time.time() - t0)
    tm = float(np.mean(ts))
    nbytes = filter_bytes(sz, kmax, degB, degAp)
    return dict(value=nbytes / tm / 1e9, unit="GB/s", cores=ocpu.threads(), kind="port",
                sample="%d of %d filter-degree steps of one application (%.1f s each run)" % (kmax, pol.deg, tm)), tm, kmax


# ---------------------------------------------------------------------------- main
def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0")); nranks = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and rank != 0:
        return
    import torch
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    from normalmodes_b200 import _lib, matvec as mvmod, pevsl
    from normalmodes_b200.create_matrix import MAT_IDS  # noqa: F401
    L = _lib.lib()
    _lib.check(L.nm_init(local))
    use_dist = nranks > 1 and a.impl == "ours"
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idbuf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            raw = C.create_string_buffer(128)
            _lib.check(L.nm_comm_unique_id(raw))
            idbuf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
        dist.broadcast(idbuf, 0)
        _lib.check(L.nm_comm_init(rank, nranks, bytes(idbuf.cpu().numpy().tobytes())))
    nr = nranks if use_dist else 1
    rk = rank if use_dist else 0
    mesh, model, fem = build_workload(a, rk, nr)
    t0 = time.time()
    fem.assemble(a.job, model)
    names = ("Ad", "B", "E", "ET", "Ap") if fem.fluidcase else ("A", "B")
    CGM = {k: fem.matrix(k) for k in names}
    t_asm = time.time() - t0
    sz = gather_sizes(CGM, fem, nr)
    log("assembly (device) + copy-out: %.1fs; global nnz A %d, B %d" % (t_asm, sz["nnzA"], sz["nnzB"]))
    t0 = time.time()
    mv = mvmod.setupmatvec(CGM, a.porder, rank=rk, nproc=nr, log=log)
    t_setup = time.time() - t0
    P = mvmod.Pevsl(mv.Gpbsiz, mv.pbsiz, mv.nfirst)
    P.setbmv_op(mv.opB); P.setbsol_chebiter(mv.chebB); P.setamv_op(mv.opA); P.set_geneig()
    t0 = time.time()
    LMIN, LMAX = P.lanbounds(3000, 5000, 1.0e-5)
    t_bounds = time.time() - t0
    lo, up = pevsl.freq_interval(a.lowfreq, a.upfreq, LMIN)
    xintv = np.array([lo, up, LMIN, LMAX])
    pol = pevsl.Pol(xintv, 0.8, 0.7)
    degAp = mv.degAp if fem.fluidcase else 0
    log("setupmatvec %.1fs, bounds of B^-1A [%.3e, %.3e] %.1fs, filter degree %d" % (t_setup, LMIN, LMAX, t_bounds, pol.deg))
    nbytes = filter_bytes(sz, pol.deg, mv.degB, degAp)
    workload = "PREM-like %d-tet mesh (builder-generated), JOB %d, pOrder %d, band %.2f-%.2f mHz, N=%d, nnz(A)=%d" % (
        mesh["ntet"], a.job, a.porder, a.lowfreq, a.upfreq, sz["N"], sz["nnzA"])
    config = dict(workload=workload, filter_degree=pol.deg, degB=mv.degB, degAp=degAp, ranks=nr,
                  l2="matrices + vectors of one application exceed L2 (CSR %.0f MB); no flush between steps" % (
                      (12 * (sz["nnzA"] + sz["nnzB"])) / 1e6))

    if a.impl == "reference":
        # The reference (Fortran + MPI + pEVSL + ParMETIS) cannot be built in this image: time the oracle's C/OpenMP
        # restatement of the same loops on the host cores, on the same matrices and polynomial.
        cops = cpu_ops_from(CGM, mv, bool(fem.fluidcase))
        cb, tm, kmax = time_cpu_sample(cops, pol, sz, mv.degB, degAp, a.cpu_seconds, steps=a.steps, warmup=min(a.warmup, 1))
        out = dict(metric="filtered_spmv_hbm_gbs", value=cb["value"], unit="GB/s", n_gpus=a.gpus, steps=a.steps,
                   warmup=a.warmup, ms_per_step=tm * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype="f64", data="synthetic", config=config, impl="reference", cpu_baseline=cb,
                   e2e=dict(value=cb["value"], unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(out), flush=True)
        return

    # ---- device-resident timing
    n = mv.pbsiz
    stream = torch.cuda.ExternalStream(L.nm_stream())
    z = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(-1, 1, generator=torch.Generator("cuda").manual_seed(12345 + rk))
    y = torch.empty_like(z); work = torch.empty(3 * max(n, 1), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step_dev():
        _lib.check(L.nm_pevsl_filter_dev(P.h, pol.h, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(work.data_ptr())))

    def barrier():
        if use_dist:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
    tw = time.time()
    for _ in range(a.warmup):
        step_dev()
    barrier()
    log("warm-up: %d filter applications, %.2f s each" % (a.warmup, (time.time() - tw) / max(a.warmup, 1)))
    l0 = L.nm_launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: the launch list covers the timed region
        e0.record(stream)
        for _ in range(a.steps):
            step_dev()
        e1.record(stream)
        barrier()
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = int(L.nm_launch_count() - l0)
    if use_dist:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    ms_per_step = ms / a.steps
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    log("device-resident: %.1f ms per application, %.0f GB/s algorithmic, %d launches" % (ms_per_step, value, launches))

    # ---- e2e: host vectors through the C ABI
    zh = torch.empty(n, dtype=torch.float64).uniform_(-1, 1).pin_memory(); yh = torch.empty(n, dtype=torch.float64).pin_memory()

    def step_host():
        _lib.check(L.nm_pevsl_filter_host(P.h, pol.h, C.c_void_p(zh.data_ptr()), C.c_void_p(yh.data_ptr())))
    step_host()
    barrier()
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    e0.record(stream)
    for _ in range(e2e_steps):
        step_host()
    e1.record(stream)
    barrier()
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    if use_dist:
        import torch.distributed as dist
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_e2e = float(t.item())
    e2e = dict(value=nbytes / (ms_e2e / e2e_steps * 1e-3) / 1e9, unit="GB/s", h2d_bytes_per_step=8 * n, d2h_bytes_per_step=8 * n,
               ms_per_step=ms_e2e / e2e_steps, steps=e2e_steps)
    log("e2e (host vectors through the C ABI): %.1f ms per application" % (ms_e2e / e2e_steps))

    # ---- roofline of the dominant kernel: the fused ChebIter step on B~ (degB launches per B-solve)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    for _ in range(3):
        _lib.check(L.nm_chebiter_solve_dev(mv.chebB, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
    torch.cuda.synchronize()
    reps = 20
    e0.record(stream)
    for _ in range(reps):
        _lib.check(L.nm_chebiter_solve_dev(mv.chebB, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr())))
    e1.record(stream)
    torch.cuda.synchronize()
    us_launch = e0.elapsed_time(e1) * 1e3 / (reps * mv.degB)
    infoB = mvmod.parcsr_info(mv.sBV)
    kind = C.c_int(); pbytes = C.c_longlong()
    _lib.check(L.nm_chebiter_pack_info(mv.chebB, C.byref(kind), C.byref(pbytes)))
    kname = ("k_spmv_kron3<EpiCheb>", "k_pack<KRON3,EpiCheb>", "k_sell<KRON3,EpiCheb>", "k_slab<3,256,EpiCheb>", "k_slabws<3,8,EpiCheb>")[kind.value]
    bytes_launch = cheb_step_bytes(infoB["nnz"], infoB["nrow"])
    fmt_bytes_launch = pbytes.value + 48 * infoB["nrow"]
    achieved = bytes_launch / (us_launch * 1e-6) / 1e9
    traffic = None                      # dram bytes per launch of this kernel from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kind.value == tj.get("kind") and nr == 1 and tj.get("n_rows") == infoB["nrow"]:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel="%s (fused ChebIter step on B~, %s; us_per_launch includes the 2 permute kernels of a solve spread over degB launches)" % (kname, infoB["format"]),
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                    us_per_launch=us_launch, algorithmic_bytes_per_launch=bytes_launch,
                    format_bytes_per_launch=fmt_bytes_launch, format_gbs=fmt_bytes_launch / (us_launch * 1e-6) / 1e9)
    comm = None
    if use_dist:
        mode = C.c_int(); ng = C.c_int(); ns = C.c_int()
        _lib.check(L.nm_parcsr_halo_info(mv.sBV, C.byref(mode), C.byref(ng), C.byref(ns)))
        for _ in range(5):
            _lib.check(L.nm_parcsr_halo_exchange_dev(mv.sBV, C.c_void_p(z.data_ptr())))
        barrier()
        e0.record(stream)
        for _ in range(200):
            _lib.check(L.nm_parcsr_halo_exchange_dev(mv.sBV, C.c_void_p(z.data_ptr())))
        e1.record(stream)
        barrier()
        comm = dict(halo=("none", "nccl send/recv", "nvlink peer window (direct stores + flags)")[mode.value],
                    us_per_exchange=e0.elapsed_time(e1) * 1e3 / 200, ghosts_rank0=ng.value, sends_rank0=ns.value,
                    exchanges_per_filter_degree=mv.degB + (degAp + 3 if fem.fluidcase else 1))
        log("halo exchange alone: %.1f us (%s), %d ghosts on rank 0" % (comm["us_per_exchange"], comm["halo"], ng.value))
    out = dict(metric="filtered_spmv_hbm_gbs", value=value, unit="GB/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
               ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", config=config, clocks=cs.summary(), e2e=e2e, gpu_launches=launches, roofline=roofline,
               setup_s=dict(assembly=t_asm, setupmatvec=t_setup, bounds=t_bounds))
    if comm:
        out["comm"] = comm
    if a.solve:
        t0 = time.time()
        r = pevsl.pnm_apply_pevsl(mv, a.lowfreq, a.upfreq, recheck=False)
        out["time_to_all_eigenpairs_s"] = time.time() - t0
        out["solve"] = dict(nev=int(r.nev), lanczos_steps=int(r.steps), t_bounds=r.t_bounds, t_cheblannr=r.t_cheblannr,
                            t_filter=r.t_filter, t_reorth=r.t_reorth, t_ritz=r.t_ritz,
                            max_res_over_lam=float((r.res2 / np.abs(r.eigval)).max()) if r.nev else None)
    if rank == 0 and nr == 1 and not a.no_cpu:
        cops = cpu_ops_from(CGM, mv, bool(fem.fluidcase))
        cb, tm, kmax = time_cpu_sample(cops, pol, sz, mv.degB, degAp, a.cpu_seconds)
        out["cpu_baseline"] = cb
    if rank == 0:
        print(json.dumps(out), flush=True)
    if use_dist:
        import torch.distributed as dist
        dist.barrier()
        _lib.check(L.nm_comm_finalize())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
