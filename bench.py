"""Benchmark of the hot path: the Chebyshev polynomial filter y = p(A B^-1) z (pEVSL ChebAv inside
pEVSL_CHEBLANNR_F90, src/mod_pevsl.f90:122) on a builder-generated PREM-like mesh.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--ntet 2000000] [--porder 2] [--degree-steps 128]

Workload (BASELINE.json config 3): PREM-like ~2 M-tet mesh, JOB 2, pOrder 2, band 0.1-1.0 mHz.  One filter application
on it is ~9000 degree steps of { ChebIter B-solve (degB fused SpMV steps) + fluid Schur term (ET, degAp fused SpMV steps
on Ap~, E) + A product fused with the three-term update } = minutes, so a "step" is a FIXED SLICE of one application:
the first --degree-steps degree steps (every degree step costs the same).  `value` = ALGORITHMIC bytes of that slice
(CSR, 8-byte values, 4-byte indices, SURVEY.md 8d: what the reference's CSR path moves) / device time, inputs resident
in HBM; `e2e` = the same through the C ABI with HOST vectors (H2D of z and D2H of y inside the timed region);
`roofline` = the dominant kernel (fused ChebIter step on B~) on the bytes it actually streams in ITS format, against the
measured HBM bandwidth; `application` = the same accounting for the whole degree step; `check` = the first degree
steps of the same application, GPU against the oracle's C port on the host, on the bench matrices (all ranks gathered);
`cpu_baseline` / `--impl reference` = the oracle's C/OpenMP port of the same loops on the box's host cores (the
Fortran + MPI + pEVSL reference cannot be built in this image).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--ntet", type=int, default=2000000)
    p.add_argument("--porder", type=int, default=2)
    p.add_argument("--job", type=int, default=2)
    p.add_argument("--lowfreq", type=float, default=0.1)
    p.add_argument("--upfreq", type=float, default=1.0)
    p.add_argument("--degree-steps", type=int, default=128,
                   help="ChebAv degree steps per bench step (0: a whole filter application)")
    p.add_argument("--check-steps", type=int, default=2, help="degree steps compared with the CPU oracle (0: no check)")
    p.add_argument("--solve", action="store_true", help="also run the full eigen-solve (time-to-all-eigenpairs)")
    p.add_argument("--cpu-seconds", type=float, default=6.0, help="target CPU work of one cpu_baseline / reference sample")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=2, help="bench steps timed through the host-vector C ABI")
    return p.parse_args(argv)


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
    """sz: dict of GLOBAL nnz/rows.  Algorithmic bytes of `deg` degree steps of y = p(A B^-1) z."""
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
def p2_node_count(mesh):
    """nvert + number of distinct mesh edges (the P2 node count; src/mod_geometry.f90:428-689 adds one node per edge)."""
    e = mesh["ele"]
    pairs = np.concatenate([e[:, [i, j]] for i, j in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))])
    pairs.sort(axis=1)
    key = pairs[:, 0].astype(np.int64) * (int(mesh["nvert"]) + 1) + pairs[:, 1]
    return int(mesh["nvert"]) + int(np.unique(key).size)


def build_workload(a, rank, nranks):
    from normalmodes_b200 import meshgen, partition
    from normalmodes_b200.create_matrix import Fem
    t0 = time.time()
    mesh = meshgen.build_mesh(a.ntet, seed=0)
    model = meshgen.build_model(mesh, a.porder, gravity=a.job >= 2)
    log("mesh: %d tets, %d vertices (%.1fs)" % (mesh["ntet"], mesh["nvert"], time.time() - t0))
    part = None
    if nranks > 1:
        # topology with a trivial partition gives the node ids (the P2 edge numbering depends on nproc only)
        nn1 = mesh["nvert"] if a.porder == 1 else p2_node_count(mesh)
        f0 = Fem(mesh, model["vs"], a.porder, nproc=nranks, part=np.zeros(nn1, dtype=np.int32), rank=0)
        assert f0.nn == nn1, (f0.nn, nn1)
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
def local_cpu_arrays(CGM, mv, fluid):
    """This rank's rows of the matrices the CPU arm multiplies by: B~ / Ap~ values read back from the device (the
    Jacobi scaling ran there), Ad / E / ET unscaled with the scalings d, dp -- exactly the reference's operands."""
    from normalmodes_b200._lib import lib, check, dptr
    B = CGM["B"]
    Bv = np.empty(B.NNZ); check(lib().nm_parcsr_get_values(mv.sBV, dptr(Bv)))
    A = CGM["Ad" if fluid else "A"]
    out = dict(B_ia=B.rowdist, B_ja=B.col, B_a=Bv, A_ia=A.rowdist, A_ja=A.col, A_a=A.val, d=B.diag)
    if fluid:
        Ap = CGM["Ap"]; Apv = np.empty(Ap.NNZ); check(lib().nm_parcsr_get_values(mv.sApV, dptr(Apv)))
        out.update(E_ia=CGM["E"].rowdist, E_ja=CGM["E"].col, E_a=CGM["E"].val,
                   ET_ia=CGM["ET"].rowdist, ET_ja=CGM["ET"].col, ET_a=CGM["ET"].val,
                   Ap_ia=Ap.rowdist, Ap_ja=Ap.col, Ap_a=Apv, dp=Ap.diag)
    return out


def cpu_ops_from_arrays(parts, mv, fluid):
    """CpuOps over the GLOBAL matrices: the ranks' row blocks (global 0-based column ids already) stacked in rank order."""
    from oracle import cpu as ocpu

    def stack(name):
        ia = [np.asarray(p[name + "_ia"], dtype=np.int64) for p in parts]
        off = np.cumsum([0] + [int(x[-1]) for x in ia])
        gia = np.concatenate([ia[0]] + [x[1:] + off[i] for i, x in enumerate(ia) if i > 0]) if len(ia) > 1 else ia[0]
        assert gia[-1] < 2 ** 31, "CPU oracle uses 32-bit row pointers"
        return (gia.astype(np.int32), np.concatenate([p[name + "_ja"] for p in parts]) if len(parts) > 1 else parts[0][name + "_ja"],
                np.concatenate([p[name + "_a"] for p in parts]) if len(parts) > 1 else parts[0][name + "_a"])
    d = np.concatenate([p["d"] for p in parts])
    kw = {}
    if fluid:
        kw = dict(E=stack("E"), ET=stack("ET"), Ap=stack("Ap"), dp=np.concatenate([p["dp"] for p in parts]),
                  boundsAp=mv.boundsAp, degAp=mv.degAp)
    return ocpu.CpuOps(stack("B"), stack("A"), d, mv.boundsB, mv.degB, **kw)


def gather_cpu_ops(CGM, mv, fluid, rank, nranks):
    """rank 0: CpuOps of the global problem.  Other ranks hand their blocks over through a node-local scratch
    directory (one box: /dev/shm), not through the communicator -- these are gigabytes of test data."""
    loc = local_cpu_arrays(CGM, mv, fluid)
    if nranks == 1:
        return cpu_ops_from_arrays([loc], mv, fluid)
    import torch.distributed as dist
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    tag = [None]
    if rank == 0:
        tag[0] = tempfile.mkdtemp(prefix="nm_bench_", dir=base)
    dist.broadcast_object_list(tag, src=0)
    d = tag[0]
    if rank != 0:
        for k, v in loc.items():
            np.save(os.path.join(d, "r%d_%s.npy" % (rank, k)), np.ascontiguousarray(v))
    dist.barrier()
    ops = None
    if rank == 0:
        parts = [loc] + [{k: np.load(os.path.join(d, "r%d_%s.npy" % (r, k))) for k in loc} for r in range(1, nranks)]
        ops = cpu_ops_from_arrays(parts, mv, fluid)
        shutil.rmtree(d, ignore_errors=True)
    return ops


def gather_vector(v_local, rank, nranks):
    """Host copy of a distributed device vector on rank 0 (rank order = global row order)."""
    h = v_local.detach().cpu().numpy()
    if nranks == 1:
        return h
    import torch.distributed as dist
    out = [None] * nranks if rank == 0 else None
    dist.gather_object(h, out, dst=0)
    return np.concatenate(out) if rank == 0 else None


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
    t_start = time.time()
    mesh, model, fem = build_workload(a, rk, nr)
    t0 = time.time()
    fem.assemble(a.job, model)
    names = ("Ad", "B", "E", "ET", "Ap") if fem.fluidcase else ("A", "B")
    CGM = {k: fem.matrix(k) for k in names}
    t_asm = time.time() - t0
    del model
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
    D = pol.deg if a.degree_steps <= 0 else min(a.degree_steps, pol.deg)
    log("setupmatvec %.1fs, bounds of B^-1A [%.3e, %.3e] %.1fs, filter degree %d, bench step = %d degree steps" % (
        t_setup, LMIN, LMAX, t_bounds, pol.deg, D))
    nbytes = filter_bytes(sz, D, mv.degB, degAp)
    # `config` names the workload only (identical in both arms and for every N); measured / derived sizes go to `detail`
    config = dict(workload="PREM-like %d-tet mesh (builder-generated, target %d), JOB %d, pOrder %d, band %.2f-%.2f mHz" % (
                      mesh["ntet"], a.ntet, a.job, a.porder, a.lowfreq, a.upfreq),
                  step="%s ChebAv degree steps of one filter application y = p(A B^-1) z; per degree step: %d fused ChebIter "
                       "steps on B~, %d on Ap~, the ET / E products and the A product fused with the three-term update" % (
                           "the first %d" % a.degree_steps if a.degree_steps > 0 else "all", mv.degB, degAp),
                  degB=mv.degB, degAp=degAp,
                  l2="every matrix of the degree step exceeds L2 at N=1; inputs are not flushed between steps (one step streams "
                     "%.0f GB in the reference's CSR layout)" % (nbytes / 1e9))
    detail = dict(N=sz["N"], Np=sz["Np"], nnzA=sz["nnzA"], nnzB=sz["nnzB"], nnzAp=sz.get("nnzAp", 0), nnzE=sz.get("nnzE", 0),
                  filter_degree=pol.deg, degree_steps_per_step=D, ranks=nr, algorithmic_bytes_per_step=nbytes,
                  spectrum=[LMIN, LMAX], interval=[lo, up])

    if a.impl == "reference":
        # The reference (Fortran + MPI + pEVSL + ParMETIS) cannot be built in this image: time the oracle's C/OpenMP
        # restatement of the same loops on the host cores, on the same matrices and polynomial.
        cops = gather_cpu_ops(CGM, mv, bool(fem.fluidcase), 0, 1)
        cb, tm, kmax = time_cpu_sample(cops, pol, sz, mv.degB, degAp, a.cpu_seconds, steps=a.steps, warmup=min(a.warmup, 1))
        out = dict(metric="filtered_spmv_hbm_gbs", value=cb["value"], unit="GB/s", n_gpus=a.gpus, steps=a.steps,
                   warmup=a.warmup, ms_per_step=tm * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype="f64", data="synthetic", config=config, detail=detail, impl="reference", cpu_baseline=cb,
                   e2e=dict(value=cb["value"], unit="GB/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(out), flush=True)
        return

    # ---- device-resident timing
    n = mv.pbsiz
    stream = torch.cuda.ExternalStream(L.nm_stream())
    z = torch.empty(n, dtype=torch.float64, device="cuda").uniform_(-1, 1, generator=torch.Generator("cuda").manual_seed(12345 + rk))
    y = torch.empty_like(z); work = torch.empty(3 * max(n, 1), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step_dev(kmax=D):
        _lib.check(L.nm_pevsl_filter_steps_dev(P.h, pol.h, int(kmax), C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()),
                                               C.c_void_p(work.data_ptr())))

    def barrier():
        if use_dist:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if not use_dist:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- check: the first degree steps of this very application, GPU against the CPU oracle, all ranks gathered
    check = None
    cops = None
    if a.check_steps > 0 and not a.no_cpu:
        t0 = time.time()
        K = min(a.check_steps, pol.deg)
        step_dev(K)
        barrier()
        yg = gather_vector(y, rk, nr); zg = gather_vector(z, rk, nr)
        cops = gather_cpu_ops(CGM, mv, bool(fem.fluidcase), rk, nr)
        if rank == 0:
            yc = cops.chebav(pol.deg, pol.mu, pol.cc, pol.dd, zg, kmax=K)
            err = float(np.abs(yg - yc).max() / np.abs(yc).max())
            check = dict(max_rel_err=err, degree_steps=K, tol=1e-10, ok=bool(err <= 1e-10),
                         against="oracle/c (C/OpenMP port of pEVSL ChebAv + ChebIter + sparsefsAV) on the host, same matrices and z, "
                                 "%d rank(s) gathered" % nr)
            log("check: %d degree steps, max |y_gpu - y_cpu| / max |y_cpu| = %.2e (%.1fs)" % (K, err, time.time() - t0))
            assert err <= 1e-8, "bench check failed: GPU and CPU filter differ by %.2e" % err
        barrier()
    tw = time.time()
    for _ in range(a.warmup):
        step_dev()
    barrier()
    log("warm-up: %d steps, %.2f s each" % (a.warmup, (time.time() - tw) / max(a.warmup, 1)))
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
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = int(L.nm_launch_count() - l0)
    ms_per_step = ms / a.steps
    value = nbytes / (ms_per_step * 1e-3) / 1e9
    log("device-resident: %.1f ms per step (%.3f ms per degree step), %.0f GB/s algorithmic, %d launches" % (
        ms_per_step, ms_per_step / D, value, launches))

    # ---- e2e: host vectors through the C ABI
    zh = torch.empty(n, dtype=torch.float64).uniform_(-1, 1).pin_memory(); yh = torch.empty(n, dtype=torch.float64).pin_memory()

    def step_host():
        _lib.check(L.nm_pevsl_filter_steps_host(P.h, pol.h, int(D), C.c_void_p(zh.data_ptr()), C.c_void_p(yh.data_ptr())))
    step_host()
    barrier()
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    e0.record(stream)
    for _ in range(e2e_steps):
        step_host()
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
    e2e = dict(value=nbytes / (ms_e2e / e2e_steps * 1e-3) / 1e9, unit="GB/s", h2d_bytes_per_step=8 * n, d2h_bytes_per_step=8 * n,
               ms_per_step=ms_e2e / e2e_steps, steps=e2e_steps)
    log("e2e (host vectors through the C ABI): %.1f ms per step" % (ms_e2e / e2e_steps))

    # ---- per-kernel times of one degree step (live, CUDA events, back to back) and the bytes each must stream
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    def timeit(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1)) * 1e3 / reps

    def pack_info(cheb):
        kind = C.c_int(); pbytes = C.c_longlong()
        _lib.check(L.nm_chebiter_pack_info(cheb, C.byref(kind), C.byref(pbytes)))
        return kind.value, pbytes.value
    knames = ("k_spmv_kron3<EpiCheb>", "(retired)", "(retired)", "k_slab<3,256,EpiCheb>",
              "k_slabws<3,8,EpiCheb>", "k_slabpers<3,8>")
    us_solveB = timeit(lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebB, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 10)
    us_launch = us_solveB / mv.degB
    infoB = mvmod.parcsr_info(mv.sBV)
    kindB, pbytesB = pack_info(mv.chebB)
    algo_launch = cheb_step_bytes(infoB["nnz"], infoB["nrow"])
    fmt_launch = pbytesB + 48 * infoB["nrow"]              # slab + r, d, x read and written once each
    parts = {"B~ ChebIter step": dict(us=us_launch, per_degree=mv.degB, format_bytes=fmt_launch, algorithmic_bytes=algo_launch)}
    if fem.fluidcase:
        npz = mv.Ap.siz(rk)
        zp = torch.empty(max(npz, 1), dtype=torch.float64, device="cuda").uniform_(-1, 1)
        yp = torch.empty_like(zp)
        us = timeit(lambda: _lib.check(L.nm_chebiter_solve_dev(mv.chebAp, C.c_void_p(zp.data_ptr()), C.c_void_p(yp.data_ptr()))), 10) / degAp
        iAp = mvmod.parcsr_info(mv.sApV); _, pbAp = pack_info(mv.chebAp)
        parts["Ap~ ChebIter step"] = dict(us=us, per_degree=degAp, format_bytes=pbAp + 48 * iAp["nrow"],
                                          algorithmic_bytes=cheb_step_bytes(iAp["nnz"], iAp["nrow"]))
        us_op = timeit(lambda: _lib.check(L.nm_op_apply_dev(mv.opA, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 5)
        iA = mvmod.parcsr_info(mv.sAdV); iE = mvmod.parcsr_info(mv.sEV); iET = mvmod.parcsr_info(mv.sETV)
        rest_fmt = (iA["fmt_bytes"] + 56 * iA["nrow"]) + (iE["fmt_bytes"] + 8 * iE["nrow"] + 8 * iE["ncol"]) + \
                   (iET["fmt_bytes"] + 8 * iET["nrow"] + 8 * iET["ncol"])
        rest_algo = filter_step_bytes(iA["nnz"], iA["nrow"]) + spmv_bytes(iE["nnz"], iE["nrow"], iE["ncol"]) + \
            spmv_bytes(iET["nnz"], iET["nrow"], iET["ncol"])
        parts["Ad + ET + E products"] = dict(us=max(us_op - us * degAp, 0.0), per_degree=1, format_bytes=rest_fmt,
                                             algorithmic_bytes=rest_algo)
    else:
        us_op = timeit(lambda: _lib.check(L.nm_op_apply_dev(mv.opA, C.c_void_p(z.data_ptr()), C.c_void_p(y.data_ptr()))), 10)
        iA = mvmod.parcsr_info(mv.sAV)
        parts["A product"] = dict(us=us_op, per_degree=1, format_bytes=iA["fmt_bytes"] + 56 * iA["nrow"],
                                  algorithmic_bytes=filter_step_bytes(iA["nnz"], iA["nrow"]))
    for v in parts.values():
        v["format_gbs"] = v["format_bytes"] / v["us"] / 1e3 if v["us"] > 0 else None
        v["frac_of_hbm"] = v["format_gbs"] / peak if v["us"] > 0 else None
    fmt_degree = sum(v["format_bytes"] * v["per_degree"] for v in parts.values())     # this rank's bytes per degree step
    us_degree = ms_per_step * 1e3 / D
    application = dict(us_per_degree_step=us_degree, format_bytes_per_degree_step_per_gpu=fmt_degree,
                       format_gbs_per_gpu=fmt_degree / us_degree / 1e3, frac_of_hbm_per_gpu=fmt_degree / us_degree / 1e3 / peak,
                       algorithmic_gbs=value, kernels=parts,
                       note="format bytes = what each kernel must stream in ITS storage format (slab / ROW3 / CSR bytes + every "
                            "vector it reads or writes once), no cache credit; value and e2e count the reference's CSR bytes")
    traffic = None                      # dram bytes per step of this kernel from the committed ncu --set full capture
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if kindB == tj.get("kind") and tj.get("n_rows") == infoB["nrow"]:
            traffic = tj["traffic_bytes_per_launch"]
    except Exception:
        pass
    achieved = fmt_launch / (us_launch * 1e-6) / 1e9
    roofline = dict(bound="hbm", kernel="%s (fused ChebIter step on B~, %s; one launch = %d steps; us_per_launch is per STEP and "
                                        "includes the 2 permute kernels of a solve)" % (knames[kindB], infoB["format"], mv.degB if kindB == 5 else 1),
                    achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                    us_per_launch=us_launch, bytes_per_launch=fmt_launch,
                    bytes_note="bytes the kernel streams per step in its own format (slab + 6 vector passes); the CSR-equivalent "
                               "(SURVEY 8d) figure is algorithmic_*",
                    algorithmic_bytes_per_launch=algo_launch, algorithmic_gbs=algo_launch / (us_launch * 1e-6) / 1e9)
    assert roofline["frac"] <= 1.05, "roofline fraction above 1: byte accounting is wrong"
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
        comm = dict(halo_products=("none", "nccl send/recv", "nvlink peer window (direct stores + flags)")[mode.value],
                    halo_chebiter="in-kernel: boundary rows stored into the peers' flag-in-data slots from the step's epilogue"
                    if kindB == 5 else "per step", us_per_standalone_exchange=e0.elapsed_time(e1) * 1e3 / 200,
                    ghosts_rank0=ng.value, sends_rank0=ns.value, standalone_exchanges_per_degree_step=3 if fem.fluidcase else 1,
                    us_per_chebiter_step_B=us_launch)
        log("halo exchange alone: %.1f us (%s), %d ghosts on rank 0" % (comm["us_per_standalone_exchange"], comm["halo_products"], ng.value))
    out = dict(metric="filtered_spmv_hbm_gbs", value=value, unit="GB/s", n_gpus=a.gpus, steps=a.steps, warmup=a.warmup,
               ms_per_step=ms_per_step, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64",
               data="synthetic", config=config, detail=detail, clocks=cs.summary(), e2e=e2e, gpu_launches=launches,
               roofline=roofline, application=application, check=check,
               value_note="CSR-equivalent (algorithmic) bytes of SURVEY 8d / time: the unit the CPU arm is measured in; it exceeds "
                          "the HBM bandwidth because B = M (x) I3 is stored in 0.3 of its CSR bytes -- roofline / application "
                          "carry the format-byte figures",
               setup_s=dict(total=time.time() - t_start, assembly=t_asm, setupmatvec=t_setup, bounds=t_bounds))
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
        if cops is None:
            cops = gather_cpu_ops(CGM, mv, bool(fem.fluidcase), 0, 1)
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
