"""Generate the committed parity fixtures from the reference's shipped demo inputs.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes, next to this file,
  inputs_<case>.npz   the demo mesh + model arrays the case uses (compressed copy of the reference's
                      demos/models/input/* binaries in the App. A layout, 0-based)
  golden.json         per case: sizes, sha256 of every CSR index array, value checksums, the
                      eigenvalues in the band from the INDEPENDENT dense / shift-invert solve
                      (oracle.solver.truth_eigs) and the oracle's filtered-Lanczos summary.
The reference ships no golden outputs of its own: these pin the oracle.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import fem, solver  # noqa: E402

DEMOS = "/root/reference/demos/models/input/"
CASES = {
    # name: (dir, basename, porder, job, lowfreq, upfreq, run_filtered_lanczos)
    "const3k_p1_j1": ("CONST3k", "CONST_1L_3k.1", 1, 1, 0.2, 2.0, True),
    "const3k_p1_j2": ("CONST3k", "CONST_1L_3k.1", 1, 2, 0.2, 1.0, True),
    "const3k_p2_j1": ("CONST3k", "CONST_1L_3k.1", 2, 1, 0.2, 0.6, False),
    "prem3k_p1_j2": ("PREM3k", "prem_3L_3k.1", 1, 2, 0.1, 1.0, True),
    # False: truth eigenvalues from the independent shift-invert solve, no oracle filtered-Lanczos run (too slow on CPU)
    "prem3k_p2_j2": ("PREM3k", "prem_3L_3k.1", 2, 2, 0.1, 0.5, False),
    # the only >= 100 k reference mesh (P1 files only, no gravity file => JOB 1); demos/models/output/Mtopo100k logs a 283 s run
    "mtopo100k_p1_j1": ("Mtopo100k", "Mtopo_6L_100k.1", 1, 1, 0.5, 1.6, False),
    # pattern + values only: the band holds a dense cluster of fluid modes, the shift-invert truth does not finish in an hour
    "rtmdwak8k_p1_j2": ("RTMDWAK8k", "RTMDWAK_3L_8k.1", 1, 2, 0.1, 0.8, None),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    gpath = os.path.join(HERE, "golden.json")
    golden = json.load(open(gpath)) if os.path.exists(gpath) and "--force" not in sys.argv else {}
    for name, (d, base, po, job, lo, up, run) in CASES.items():
        if name in golden and (run is None or "truth_eigs" in golden[name]):
            continue
        t0 = time.time()
        mesh = fem.read_mesh(DEMOS + d + "/", base)
        model = fem.read_model(DEMOS + d + "/", base, po, job, mesh["ntet"])
        arrs = dict(ele=mesh["ele"].astype(np.int32), neigh=mesh["neigh"].astype(np.int32), node=mesh["node"],
                    vp=model["vp"], vs=model["vs"], rho=model["rho"])
        if model["g0"] is not None:
            arrs["g0"] = model["g0"]
        np.savez_compressed(os.path.join(HERE, "inputs_%s.npz" % name), **arrs)
        mats, topo, num, geo = fem.assemble(mesh, model, po, job)
        g = dict(dir=d, basename=base, porder=po, job=job, lowfreq=lo, upfreq=up, N=num["N"], Np=num["Np"],
                 ntet=mesh["ntet"], nvert=mesh["nvert"], matrices={})
        for k, m in mats.items():
            g["matrices"][k] = dict(shape=list(m["shape"]), nnz=int(m["ja"].size),
                                    ia_sha256=sha(m["ia"].astype(np.int32)), ja_sha256=sha(m["ja"].astype(np.int32)),
                                    sum=float(m["a"].sum()), abssum=float(np.abs(m["a"]).sum()),
                                    sample=[float(x) for x in m["a"][:: max(1, m["a"].size // 7)][:8]])
        # independent truth in the band (interval edges exactly as the reference computes them)
        a, b = solver.freq_interval(lo, up, 0.0)
        g["interval"] = [a, b]
        w = []
        if run is not None:                       # None: pattern + values only (large fluid P2 truth is too slow)
            w = solver.truth_eigs(mats, a, b)
            g["truth_eigs"] = [float(x) for x in w]
        if run:
            ops, lam, Y, res, info = solver.solve(mats, po, lo, up)
            g["oracle_lanczos"] = dict(nev=int(len(lam)), steps=int(info["steps"]), deg=int(info["deg"]),
                                       max_rel_err_vs_truth=float(np.max(np.abs(lam - w) / np.abs(w))) if len(lam) == len(w) else None,
                                       max_res_over_lam=float((res / np.abs(lam)).max()),
                                       boundsB=[float(x) for x in ops.boundsB],
                                       bounds=[float(info["xintv"][2]), float(info["xintv"][3])])
        golden[name] = g
        print(name, "N", g["N"], "Np", g["Np"], "truth", len(w), g.get("oracle_lanczos"), "%.1fs" % (time.time() - t0),
              flush=True)
        json.dump(golden, open(gpath, "w"), indent=1)


if __name__ == "__main__":
    main()
