"""Host-side mirror of the reference's raw-binary readers and MPI-IO result writers.

    read_mesh / read_model      src/mod_geometry.f90:83-135,179-220,1268-1349 ; src/mod_cg_models.f90:301-429
    output_names                src/mod_para.f90:271-289
    save_eigenvectors           src/mod_pevsl.f90:188-201  (pnm_save_dreal :283-303)
    save_vlist_vstat            src/mod_pevsl.f90:225-243  (pnm_save_integer :262-281)

All files are little-endian raw binary without headers.  Every rank writes ITS slice at its byte offset, exactly as
the reference's `mpi_file_set_view(fid, disp, ...)` + `mpi_file_write` do, so N processes (one per GPU) can call the
writers concurrently on a shared file system; a single process passes every rank's slice in turn.
"""
import os

import numpy as np


def read_mesh(prefix):
    """`<prefix>_mesh.header` (text: Ntet Nvert), `_ele.dat` / `_neigh.dat` int32 [Ntet][4] 1-based (-1 = boundary),
    `_node.dat` fp64 [Nvert][3].  Returns 0-based arrays in the layout of normalmodes_b200.meshgen.build_mesh."""
    with open(prefix + "_mesh.header") as f:
        ntet, nvert = [int(v) for v in f.read().split()[:2]]
    ele = np.fromfile(prefix + "_ele.dat", dtype="<i4")
    neigh = np.fromfile(prefix + "_neigh.dat", dtype="<i4")
    node = np.fromfile(prefix + "_node.dat", dtype="<f8")
    if ele.size != 4 * ntet or neigh.size != 4 * ntet or node.size != 3 * nvert:
        raise ValueError("mesh files of %s do not match the header (Ntet %d, Nvert %d)" % (prefix, ntet, nvert))
    ele = ele.reshape(ntet, 4).astype(np.int64) - 1
    neigh = neigh.reshape(ntet, 4).astype(np.int64)
    neigh = np.where(neigh > 0, neigh - 1, -1)
    return dict(ntet=ntet, nvert=nvert, ele=ele, neigh=neigh, node=node.reshape(nvert, 3))


def read_model(prefix, ntet, porder, job):
    """`_vp_pod_<p>_true.dat`, `_vs_...`, `_rho_...` fp64 [Ntet][pNp]; JOB >= 2: `_pod_<p>_potential_acceleration_true.dat`
    fp64 [Ntet][pNp][3] (m/s^2)."""
    pnp = 4 if porder == 1 else 10
    out = {}
    for k in ("vp", "vs", "rho"):
        a = np.fromfile("%s_%s_pod_%d_true.dat" % (prefix, k, porder), dtype="<f8")
        if a.size != ntet * pnp:
            raise ValueError("%s model file has %d values, expected %d" % (k, a.size, ntet * pnp))
        out[k] = a.reshape(ntet, pnp)
    out["g0"] = None
    if job >= 2:
        g = np.fromfile("%s_pod_%d_potential_acceleration_true.dat" % (prefix, porder), dtype="<f8")
        if g.size != ntet * pnp * 3:
            raise ValueError("gravity file has %d values, expected %d" % (g.size, ntet * pnp * 3))
        out["g0"] = g.reshape(ntet, pnp, 3)
    return out


def _list_directed_real4(v):
    """Text of `write(s,*) real(v,4)` as gfortran prints it (list-directed REAL(4) = 1PG16.9E2-like: 9 significant
    digits in total; F form with 9 - (integer digits) decimals for 0.1 <= |x| < 1e9, otherwise d.ddddddddE+ee).
    Intel Fortran prints 7 digits -- the file name is compiler-dependent in the reference too (src/mod_para.f90:271-289)."""
    x = float(np.float32(v))
    if x == 0.0:
        return "0.00000000"
    ax = abs(x)
    if ax < 0.1 or ax >= 1.0e9:
        m, e = ("%.8E" % x).split("E")
        return "%sE%s%02d" % (m, e[0], abs(int(e)))
    e10 = int(np.floor(np.log10(ax)))
    if float("%.9g" % ax) >= 10.0 ** (e10 + 1):             # rounding carried into the next decade (9.9999999999 -> 10.0)
        e10 += 1
    nint = max(e10 + 1, 0)                                  # digits in front of the decimal point (0.2 -> none counted)
    dec = 9 - nint if nint > 0 else 9
    return "%.*f" % (dec, x)


def output_names(outputdir, basename, job, porder, nproc, lowfreq, upfreq):
    """pin%fvlist / fvstat / fvdata of src/mod_para.f90:271-289."""
    stem = "%s%s_pod%d_np%d" % (outputdir, basename, porder, nproc)
    fvdata = "%s%s_JOB%d_pod%d_np%d_%s_%s" % (outputdir, basename, job, porder, nproc, _list_directed_real4(lowfreq),
                                               _list_directed_real4(upfreq))
    return dict(fvlist=stem + "_vlist.dat", fvstat=stem + "_vstat.dat", fvdata=fvdata)


def _write_at(fname, byte_offset, arr):
    """One rank's mpi_file_set_view(disp) + mpi_file_write: create if missing, never truncate, write at the offset."""
    fd = os.open(fname, os.O_WRONLY | os.O_CREAT, 0o644)
    try:
        os.pwrite(fd, arr.tobytes(), int(byte_offset))
    finally:
        os.close(fd)


def save_eigenvectors(fvdata, eigvec, diag, sizdist, rank):
    """Eigenvector i (ascending eigenvalue, 1-based file index) of this rank's rows in PHYSICAL coordinates
    x = d * y (`EIGVEC(...)*mymatvec%B%diag`, :199) at byte offset sizdist[rank]*8 of `<fvdata>_<i>.dat`.
    eigvec: [nev][n_local] (rows = pairs, as pnm_apply_pevsl returns them), diag: the Jacobi scaling of B."""
    eigvec = np.asarray(eigvec, dtype=np.float64)
    diag = np.asarray(diag, dtype=np.float64)
    names = []
    for i in range(eigvec.shape[0]):
        fname = "%s_%d.dat" % (fvdata, i + 1)
        _write_at(fname, int(sizdist[rank]) * 8, (eigvec[i] * diag).astype("<f8"))
        names.append(fname)
    return names


def save_vlist_vstat(names, vlist_local, vtxdist, rank, vstat_local=None):
    """`unstrM%new%vlist` (original 1-based node id of each row block of this rank) at byte offset vtxdist[rank]*4 of
    fvlist; the node status likewise into fvstat when the model has a fluid-solid boundary (:225-243)."""
    _write_at(names["fvlist"], int(vtxdist[rank]) * 4, np.asarray(vlist_local).astype("<i4"))
    if vstat_local is not None:
        _write_at(names["fvstat"], int(vtxdist[rank]) * 4, np.asarray(vstat_local).astype("<i4"))


def read_eigenvector(fvdata, i, n_global):
    a = np.fromfile("%s_%d.dat" % (fvdata, i), dtype="<f8")
    if a.size != n_global:
        raise ValueError("eigenvector file holds %d values, expected %d" % (a.size, n_global))
    return a
