/* pevsl_f90.h -- the Fortran-callable pEVSL entry points NormalModes links against, re-implemented
 * on a B200 by libnm_b200.so.  This is the drop-in boundary of the hot path: the reference's Fortran
 * host (mod_matvec.f90 / mod_pevsl.f90) calls exactly these symbols (GNU/Intel external-procedure
 * convention: lowercase + trailing underscore, every argument by reference, Fortran `integer` =
 * int32_t, `integer*8` handles = uintptr_t, `double precision` = double, arrays contiguous).
 * All file:line citations are into /root/reference/src.
 *
 * Errors: none of the reference's calls has a status argument; its own convention is
 * `print*,'Error...'; stop`.  These wrappers print "[nm_b200] Error: ..." and exit(1).
 * The status-returning twin of every entry point is in nm_b200.h.
 */
#ifndef NM_PEVSL_F90_H
#define NM_PEVSL_F90_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t nm_fint;                 /* Fortran default integer / MPI_Fint */
typedef uintptr_t nm_handle;             /* Fortran integer*8 opaque handle     */
/* callback signature of sparseAV / sparsefsAV / sparseBV / sparseApV (mod_matvec.f90:445,461,485,498) */
typedef void (*nm_f90_matvec)(double* x, double* y, void* data);

/* mod_matvec.f90:75,152 ; mod_pevsl.f90:54.  `comm` is an opaque rank/size source here (ranks = GPUs,
 * communicator set up by nm_comm_init, see nm_b200.h). */
void pevsl_start_f90_(nm_fint* comm, nm_handle* pevsl_out);
/* mod_matvec.f90:95 ; mod_pevsl.f90:220 */
void pevsl_finish_f90_(nm_handle* pevsl);
/* mod_matvec.f90:78,155 ; mod_pevsl.f90:57.  nfirst = -1: undefined */
void pevsl_setprobsizes_f90_(nm_handle* pevsl, nm_fint* N_global, nm_fint* n_local, nm_fint* nfirst);
/* mod_matvec.f90:69-71,117-119,146-148,196-198,218-220,242-244.
 * row_starts/col_starts: [nranks+1] global offsets; ia: [n_local+1] 0-based; ja: 0-based GLOBAL ids.
 * Rectangular allowed (E: rows Ad%sizdist, cols Ap%sizdist).  Collective.  The arrays are copied. */
void pevsl_parcsrcreate_f90_(nm_fint* nrow_glob, nm_fint* ncol_glob, nm_fint* row_starts, nm_fint* col_starts,
                             nm_fint* ia, nm_fint* ja, double* a, nm_fint* comm, nm_handle* mat_out);
/* mod_matvec.f90:453,471,495,507,509,515.  x: n_cols_local, y: n_rows_local (host).  Collective. */
void pevsl_parcsrmatvec_f90_(double* x, double* y, nm_handle* mat);
/* mod_matvec.f90:80,157 ; mod_pevsl.f90:69,77,79 */
void pevsl_setamv_f90_(nm_handle* pevsl, nm_f90_matvec f, void* data);
void pevsl_setbmv_f90_(nm_handle* pevsl, nm_f90_matvec f, void* data);
/* mod_matvec.f90:85,162 ; mod_pevsl.f90:84 */
void pevsl_lanbounds_f90_(nm_handle* pevsl, nm_fint* mlan, nm_fint* lanstep, double* tol, double* lmin_out,
                          double* lmax_out);
/* mod_matvec.f90:93,174 */
void pevsl_setup_chebiter_f90_(double* lmin, double* lmax, nm_fint* deg, nm_handle* mat, nm_handle* cheb_out);
/* mod_matvec.f90:480,512.  type must be 2 (the only one the reference uses, :84,161) */
void pevsl_chebiter_f90_(nm_fint* type, double* b, double* x, nm_handle* cheb);
/* mod_pevsl.f90:73 */
void pevsl_setbsol_chebiter_f90_(nm_handle* pevsl, nm_fint* type, nm_handle* cheb);
/* mod_pevsl.f90:82 */
void pevsl_set_geneig_f90_(nm_handle* pevsl);
/* mod_pevsl.f90:115.  xintv = [a, b, lmin, lmax] */
void pevsl_findpol_f90_(double* xintv, double* thresh_int, double* thresh_ext, nm_handle* pol_out);
/* mod_pevsl.f90:122 */
void pevsl_cheblannr_f90_(nm_handle* pevsl, double* xintv, nm_fint* maxit, double* tol, nm_handle* pol);
/* mod_pevsl.f90:124 */
void pevsl_get_nev_f90_(nm_handle* pevsl, nm_fint* nev_out);
/* mod_pevsl.f90:130.  vals[nev], vecs[ld*nev] column-major */
void pevsl_copy_result_f90_(nm_handle* pevsl, double* vals, double* vecs, nm_fint* ld);
/* mod_pevsl.f90:132,134 */
void pevsl_chebiterstatsprint_f90_(nm_handle* cheb);
/* mod_pevsl.f90:215 */
void pevsl_freepol_f90_(nm_handle* pol);

/* ---- device-resident operator registration (what a GPU mod_matvec calls INSTEAD of setamv/setbmv with
 * host callbacks, so that cheblannr never bounces vectors through host memory).
 * The callbacks they replace: sparseAV (mod_matvec.f90:445-458), sparsefsAV (:498-520), sparseBV (:461-472),
 * sparseApV (:485-496). */
void nm_setamv_solid_f90_(nm_handle* pevsl, nm_handle* A, double* diag);
void nm_setamv_fluidsolid_f90_(nm_handle* pevsl, nm_handle* Ad, nm_handle* E, nm_handle* ET, nm_handle* chebAp,
                               double* diag, double* pdiag);
void nm_setamv_parcsr_f90_(nm_handle* pevsl, nm_handle* mat);
void nm_setbmv_parcsr_f90_(nm_handle* pevsl, nm_handle* mat);

#ifdef __cplusplus
}
#endif
#endif
