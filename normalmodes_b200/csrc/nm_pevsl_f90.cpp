// Fortran-callable pEVSL entry points (include/pevsl_f90.h): thin by-reference wrappers over the
// status-returning C API.  Error convention follows the reference (print 'Error...'; stop).
#include "nm_internal.h"
#include "../../include/pevsl_f90.h"
#include "../../include/nm_b200.h"

static void f90_check(int rc, const char* who) {
  if (rc != NM_OK) {
    fprintf(stderr, "[nm_b200] Error in %s: %s\n", who, nm_last_error_message());
    fflush(stderr);
    exit(1);
  }
}
#define F90(call) f90_check((call), __func__)

extern "C" {

void pevsl_start_f90_(nm_fint* comm, nm_handle* pevsl_out) {
  (void)comm;                                   // ranks = GPUs; communicator comes from nm_comm_init
  void* h = nullptr;
  F90(nm_pevsl_create(&h));
  *pevsl_out = (nm_handle)h;
}
void pevsl_finish_f90_(nm_handle* pevsl) {
  F90(nm_pevsl_free((void*)*pevsl));
  *pevsl = 0;
}
void pevsl_setprobsizes_f90_(nm_handle* pevsl, nm_fint* N, nm_fint* n, nm_fint* nfirst) {
  F90(nm_pevsl_setprobsizes((void*)*pevsl, *N, *n, *nfirst));
}
void pevsl_parcsrcreate_f90_(nm_fint* nrow_glob, nm_fint* ncol_glob, nm_fint* row_starts, nm_fint* col_starts,
                             nm_fint* ia, nm_fint* ja, double* a, nm_fint* comm, nm_handle* mat_out) {
  (void)comm;
  void* h = nullptr;
  F90(nm_parcsr_create(*nrow_glob, *ncol_glob, row_starts, col_starts, ia, ja, a, &h));
  *mat_out = (nm_handle)h;
}
void pevsl_parcsrmatvec_f90_(double* x, double* y, nm_handle* mat) { F90(nm_parcsr_matvec((void*)*mat, x, y)); }
void pevsl_setamv_f90_(nm_handle* pevsl, nm_f90_matvec f, void* data) {
  F90(nm_pevsl_setamv_callback((void*)*pevsl, f, data));
}
void pevsl_setbmv_f90_(nm_handle* pevsl, nm_f90_matvec f, void* data) {
  F90(nm_pevsl_setbmv_callback((void*)*pevsl, f, data));
}
void pevsl_lanbounds_f90_(nm_handle* pevsl, nm_fint* mlan, nm_fint* lanstep, double* tol, double* lmin, double* lmax) {
  F90(nm_pevsl_lanbounds((void*)*pevsl, *mlan, *lanstep, *tol, lmin, lmax));
}
void pevsl_setup_chebiter_f90_(double* lmin, double* lmax, nm_fint* deg, nm_handle* mat, nm_handle* cheb_out) {
  void* h = nullptr;
  F90(nm_chebiter_create(*lmin, *lmax, *deg, (void*)*mat, &h));
  *cheb_out = (nm_handle)h;
}
void pevsl_chebiter_f90_(nm_fint* type, double* b, double* x, nm_handle* cheb) {
  if (*type != 2) {
    nm_record_error("pevsl_chebiter_f90: only type 2 is implemented (the one NormalModes uses)");
    f90_check(NM_ERR, __func__);
  }
  F90(nm_chebiter_solve_host((void*)*cheb, b, x));
}
void pevsl_setbsol_chebiter_f90_(nm_handle* pevsl, nm_fint* type, nm_handle* cheb) {
  if (*type != 2) {
    nm_record_error("pevsl_setbsol_chebiter_f90: only type 2 is implemented");
    f90_check(NM_ERR, __func__);
  }
  F90(nm_pevsl_setbsol_chebiter((void*)*pevsl, (void*)*cheb));
}
void pevsl_set_geneig_f90_(nm_handle* pevsl) { F90(nm_pevsl_set_geneig((void*)*pevsl)); }
void pevsl_findpol_f90_(double* xintv, double* thresh_int, double* thresh_ext, nm_handle* pol_out) {
  void* h = nullptr;
  F90(nm_findpol_create(xintv, *thresh_int, *thresh_ext, &h));
  *pol_out = (nm_handle)h;
}
void pevsl_cheblannr_f90_(nm_handle* pevsl, double* xintv, nm_fint* maxit, double* tol, nm_handle* pol) {
  F90(nm_pevsl_cheblannr((void*)*pevsl, xintv, *maxit, *tol, (void*)*pol));
}
void pevsl_get_nev_f90_(nm_handle* pevsl, nm_fint* nev_out) {
  int nev = 0;
  F90(nm_pevsl_get_nev((void*)*pevsl, &nev));
  *nev_out = nev;
}
void pevsl_copy_result_f90_(nm_handle* pevsl, double* vals, double* vecs, nm_fint* ld) {
  F90(nm_pevsl_copy_result((void*)*pevsl, vals, vecs, *ld, nullptr));
}
void pevsl_chebiterstatsprint_f90_(nm_handle* cheb) {
  long long nsolve = 0, nmv = 0;
  int deg = 0, rank = 0;
  double lb = 0, ub = 0;
  F90(nm_chebiter_stats((void*)*cheb, &nsolve, &nmv, &deg, &lb, &ub));
  nm_device_info(nullptr, nullptr, &rank, nullptr);
  if (rank == 0)
    printf(" ChebIter: deg %d on [%.6e, %.6e]: %lld solves, %lld matvecs\n", deg, lb, ub, nsolve, nmv);
}
void pevsl_freepol_f90_(nm_handle* pol) {
  F90(nm_pol_free((void*)*pol));
  *pol = 0;
}

// ---- device-resident operator registration
void nm_setamv_solid_f90_(nm_handle* pevsl, nm_handle* A, double* diag) {
  void* op = nullptr;
  F90(nm_op_create_solid((void*)*A, diag, &op));
  F90(nm_pevsl_setamv_op((void*)*pevsl, op));
  F90(nm_pevsl_adopt_op((void*)*pevsl, op));
}
void nm_setamv_fluidsolid_f90_(nm_handle* pevsl, nm_handle* Ad, nm_handle* E, nm_handle* ET, nm_handle* chebAp,
                               double* diag, double* pdiag) {
  void* op = nullptr;
  F90(nm_op_create_fluidsolid((void*)*Ad, (void*)*E, (void*)*ET, (void*)*chebAp, diag, pdiag, &op));
  F90(nm_pevsl_setamv_op((void*)*pevsl, op));
  F90(nm_pevsl_adopt_op((void*)*pevsl, op));
}
void nm_setamv_parcsr_f90_(nm_handle* pevsl, nm_handle* mat) {
  void* op = nullptr;
  F90(nm_op_create_csr((void*)*mat, &op));
  F90(nm_pevsl_setamv_op((void*)*pevsl, op));
  F90(nm_pevsl_adopt_op((void*)*pevsl, op));
}
void nm_setbmv_parcsr_f90_(nm_handle* pevsl, nm_handle* mat) {
  void* op = nullptr;
  F90(nm_op_create_csr((void*)*mat, &op));
  F90(nm_pevsl_setbmv_op((void*)*pevsl, op));
  F90(nm_pevsl_adopt_op((void*)*pevsl, op));
}

}  // extern "C"
