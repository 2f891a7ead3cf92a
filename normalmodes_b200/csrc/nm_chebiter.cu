// Fixed-degree Chebyshev iteration x = q_deg(M) b ~ M^-1 b: replaces pEVSL's
// EXTERNAL/ITERSOL/chebiter.c as used through pevsl_setup_chebiter_f90 / pevsl_chebiter_f90
// (src/mod_matvec.f90:88-93,167-174,480,512) and as the B-solve registered by
// pevsl_setbsol_chebiter_f90 (src/mod_pevsl.f90:72-73).  Saad, Iterative Methods, Alg. 12.1 with a
// zero initial guess; no inner products.  Each step is ONE kernel: the SpMV r -= M d with the
// x += d and d = a d + b r updates in its epilogue (EpiCheb), so a step streams M once and touches
// each vector once.
#include "nm_spmv.cuh"

NmChebIter* nm_chebiter_build(double lb, double ub, int deg, NmParcsr* M) {
  NM_REQUIRE(M && M->nrow == M->ncol, "setup_chebiter: square matrix required");
  NM_REQUIRE(deg >= 1, "setup_chebiter: degree %d < 1", deg);
  NM_REQUIRE(ub > lb && lb > 0.0, "setup_chebiter: need 0 < lmin < lmax (got %g, %g)", lb, ub);
  std::unique_ptr<NmChebIter> C(new NmChebIter());
  C->M = M; C->lb = lb; C->ub = ub; C->deg = deg;
  const double theta = (ub + lb) / 2.0, delta = (ub - lb) / 2.0;
  const double sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
  C->theta = theta;
  for (int k = 0; k < deg; ++k) {
    const double rho1 = 1.0 / (2.0 * sigma1 - rho);
    C->ak.push_back(rho1 * rho);
    C->bk.push_back(2.0 * rho1 / delta);
    rho = rho1;
  }
  const size_t n = M->nrow > 0 ? M->nrow : 1;
  C->r.alloc(n); C->d0.alloc(n); C->d1.alloc(n);
  return C.release();
}

void nm_chebiter_solve(NmChebIter& C, const double* b, double* x) {
  NM_REQUIRE(b != x, "chebiter: b and x must not alias");
  NmParcsr& M = *C.M;
  double* dbuf[2] = {C.d0.p, C.d1.p};
  const double* din = b;
  for (int k = 0; k < C.deg; ++k) {
    EpiCheb e;
    e.first = (k == 0); e.last = (k == C.deg - 1);
    e.r_in = e.first ? b : C.r.p;
    e.d_in = din;
    e.r_out = C.r.p;
    e.d_out = dbuf[k & 1];
    e.x = x;
    e.inv_theta = 1.0 / C.theta; e.ak = C.ak[k]; e.bk = C.bk[k];
    nm_spmv_epi(M, din, e);
    din = dbuf[k & 1];
  }
  C.nsolve++;
  C.nmatvec += C.deg;
}

// ---------------------------------------------------------------- C ABI
extern "C" int nm_chebiter_create(double lmin, double lmax, int deg, void* mat, void** out) {
  NM_API_BEGIN
  *out = nm_chebiter_build(lmin, lmax, deg, (NmParcsr*)mat);
  NM_API_END
}
extern "C" int nm_chebiter_free(void* h) {
  NM_API_BEGIN
  if (h) { NM_CUDA(cudaStreamSynchronize(nm_ctx().stream)); delete (NmChebIter*)h; }
  NM_API_END
}
// x = q(M) b with HOST vectors (pevsl_chebiter_f90 contract).
extern "C" int nm_chebiter_solve_host(void* h, const double* b, double* x) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  NmCtx& c = nm_ctx();
  const int n = C.M->nrow;
  DBuf<double> db(n > 0 ? n : 1), dx(n > 0 ? n : 1);
  if (n) NM_CUDA(cudaMemcpyAsync(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  nm_chebiter_solve(C, db.p, dx.p);
  if (n) NM_CUDA(cudaMemcpyAsync(x, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_API_END
}
extern "C" int nm_chebiter_solve_dev(void* h, const double* b_dev, double* x_dev) {
  NM_API_BEGIN
  nm_chebiter_solve(*(NmChebIter*)h, b_dev, x_dev);
  NM_API_END
}
extern "C" int nm_chebiter_stats(void* h, long long* nsolve, long long* nmatvec, int* deg, double* lmin, double* lmax) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  if (nsolve) *nsolve = C.nsolve;
  if (nmatvec) *nmatvec = C.nmatvec;
  if (deg) *deg = C.deg;
  if (lmin) *lmin = C.lb;
  if (lmax) *lmax = C.ub;
  NM_API_END
}
