// The operators pEVSL calls back into, as device-resident objects: sparseAV (src/mod_matvec.f90:445-458),
// sparsefsAV (:498-520), sparseBV (:461-472), sparseApV (:485-496), plus ChebAv -- the polynomial
// filter y = p(A B^-1) z of pEVSL's cheblanNr.c -- with the three-term update fused into the A product.
#include "nm_spmv.cuh"
#include <algorithm>

// ---------------------------------------------------------------- construction
static NmOp* op_new(int kind, int n) {
  NmOp* op = new NmOp();
  op->kind = kind; op->n = n;
  return op;
}

// w = D A D v : A is the UNSCALED stiffness handle, d the Jacobi scaling (host, local rows).
static NmOp* op_solid(NmParcsr* A, const double* d_host) {
  NM_REQUIRE(A->nrow == A->ncol, "solid operator: A must be square");
  std::unique_ptr<NmOp> op(op_new(NM_OP_SOLID, A->nrow));
  DBuf<double> d(std::max(A->nrow, 1));
  d.upload(d_host, A->nrow);
  op->As.reset(nm_parcsr_scaled_copy(*A, d.p, d.p));
  return op.release();
}

// w = D [Ad + E Dp Ap~^-1 Dp ET] D v (sparsefsAV): scaled copies D Ad D, D E Dp, Dp ET D.
static NmOp* op_fluidsolid(NmParcsr* Ad, NmParcsr* E, NmParcsr* ET, NmChebIter* chebAp, const double* d_host,
                           const double* dp_host) {
  NM_REQUIRE(Ad->nrow == Ad->ncol && E->nrow == Ad->nrow && ET->ncol == Ad->ncol && E->ncol == ET->nrow &&
                 chebAp->M->nrow == ET->nrow,
             "fluid-solid operator: inconsistent block sizes");
  std::unique_ptr<NmOp> op(op_new(NM_OP_FLUIDSOLID, Ad->nrow));
  const int n = Ad->nrow, np = ET->nrow;
  DBuf<double> d(std::max(n, 1)), dp(std::max(np, 1));
  d.upload(d_host, n); dp.upload(dp_host, np);
  op->As.reset(nm_parcsr_scaled_copy(*Ad, d.p, d.p));
  op->Es.reset(nm_parcsr_scaled_copy(*E, d.p, dp.p));
  op->ETs.reset(nm_parcsr_scaled_copy(*ET, dp.p, d.p));
  op->chebAp = chebAp;
  op->x1.alloc(std::max(np, 1)); op->y0.alloc(std::max(np, 1)); op->w1.alloc(std::max(n, 1));
  return op.release();
}

// ---------------------------------------------------------------- application
static void op_callback(NmOp& op, const double* x, double* y) {
  // Host callback (the reference's Fortran sparseAV/sparseBV): bounce through host memory.
  NmCtx& c = nm_ctx();
  op.hx.resize(op.n); op.hy.resize(op.n);
  NM_CUDA(cudaMemcpyAsync(op.hx.data(), x, op.n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  op.fn(op.hx.data(), op.hy.data(), op.fn_data);
  NM_CUDA(cudaMemcpyAsync(y, op.hy.data(), op.n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
}

// pressure part of the fluid-solid operator: w1 = E' Ap~^-1 ET' x
static void op_fluid_term(NmOp& op, const double* x) {
  nm_spmv(*op.ETs, x, op.x1.p);
  nm_chebiter_solve(*op.chebAp, op.x1.p, op.y0.p);
  nm_spmv(*op.Es, op.y0.p, op.w1.p);
}

void nm_op_apply(NmOp& op, const double* x, double* y) {
  op.napply++;
  switch (op.kind) {
    case NM_OP_CSR: nm_spmv(*op.M, x, y); break;
    case NM_OP_SOLID: nm_spmv(*op.As, x, y); break;
    case NM_OP_FLUIDSOLID:
      op_fluid_term(op, x);
      nm_spmv_epi(*op.As, x, EpiStorePlus{y, op.w1.p});
      break;
    case NM_OP_CALLBACK: op_callback(op, x, y); break;
    default: NM_REQUIRE(false, "unknown operator kind %d", op.kind);
  }
}

void nm_op_apply_filter(NmOp& op, const double* w, const double* vk, const double* vkm1, double* vout, double* y,
                        double t, double cc, double mu, double mu0, int first) {
  op.napply++;
  EpiFilter e{vk, vkm1, vout, y, nullptr, t, cc, mu, mu0, first};
  switch (op.kind) {
    case NM_OP_CSR: nm_spmv_epi(*op.M, w, e); break;
    case NM_OP_SOLID: nm_spmv_epi(*op.As, w, e); break;
    case NM_OP_FLUIDSOLID:
      op_fluid_term(op, w);
      e.add = op.w1.p;
      nm_spmv_epi(*op.As, w, e);
      break;
    case NM_OP_CALLBACK:
      if (op.w1.n < (size_t)op.n) op.w1.alloc(std::max(op.n, 1));
      op_callback(op, w, op.w1.p);
      nm_filter_update(vkm1, vout, vk, op.w1.p, y, t, cc, mu, mu0, first, op.n);
      break;
    default: NM_REQUIRE(false, "unknown operator kind %d", op.kind);
  }
}

// ---------------------------------------------------------------- ChebAv
// y = sum_k mu_k T_k((A B^-1 - cc)/dd) z.  work: 3n doubles (two recurrence vectors + w = B^-1 vk).
void nm_filter_apply(NmPevsl& P, const NmPol& pol, const double* z, double* y, double* work, int kmax) {
  const size_t n = P.n;
  NM_REQUIRE(P.A, "filter: no A operator registered");
  double* va = work;
  double* vb = work + n;
  double* w = work + 2 * n;
  const double* vk = z;          // v_1 = z is only read
  const double* vkm1 = nullptr;  // v_0 = 0
  NM_REQUIRE(pol.deg >= 1, "filter: polynomial degree %d < 1", pol.deg);
  // kmax > 0: the sum truncated after kmax degree steps (a slice of one application: the unit bench.py times on
  // meshes where a whole application takes minutes); every degree step costs the same
  const int m = (kmax > 0 && kmax < pol.deg) ? kmax : pol.deg;
  for (int k = 1; k <= m; ++k) {
    // output slot: k=1 -> va, k=2 -> vb (v_{k-1} = z must survive), then in place over v_{k-1}
    double* vout = (k == 1) ? va : (k == 2 ? vb : const_cast<double*>(vkm1));
    const double t = (k == 1 ? 1.0 : 2.0) / pol.dd;
    const double* src = vk;
    if (P.geneig) {
      NM_REQUIRE(P.bsol, "filter: generalised problem without a B solver (pevsl_setbsol_chebiter_f90)");
      nm_chebiter_solve(*P.bsol, vk, w);
      src = w;
    }
    nm_op_apply_filter(*P.A, src, vk, vkm1, vout, y, t, pol.cc, pol.mu[k], pol.mu[0], k == 1);
    vkm1 = vk; vk = vout;
  }
  if (m == pol.deg) P.n_filter_apply++;
}

// ---------------------------------------------------------------- C ABI: operators
extern "C" int nm_op_create_csr(void* mat, void** out) {
  NM_API_BEGIN
  NmParcsr* M = (NmParcsr*)mat;
  NM_REQUIRE(M->nrow == M->ncol, "nm_op_create_csr: square matrix required");
  NmOp* op = op_new(NM_OP_CSR, M->nrow);
  op->M = M;
  *out = op;
  NM_API_END
}
extern "C" int nm_op_create_solid(void* A, const double* d, void** out) {
  NM_API_BEGIN
  *out = op_solid((NmParcsr*)A, d);
  NM_API_END
}
extern "C" int nm_op_create_fluidsolid(void* Ad, void* E, void* ET, void* chebAp, const double* d, const double* dp,
                                       void** out) {
  NM_API_BEGIN
  *out = op_fluidsolid((NmParcsr*)Ad, (NmParcsr*)E, (NmParcsr*)ET, (NmChebIter*)chebAp, d, dp);
  NM_API_END
}
extern "C" int nm_op_create_callback(int n, nm_matvec_fn fn, void* data, void** out) {
  NM_API_BEGIN
  NmOp* op = op_new(NM_OP_CALLBACK, n);
  op->fn = fn; op->fn_data = data;
  *out = op;
  NM_API_END
}
extern "C" int nm_op_free(void* h) {
  NM_API_BEGIN
  if (h) { NM_CUDA(cudaStreamSynchronize(nm_ctx().stream)); delete (NmOp*)h; }
  NM_API_END
}
// y = Op x with HOST vectors.
extern "C" int nm_op_apply_host(void* h, const double* x, double* y) {
  NM_API_BEGIN
  NmOp& op = *(NmOp*)h;
  NmCtx& c = nm_ctx();
  nm_ensure_init();
  const int n = op.n;
  DBuf<double> dx(std::max(n, 1)), dy(std::max(n, 1));
  if (n) NM_CUDA(cudaMemcpyAsync(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  nm_op_apply(op, dx.p, dy.p);
  if (n) NM_CUDA(cudaMemcpyAsync(y, dy.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_API_END
}
extern "C" int nm_op_apply_dev(void* h, const double* x_dev, double* y_dev) {
  NM_API_BEGIN
  nm_op_apply(*(NmOp*)h, x_dev, y_dev);
  NM_API_END
}
