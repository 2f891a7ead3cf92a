/* ORACLE (test infrastructure, NOT product code): plain C + OpenMP restatement of the CPU hot loops of
 * the reference stack, used (a) as a second checker beside oracle/solver.py and (b) as the CPU baseline
 * timed on the GPU box's host cores by bench.py (`cpu_baseline`, `--impl reference`).
 *
 * Parity status: UNPINNED -- the loops below restate pEVSL (fork js1019/pEVSL, source absent from
 * /root/reference): parcsr matvec (called from src/mod_matvec.f90:453,471,495,507,509,515), ChebIter
 * (:480,512; Saad Alg. 12.1) and ChebAv (inside pEVSL_CHEBLANNR_F90, src/mod_pevsl.f90:122), with the
 * reference's own operator definitions sparseAV (:445-458) and sparsefsAV (:498-520).  Shared-memory
 * threads stand in for the reference's MPI ranks (same row-block decomposition, no halo copies needed).
 *
 *   gcc -O3 -march=native -fopenmp -shared -fPIC nm_cpu.c -o libnm_cpu.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef struct {
  int nrow;
  const int* ia;
  const int* ja;
  const double* a;
} csr_t;

int nmcpu_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to every rank: the caller sets the team size from its affinity mask instead */
void nmcpu_set_threads(int n) { if (n >= 1) omp_set_num_threads(n); }

/* y = A x */
void nmcpu_spmv(int nrow, const int* ia, const int* ja, const double* a, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nrow; ++i) {
    double s = 0.0;
    for (int p = ia[i]; p < ia[i + 1]; ++p) s += a[p] * x[ja[p]];
    y[i] = s;
  }
}

/* x = q_deg(M) b, Chebyshev iteration on [lb,ub], zero initial guess; work: 3n doubles */
void nmcpu_chebiter(int n, const int* ia, const int* ja, const double* a, double lb, double ub, int deg,
                    const double* b, double* x, double* work) {
  double* r = work; double* d = work + n; double* w = work + 2 * (size_t)n;
  const double theta = (ub + lb) / 2.0, delta = (ub - lb) / 2.0, sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) { r[i] = b[i]; d[i] = b[i] / theta; x[i] = 0.0; }
  for (int k = 0; k < deg; ++k) {
    const double rho1 = 1.0 / (2.0 * sigma1 - rho);
    const double ak = rho1 * rho, bk = 2.0 * rho1 / delta;
    nmcpu_spmv(n, ia, ja, a, d, w);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      x[i] += d[i];
      r[i] -= w[i];
      d[i] = ak * d[i] + bk * r[i];
    }
    rho = rho1;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) x[i] += d[i];
}

/* The operator set of setupmatvec: B~ (scaled), d; solid: A (unscaled) ; fluid: Ad, E, ET (unscaled), Ap~ (scaled), dp */
typedef struct {
  int n, np, fluid;
  csr_t B, A, E, ET, Ap;
  const double* d; const double* dp;
  double lbB, ubB, lbAp, ubAp;
  int degB, degAp;
} ops_t;

/* w = sparseAV(v) or sparsefsAV(v); work: 2n + 5np doubles (+3np chebiter) */
static void apply_A(const ops_t* o, const double* v, double* w, double* work) {
  const int n = o->n, np = o->np;
  double* v0 = work; double* w0 = work + n;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) v0[i] = v[i] * o->d[i];
  nmcpu_spmv(n, o->A.ia, o->A.ja, o->A.a, v0, w0);
  if (!o->fluid) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) w[i] = w0[i] * o->d[i];
    return;
  }
  double* x0 = work + 2 * (size_t)n; double* y0 = x0 + np; double* cw = y0 + np;   /* cw: 3np */
  nmcpu_spmv(np, o->ET.ia, o->ET.ja, o->ET.a, v0, x0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < np; ++i) x0[i] *= o->dp[i];
  nmcpu_chebiter(np, o->Ap.ia, o->Ap.ja, o->Ap.a, o->lbAp, o->ubAp, o->degAp, x0, y0, cw);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < np; ++i) y0[i] *= o->dp[i];
  nmcpu_spmv(n, o->E.ia, o->E.ja, o->E.a, y0, w);            /* w1 */
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) w[i] = (w0[i] + w[i]) * o->d[i];
}

void nmcpu_apply_A(const ops_t* o, const double* v, double* w) {
  double* work = (double*)malloc(sizeof(double) * (2 * (size_t)o->n + 5 * (size_t)o->np + 8));
  apply_A(o, v, w, work);
  free(work);
}

/* y = sum_{k<=kmax} mu_k T_k((A B^-1 - cc)/dd) z : ChebAv truncated after kmax <= deg degree steps
 * (kmax = deg is the full filter; a smaller kmax is the bounded sample bench.py times). */
void nmcpu_chebav(const ops_t* o, int deg, const double* mu, double cc, double dd, int kmax, const double* z,
                  double* y) {
  const int n = o->n, np = o->np;
  double* buf = (double*)malloc(sizeof(double) * (9 * (size_t)n + 5 * (size_t)np + 8));
  double* vk = buf; double* vkp1 = buf + n; double* vkm1 = buf + 2 * (size_t)n; double* w2 = buf + 3 * (size_t)n;
  double* cw = buf + 4 * (size_t)n;                         /* 3n chebiter work */
  double* aw = buf + 7 * (size_t)n;                         /* 2n + 5np operator work */
  if (kmax > deg) kmax = deg;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) { vk[i] = z[i]; vkm1[i] = 0.0; y[i] = mu[0] * z[i]; }
  for (int k = 1; k <= kmax; ++k) {
    const double t = (k == 1 ? 1.0 : 2.0) / dd, s = mu[k];
    nmcpu_chebiter(n, o->B.ia, o->B.ja, o->B.a, o->lbB, o->ubB, o->degB, vk, w2, cw);
    apply_A(o, w2, vkp1, aw);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      vkp1[i] = t * (vkp1[i] - cc * vk[i]) - vkm1[i];
      y[i] += s * vkp1[i];
    }
    double* tmp = vkm1; vkm1 = vk; vk = vkp1; vkp1 = tmp;
  }
  free(buf);
}
