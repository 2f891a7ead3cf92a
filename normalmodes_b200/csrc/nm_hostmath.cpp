// Host-side dense maths of the solve: symmetric tridiagonal eigenproblem (pEVSL SymmTridEig,
// LAPACK there) and the filter-polynomial construction (pEVSL find_pol / chebpoly.c as called from
// src/mod_pevsl.f90:108-115).  O(k^2)-O(k^3) scalar work on k <= MAXIT numbers; not HBM-relevant.
#include "nm_internal.h"
#include <algorithm>
#include <numeric>

// ---------------------------------------------------------------- implicit QL (EISPACK tql1/tql2 scheme)
// d[0..k-1] diagonal, e[0..k-2] sub-diagonal.  w: ascending eigenvalues.  Z (optional): k*k
// column-major, column j = eigenvector of w[j].  Returns 0, or l+1 if eigenvalue l failed to converge.
// lastrow (optional, length k): bottom component of every eigenvector, obtained by applying the
// rotations to the last row of the identity only -- O(k^2), enough for Lanczos residual estimates.
int nm_tridiag_eig_ex(int k, const double* din, const double* ein, double* w, double* Z, double* lastrow) {
  if (k <= 0) return 0;
  std::vector<double> d(din, din + k), e(k, 0.0);
  for (int i = 0; i + 1 < k; ++i) e[i] = ein[i];
  if (Z) {
    std::fill(Z, Z + (size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) Z[(size_t)i * k + i] = 1.0;
  }
  std::vector<double> lr;
  if (lastrow) { lr.assign(k, 0.0); lr[k - 1] = 1.0; }
  double f = 0.0, tst1 = 0.0;
  for (int l = 0; l < k; ++l) {
    int iter = 0;
    double h = fabs(d[l]) + fabs(e[l]);
    if (tst1 < h) tst1 = h;
    int m = l;
    while (m < k) {
      if (tst1 + fabs(e[m]) == tst1) break;     // e[k-1] == 0 always stops the scan
      ++m;
    }
    if (m != l) {
      double tst2;
      do {
        if (iter++ == 60) return l + 1;
        const int l1 = l + 1;
        double g = d[l];
        double p = (d[l1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        const double sr = p >= 0 ? fabs(r) : -fabs(r);
        d[l] = e[l] / (p + sr);
        d[l1] = e[l] * (p + sr);
        const double dl1 = d[l1];
        h = g - d[l];
        for (int i = l1 + 1; i < k; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
        const double el1 = e[l1];
        for (int i = m - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          if (Z) {
            double* zi = Z + (size_t)i * k;
            double* zi1 = Z + (size_t)(i + 1) * k;
            for (int q = 0; q < k; ++q) {
              const double hh = zi1[q];
              zi1[q] = s * zi[q] + c * hh;
              zi[q] = c * zi[q] - s * hh;
            }
          }
          if (lastrow) {
            const double hh = lr[i + 1];
            lr[i + 1] = s * lr[i] + c * hh;
            lr[i] = c * lr[i] - s * hh;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
        tst2 = tst1 + fabs(e[l]);
      } while (tst2 > tst1);
    }
    d[l] += f;
  }
  std::vector<int> idx(k);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  for (int i = 0; i < k; ++i) w[i] = d[idx[i]];
  if (Z) {
    std::vector<double> T((size_t)k * k);
    for (int j = 0; j < k; ++j) std::copy(Z + (size_t)idx[j] * k, Z + (size_t)idx[j] * k + k, T.begin() + (size_t)j * k);
    std::copy(T.begin(), T.end(), Z);
  }
  if (lastrow) for (int j = 0; j < k; ++j) lastrow[j] = lr[idx[j]];
  return 0;
}

int nm_tridiag_eig(int k, const double* d, const double* e, double* w, double* Z) {
  return nm_tridiag_eig_ex(k, d, e, w, Z, nullptr);
}

// ---------------------------------------------------------------- small dense symmetric-definite pencil
// H c = w G c for the Rayleigh-Ritz refinement of the accepted Ritz vectors (nm_lanczos.cu): H, G are m x m,
// column-major, symmetric, G ~ I (the vectors are B-orthonormal to rounding) and H nearly diagonal (they are nearly
// eigenvectors) -- the case where the cyclic Jacobi method converges in two or three sweeps and is as accurate as
// anything (the image has no LAPACK).  G = L L^T (Cholesky), H' = L^-1 H L^-T, Jacobi on H', C = L^-T C'.
// Output: w ascending, C column-major with C^T G C = I.  Returns 0, 1 (G not positive definite) or 2 (no convergence).
int nm_sym_geneig(int m, const double* Hin, const double* Gin, double* w, double* C) {
  if (m <= 0) return 0;
  const size_t M = (size_t)m;
  std::vector<double> L(M * M, 0.0), A(M * M);
  for (int j = 0; j < m; ++j)                                   // Cholesky, lower, column by column
    for (int i = j; i < m; ++i) {
      double sum = 0.5 * (Gin[i + j * M] + Gin[j + i * M]);
      for (int k = 0; k < j; ++k) sum -= L[i + k * M] * L[j + k * M];
      if (i == j) {
        if (!(sum > 0.0)) return 1;
        L[j + j * M] = sqrt(sum);
      } else {
        L[i + j * M] = sum / L[j + j * M];
      }
    }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < m; ++i) A[i + j * M] = 0.5 * (Hin[i + j * M] + Hin[j + i * M]);
  for (int j = 0; j < m; ++j)                                   // A <- L^-1 A (forward substitution on every column)
    for (int i = 0; i < m; ++i) {
      double sum = A[i + j * M];
      for (int k = 0; k < i; ++k) sum -= L[i + k * M] * A[k + j * M];
      A[i + j * M] = sum / L[i + i * M];
    }
  for (int i = 0; i < m; ++i)                                   // A <- A L^-T (the same on every row)
    for (int j = 0; j < m; ++j) {
      double sum = A[i + j * M];
      for (int k = 0; k < j; ++k) sum -= A[i + k * M] * L[j + k * M];
      A[i + j * M] = sum / L[j + j * M];
    }
  for (int j = 0; j < m; ++j)
    for (int i = 0; i < j; ++i) { const double t = 0.5 * (A[i + j * M] + A[j + i * M]); A[i + j * M] = A[j + i * M] = t; }
  std::vector<double> V(M * M, 0.0);
  for (int i = 0; i < m; ++i) V[i + i * M] = 1.0;
  // rotate (p, q) only while |a_pq| > eps sqrt(|a_pp a_qq|) (the relative criterion that gives small eigenvalues to
  // full relative accuracy); converged when a whole sweep rotates nothing
  bool converged = false;
  const double eps = 2.220446049250313e-16;
  for (int sweep = 0; sweep < 60 && !converged; ++sweep) {
    long rotations = 0;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[p + q * M];
        const double app = A[p + p * M], aqq = A[q + q * M];
        if (fabs(apq) <= eps * sqrt(fabs(app * aqq))) { A[p + q * M] = A[q + p * M] = 0.0; continue; }
        ++rotations;
        const double theta = (aqq - app) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < m; ++k) {                            // columns p, q
          const double akp = A[k + p * M], akq = A[k + q * M];
          A[k + p * M] = c * akp - sn * akq;
          A[k + q * M] = sn * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {                            // rows p, q
          const double apk = A[p + k * M], aqk = A[q + k * M];
          A[p + k * M] = c * apk - sn * aqk;
          A[q + k * M] = sn * apk + c * aqk;
        }
        A[p + q * M] = A[q + p * M] = 0.0;
        for (int k = 0; k < m; ++k) {
          const double vkp = V[k + p * M], vkq = V[k + q * M];
          V[k + p * M] = c * vkp - sn * vkq;
          V[k + q * M] = sn * vkp + c * vkq;
        }
      }
    converged = rotations == 0;
  }
  if (!converged) return 2;
  for (int j = 0; j < m; ++j)                                   // V <- L^-T V (back substitution)
    for (int i = m - 1; i >= 0; --i) {
      double sum = V[i + j * M];
      for (int k = i + 1; k < m; ++k) sum -= L[k + i * M] * V[k + j * M];
      V[i + j * M] = sum / L[i + i * M];
    }
  std::vector<int> idx(m);
  std::iota(idx.begin(), idx.end(), 0);
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[a + a * M] < A[b + b * M]; });
  for (int j = 0; j < m; ++j) {
    w[j] = A[idx[j] + idx[j] * M];
    std::copy(V.begin() + idx[j] * M, V.begin() + (idx[j] + 1) * M, C + j * M);
  }
  return 0;
}

// ---------------------------------------------------------------- find_pol
static void dampcf(int m, int damping, std::vector<double>& jac) {
  jac.assign(m + 1, 1.0);
  jac[0] = 0.5;                                   // half weight of the zeroth Chebyshev term
  const double dm = (double)m;
  for (int k = 1; k <= m; ++k) {
    if (damping == 1) {                           // Jackson
      const double thJ = M_PI / (dm + 2.0), a1 = 1.0 / (dm + 2.0), a2 = sin(thJ);
      jac[k] = a1 * sin((k + 1) * thJ) / a2 + (1.0 - (k + 1) * a1) * cos(k * thJ);
    } else if (damping == 2) {                    // Lanczos sigma
      const double thL = M_PI / (dm + 1.0);
      jac[k] = sin(k * thL) / (k * thL);
    }
  }
}

static double chebx(const std::vector<double>& mu, int m, double x) {
  double vkm1 = 0.0, vk = 1.0, y = mu[0];
  for (int k = 1; k <= m; ++k) {
    const double s = (k == 1) ? 1.0 : 2.0;
    const double vkp1 = s * x * vk - vkm1;
    y += mu[k] * vkp1;
    vkm1 = vk; vk = vkp1;
  }
  return y;
}

// centre thc of the damped delta expansion with p(cos tha) = p(cos thb): safeguarded Newton
static double rootchb(int m, const std::vector<double>& jac, double tha, double thb) {
  std::vector<double> w(m + 1);
  for (int j = 0; j <= m; ++j) w[j] = jac[j] * (cos(j * tha) - cos(j * thb));
  auto f = [&](double t) { double s = 0; for (int j = 0; j <= m; ++j) s += w[j] * cos(j * t); return s; };
  auto df = [&](double t) { double s = 0; for (int j = 0; j <= m; ++j) s -= w[j] * j * sin(j * t); return s; };
  double lo = thb, hi = tha, flo = f(lo), fhi = f(hi);
  double thc = 0.5 * (tha + thb);
  const double tol = fabs(tha - thb) * 1.0e-13;
  if (flo * fhi > 0) return thc;
  for (int it = 0; it < 200; ++it) {
    const double fv = f(thc);
    if (fv == 0.0) break;
    if ((fv > 0) == (flo > 0)) { lo = thc; flo = fv; } else { hi = thc; fhi = fv; }
    const double dv = df(thc);
    double tn = dv != 0.0 ? thc - fv / dv : 0.5 * (lo + hi);
    if (!(lo < tn && tn < hi)) tn = 0.5 * (lo + hi);
    if (fabs(tn - thc) < tol) { thc = tn; break; }
    thc = tn;
  }
  return thc;
}

void nm_findpol(const double xintv[4], double thresh_int, double thresh_ext, NmPol& pol) {
  const int max_deg = 10000, min_deg = 2, damping = 2;
  const double intvtol = 1.0e-9;
  const double a = xintv[0], b = xintv[1], lmin = xintv[2], lmax = xintv[3];
  NM_REQUIRE(lmax > lmin, "findpol: empty spectrum interval [%g, %g]", lmin, lmax);
  NM_REQUIRE(b > a, "findpol: empty target interval [%g, %g]", a, b);
  NM_REQUIRE(b > lmin && a < lmax, "findpol: [%g, %g] does not intersect the spectrum [%g, %g]", a, b, lmin, lmax);
  for (int i = 0; i < 4; ++i) pol.intv[i] = xintv[i];
  const double cc = (lmax + lmin) / 2.0, dd = (lmax - lmin) / 2.0;
  pol.cc = cc; pol.dd = dd;
  const double aa = std::max(a, lmin), bb = std::min(b, lmax);
  double itv[2] = {std::max(-1.0, (aa - cc) / dd), std::min(1.0, (bb - cc) / dd)};
  const double tha = acos(itv[0]), thb = acos(itv[1]);
  std::vector<double> jac, mu;
  int m = min_deg;
  if (aa - intvtol <= lmin || bb + intvtol >= lmax) {
    const bool left = aa - intvtol <= lmin;
    const double thc = left ? tha : thb;
    const double xin = left ? itv[0] : itv[1], xout = left ? itv[1] : itv[0];
    double t = 1.0, v = 0.0;
    for (m = min_deg; m < max_deg; ++m) {
      dampcf(m, damping, jac);
      mu.resize(m + 1);
      for (int j = 0; j <= m; ++j) mu[j] = cos(j * thc) * jac[j];
      t = chebx(mu, m, xin); v = chebx(mu, m, xout);
      if (v <= t * thresh_ext) break;
    }
    if (m == max_deg) m = max_deg - 1;
    for (auto& x : mu) x /= t;
    pol.deg = m; pol.mu = mu; pol.gam = xin; pol.bar = v / t; pol.type = left ? 1 : 3;
    return;
  }
  double t = 1.0, va = 0.0, vb = 0.0, gam = 0.0;
  for (m = min_deg; m < max_deg; ++m) {
    dampcf(m, damping, jac);
    const double thc = rootchb(m, jac, tha, thb);
    mu.resize(m + 1);
    for (int j = 0; j <= m; ++j) mu[j] = cos(j * thc) * jac[j];
    gam = cos(thc);
    t = chebx(mu, m, gam); va = chebx(mu, m, itv[0]); vb = chebx(mu, m, itv[1]);
    if (va <= t * thresh_int && vb <= t * thresh_int) break;
  }
  if (m == max_deg) m = max_deg - 1;
  for (auto& x : mu) x /= t;
  pol.deg = m; pol.mu = mu; pol.gam = gam; pol.bar = std::min(va, vb) / t; pol.type = 2;
}

// ---------------------------------------------------------------- C ABI (host maths, no GPU needed)
extern "C" int nm_tridiag_eig_host(int k, const double* d, const double* e, double* w, double* Z, double* lastrow) {
  NM_API_BEGIN
  int rc = nm_tridiag_eig_ex(k, d, e, w, Z, lastrow);
  NM_REQUIRE(rc == 0, "tridiagonal QL failed to converge for eigenvalue %d", rc - 1);
  NM_API_END
}

extern "C" int nm_sym_geneig_host(int m, const double* H, const double* G, double* w, double* Cout) {
  NM_API_BEGIN
  int rc = nm_sym_geneig(m, H, G, w, Cout);
  NM_REQUIRE(rc == 0, "dense symmetric-definite eigensolver failed (%s)", rc == 1 ? "G not positive definite" : "Jacobi did not converge");
  NM_API_END
}

extern "C" int nm_findpol_create(const double* xintv, double thresh_int, double thresh_ext, void** out) {
  NM_API_BEGIN
  std::unique_ptr<NmPol> p(new NmPol());
  nm_findpol(xintv, thresh_int, thresh_ext, *p);
  *out = p.release();
  NM_API_END
}
extern "C" int nm_pol_info(void* h, int* deg, double* cc, double* dd, double* gam, double* bar, int* type) {
  NM_API_BEGIN
  NmPol& p = *(NmPol*)h;
  if (deg) *deg = p.deg;
  if (cc) *cc = p.cc;
  if (dd) *dd = p.dd;
  if (gam) *gam = p.gam;
  if (bar) *bar = p.bar;
  if (type) *type = p.type;
  NM_API_END
}
extern "C" int nm_pol_coeffs(void* h, double* mu /* deg+1 */) {
  NM_API_BEGIN
  NmPol& p = *(NmPol*)h;
  std::copy(p.mu.begin(), p.mu.end(), mu);
  NM_API_END
}
extern "C" int nm_pol_free(void* h) {
  NM_API_BEGIN
  delete (NmPol*)h;
  NM_API_END
}
