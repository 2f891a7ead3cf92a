// Element-local stiffness / mass integration and scatter into the reference's CSR pattern (K9):
// CGE3D_ISO (src/mod_cg_create_matrix.f90:980-1266) and CGFSE3D_ISO (:103-977) on the device.
// One warp integrates one element with its dense pNp x pNp operands in shared memory (a thread block
// = a batch of NM_ASM_WARPS elements); every local entry is added to its CSR slot, found by a binary
// search of the owned row (the reference's findorder, :1223-1263), with an fp64 atomicAdd.  Only rows
// of owned nodes are written, so no values are communicated (as in the reference, :414,711,1226).
#include "nm_fem.h"
#include <algorithm>

#define NM_ASM_WARPS 4
#define NM_TOL 1.0e-8            // pin%TOL (src/mod_para.f90:50)
#define NM_EPS0 1.0e-15          // src/mod_cg_create_matrix.f90:30

// ---------------------------------------------------------------- reference element (host)
// Lagrange basis on the reference's equispaced nodes of the tetrahedron (-1,-1,-1),(1,-1,-1),(-1,1,-1),
// (-1,-1,1) (src/mod_geometry.f90:2307-2523; node order t outer, s, r inner).  Mathematically the same
// MassM = invV^T invV, Drst = D_a invV and MassF = (V2D V2D^T)^-1 the reference builds from its
// orthonormal basis; here from a monomial Vandermonde matrix and exact Gauss quadrature.
struct RefElem {
  int pNp, Nfp;
  double M[100], D[3][100], MF[4][36];
  int Fmask[4][6], vord[4];
};

static void gauss_legendre(int n, double* x, double* w) {      // on [0,1]
  for (int i = 0; i < n; ++i) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1, p2 = 0;
      for (int j = 1; j <= n; ++j) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1);
      double z1 = z; z = z1 - p1 / pp;
      if (fabs(z - z1) < 1e-16) break;
    }
    x[i] = 0.5 * (1 - z); w[i] = 1.0 / ((1 - z * z) * pp * pp);
  }
}

static void invert(int n, std::vector<double>& a) {            // Gauss-Jordan with partial pivoting
  std::vector<double> inv((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (fabs(a[(size_t)r * n + c]) > fabs(a[(size_t)piv * n + c])) piv = r;
    for (int k = 0; k < n; ++k) { std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]); std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]); }
    const double d = 1.0 / a[(size_t)c * n + c];
    for (int k = 0; k < n; ++k) { a[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
    for (int r = 0; r < n; ++r) {
      if (r == c) continue;
      const double f = a[(size_t)r * n + c];
      if (f == 0) continue;
      for (int k = 0; k < n; ++k) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
    }
  }
  a = inv;
}

static void build_refelem(int porder, RefElem& R) {
  const int pNp = porder == 1 ? 4 : 10, Nfp = porder == 1 ? 3 : 6;
  R.pNp = pNp; R.Nfp = Nfp;
  double r[10], s[10], t[10];
  int n = 0;
  for (int it = 0; it <= porder; ++it)
    for (int is = 0; is <= porder - it; ++is)
      for (int ir = 0; ir <= porder - it - is; ++ir) {
        r[n] = -1.0 + 2.0 * ir / porder; s[n] = -1.0 + 2.0 * is / porder; t[n] = -1.0 + 2.0 * it / porder; ++n;
      }
  // monomials r^a s^b t^c, a+b+c <= porder
  int ea[10], eb[10], ec[10], nm = 0;
  for (int c = 0; c <= porder; ++c) for (int b = 0; b <= porder - c; ++b) for (int a = 0; a <= porder - b - c; ++a) { ea[nm] = a; eb[nm] = b; ec[nm] = c; ++nm; }
  auto mono = [&](int k, double x, double y, double z) { return pow(x, ea[k]) * pow(y, eb[k]) * pow(z, ec[k]); };
  std::vector<double> C((size_t)pNp * pNp);                    // V[m][k] = mono_k(node m);  phi_n = sum_k C[k][n] mono_k
  for (int m = 0; m < pNp; ++m) for (int k = 0; k < pNp; ++k) C[(size_t)m * pNp + k] = mono(k, r[m], s[m], t[m]);
  invert(pNp, C);
  auto phi = [&](int nn_, double x, double y, double z) { double v = 0; for (int k = 0; k < pNp; ++k) v += C[(size_t)k * pNp + nn_] * mono(k, x, y, z); return v; };
  auto dphi = [&](int nn_, int dir, double x, double y, double z) {
    double v = 0;
    for (int k = 0; k < pNp; ++k) {
      int e[3] = {ea[k], eb[k], ec[k]};
      if (e[dir] == 0) continue;
      double c = e[dir]; e[dir]--;
      v += C[(size_t)k * pNp + nn_] * c * pow(x, e[0]) * pow(y, e[1]) * pow(z, e[2]);
    }
    return v;
  };
  // mass matrix: Duffy-collapsed Gauss quadrature on the unit simplex mapped to the reference tet (volume 4/3)
  const int nq = 6;
  double gx[nq], gw[nq];
  gauss_legendre(nq, gx, gw);
  for (int i = 0; i < pNp * pNp; ++i) R.M[i] = 0;
  for (int a = 0; a < nq; ++a) for (int b = 0; b < nq; ++b) for (int c = 0; c < nq; ++c) {
    const double u = gx[a], v = gx[b], w = gx[c];
    const double X = u, Y = v * (1 - u), Z = w * (1 - u) * (1 - v);           // unit simplex
    const double jac = (1 - u) * (1 - u) * (1 - v) * gw[a] * gw[b] * gw[c] * 8.0;   // d(r,s,t) = 8 d(X,Y,Z)
    const double x = 2 * X - 1, y = 2 * Y - 1, z = 2 * Z - 1;
    double ph[10];
    for (int m = 0; m < pNp; ++m) ph[m] = phi(m, x, y, z);
    for (int m = 0; m < pNp; ++m) for (int k = 0; k < pNp; ++k) R.M[m * pNp + k] += jac * ph[m] * ph[k];
  }
  for (int dir = 0; dir < 3; ++dir)
    for (int m = 0; m < pNp; ++m) for (int k = 0; k < pNp; ++k) R.D[dir][m * pNp + k] = dphi(k, dir, r[m], s[m], t[m]);
  // faces (src/mod_geometry.f90:2392-2405): 0: r+s+t=-1 ; 1: r=-1 ; 2: s=-1 ; 3: t=-1
  int cnt[4] = {0, 0, 0, 0};
  for (int m = 0; m < pNp; ++m) {
    if (fabs(1 + r[m]) <= NM_TOL) R.Fmask[1][cnt[1]++] = m;
    if (fabs(1 + s[m]) <= NM_TOL) R.Fmask[2][cnt[2]++] = m;
    if (fabs(1 + t[m]) <= NM_TOL) R.Fmask[3][cnt[3]++] = m;
    if (fabs(1 + r[m] + s[m] + t[m]) <= NM_TOL) R.Fmask[0][cnt[0]++] = m;
  }
  // face mass matrices on the reference triangle (-1,-1),(1,-1),(-1,1) in the face's coordinate pair (:2491-2510)
  for (int f = 0; f < 4; ++f) {
    double fa[6], fb[6];
    for (int k = 0; k < Nfp; ++k) {
      const int m = R.Fmask[f][k];
      fa[k] = (f <= 1) ? s[m] : r[m];
      fb[k] = (f <= 2) ? t[m] : s[m];
      if (f == 2) { fa[k] = r[m]; fb[k] = t[m]; }
      if (f == 3) { fa[k] = r[m]; fb[k] = s[m]; }
    }
    int pa[6], pb[6], n2 = 0;
    for (int b = 0; b <= porder; ++b) for (int a = 0; a <= porder - b; ++a) { pa[n2] = a; pb[n2] = b; ++n2; }
    std::vector<double> C2((size_t)Nfp * Nfp);
    for (int m = 0; m < Nfp; ++m) for (int k = 0; k < Nfp; ++k) C2[(size_t)m * Nfp + k] = pow(fa[m], pa[k]) * pow(fb[m], pb[k]);
    invert(Nfp, C2);
    for (int i = 0; i < Nfp * Nfp; ++i) R.MF[f][i] = 0;
    for (int a = 0; a < nq; ++a) for (int b = 0; b < nq; ++b) {
      const double u = gx[a], v = gx[b];
      const double X = u, Y = v * (1 - u);
      const double jac = (1 - u) * gw[a] * gw[b] * 4.0;
      const double x = 2 * X - 1, y = 2 * Y - 1;
      double ph[6];
      for (int m = 0; m < Nfp; ++m) { ph[m] = 0; for (int k = 0; k < Nfp; ++k) ph[m] += C2[(size_t)k * Nfp + m] * pow(x, pa[k]) * pow(y, pb[k]); }
      for (int m = 0; m < Nfp; ++m) for (int k = 0; k < Nfp; ++k) R.MF[f][m * Nfp + k] += jac * ph[m] * ph[k];
    }
  }
  const int v1[4] = {0, 1, 2, 3}, v2[4] = {0, 2, 5, 9};        // local ids of the 4 vertices (:2514-2518)
  for (int k = 0; k < 4; ++k) R.vord[k] = porder == 1 ? v1[k] : v2[k];
}

// ---------------------------------------------------------------- device-side views
struct DevMat {
  const int* ia; const int* ja; double* val;
  int row0, nrow;
};
struct DevFem {
  int pNp, Nfp, selfG, fluidcase, purefluid;
  const int* lelist; int nle;
  const int* ele; const int* neigh; const double* node; const int* t2n;
  const int* vstt; const int* vnum; const int* pstt;
  const double* vp; const double* vs; const double* rho; const double* g0;
  DevMat mat[NM_NMAT];
  int* err;
};
__constant__ RefElem c_ref;

// position of global column `col` in local row `lrow` (columns sorted ascending) or -1
__device__ __forceinline__ int find_pos(const DevMat& A, int lrow, int col) {
  int lo = A.ia[lrow], hi = A.ia[lrow + 1] - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int c = A.ja[mid];
    if (c == col) return mid;
    if (c < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
__device__ __forceinline__ void add_entry(const DevFem& F, int which, int grow, int gcol, double v) {
  const DevMat& A = F.mat[which];
  const int lrow = grow - A.row0;
  if (lrow < 0 || lrow >= A.nrow) return;                    // row owned by another rank
  const int p = find_pos(A, lrow, gcol);
  if (p < 0) { atomicAdd(F.err, 1); return; }                // "error: can not find the id" (src/mod_utility.f90:1011)
  atomicAdd(A.val + p, v);
}

struct Geo {
  double invJ[3][3], detJ, nrm[4][3], sJ[4], X[4][3];
};
__device__ void element_geometry(const DevFem& F, int e, Geo& G) {      // src/mod_geometry.f90:2216-2304,1401-1428
  for (int k = 0; k < 4; ++k) for (int c = 0; c < 3; ++c) G.X[k][c] = F.node[3 * (size_t)F.ele[4 * e + k] + c];
  double B[3][3];
  for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) B[a][c] = G.X[a + 1][c] - G.X[0][c];
  const double det = B[0][0] * (B[1][1] * B[2][2] - B[2][1] * B[1][2]) + B[0][1] * (B[2][0] * B[1][2] - B[1][0] * B[2][2]) +
                     B[0][2] * (B[1][0] * B[2][1] - B[2][0] * B[1][1]);
  const double id = 2.0 / det;                               // invJ = 2 * inv(B)
  G.invJ[0][0] = (B[1][1] * B[2][2] - B[1][2] * B[2][1]) * id; G.invJ[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id; G.invJ[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
  G.invJ[1][0] = (B[1][2] * B[2][0] - B[1][0] * B[2][2]) * id; G.invJ[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id; G.invJ[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
  G.invJ[2][0] = (B[1][0] * B[2][1] - B[1][1] * B[2][0]) * id; G.invJ[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id; G.invJ[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
  G.detJ = det / 8.0;
  for (int f = 0; f < 4; ++f) {
    int o[3], k = 0;
    for (int j = 0; j < 4; ++j) if (j != f) o[k++] = j;
    double a[3], b[3], c[3];
    for (int d = 0; d < 3; ++d) { a[d] = G.X[o[1]][d] - G.X[o[0]][d]; b[d] = G.X[o[2]][d] - G.X[o[0]][d]; }
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
    const double len = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    double dotp = 0;
    for (int d = 0; d < 3; ++d) { c[d] /= len; dotp += (G.X[f][d] - G.X[o[0]][d]) * c[d]; }
    const double sg = dotp > 0 ? -1.0 : 1.0;
    for (int d = 0; d < 3; ++d) G.nrm[f][d] = sg * c[d];
    G.sJ[f] = len / 4.0;
  }
}

// physical coordinates of local node m: X0 + (B/2)^T (ref + 1)
__device__ __forceinline__ void node_coord(const Geo& G, int pNp, int m, double* x) {
  // reference coordinates of the equispaced nodes, regenerated from the index (same order as build_refelem)
  const int porder = pNp == 4 ? 1 : 2;
  int n = 0, ir_ = 0, is_ = 0, it_ = 0;
  for (int it = 0; it <= porder; ++it) for (int is = 0; is <= porder - it; ++is) for (int ir = 0; ir <= porder - it - is; ++ir) { if (n == m) { ir_ = ir; is_ = is; it_ = it; } ++n; }
  const double r = 2.0 * ir_ / porder, s = 2.0 * is_ / porder, t = 2.0 * it_ / porder;   // ref + 1
  for (int d = 0; d < 3; ++d)
    x[d] = G.X[0][d] + 0.5 * ((G.X[1][d] - G.X[0][d]) * r + (G.X[2][d] - G.X[0][d]) * s + (G.X[3][d] - G.X[0][d]) * t);
}

// least-squares gradient of ncomp nodal fields (:1063-1078, :197-218): grad[i][c] = d f_c / d x_i
__device__ void grad_ls(const double (*nods)[3], const double* f, int ncomp, int pNp, double* grad) {
  double mean[3] = {0, 0, 0}, fm[3] = {0, 0, 0};
  for (int m = 0; m < pNp; ++m) { for (int d = 0; d < 3; ++d) mean[d] += nods[m][d]; for (int c = 0; c < ncomp; ++c) fm[c] += f[m * ncomp + c]; }
  for (int d = 0; d < 3; ++d) mean[d] /= pNp;
  for (int c = 0; c < ncomp; ++c) fm[c] /= pNp;
  double nd[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, rhs[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int m = 0; m < pNp; ++m) {
    double dm[3];
    for (int d = 0; d < 3; ++d) dm[d] = nods[m][d] - mean[d];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) nd[i][j] += dm[i] * dm[j];
      for (int c = 0; c < ncomp; ++c) rhs[i][c] += dm[i] * (f[m * ncomp + c] - fm[c]);
    }
  }
  const double det = nd[0][0] * (nd[1][1] * nd[2][2] - nd[1][2] * nd[2][1]) - nd[0][1] * (nd[1][0] * nd[2][2] - nd[1][2] * nd[2][0]) +
                     nd[0][2] * (nd[1][0] * nd[2][1] - nd[1][1] * nd[2][0]);
  double inv[3][3];
  inv[0][0] = (nd[1][1] * nd[2][2] - nd[1][2] * nd[2][1]) / det; inv[0][1] = (nd[0][2] * nd[2][1] - nd[0][1] * nd[2][2]) / det; inv[0][2] = (nd[0][1] * nd[1][2] - nd[0][2] * nd[1][1]) / det;
  inv[1][0] = (nd[1][2] * nd[2][0] - nd[1][0] * nd[2][2]) / det; inv[1][1] = (nd[0][0] * nd[2][2] - nd[0][2] * nd[2][0]) / det; inv[1][2] = (nd[0][2] * nd[1][0] - nd[0][0] * nd[1][2]) / det;
  inv[2][0] = (nd[1][0] * nd[2][1] - nd[1][1] * nd[2][0]) / det; inv[2][1] = (nd[0][1] * nd[2][0] - nd[0][0] * nd[2][1]) / det; inv[2][2] = (nd[0][0] * nd[1][1] - nd[0][1] * nd[1][0]) / det;
  for (int i = 0; i < 3; ++i) for (int c = 0; c < ncomp; ++c) grad[i * ncomp + c] = inv[i][0] * rhs[0][c] + inv[i][1] * rhs[1][c] + inv[i][2] * rhs[2][c];
}

// ---------------------------------------------------------------- the element kernel
// Shared memory per warp (doubles): D[3], OP1[3], OP2[3], MD[3] (pNp^2 each) + per-node vectors.
template <int PNP>
__global__ void __launch_bounds__(NM_ASM_WARPS * 32)
k_assemble(DevFem F) {
  constexpr int NN = PNP * PNP;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int le = blockIdx.x * NM_ASM_WARPS + warp;
  if (le >= F.nle) return;
  const int e = F.lelist[le];
  double* W = smem + (size_t)warp * (12 * NN + 16 * PNP);
  double* D = W;                 // [3][NN]  D_i[k][n] = d phi_n / d x_i at node k
  double* OP1 = D + 3 * NN;      // Ll D_i   (fluid: unused)
  double* OP2 = OP1 + 3 * NN;    // Lm D_i
  double* MD = OP2 + 3 * NN;     // M D_i
  double* lam = MD + 3 * NN;     // per-node vectors
  double* mu = lam + PNP;
  double* rho = mu + PNP;
  double* gk = rho + PNP;        // [PNP][3] gravity (km/s^2)
  double* nod = gk + 3 * PNP;    // [PNP][3] node coordinates
  double* aux = nod + 3 * PNP;   // [PNP][4] scratch (normalg*N2*rho etc.)
  __shared__ Geo sgeo[NM_ASM_WARPS];
  __shared__ double sgrad[NM_ASM_WARPS][12];
  Geo& G = sgeo[warp];
  if (lane == 0) element_geometry(F, e, G);
  __syncwarp();
  const double* M = c_ref.M;
  double mumax = 0.0;
  for (int m = lane; m < PNP; m += 32) {
    const double r = F.rho[(size_t)e * PNP + m], vs = F.vs[(size_t)e * PNP + m], vp = F.vp[(size_t)e * PNP + m];
    rho[m] = r; mu[m] = r * vs * vs; lam[m] = r * vp * vp - 2.0 * mu[m];              // :1019-1020
    double x[3];
    node_coord(G, PNP, m, x);
    for (int d = 0; d < 3; ++d) { nod[3 * m + d] = x[d]; gk[3 * m + d] = F.selfG ? F.g0[((size_t)e * PNP + m) * 3 + d] / 1.0e3 : 0.0; }
  }
  __syncwarp();
  for (int m = 0; m < PNP; ++m) mumax = fmax(mumax, mu[m]);
  const bool solid = !F.fluidcase || mumax >= NM_TOL;                                    // :279
  // D_i = sum_a invJ[i][a] Drst[a]
  for (int t = lane; t < 3 * NN; t += 32) {
    const int i = t / NN, mn = t % NN;
    D[t] = G.invJ[i][0] * c_ref.D[0][mn] + G.invJ[i][1] * c_ref.D[1][mn] + G.invJ[i][2] * c_ref.D[2][mn];
  }
  __syncwarp();
  // MD_i = M D_i ; OP1_i = Ll D_i ; OP2_i = Lm D_i with Ll = (M diag(lam) + diag(lam) M)/2 (:1112-1121)
  for (int t = lane; t < 3 * NN; t += 32) {
    const int i = t / NN, m = (t % NN) / PNP, n = t % PNP;
    double s0 = 0, s1 = 0, s2 = 0;
    for (int k = 0; k < PNP; ++k) {
      const double mk = M[m * PNP + k], d = D[i * NN + k * PNP + n];
      s0 += mk * d;
      s1 += (mk * lam[k] + M[k * PNP + m] * lam[m]) / 2.0 * d;
      s2 += (mk * mu[k] + M[k * PNP + m] * mu[m]) / 2.0 * d;
    }
    MD[t] = s0; OP1[t] = s1; OP2[t] = s2;
  }
  __syncwarp();
  const int* nd = F.t2n + (size_t)e * PNP;
  const double detJ = G.detJ;
  const int A_ = NM_MAT_A, B_ = NM_MAT_B;
  if (solid) {
    // gravity: least-squares gradient of g (:1063-1078)
    if (F.selfG && lane == 0) grad_ls((const double(*)[3])nod, gk, 3, PNP, sgrad[warp]);
    __syncwarp();
    const double* dg = sgrad[warp];                          // dg[i*3+c] = d g_c / d x_i
    double rhoavg = 0;
    for (int m = 0; m < PNP; ++m) rhoavg += rho[m];
    rhoavg /= PNP;
    for (int t = lane; t < NN; t += 32) {
      const int m = t / PNP, n = t % PNP;
      // G1_ij[m][n] = sum_k D_i[k][m] OP1_j[k][n] ; G2 likewise with OP2
      double G1[3][3], G2[3][3], G1t[3][3], G2t[3][3];       // *t: entry [n][m]
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double a = 0, b = 0, at = 0, bt = 0;
        for (int k = 0; k < PNP; ++k) {
          const double dim = D[i * NN + k * PNP + m], din = D[i * NN + k * PNP + n];
          a += dim * OP1[j * NN + k * PNP + n]; b += dim * OP2[j * NN + k * PNP + n];
          at += din * OP1[j * NN + k * PNP + m]; bt += din * OP2[j * NN + k * PNP + m];
        }
        G1[i][j] = a; G2[i][j] = b; G1t[i][j] = at; G2t[i][j] = bt;
      }
      const double OPs = G2[0][0] + G2[1][1] + G2[2][2], OPst = G2t[0][0] + G2t[1][1] + G2t[2][2];
      const int rm = F.vstt[nd[m]], cn = F.vstt[nd[n]];      // solid-side triples (roff = coff = 0)
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        double X;
        if (i == j) X = ((OPs + G2[i][i] + G1[i][i]) + (OPst + G2t[i][i] + G1t[i][i])) / 2.0;          // :1187-1196
        else X = ((G1t[j][i] + G2t[i][j]) + (G1[i][j] + G2[j][i])) / 2.0;                              // :1201-1212
        if (F.selfG) {                                                                                  // OPrho (:1132-1177)
          double b1 = 0, b2 = 0;
          for (int k = 0; k < PNP; ++k) {
            b1 += D[j * NN + k * PNP + m] * gk[3 * k + i] * M[k * PNP + n];
            b2 += M[m * PNP + k] * gk[3 * k + j] * D[i * NN + k * PNP + n];
          }
          double O = (MD[i * NN + n * PNP + m] * gk[3 * n + j] + gk[3 * m + i] * MD[j * NN + m * PNP + n]) / 2.0;
          O -= M[m * PNP + n] * (dg[i * 3 + j] + dg[j * 3 + i]) / 2.0;
          O -= (b1 + b2) / 2.0;
          X += O * (rho[m] + rho[n]) / 2.0;
        }
        add_entry(F, A_, rm + i, cn + j, X * detJ);
      }
      // mass: CGE3D_ISO (sym, :1089-1091) vs CGFSE3D_ISO solid branch (mean rho, :293)
      const double mm = (F.fluidcase ? M[m * PNP + n] * rhoavg : (rho[m] * M[m * PNP + n] + M[m * PNP + n] * rho[n]) / 2.0) * detJ;
      for (int p = 0; p < 3; ++p) add_entry(F, B_, rm + p, cn + p, mm);
    }
    return;
  }
  // ------------------------------------------------------------ fluid element (:454-801)
  double rhoavg = 0;
  for (int m = 0; m < PNP; ++m) rhoavg += rho[m];
  rhoavg /= PNP;
  // aux[m][0..2] = normalg ; aux[m][3] = N2*rho  (selfG)
  if (F.selfG) {
    if (lane == 0) {
      double f[10];
      for (int m = 0; m < PNP; ++m) f[m] = rho[m] - rhoavg;
      grad_ls((const double(*)[3])nod, f, 1, PNP, sgrad[warp]);                                        // drho0 (:216-218)
    }
    __syncwarp();
    for (int m = lane; m < PNP; m += 32) {
      const double g0 = gk[3 * m], g1 = gk[3 * m + 1], g2 = gk[3 * m + 2];
      const double normg = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
      double N2 = sgrad[warp][0] * g0 + sgrad[warp][1] * g1 + sgrad[warp][2] * g2;
      const double den = fmax(normg, NM_EPS0);
      aux[4 * m] = g0 / den; aux[4 * m + 1] = g1 / den; aux[4 * m + 2] = g2 / den;
      N2 = N2 / rho[m] - normg * normg / lam[m] * rho[m];                                               // :246
      if (normg < NM_EPS0 || F.purefluid) N2 = 0.0;
      aux[4 * m + 3] = N2 * rho[m];
    }
  } else {
    for (int m = lane; m < PNP; m += 32) aux[4 * m] = aux[4 * m + 1] = aux[4 * m + 2] = aux[4 * m + 3] = 0.0;
  }
  __syncwarp();
  int v6[4], cnt6 = 0;
  for (int k = 0; k < 4; ++k) { v6[k] = F.vnum[nd[c_ref.vord[k]]] == 6; cnt6 += v6[k]; }
  for (int t = lane; t < NN; t += 32) {
    const int m = t / PNP, n = t % PNP;
    const double Mmn = M[m * PNP + n];
    const int fom = F.vnum[nd[m]] - 3, fon = F.vnum[nd[n]] - 3;
    const int rm = F.vstt[nd[m]] + fom, cn = F.vstt[nd[n]] + fon;   // fluid-side triples
    const int pm = F.pstt[nd[m]], pn = F.pstt[nd[n]];
    // which faces contain both m and n (face terms are Fmask x Fmask blocks)
    double FTpp = -(1.0 / sqrt(lam[m])) * Mmn * (1.0 / sqrt(lam[n])) * detJ;                            // :470-472,638-639
    double FTuu_face[3] = {0, 0, 0};
    for (int f = 0; f < 4; ++f) {
      int km = -1, kn = -1;
      for (int k = 0; k < c_ref.Nfp; ++k) { if (c_ref.Fmask[f][k] == m) km = k; if (c_ref.Fmask[f][k] == n) kn = k; }
      if (km < 0 || kn < 0) continue;
      double surfrho = 0, sgn_abs = 0, sgn = 0;
      for (int k = 0; k < c_ref.Nfp; ++k) {
        const int q = c_ref.Fmask[f][k];
        surfrho += rho[q];
        const double gn = gk[3 * q] * G.nrm[f][0] + gk[3 * q + 1] * G.nrm[f][1] + gk[3 * q + 2] * G.nrm[f][2];
        sgn_abs += sqrt(gn * gn); sgn += gn;
      }
      surfrho /= c_ref.Nfp; sgn_abs /= c_ref.Nfp; sgn /= c_ref.Nfp;
      const double mf = c_ref.MF[f][km * c_ref.Nfp + kn];
      if (F.neigh[4 * e + f] < 0) {                                                                     // free surface (:644-666)
        if (sgn_abs == 0.0) { atomicAdd(F.err + 1, 1); continue; }
        FTpp += -(mf / sgn_abs / surfrho) * G.sJ[f];
      } else if (cnt6 - v6[f] < 3) {                                                                    // interior face (:667-703)
        const double sp = mf * (sgn * surfrho);
        for (int j = 0; j < 3; ++j) FTuu_face[j] += sp * (G.sJ[f] * G.nrm[f][j] * G.nrm[f][j]);
      }
    }
    add_entry(F, NM_MAT_AP, pm, pn, FTpp);                                                              // :715-728
    const double rinl_n = rho[n] / lam[n];
    for (int i = 0; i < 3; ++i) {
      // FTup_i[m][n] = (M D_i)[m][n] - M[m][n]*rho[n]/lam[n]*g_i[n]   (:521-525,550-552)
      double up = MD[i * NN + m * PNP + n];
      if (F.selfG) up -= Mmn * rinl_n * gk[3 * n + i];
      up *= detJ;
      add_entry(F, NM_MAT_E, rm + i, pn, up);                                                           // :781-794
      // ET row of pressure node n?  ET(p_m', u_q n') = FTup_q[n'][m'] : emit with roles swapped
      add_entry(F, NM_MAT_ET, pn, rm + i, up);                                                          // :730-746
      for (int j = 0; j < 3; ++j) {
        double O = 0.0;
        if (F.selfG) {
          const double Ni_m = aux[4 * m + i], Rj_n = aux[4 * n + j] * aux[4 * n + 3];
          if (i == j) O = (Ni_m * Mmn * Rj_n + aux[4 * m + j] * aux[4 * m + 3] * Mmn * aux[4 * n + i]) / 2.0;      // :497-511
          else O = (Ni_m * Mmn * Rj_n + aux[4 * m + i] * aux[4 * m + 3] * Mmn * aux[4 * n + j]) / 2.0;
        }
        O *= detJ;
        if (i == j) O += FTuu_face[j];
        if (O != 0.0 || F.selfG || i == j) add_entry(F, A_, rm + i, cn + j, O);                         // :749-765
      }
      if (i == 0) add_entry(F, B_, rm, cn, Mmn * rhoavg * detJ);   // :767-779; components 1,2: k_mass_replicate
    }
  }
  // ------------------------------------------------------------ fluid-solid interface face (:804-952)
  if (cnt6 == 3) {
    int fc = 0;
    for (int k = 0; k < 4; ++k) if (!v6[k]) fc = k;
    const int Nfp = c_ref.Nfp;
    double rhof = 0;
    for (int k = 0; k < Nfp; ++k) rhof += rho[c_ref.Fmask[fc][k]];
    rhof /= Nfp;
    const double sj = G.sJ[fc];
    for (int t = lane; t < Nfp * Nfp; t += 32) {
      const int a = t / Nfp, b = t % Nfp;
      const int ma = c_ref.Fmask[fc][a], mb = c_ref.Fmask[fc][b];
      const int na = nd[ma], nb = nd[mb];
      const double mf = c_ref.MF[fc][a * Nfp + b];
      for (int i = 0; i < 3; ++i) {
        if (F.selfG)
          for (int j = 0; j < 3; ++j) {                                                                 // :829-892
            const double scm = sj * mf * (G.nrm[fc][i] * gk[3 * mb + j] + G.nrm[fc][j] * gk[3 * ma + i]) / 2.0 * rhof;
            add_entry(F, A_, F.vstt[na] + i, F.vstt[nb] + j, -scm);
          }
        // ET(p_a, u_i b [solid side]) -= n_i sJac MassF(b,a) ; E(u_i a [solid side], p_b) -= n_i sJac MassF(a,b)
        add_entry(F, NM_MAT_ET, F.pstt[na], F.vstt[nb] + i, -c_ref.MF[fc][b * Nfp + a] * G.nrm[fc][i] * sj);   // :895-924
        add_entry(F, NM_MAT_E, F.vstt[na] + i, F.pstt[nb], -mf * G.nrm[fc][i] * sj);                            // :926-948
      }
    }
  }
}

// B = M (x) I3: the reference adds the SAME value to the three component rows of a node in the same element
// order (:1247-1259, :767-779), so those rows are bit-identical.  The element kernel accumulates component 0
// only; this kernel copies it to components 1 and 2 (which also keeps the three rows identical under the
// unordered fp64 atomics, so the parcsr layer can store M once).
__global__ void k_mass_replicate(DevMat B, int* err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (3 * b + 2 >= B.nrow) return;
  const int s0 = B.ia[3 * b], s1 = B.ia[3 * b + 1], s2 = B.ia[3 * b + 2], s3 = B.ia[3 * b + 3];
  const int len = s1 - s0;
  if (s2 - s1 != len || s3 - s2 != len) { atomicAdd(err, 1); return; }
  for (int u = 0; u < len; ++u) {
    const int c = B.ja[s0 + u];
    if (B.ja[s1 + u] != c + 1 || B.ja[s2 + u] != c + 2) { atomicAdd(err, 1); return; }
    const double v = B.val[s0 + u];
    B.val[s1 + u] = v;
    B.val[s2 + u] = v;
  }
}

// ---------------------------------------------------------------- host driver
void nm_fem_assemble(NmFem& F, int job, const double* vp, const double* vs, const double* rho, const double* g0) {
  nm_ensure_init();
  NmCtx& c = nm_ctx();
  const int pNp = F.pNp;
  const bool selfG = job >= 2;
  NM_REQUIRE(!selfG || g0, "JOB >= 2 needs the reference gravity g0 (src/mod_cg_models.f90:402-420)");
  RefElem R;
  build_refelem(F.porder, R);
  NM_CUDA(cudaMemcpyToSymbol(c_ref, &R, sizeof(RefElem)));
  const size_t ne = F.ntet;
  DBuf<int> d_le, d_ele, d_neigh, d_t2n, d_vstt, d_vnum, d_pstt, d_err(2);
  DBuf<double> d_node, d_vp, d_vs, d_rho, d_g0;
  d_le.from_host(F.lelist); d_ele.from_host(F.ele); d_neigh.from_host(F.neigh); d_t2n.from_host(F.t2n);
  d_vstt.from_host(F.vstt); d_vnum.from_host(F.vnum); d_pstt.from_host(F.pstt);
  d_node.from_host(F.node);
  d_vp.alloc(ne * pNp); d_vp.upload(vp, ne * pNp);
  d_vs.alloc(ne * pNp); d_vs.upload(vs, ne * pNp);
  d_rho.alloc(ne * pNp); d_rho.upload(rho, ne * pNp);
  if (selfG) { d_g0.alloc(ne * pNp * 3); d_g0.upload(g0, ne * pNp * 3); }
  d_err.zero();
  DevFem D;
  memset(&D, 0, sizeof(D));
  D.pNp = pNp; D.Nfp = R.Nfp; D.selfG = selfG; D.fluidcase = F.fluidcase; D.purefluid = F.purefluid;
  D.lelist = d_le.p; D.nle = (int)F.lelist.size();
  D.ele = d_ele.p; D.neigh = d_neigh.p; D.node = d_node.p; D.t2n = d_t2n.p;
  D.vstt = d_vstt.p; D.vnum = d_vnum.p; D.pstt = d_pstt.p;
  D.vp = d_vp.p; D.vs = d_vs.p; D.rho = d_rho.p; D.g0 = d_g0.p; D.err = d_err.p;
  DBuf<int> d_ia[NM_NMAT], d_ja[NM_NMAT];
  DBuf<double> d_val[NM_NMAT];
  for (int k = 0; k < NM_NMAT; ++k) {
    NmPattern& P = F.pat[k];
    if (!P.present) continue;
    d_ia[k].from_host(P.ia);
    d_ja[k].alloc(std::max<size_t>(P.ja.size(), 1)); d_ja[k].upload(P.ja.data(), P.ja.size());
    d_val[k].alloc(std::max<size_t>(P.ja.size(), 1)); d_val[k].zero();
    D.mat[k].ia = d_ia[k].p; D.mat[k].ja = d_ja[k].p; D.mat[k].val = d_val[k].p;
    D.mat[k].row0 = P.rowdist[F.rank]; D.mat[k].nrow = P.nrow;
  }
  if (D.nle > 0) {
    const size_t smem = (size_t)NM_ASM_WARPS * (12 * pNp * pNp + 16 * pNp) * sizeof(double);
    const int grid = nm_div_up(D.nle, NM_ASM_WARPS);
    if (pNp == 4) {
      k_assemble<4><<<grid, NM_ASM_WARPS * 32, smem, c.stream>>>(D);
    } else {
      NM_CUDA(cudaFuncSetAttribute(k_assemble<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_assemble<10><<<grid, NM_ASM_WARPS * 32, smem, c.stream>>>(D);
    }
    c.launches++;
    NM_CUDA(cudaGetLastError());
  }
  DBuf<int> d_err3(1);
  d_err3.zero();
  if (F.pat[NM_MAT_B].present && F.pat[NM_MAT_B].nrow > 0) {
    NM_REQUIRE(F.pat[NM_MAT_B].nrow % 3 == 0, "assembly: B rows are not node triples");
    k_mass_replicate<<<nm_div_up(F.pat[NM_MAT_B].nrow / 3, 128), 128, 0, c.stream>>>(D.mat[NM_MAT_B], d_err3.p);
    c.launches++;
  }
  int herr[2] = {0, 0}, herr3 = 0;
  d_err.download(herr, 2);
  d_err3.download(&herr3, 1);
  NM_REQUIRE(herr3 == 0, "assembly: %d node triples of B do not share one column list (M (x) I3 pattern expected)", herr3);
  NM_REQUIRE(herr[0] == 0, "assembly: %d element entries have no slot in the CSR pattern (error: can not find the id)", herr[0]);
  NM_REQUIRE(herr[1] == 0, "assembly: fluid free-surface face without gravity is undefined in the reference "
                           "(src/mod_cg_create_matrix.f90:656-657); use JOB 2 for models with a fluid surface");
  for (int k = 0; k < NM_NMAT; ++k) {
    NmPattern& P = F.pat[k];
    if (!P.present) continue;
    P.val.resize(P.ja.size());
    d_val[k].download(P.val.data(), P.val.size());
  }
}

// ---------------------------------------------------------------- C ABI
// reference-element matrices (host only): M [pNp*pNp], D [3*pNp*pNp], MF [4*Nfp*Nfp], Fmask [4*Nfp]
extern "C" int nm_refelem_get(int porder, double* M, double* D, double* MF, int* Fmask) {
  NM_API_BEGIN
  NM_REQUIRE(porder == 1 || porder == 2, "pOrder must be 1 or 2");
  RefElem R;
  build_refelem(porder, R);
  const int n2 = R.pNp * R.pNp, f2 = R.Nfp * R.Nfp;
  if (M) std::copy(R.M, R.M + n2, M);
  if (D) for (int a = 0; a < 3; ++a) std::copy(R.D[a], R.D[a] + n2, D + a * n2);
  if (MF) for (int f = 0; f < 4; ++f) std::copy(R.MF[f], R.MF[f] + f2, MF + f * f2);
  if (Fmask) for (int f = 0; f < 4; ++f) std::copy(R.Fmask[f], R.Fmask[f] + R.Nfp, Fmask + f * R.Nfp);
  NM_API_END
}
extern "C" int nm_fem_create(int ntet, int nvert, const int* ele, const int* neigh, const double* node, int porder,
                             const double* vs, int nproc, const int* part, int rank, void** out) {
  NM_API_BEGIN
  *out = nm_fem_build(ntet, nvert, ele, neigh, node, porder, vs, nproc, part, rank);
  NM_API_END
}
extern "C" int nm_fem_free(void* h) {
  NM_API_BEGIN
  delete (NmFem*)h;
  NM_API_END
}
extern "C" int nm_fem_info(void* h, int* nn, int* N, int* Np, int* fluidcase, int* nle) {
  NM_API_BEGIN
  NmFem& F = *(NmFem*)h;
  if (nn) *nn = F.nn;
  if (N) *N = F.N;
  if (Np) *Np = F.Np;
  if (fluidcase) *fluidcase = F.fluidcase ? (F.purefluid ? 2 : 1) : 0;
  if (nle) *nle = (int)F.lelist.size();
  NM_API_END
}
// which: 0 A/Ad, 1 B, 2 E, 3 ET, 4 Ap.  Sizes first (pointers may be NULL), then the arrays.
extern "C" int nm_fem_matrix_sizes(void* h, int which, int* present, int* nrow_local, long long* nnz_local) {
  NM_API_BEGIN
  NmFem& F = *(NmFem*)h;
  NM_REQUIRE(which >= 0 && which < NM_NMAT, "matrix id %d out of range", which);
  NmPattern& P = F.pat[which];
  if (present) *present = P.present;
  if (nrow_local) *nrow_local = P.nrow;
  if (nnz_local) *nnz_local = (long long)P.ja.size();
  NM_API_END
}
extern "C" int nm_fem_matrix_get(void* h, int which, int* rowdist, int* coldist, int* ia, int* ja, double* val) {
  NM_API_BEGIN
  NmFem& F = *(NmFem*)h;
  NM_REQUIRE(which >= 0 && which < NM_NMAT && F.pat[which].present, "matrix %d not present", which);
  NmPattern& P = F.pat[which];
  if (rowdist) std::copy(P.rowdist.begin(), P.rowdist.end(), rowdist);
  if (coldist) std::copy(P.coldist.begin(), P.coldist.end(), coldist);
  if (ia) std::copy(P.ia.begin(), P.ia.end(), ia);
  if (ja) std::copy(P.ja.begin(), P.ja.end(), ja);
  if (val) {
    NM_REQUIRE(P.val.size() == P.ja.size(), "matrix %d has no values yet (call nm_fem_assemble)", which);
    std::copy(P.val.begin(), P.val.end(), val);
  }
  NM_API_END
}
// node-level numbering arrays, each [nn] (NULL to skip): vstat, vnum, pnum, vstt, pstt, order
extern "C" int nm_fem_numbering(void* h, int* vstat, int* vnum, int* pnum, int* vstt, int* pstt, int* order) {
  NM_API_BEGIN
  NmFem& F = *(NmFem*)h;
  if (vstat) std::copy(F.vstat.begin(), F.vstat.end(), vstat);
  if (vnum) std::copy(F.vnum.begin(), F.vnum.end(), vnum);
  if (pnum) std::copy(F.pnum.begin(), F.pnum.end(), pnum);
  if (vstt) std::copy(F.vstt.begin(), F.vstt.end(), vstt);
  if (pstt) std::copy(F.pstt.begin(), F.pstt.end(), pstt);
  if (order) std::copy(F.order.begin(), F.order.end(), order);
  NM_API_END
}
extern "C" int nm_fem_t2n(void* h, int* t2n /* [ntet][pNp] */) {
  NM_API_BEGIN
  NmFem& F = *(NmFem*)h;
  std::copy(F.t2n.begin(), F.t2n.end(), t2n);
  NM_API_END
}
extern "C" int nm_fem_assemble_values(void* h, int job, const double* vp, const double* vs, const double* rho,
                                      const double* g0) {
  NM_API_BEGIN
  nm_fem_assemble(*(NmFem*)h, job, vp, vs, rho, g0);
  NM_API_END
}
