// Lanczos drivers on the device: spectral bounds (pEVSL LanTrbounds, called at
// src/mod_matvec.f90:85,162 and src/mod_pevsl.f90:84) and the Chebyshev-filtered non-restarted
// Lanczos with full CGS-DGKS2 reorthogonalisation in the B inner product (pEVSL ChebLanNr,
// src/mod_pevsl.f90:122), Ritz extraction, acceptance and residuals.
#include "nm_internal.h"
#include <algorithm>
#include <chrono>

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- Lanczos basis (chunked columns)
// pEVSL allocates n x (maxit+1) up front; maxit = 9624 (src/mod_pevsl.f90:117-119) would need hundreds of
// GB at PREM scale, so columns are allocated in chunks as the iteration proceeds.
struct Basis {
  size_t n = 0;
  int cpc = 64;                                  // columns per chunk
  std::vector<DBuf<double>> chunks;
  void init(size_t n_) {
    n = n_ ? n_ : 1;
    long long c = (256ll << 20) / (long long)(8 * n);   // ~256 MB per chunk
    cpc = (int)std::max(16ll, std::min(512ll, c));
  }
  double* col(int j) {
    while ((int)chunks.size() * cpc <= j) chunks.emplace_back((size_t)cpc * n);
    return chunks[j / cpc].p + (size_t)(j % cpc) * n;
  }
};

// ---------------------------------------------------------------- reorthogonalisation kernels (K6)
// c = V^T z : every CTA owns GT_ROWS rows x GT_COLS columns; z is loaded once per row and reused
// for GT_COLS columns, V is streamed exactly once.  Per-tile partials are reduced in a fixed order
// by k_reduce_partials (deterministic, no fp64 atomics).
#define GT_THREADS 256
#define GT_RPT 4
#define GT_ROWS (GT_THREADS * GT_RPT)
#define GT_COLS 8

__global__ void __launch_bounds__(GT_THREADS)
k_gemvT(const double* __restrict__ V, size_t n, int nc, const double* __restrict__ z, double* __restrict__ partial,
        int ntiles) {
  __shared__ double sm[GT_THREADS / 32][GT_COLS];
  const int tile = blockIdx.x;
  const int j0 = blockIdx.y * GT_COLS;
  const size_t r0 = (size_t)tile * GT_ROWS + threadIdx.x;
  double zr[GT_RPT];
#pragma unroll
  for (int q = 0; q < GT_RPT; ++q) {
    const size_t i = r0 + (size_t)q * GT_THREADS;
    zr[q] = i < n ? z[i] : 0.0;
  }
  double acc[GT_COLS];
#pragma unroll
  for (int jj = 0; jj < GT_COLS; ++jj) {
    acc[jj] = 0.0;
    const int j = j0 + jj;
    if (j < nc) {
      const double* __restrict__ col = V + (size_t)j * n;
#pragma unroll
      for (int q = 0; q < GT_RPT; ++q) {
        const size_t i = r0 + (size_t)q * GT_THREADS;
        if (i < n) acc[jj] += col[i] * zr[q];
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < GT_COLS; ++jj) {
    double v = acc[jj];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5][jj] = v;
  }
  __syncthreads();
  if (threadIdx.x < GT_COLS && j0 + threadIdx.x < nc) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < GT_THREADS / 32; ++w) s += sm[w][threadIdx.x];
    partial[(size_t)(j0 + threadIdx.x) * ntiles + tile] = s;
  }
}

// c[j] = sum_t partial[j][t], one warp per column
__global__ void k_reduce_partials(const double* __restrict__ partial, int ntiles, int nc, double* __restrict__ c) {
  const int j = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  if (j < nc)
    for (int t = lane; t < ntiles; t += 32) s += partial[(size_t)j * ntiles + t];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (j < nc && lane == 0) c[j] = s;
}

// z -= Z c : one thread per row, Z streamed once, c in shared memory
__global__ void __launch_bounds__(256)
k_gemvN(const double* __restrict__ Zc, size_t n, int nc, const double* __restrict__ c, double* __restrict__ z) {
  extern __shared__ double sc[];
  for (int j = threadIdx.x; j < nc; j += blockDim.x) sc[j] = c[j];
  __syncthreads();
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int j = 0;
  for (; j + 8 <= nc; j += 8) {
    const double* p = Zc + (size_t)j * n + i;
    const double v0 = p[0], v1 = p[n], v2 = p[2 * n], v3 = p[3 * n];
    const double v4 = p[4 * n], v5 = p[5 * n], v6 = p[6 * n], v7 = p[7 * n];
    a0 += v0 * sc[j] + v4 * sc[j + 4];
    a1 += v1 * sc[j + 1] + v5 * sc[j + 5];
    a2 += v2 * sc[j + 2] + v6 * sc[j + 6];
    a3 += v3 * sc[j + 3] + v7 * sc[j + 7];
  }
  for (; j < nc; ++j) a0 += Zc[(size_t)j * n + i] * sc[j];
  z[i] -= (a0 + a1) + (a2 + a3);
}

// z -= Z (V^T z), k columns, `passes` times (CGS_DGKS2 with NGS_MAX passes, no norm test).
static void reorth(Basis& V, Basis& Z, int k, double* z, double* c_dev, int passes) {
  NmCtx& ctx = nm_ctx();
  const size_t n = V.n;
  const int ntiles = nm_div_up((long long)n, GT_ROWS);
  for (int pass = 0; pass < passes; ++pass) {
    for (int j0 = 0; j0 < k; j0 += V.cpc) {
      const int nc = std::min(V.cpc, k - j0);
      double* partial = nm_red_scratch((size_t)V.cpc * ntiles);
      dim3 grid(ntiles, nm_div_up(nc, GT_COLS));
      k_gemvT<<<grid, GT_THREADS, 0, ctx.stream>>>(V.col(j0), n, nc, z, partial, ntiles);
      k_reduce_partials<<<nm_div_up(nc, 8), 256, 0, ctx.stream>>>(partial, ntiles, nc, c_dev + j0);
      ctx.launches += 2;
    }
    nm_allreduce_sum(c_dev, k);
    for (int j0 = 0; j0 < k; j0 += Z.cpc) {
      const int nc = std::min(Z.cpc, k - j0);
      k_gemvN<<<nm_div_up((long long)n, 256), 256, nc * sizeof(double), ctx.stream>>>(Z.col(j0), n, nc, c_dev + j0, z);
      ctx.launches++;
    }
  }
}

// ---------------------------------------------------------------- Ritz extraction (K7): U = V S
// A real dense contraction -> fp64 tensor cores (DMMA, mma.sync m8n8k4 f64; tcgen05 has no fp64 kind).
// CTA tile 64 rows x 32 columns, 4 warps (16 rows each), K staged through shared memory 16 at a time.
#define RG_BM 64
#define RG_BN 32
#define RG_BK 16
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
// U[n x ns] (ld n) += Vc[n x kc] (ld n) * S[kc x ns] (ld lds)
__global__ void __launch_bounds__(128)
k_ritz_gemm(const double* __restrict__ Vc, size_t n, int kc, const double* __restrict__ S, int lds, int ns,
            double* __restrict__ U, int accumulate) {
  __shared__ double sA[RG_BK][RG_BM + 1];       // [k][row]
  __shared__ double sB[RG_BN][RG_BK + 1];       // [col][k]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const size_t row0 = (size_t)blockIdx.x * RG_BM;
  const int col0 = blockIdx.y * RG_BN;
  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int k0 = 0; k0 < kc; k0 += RG_BK) {
    for (int t = threadIdx.x; t < RG_BK * RG_BM; t += 128) {
      const int kk = t / RG_BM, r = t % RG_BM;
      const size_t gi = row0 + r;
      sA[kk][r] = (gi < n && k0 + kk < kc) ? Vc[(size_t)(k0 + kk) * n + gi] : 0.0;
    }
    for (int t = threadIdx.x; t < RG_BN * RG_BK; t += 128) {
      const int cc = t / RG_BK, kk = t % RG_BK;
      sB[cc][kk] = (col0 + cc < ns && k0 + kk < kc) ? S[(size_t)(col0 + cc) * lds + k0 + kk] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < RG_BK; k4 += 4) {
      double af[2], bf[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = sA[k4 + tg][warp * 16 + a * 8 + g];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = sB[b * 8 + g][k4 + tg];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const size_t gi = row0 + warp * 16 + a * 8 + g;
        const int gc = col0 + b * 8 + tg * 2 + e;
        if (gi < n && gc < ns) {
          double* p = U + (size_t)gc * n + gi;
          *p = accumulate ? (*p + acc[a][b][e]) : acc[a][b][e];
        }
      }
}

static void ritz_vectors(Basis& V, int kdim, const double* S_dev, int lds, int ns, double* U) {
  NmCtx& ctx = nm_ctx();
  const size_t n = V.n;
  for (int j0 = 0; j0 < kdim; j0 += V.cpc) {
    const int kc = std::min(V.cpc, kdim - j0);
    dim3 grid(nm_div_up((long long)n, RG_BM), nm_div_up(ns, RG_BN));
    k_ritz_gemm<<<grid, 128, 0, ctx.stream>>>(V.col(j0), n, kc, S_dev + j0, lds, ns, U, j0 > 0);
    ctx.launches++;
  }
}

// ---------------------------------------------------------------- Gram blocks G = X^T Y on fp64 tensor cores
// The block inner products of the Rayleigh-Ritz refinement (and of any block orthogonalisation): X (n x p) and
// Y (n x q) column-major with leading dimension n, G (p x q) column-major.  A real dense contraction with the long
// dimension n as the MMA's k: mma.sync m8n8k4 f64, A fragment = 4 consecutive rows of 8 columns of X (each lane one
// double, 4 lanes = one 32-byte sector), B fragment likewise from Y; a CTA of 4 warps owns a 32 x 32 tile of G over one
// of GR_SLABS row slabs, partial tiles are summed over the slabs in a fixed order (deterministic, no fp64 atomics).
#define GR_SLABS 64
__global__ void __launch_bounds__(128)
k_gram_dmma(const double* __restrict__ X, const double* __restrict__ Y, size_t n, int p, int q, double* __restrict__ partial) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tg = lane & 3;
  const int i0 = blockIdx.y * 32 + (warp >> 1) * 16, j0 = blockIdx.z * 32 + (warp & 1) * 16;
  const size_t rows_per = ((n + GR_SLABS - 1) / GR_SLABS + 3) & ~(size_t)3;
  const size_t r0 = (size_t)blockIdx.x * rows_per, r1 = r0 + rows_per < n ? r0 + rows_per : n;
  double acc[2][2][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const double* xa[2]; const double* yb[2];
  bool xok[2], yok[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) { const int col = i0 + 8 * a + g; xok[a] = col < p; xa[a] = X + (size_t)(xok[a] ? col : 0) * n; }
#pragma unroll
  for (int b = 0; b < 2; ++b) { const int col = j0 + 8 * b + g; yok[b] = col < q; yb[b] = Y + (size_t)(yok[b] ? col : 0) * n; }
  for (size_t r = r0; r < r1; r += 4) {
    const size_t rr = r + tg;
    const bool in = rr < r1;
    double af[2], bf[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) af[a] = (in && xok[a]) ? xa[a][rr] : 0.0;
#pragma unroll
    for (int b = 0; b < 2; ++b) bf[b] = (in && yok[b]) ? yb[b][rr] : 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
  }
  double* out = partial + (size_t)blockIdx.x * p * q;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gi = i0 + 8 * a + g, gj = j0 + 8 * b + 2 * tg + e;
        if (gi < p && gj < q) out[(size_t)gj * p + gi] = acc[a][b][e];
      }
}
__global__ void k_gram_reduce(const double* __restrict__ partial, size_t pq, double* __restrict__ G) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= pq) return;
  double s = 0.0;
  for (int t = 0; t < GR_SLABS; ++t) s += partial[(size_t)t * pq + i];
  G[i] = s;
}
// G_dev (p x q, column-major) = X^T Y, summed over the ranks
static void gram(const double* X, const double* Y, size_t n, int p, int q, double* G_dev) {
  NmCtx& ctx = nm_ctx();
  const size_t pq = (size_t)p * q;
  double* partial = nm_red_scratch((size_t)GR_SLABS * pq);
  dim3 grid(GR_SLABS, nm_div_up(p, 32), nm_div_up(q, 32));
  k_gram_dmma<<<grid, 128, 0, ctx.stream>>>(X, Y, n, p, q, partial);
  k_gram_reduce<<<nm_div_up((long long)pq, 256), 256, 0, ctx.stream>>>(partial, pq, G_dev);
  ctx.launches += 2;
  nm_allreduce_sum(G_dev, pq);
}

// out = a*x  (scale-copy)
__global__ void k_scale_copy(double* out, const double* x, double a, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = a * x[i];
}
static void scale_copy(double* out, const double* x, double a, size_t n) {
  NmCtx& ctx = nm_ctx();
  if (!n) return;
  int grid = std::min<long long>((n + 255) / 256, (long long)ctx.sm_count * 8);
  k_scale_copy<<<grid, 256, 0, ctx.stream>>>(out, x, a, n);
  ctx.launches++;
}

// ---------------------------------------------------------------- LanTrbounds
void nm_lanbounds(NmPevsl& P, int mlan, int lanstep, double tol, double* lmin_out, double* lmax_out) {
  nm_ensure_init();
  NM_REQUIRE(P.A, "lanbounds: no A operator registered (pevsl_setamv_f90)");
  const bool gen = P.geneig;
  if (gen) NM_REQUIRE(P.bsol && P.B, "lanbounds: generalised problem needs setbmv + setbsol_chebiter");
  const size_t n = P.n;
  const int m = std::max(2, std::min(std::min(mlan, lanstep), P.N));       // operator applications at most
  // Basis columns kept at most: pEVSL's LanTrbounds is a thick-restart Lanczos with mlan basis vectors; the full basis of
  // a 2 M-tet mesh (67 MB per column, V and Z in the generalised case) must not be allowed to grow to mlan = 3000
  // columns.  When the cap is reached the iteration restarts from the sum of the two extreme Ritz vectors (explicit
  // restart: the new Krylov space starts where the old one's extremes were), a quarter of the free memory at most.
  size_t free_b = 0, total_b = 0;
  NM_CUDA(cudaMemGetInfo(&free_b, &total_b));
  long long cap_ll = (long long)(free_b / 4) / (long long)(8 * std::max<size_t>(n, 1) * (gen ? 2 : 1));
  int cap = (int)std::max(32ll, std::min((long long)m, cap_ll));
  cap = std::max(8, std::min(cap, nm_env_int("NM_LANBOUNDS_MAXCOLS", cap)));
  Basis V, Z;
  V.init(n); Z.init(n);
  Basis& Zr = gen ? Z : V;
  DBuf<double> w(std::max<size_t>(n, 1)), cbuf(cap + 2);
  std::vector<double> dT, eT, th, lr;
  auto normalise_start = [&]() {
    if (gen) {
      nm_op_apply(*P.B, V.col(0), Z.col(0));
      const double t = 1.0 / sqrt(nm_vec_dot(V.col(0), Z.col(0), n));
      nm_vec_scale(V.col(0), t, n); nm_vec_scale(Z.col(0), t, n);
    } else {
      const double t = 1.0 / sqrt(nm_vec_dot(V.col(0), V.col(0), n));
      nm_vec_scale(V.col(0), t, n);
    }
  };
  nm_vec_random(V.col(0), n, P.seed + 17, (unsigned long long)(P.nfirst >= 0 ? P.nfirst : 0));
  normalise_start();
  double lmin = 0, lmax = 0, beta = 0;
  bool breakdown = false;
  const int check_every = 10;
  int k = 0;                                                    // column of the current basis
  for (int step = 0; step < m; ++step) {
    double* v = V.col(k);
    double* znew = gen ? Z.col(k + 1) : w.p;
    nm_op_apply(*P.A, v, znew);
    if (k > 0) nm_vec_axpy(znew, -beta, Zr.col(k - 1), n);
    const double alpha = nm_vec_dot(v, znew, n);
    dT.push_back(alpha);
    reorth(V, Zr, k + 1, znew, cbuf.p, 2);
    if (gen) {
      double* vnew = V.col(k + 1);
      nm_chebiter_solve(*P.bsol, znew, vnew);
      beta = sqrt(fabs(nm_vec_dot(vnew, znew, n)));
      NM_REQUIRE(beta > 0 && std::isfinite(beta), "lanbounds: breakdown (beta = %g) at step %d", beta, step);
      nm_vec_scale(vnew, 1.0 / beta, n); nm_vec_scale(znew, 1.0 / beta, n);
    } else {
      beta = sqrt(nm_vec_dot(znew, znew, n));
      NM_REQUIRE(std::isfinite(beta), "lanbounds: non-finite beta at step %d", step);
      // exact breakdown (invariant subspace: B~ = I from a lumped mass matrix, a tiny matrix, ...): the Ritz values of
      // the T built so far are exact eigenvalues -- fall through to the bounds with r1 = r2 = 0 instead of skipping them
      breakdown = beta <= 1e-300 || beta * beta <= 1e-30 * fabs(alpha) * fabs(alpha);
      if (!breakdown) scale_copy(V.col(k + 1), znew, 1.0 / beta, n);
      else beta = 0.0;
    }
    eT.push_back(beta);
    const int kk = ++k;
    const bool last = step == m - 1, full = kk == cap && !last;
    if (!breakdown && !full && kk % check_every && !last) continue;
    th.resize(kk); lr.resize(kk);
    std::vector<double> S;
    if (full) S.resize((size_t)kk * kk);
    int rc = nm_tridiag_eig_ex(kk, dT.data(), eT.data(), th.data(), full ? S.data() : nullptr, lr.data());
    NM_REQUIRE(rc == 0, "lanbounds: tridiagonal eigensolver failed");
    const double r1 = fabs(beta * lr[0]), r2 = fabs(beta * lr[kk - 1]);
    lmin = th[0] - r1; lmax = th[kk - 1] + r2;
    if (breakdown || r1 + r2 < tol * (fabs(lmin) + fabs(lmax))) break;
    if (full) {
      // explicit restart: v0 <- y_min + y_max (Ritz vectors of the two ends), normalised; T starts over
      std::vector<double> comb(kk);
      for (int i = 0; i < kk; ++i) comb[i] = S[i] + S[(size_t)(kk - 1) * kk + i];
      DBuf<double> dS(kk);
      dS.upload(comb.data(), kk);
      ritz_vectors(V, kk, dS.p, kk, 1, w.p);
      nm_vec_copy(V.col(0), w.p, n);
      normalise_start();
      dT.clear(); eT.clear();
      k = 0; beta = 0.0;
    }
  }
  NM_CUDA(cudaStreamSynchronize(nm_ctx().stream));
  nm_check_device_status();
  *lmin_out = lmin; *lmax_out = lmax;
}

// ---------------------------------------------------------------- ChebLanNr
void nm_cheblannr(NmPevsl& P, const double xintv[4], int maxit, double tol, const NmPol& pol) {
  nm_ensure_init();
  NmCtx& ctx = nm_ctx();
  NM_REQUIRE(P.A && P.B && P.bsol && P.geneig,
             "cheblannr: needs setamv, setbmv, setbsol_chebiter and set_geneig (src/mod_pevsl.f90:69-82)");
  const double t_begin = now_s();
  const size_t n = P.n;
  const double aa = xintv[0], bb = xintv[1];
  const double bar = pol.bar;
  const int Ntest = 30, cycle = 20, NGS_MAX = 2;
  const double DBL_EPS_MULT = 10.0, orthTol = 1e-14;
  maxit = std::min(P.N, maxit);
  NM_REQUIRE(maxit >= 1, "cheblannr: maxit < 1");
  Basis V, Z;
  V.init(n); Z.init(n);
  DBuf<double> work(3 * std::max<size_t>(n, 1)), cbuf(maxit + 2);
  std::vector<double> dT, eT, th;
  // start vector: random, B-normalised
  nm_vec_random(V.col(0), n, P.seed, (unsigned long long)(P.nfirst >= 0 ? P.nfirst : 0));
  nm_op_apply(*P.B, V.col(0), Z.col(0));
  {
    const double t = 1.0 / sqrt(nm_vec_dot(V.col(0), Z.col(0), n));
    nm_vec_scale(V.col(0), t, n); nm_vec_scale(Z.col(0), t, n);
  }
  double beta = 0.0, wn = 0.0, tr0 = 0.0;
  int nwn = 0, kdim = 0;
  double t_filter = 0, t_reorth = 0;
  cudaEvent_t ev[4];
  for (auto& e : ev) NM_CUDA(cudaEventCreate(&e));
  for (int k = 0; k < maxit; ++k) {
    double* v = V.col(k);
    double* z = Z.col(k);
    double* vnew = V.col(k + 1);
    double* znew = Z.col(k + 1);
    NM_CUDA(cudaEventRecord(ev[0], ctx.stream));
    nm_filter_apply(P, pol, z, znew, work.p);
    NM_CUDA(cudaEventRecord(ev[1], ctx.stream));
    if (k > 0) nm_vec_axpy(znew, -beta, Z.col(k - 1), n);
    const double alpha = nm_vec_dot(v, znew, n);
    dT.push_back(alpha);
    wn += fabs(alpha);
    nm_vec_axpy(znew, -alpha, z, n);
    NM_CUDA(cudaEventRecord(ev[2], ctx.stream));
    reorth(V, Z, k + 1, znew, cbuf.p, NGS_MAX);
    NM_CUDA(cudaEventRecord(ev[3], ctx.stream));
    nm_chebiter_solve(*P.bsol, znew, vnew);
    beta = sqrt(fabs(nm_vec_dot(vnew, znew, n)));
    NM_REQUIRE(std::isfinite(beta), "cheblannr: non-finite beta at step %d", k);
    {
      float ms;
      NM_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[1])); t_filter += ms * 1e-3;
      NM_CUDA(cudaEventElapsedTime(&ms, ev[2], ev[3])); t_reorth += ms * 1e-3;
    }
    wn += 2.0 * beta;
    nwn += 3;
    if (beta * nwn < orthTol * wn) {
      // lucky breakdown: new random direction, B-orthogonalised against the basis
      nm_vec_random(vnew, n, P.seed + 1000 + k, (unsigned long long)(P.nfirst >= 0 ? P.nfirst : 0));
      reorth(Z, V, k + 1, vnew, cbuf.p, NGS_MAX);
      nm_op_apply(*P.B, vnew, znew);
      beta = sqrt(nm_vec_dot(vnew, znew, n));
      nm_vec_scale(vnew, 1.0 / beta, n); nm_vec_scale(znew, 1.0 / beta, n);
      beta = 0.0;
    } else {
      nm_vec_scale(vnew, 1.0 / beta, n); nm_vec_scale(znew, 1.0 / beta, n);
    }
    eT.push_back(beta);
    kdim = k + 1;
    if ((k < Ntest || (k - Ntest) % cycle != 0) && k != maxit - 1) continue;
    th.resize(kdim);
    int rc = nm_tridiag_eig(kdim, dT.data(), eT.data(), th.data(), nullptr);
    NM_REQUIRE(rc == 0, "cheblannr: tridiagonal eigensolver failed");
    double tr1 = 0.0;
    for (int i = 0; i < kdim; ++i)
      if (th[i] + DBL_EPS_MULT * 2.220446049250313e-16 >= bar) tr1 += th[i];
    bool done = fabs(tr1 - tr0) < tol * fabs(tr1);
    if (done && P.ritz_tol > 0.0) {
      // optional gate on top of pEVSL's trace test: every wanted Ritz pair's Lanczos residual estimate
      // |beta_k s_ki| must be below ritz_tol (the pairs next to the band edges converge last)
      std::vector<double> lastrow(kdim);
      rc = nm_tridiag_eig_ex(kdim, dT.data(), eT.data(), th.data(), nullptr, lastrow.data());
      NM_REQUIRE(rc == 0, "cheblannr: tridiagonal eigensolver failed");
      for (int i = 0; i < kdim; ++i)
        if (th[i] >= bar && fabs(beta * lastrow[i]) > P.ritz_tol) { done = false; break; }
    }
    if (done) break;
    tr0 = tr1;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  // ---- Ritz pairs
  const double t_r0 = now_s();
  th.resize(kdim);
  std::vector<double> S((size_t)kdim * kdim);
  int rc = nm_tridiag_eig(kdim, dT.data(), eT.data(), th.data(), S.data());
  NM_REQUIRE(rc == 0, "cheblannr: tridiagonal eigensolver (vectors) failed");
  std::vector<int> sel;
  for (int i = 0; i < kdim; ++i)
    if (th[i] >= bar) sel.push_back(i);
  const int ns = (int)sel.size();
  P.nev = 0; P.lam.clear(); P.res.clear();
  P.last_steps = kdim; P.last_deg = pol.deg;
  if (ns > 0) {
    std::vector<double> Ssel((size_t)kdim * ns);
    for (int c = 0; c < ns; ++c) std::copy(S.begin() + (size_t)sel[c] * kdim, S.begin() + (size_t)(sel[c] + 1) * kdim, Ssel.begin() + (size_t)c * kdim);
    DBuf<double> dS(Ssel.size());
    dS.upload(Ssel.data(), Ssel.size());
    const size_t nn = std::max<size_t>(n, 1);
    DBuf<double> U((size_t)ns * nn);
    ritz_vectors(V, kdim, dS.p, kdim, ns, U.p);
    V.chunks.clear(); Z.chunks.clear();                         // the Lanczos bases are done: their memory serves A U, B U
    // accepted pairs: B-normalise, Rayleigh quotient, acceptance in [a, b] (pEVSL ChebLanNr); A u and B u of the
    // accepted ones are kept (columns nk of AU, BU; u compacted to column nk of U)
    DBuf<double> AU((size_t)ns * nn), BU((size_t)ns * nn);
    int nk = 0;
    for (int c = 0; c < ns; ++c) {
      double* u = U.p + (size_t)c * n;
      double* w2 = BU.p + (size_t)nk * n;
      double* wk = AU.p + (size_t)nk * n;
      nm_op_apply(*P.B, u, w2);
      double t = sqrt(nm_vec_dot(u, w2, n));
      NM_REQUIRE(t > 0.0, "cheblannr: zero Ritz vector");
      t = 1.0 / t;
      nm_vec_scale(u, t, n); nm_vec_scale(w2, t, n);
      nm_op_apply(*P.A, u, wk);
      const double lam = nm_vec_dot(wk, u, n);
      if (lam < aa - DBL_EPS_MULT * 2.220446049250313e-16 || lam > bb + DBL_EPS_MULT * 2.220446049250313e-16) continue;
      if (c != nk) nm_vec_copy(U.p + (size_t)nk * n, u, n);
      P.lam.push_back(lam);
      ++nk;
    }
    P.nev = nk;
    P.Y.alloc(std::max<size_t>((size_t)nk * n, 1));
    const bool refine = nm_env_int("NM_RITZ_REFINE", 1) != 0;
    bool refined = false;
    if (nk > 1 && refine) {
      // Rayleigh-Ritz on the span of the accepted vectors: what is left in a Ritz vector of a non-restarted Lanczos
      // stopped by the trace test is mostly a mixture of OTHER wanted eigenvectors (multiplet members resolve last);
      // the nk x nk projected pencil (U^T A U) c = lam (U^T B U) c removes exactly that.  Two Gram blocks and three
      // n x nk x nk products on the fp64 tensor cores; A U and B U rotate with U, so no further operator applications.
      DBuf<double> HG((size_t)2 * nk * nk);
      gram(U.p, AU.p, n, nk, nk, HG.p);
      gram(U.p, BU.p, n, nk, nk, HG.p + (size_t)nk * nk);
      std::vector<double> hHG((size_t)2 * nk * nk), w(nk), Cm((size_t)nk * nk);
      HG.download(hHG.data(), hHG.size());
      const int rc2 = nm_sym_geneig(nk, hHG.data(), hHG.data() + (size_t)nk * nk, w.data(), Cm.data());
      if (rc2 == 0 && w[0] >= aa - 1e-8 * fabs(aa) && w[nk - 1] <= bb + 1e-8 * fabs(bb)) {
        DBuf<double> dC(Cm.size());
        dC.upload(Cm.data(), Cm.size());
        auto rotate = [&](DBuf<double>& X, double* out) {       // out = X C  (n x nk by nk x nk)
          dim3 grid(nm_div_up((long long)n, RG_BM), nm_div_up(nk, RG_BN));
          k_ritz_gemm<<<grid, 128, 0, ctx.stream>>>(X.p, n, nk, dC.p, nk, nk, out, 0);
          ctx.launches++;
        };
        rotate(U, P.Y.p);
        DBuf<double> T((size_t)nk * nn);
        rotate(AU, T.p); std::swap(AU.p, T.p); std::swap(AU.n, T.n);
        rotate(BU, T.p); std::swap(BU.p, T.p); std::swap(BU.n, T.n);
        for (int i = 0; i < nk; ++i) P.lam[i] = w[i];
        refined = true;
      }
    }
    if (!refined)
      for (int i = 0; i < nk; ++i) nm_vec_copy(P.Y.p + (size_t)i * n, U.p + (size_t)i * n, n);
    // residuals ||A y - lam B y||_2 from the (rotated) products
    for (int i = 0; i < nk; ++i) {
      double* wk = AU.p + (size_t)i * n;
      nm_vec_axpy(wk, -P.lam[i], BU.p + (size_t)i * n, n);
      P.res.push_back(sqrt(nm_vec_dot(wk, wk, n)));
    }
    NM_CUDA(cudaStreamSynchronize(ctx.stream));
  }
  nm_check_device_status();
  P.t_filter = t_filter; P.t_reorth = t_reorth; P.t_ritz = now_s() - t_r0; P.t_total = now_s() - t_begin;
}

// ---------------------------------------------------------------- diagnostics: the dense kernels of the Lanczos phase alone
// One CGS pass (c = V^T z, z -= Z c) over k basis columns, the Ritz product U = V S (k x ns) and the Gram block
// U^T W (ns x ns) on random data of local length n: microseconds per call (CUDA events, 3 repetitions after one
// warm-up).  Used by tools/lanczos_kernels.py for the live numbers and as the ncu target of K6/K7.
extern "C" int nm_diag_lanczos_kernels(long long n_, int k, int ns, double* us_out /* 4: gemvT+reduce, gemvN, ritz, gram */) {
  NM_API_BEGIN
  nm_ensure_init();
  NmCtx& ctx = nm_ctx();
  const size_t n = (size_t)n_;
  NM_REQUIRE(n > 0 && k > 0 && ns > 0, "nm_diag_lanczos_kernels: bad sizes");
  Basis V;
  V.init(n);
  for (int j = 0; j < k; ++j) nm_vec_random(V.col(j), n, 99, (unsigned long long)j * n);
  DBuf<double> z(n), cbuf(k + 1), U((size_t)ns * n), S((size_t)k * ns), G((size_t)ns * ns);
  nm_vec_random(z.p, n, 7, 0);
  nm_vec_random(S.p, (size_t)k * ns, 8, 0);
  const int ntiles = nm_div_up((long long)n, GT_ROWS);
  cudaEvent_t e0, e1;
  NM_CUDA(cudaEventCreate(&e0)); NM_CUDA(cudaEventCreate(&e1));
  auto timed = [&](int which) {
    for (int rep = -1; rep < 3; ++rep) {
      if (rep == 0) NM_CUDA(cudaEventRecord(e0, ctx.stream));
      if (which == 0) {
        for (int j0 = 0; j0 < k; j0 += V.cpc) {
          const int nc = std::min(V.cpc, k - j0);
          double* partial = nm_red_scratch((size_t)V.cpc * ntiles);
          dim3 grid(ntiles, nm_div_up(nc, GT_COLS));
          k_gemvT<<<grid, GT_THREADS, 0, ctx.stream>>>(V.col(j0), n, nc, z.p, partial, ntiles);
          k_reduce_partials<<<nm_div_up(nc, 8), 256, 0, ctx.stream>>>(partial, ntiles, nc, cbuf.p + j0);
        }
      } else if (which == 1) {
        for (int j0 = 0; j0 < k; j0 += V.cpc) {
          const int nc = std::min(V.cpc, k - j0);
          k_gemvN<<<nm_div_up((long long)n, 256), 256, nc * sizeof(double), ctx.stream>>>(V.col(j0), n, nc, cbuf.p + j0, z.p);
        }
      } else if (which == 2) {
        ritz_vectors(V, k, S.p, k, ns, U.p);
      } else {
        gram(U.p, U.p, n, ns, ns, G.p);
      }
    }
    NM_CUDA(cudaEventRecord(e1, ctx.stream));
    NM_CUDA(cudaStreamSynchronize(ctx.stream));
    float ms;
    NM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    return (double)ms * 1e3 / 3.0;
  };
  for (int w = 0; w < 4; ++w) us_out[w] = timed(w);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  NM_API_END
}

// ---------------------------------------------------------------- C ABI: solver context
extern "C" int nm_pevsl_set_ritz_tol(void* h, double tol) {
  NM_API_BEGIN
  ((NmPevsl*)h)->ritz_tol = tol;
  NM_API_END
}
extern "C" int nm_pevsl_create(void** out) {
  NM_API_BEGIN
  nm_ensure_init();
  *out = new NmPevsl();
  NM_API_END
}
extern "C" int nm_pevsl_free(void* h) {
  NM_API_BEGIN
  if (h) { if (nm_ctx().ready) NM_CUDA(cudaStreamSynchronize(nm_ctx().stream)); delete (NmPevsl*)h; }
  NM_API_END
}
extern "C" int nm_pevsl_setprobsizes(void* h, int N, int n, int nfirst) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  NM_REQUIRE(N >= n && n >= 0, "setprobsizes: N = %d, n = %d", N, n);
  P.N = N; P.n = n; P.nfirst = nfirst;
  NM_API_END
}
extern "C" int nm_pevsl_set_nfirst(void* h, int nfirst) {
  NM_API_BEGIN
  ((NmPevsl*)h)->nfirst = nfirst;
  NM_API_END
}
static NmOp* own_callback(NmPevsl& P, nm_matvec_fn fn, void* data) {
  NmOp* op = new NmOp();
  op->kind = NM_OP_CALLBACK; op->n = P.n; op->fn = fn; op->fn_data = data;
  P.owned_ops.emplace_back(op);
  return op;
}
extern "C" int nm_pevsl_setamv_callback(void* h, nm_matvec_fn fn, void* data) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  P.A = own_callback(P, fn, data);
  NM_API_END
}
extern "C" int nm_pevsl_setbmv_callback(void* h, nm_matvec_fn fn, void* data) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  P.B = own_callback(P, fn, data);
  NM_API_END
}
extern "C" int nm_pevsl_setamv_op(void* h, void* op) {
  NM_API_BEGIN
  ((NmPevsl*)h)->A = (NmOp*)op;
  NM_API_END
}
extern "C" int nm_pevsl_setbmv_op(void* h, void* op) {
  NM_API_BEGIN
  ((NmPevsl*)h)->B = (NmOp*)op;
  NM_API_END
}
// hand ownership of an operator to the context (freed with it)
extern "C" int nm_pevsl_adopt_op(void* h, void* op) {
  NM_API_BEGIN
  ((NmPevsl*)h)->owned_ops.emplace_back((NmOp*)op);
  NM_API_END
}
extern "C" int nm_pevsl_setbsol_chebiter(void* h, void* cheb) {
  NM_API_BEGIN
  ((NmPevsl*)h)->bsol = (NmChebIter*)cheb;
  NM_API_END
}
extern "C" int nm_pevsl_set_geneig(void* h) {
  NM_API_BEGIN
  ((NmPevsl*)h)->geneig = true;
  NM_API_END
}
extern "C" int nm_pevsl_set_seed(void* h, unsigned long long seed) {
  NM_API_BEGIN
  ((NmPevsl*)h)->seed = seed;
  NM_API_END
}
extern "C" int nm_pevsl_lanbounds(void* h, int mlan, int lanstep, double tol, double* lmin, double* lmax) {
  NM_API_BEGIN
  nm_lanbounds(*(NmPevsl*)h, mlan, lanstep, tol, lmin, lmax);
  NM_API_END
}
extern "C" int nm_pevsl_cheblannr(void* h, const double* xintv, int maxit, double tol, void* pol) {
  NM_API_BEGIN
  nm_cheblannr(*(NmPevsl*)h, xintv, maxit, tol, *(NmPol*)pol);
  NM_API_END
}
extern "C" int nm_pevsl_get_nev(void* h, int* nev) {
  NM_API_BEGIN
  *nev = ((NmPevsl*)h)->nev;
  NM_API_END
}
// vals[nev], vecs[ld*nev] column-major (pevsl_copy_result_f90, src/mod_pevsl.f90:130); res optional.
extern "C" int nm_pevsl_copy_result(void* h, double* vals, double* vecs, int ld, double* res) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  NM_REQUIRE(ld >= P.n, "copy_result: ld = %d < n_local = %d", ld, P.n);
  for (int i = 0; i < P.nev; ++i) {
    if (vals) vals[i] = P.lam[i];
    if (res) res[i] = P.res[i];
  }
  if (vecs && P.nev && P.n)
    NM_CUDA(cudaMemcpy2D(vecs, (size_t)ld * sizeof(double), P.Y.p, (size_t)P.n * sizeof(double),
                         (size_t)P.n * sizeof(double), P.nev, cudaMemcpyDeviceToHost));
  NM_API_END
}
extern "C" int nm_pevsl_stats(void* h, int* steps, int* deg, double* t_total, double* t_filter, double* t_reorth,
                              double* t_ritz, long long* n_filter) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  if (steps) *steps = P.last_steps;
  if (deg) *deg = P.last_deg;
  if (t_total) *t_total = P.t_total;
  if (t_filter) *t_filter = P.t_filter;
  if (t_reorth) *t_reorth = P.t_reorth;
  if (t_ritz) *t_ritz = P.t_ritz;
  if (n_filter) *n_filter = P.n_filter_apply;
  NM_API_END
}
// y = p(A B^-1) z with HOST vectors: one application of the polynomial filter (bench e2e path).
extern "C" int nm_pevsl_filter_host(void* h, void* pol, const double* z, double* y) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  NmCtx& ctx = nm_ctx();
  const size_t n = P.n;
  DBuf<double> dz(std::max<size_t>(n, 1)), dy(std::max<size_t>(n, 1)), work(3 * std::max<size_t>(n, 1));
  if (n) NM_CUDA(cudaMemcpyAsync(dz.p, z, n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  nm_filter_apply(P, *(NmPol*)pol, dz.p, dy.p, work.p);
  if (n) NM_CUDA(cudaMemcpyAsync(y, dy.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  NM_CUDA(cudaStreamSynchronize(ctx.stream));
  NM_API_END
}
extern "C" int nm_pevsl_filter_dev(void* h, void* pol, const double* z_dev, double* y_dev, double* work_dev) {
  NM_API_BEGIN
  nm_filter_apply(*(NmPevsl*)h, *(NmPol*)pol, z_dev, y_dev, work_dev);
  NM_API_END
}
// The same, truncated after kmax degree steps: y = sum_{k<=kmax} mu_k T_k((A B^-1 - c)/d) z (kmax <= 0: all of them).
extern "C" int nm_pevsl_filter_steps_host(void* h, void* pol, int kmax, const double* z, double* y) {
  NM_API_BEGIN
  NmPevsl& P = *(NmPevsl*)h;
  NmCtx& ctx = nm_ctx();
  const size_t n = P.n;
  if (P.fwork.n < 5 * std::max<size_t>(n, 1)) P.fwork.alloc(5 * std::max<size_t>(n, 1));
  double* dz = P.fwork.p; double* dy = dz + n; double* work = dy + n;
  if (n) NM_CUDA(cudaMemcpyAsync(dz, z, n * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  nm_filter_apply(P, *(NmPol*)pol, dz, dy, work, kmax);
  if (n) NM_CUDA(cudaMemcpyAsync(y, dy, n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  NM_CUDA(cudaStreamSynchronize(ctx.stream));
  nm_check_device_status();
  NM_API_END
}
extern "C" int nm_pevsl_filter_steps_dev(void* h, void* pol, int kmax, const double* z_dev, double* y_dev, double* work_dev) {
  NM_API_BEGIN
  nm_filter_apply(*(NmPevsl*)h, *(NmPol*)pol, z_dev, y_dev, work_dev, kmax);
  NM_API_END
}
