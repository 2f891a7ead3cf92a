// Sparse mat-vec kernels for sm_100a, templated on the EPILOGUE so that the Chebyshev
// recurrences of the hot path read the matrix once and write the three-term vector update in
// the same pass (north_star items 1-2; replaces pEVSL parcsrmatvec + ChebIter/ChebAv AXPYs called
// from src/mod_matvec.f90:453,471,480,495,507-515).
//
// Three storage formats, chosen once in nm_parcsr_build:
//   CSR    one subwarp per row; values and column ids streamed coalesced, row sum by shuffles.
//   ROW3   the 3 rows of a node share one column list made of aligned column triples
//          (src/mod_cg_create_matrix.f90:1373-1388): values stay in CSR order (3 coalesced
//          streams), ONE block-column id per 9 values, x gathered once per column for 3 rows.
//   KRON3  B = M (x) I3 (src/mod_cg_create_matrix.f90:1247-1259,1417-1434): scalar values and one
//          block-column id per 3x3 block -> 12 bytes per 3 non-zeros of each of the 3 rows.
// HBM-bound integer/fp64 streaming work: no tensor cores; the x gathers are served by L1/L2.
#pragma once
#include "nm_internal.h"

#define NM_SPMV_THREADS 256

template <int W>
__device__ __forceinline__ double nm_subwarp_sum(double v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, W);
  return v;
}

__device__ __forceinline__ double nm_ldx(const double* __restrict__ x, const double* __restrict__ xg, int ncol, int c) {
  return c < ncol ? __ldg(x + c) : __ldg(xg + (c - ncol));
}

// ---------------------------------------------------------------- epilogues: called once per row
struct EpiStore {            // y = A x
  double* y;
  __device__ __forceinline__ void operator()(int row, double acc) const { y[row] = acc; }
};
struct EpiAdd {              // y += A x
  double* y;
  __device__ __forceinline__ void operator()(int row, double acc) const { y[row] += acc; }
};
struct EpiStorePlus {        // y = A x + add
  double* y;
  const double* add;
  __device__ __forceinline__ void operator()(int row, double acc) const { y[row] = acc + add[row]; }
};
struct EpiScaleStore {       // y = s_row * (A x)  (row scaling vector)
  double* y;
  const double* s;
  __device__ __forceinline__ void operator()(int row, double acc) const { y[row] = acc * s[row]; }
};
// One step of the Chebyshev iteration x = q(M) b (Saad Alg. 12.1), fused with r -= M d:
//   first : d0 = b/theta (gathered vector is b, so M d0 = acc/theta); x = d0
//   else  : x += d_in
//   r_out = r_in - M d ; d_out = ak*d + bk*r_out ; last: x += d_out as well (and r, d not stored)
struct EpiCheb {
  const double* r_in;      // b on the first step
  const double* d_in;      // b on the first step
  double* r_out;
  double* d_out;
  double* x;
  double inv_theta, ak, bk;
  int first, last;
  __device__ __forceinline__ void operator()(int row, double acc) const {
    double d, rn, xn;
    if (first) {
      const double b = r_in[row];
      d = b * inv_theta;
      rn = b - acc * inv_theta;
      xn = d;
    } else {
      d = d_in[row];
      rn = r_in[row] - acc;
      xn = x[row] + d;
    }
    const double dn = ak * d + bk * rn;
    if (last) {
      x[row] = xn + dn;
    } else {
      x[row] = xn;
      r_out[row] = rn;
      d_out[row] = dn;
    }
  }
};
// ChebAv three-term update fused with the A product (u = A w [+ add]):
//   v+ = t*(u - cc*vk) - vkm1 ; y (+)= mu*v+ ; vout may alias vkm1 (only this row reads it)
struct EpiFilter {
  const double* vk;
  const double* vkm1;
  double* vout;
  double* y;
  const double* add;       // optional extra term of the operator (E Ap^-1 ET w), may be null
  double t, cc, mu, mu0;
  int first;
  __device__ __forceinline__ void operator()(int row, double acc) const {
    if (add) acc += add[row];
    const double v = vk[row];
    double vn = t * (acc - cc * v);
    if (!first) vn -= vkm1[row];
    vout[row] = vn;
    y[row] = first ? (mu0 * v + mu * vn) : (y[row] + mu * vn);
  }
};

// ---------------------------------------------------------------- kernels
template <int W, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS)
k_spmv_csr(int nrow, int ncol, const int* __restrict__ ia, const int* __restrict__ ja, const double* __restrict__ a,
           const double* __restrict__ x, const double* __restrict__ xg, Epi epi) {
  const int lane = threadIdx.x & (W - 1);
  const int row = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / W);
  double acc = 0.0;
  if (row < nrow) {
    const int s = ia[row], e = ia[row + 1];
    for (int p = s + lane; p < e; p += W) acc += a[p] * nm_ldx(x, xg, ncol, ja[p]);
  }
  acc = nm_subwarp_sum<W>(acc);
  if (row < nrow && lane == 0) epi(row, acc);
}

template <int W, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS)
k_spmv_row3(int nbrow, int ncol, const int* __restrict__ bia, const int* __restrict__ bja,
            const double* __restrict__ a, const double* __restrict__ x, const double* __restrict__ xg, Epi epi) {
  const int lane = threadIdx.x & (W - 1);
  const int br = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / W);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (br < nbrow) {
    const int s = bia[br];
    const int len = 3 * (bia[br + 1] - s);                 // entries per row
    const double* __restrict__ v0 = a + 9ll * s;
    const double* __restrict__ v1 = v0 + len;
    const double* __restrict__ v2 = v1 + len;
    for (int u = lane; u < len; u += W) {
      const int q = u / 3;
      const int c = 3 * bja[s + q] + (u - 3 * q);
      const double xv = nm_ldx(x, xg, ncol, c);
      a0 += v0[u] * xv;
      a1 += v1[u] * xv;
      a2 += v2[u] * xv;
    }
  }
  a0 = nm_subwarp_sum<W>(a0);
  a1 = nm_subwarp_sum<W>(a1);
  a2 = nm_subwarp_sum<W>(a2);
  if (br < nbrow && lane < 3) epi(3 * br + lane, lane == 0 ? a0 : (lane == 1 ? a1 : a2));
}

template <int W, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS)
k_spmv_kron3(int nbrow, int ncol, const int* __restrict__ bia, const int* __restrict__ bja,
             const double* __restrict__ mval, const double* __restrict__ x, const double* __restrict__ xg, Epi epi) {
  const int lane = threadIdx.x & (W - 1);
  const int br = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) / W);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (br < nbrow) {
    const int s = bia[br], e = bia[br + 1];
    for (int p = s + lane; p < e; p += W) {
      const double m = mval[p];
      const int c = 3 * bja[p];
      a0 += m * nm_ldx(x, xg, ncol, c);
      a1 += m * nm_ldx(x, xg, ncol, c + 1);
      a2 += m * nm_ldx(x, xg, ncol, c + 2);
    }
  }
  a0 = nm_subwarp_sum<W>(a0);
  a1 = nm_subwarp_sum<W>(a1);
  a2 = nm_subwarp_sum<W>(a2);
  if (br < nbrow && lane < 3) epi(3 * br + lane, lane == 0 ? a0 : (lane == 1 ? a1 : a2));
}

// ---------------------------------------------------------------- host-side dispatch
template <int W, class Epi>
static inline void nm_spmv_launch_w(NmParcsr& M, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  const double* xg = M.halo.xg.p ? M.halo.xg.p : x;
  const int per_block = NM_SPMV_THREADS / W;
  if (M.format == NM_FMT_CSR) {
    if (M.nrow == 0) return;
    k_spmv_csr<W, Epi><<<nm_div_up(M.nrow, per_block), NM_SPMV_THREADS, 0, c.stream>>>(
        M.nrow, M.ncol, M.ia.p, M.ja.p, M.a.p, x, xg, epi);
  } else if (M.format == NM_FMT_ROW3) {
    if (M.nbrow == 0) return;
    k_spmv_row3<W, Epi><<<nm_div_up(M.nbrow, per_block), NM_SPMV_THREADS, 0, c.stream>>>(
        M.nbrow, M.ncol, M.bia.p, M.bja.p, M.a.p, x, xg, epi);
  } else {
    if (M.nbrow == 0) return;
    k_spmv_kron3<W, Epi><<<nm_div_up(M.nbrow, per_block), NM_SPMV_THREADS, 0, c.stream>>>(
        M.nbrow, M.ncol, M.bia.p, M.bja.p, M.mval.p, x, xg, epi);
  }
  c.launches++;
}

// Halo exchange (if any) + SpMV with the given epilogue.  x: device, owned part only.
template <class Epi>
static inline void nm_spmv_epi(NmParcsr& M, const double* x, const Epi& epi) {
  if (M.halo.nghost > 0) nm_halo_exchange(M, x);
  const double r = M.avg_row;              // entries one subwarp walks through
  if (r <= 6.0) nm_spmv_launch_w<4, Epi>(M, x, epi);
  else if (r <= 12.0) nm_spmv_launch_w<8, Epi>(M, x, epi);
  else if (r <= 48.0) nm_spmv_launch_w<16, Epi>(M, x, epi);
  else nm_spmv_launch_w<32, Epi>(M, x, epi);
}
