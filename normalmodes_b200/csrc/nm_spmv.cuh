// Sparse mat-vec kernels for sm_100a, templated on the EPILOGUE so that the Chebyshev
// recurrences of the hot path read the matrix once and write the three-term vector update in
// the same pass (north_star items 1-2; replaces pEVSL parcsrmatvec + ChebIter/ChebAv AXPYs called
// from src/mod_matvec.f90:453,471,480,495,507-515).
//
// Three storage formats, chosen once in nm_parcsr_build:
//   CSR    scalar rows.
//   ROW3   the 3 rows of a node share one column list made of aligned column triples
//          (src/mod_cg_create_matrix.f90:1373-1388): values stay in CSR order (3 coalesced
//          streams), ONE block-column id per 9 values, x gathered once per column for 3 rows.
//   KRON3  B = M (x) I3 (src/mod_cg_create_matrix.f90:1247-1259,1417-1434): scalar values and one
//          block-column id per 3x3 block -> 12 bytes per 3 non-zeros of each of the 3 rows.
//
// The Chebyshev iterations (the hot loop) run on k_slabws / k_slabpers / k_slab of nm_slab.cuh; the kernels in this
// file are the general products (A, Ad, E, ET with the fused filter epilogue): one subwarp per (block-)row straight
// from global memory, also the fallback of the iterations for matrices the slab packer refuses.  The mbarrier / TMA /
// cp.async / flag-in-data helpers shared with nm_slab.cuh live here as well.
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

// ---------------------------------------------------------------- epilogues
// load(row): the vector operands of the row (issued before the gather phase); apply(row, acc, in): the update.
struct EpiStore {            // y = A x
  double* y;
  struct In {};
  __device__ __forceinline__ In load(int) const { return In(); }
  __device__ __forceinline__ void apply(int row, double acc, const In&) const { y[row] = acc; }
};
struct EpiAdd {              // y += A x
  double* y;
  struct In { double y; };
  __device__ __forceinline__ In load(int row) const { return In{y[row]}; }
  __device__ __forceinline__ void apply(int row, double acc, const In& in) const { y[row] = in.y + acc; }
};
struct EpiStorePlus {        // y = A x + add
  double* y;
  const double* add;
  struct In { double a; };
  __device__ __forceinline__ In load(int row) const { return In{add[row]}; }
  __device__ __forceinline__ void apply(int row, double acc, const In& in) const { y[row] = acc + in.a; }
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
  struct In { double r, d, x; };
  __device__ __forceinline__ In load(int row) const {
    In in;
    in.r = r_in[row];
    in.d = first ? 0.0 : d_in[row];
    in.x = first ? 0.0 : x[row];
    return in;
  }
  __device__ __forceinline__ void apply(int row, double acc, const In& in) const {
    double d, rn, xn;
    if (first) {
      const double b = in.r;
      d = b * inv_theta;
      rn = b - acc * inv_theta;
      xn = d;
    } else {
      d = in.d;
      rn = in.r - acc;
      xn = in.x + d;
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
  // same update, returning the new direction d_out[row] (what the next step gathers): used by the fused multi-GPU
  // step, which stores it straight into the peers' ghost buffers
  __device__ __forceinline__ double apply_dn(int row, double acc, const In& in) const {
    double d, rn, xn;
    if (first) {
      const double b = in.r;
      d = b * inv_theta;
      rn = b - acc * inv_theta;
      xn = d;
    } else {
      d = in.d;
      rn = in.r - acc;
      xn = in.x + d;
    }
    const double dn = ak * d + bk * rn;
    if (last) {
      x[row] = xn + dn;
    } else {
      x[row] = xn;
      r_out[row] = rn;
      d_out[row] = dn;
    }
    return dn;
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
  struct In { double v, vm, y, a; };
  __device__ __forceinline__ In load(int row) const {
    In in;
    in.v = vk[row];
    in.vm = first ? 0.0 : vkm1[row];
    in.y = first ? 0.0 : y[row];
    in.a = add ? add[row] : 0.0;
    return in;
  }
  __device__ __forceinline__ void apply(int row, double acc, const In& in) const {
    if (add) acc += in.a;
    double vn = t * (acc - cc * in.v);
    if (!first) vn -= in.vm;
    vout[row] = vn;
    y[row] = first ? (mu0 * in.v + mu * vn) : (in.y + mu * vn);
  }
};

// ---------------------------------------------------------------- flag-in-data halo slots
// One ghost value = 16 bytes {lo32, tag, hi32, tag} written with two 8-byte words in ONE vector store (each 8-byte
// word lands atomically, the scheme of NCCL's LL protocol): the reader polls the slot itself until both tags match,
// so the exchange needs no fence, no arrival flag and no atomic -- its latency is one NVLink store flight.
__device__ __forceinline__ void nm_ll_store(unsigned long long* slot, double v, unsigned tag) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (u & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (u >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
// variants for measurements (NM_DEBUG_LL bits 2-4): 0 volatile = relaxed.sys (default), 1 ld.global.cg (weak, L2),
// 2 ld.relaxed.gpu
__device__ __forceinline__ void nm_ll_load_raw(const unsigned long long* slot, int mode, unsigned& a, unsigned& fa, unsigned& b, unsigned& fb) {
  if (mode == 1)
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(slot) : "memory");
  else if (mode == 2)
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(slot) : "memory");
  else
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(slot) : "memory");
}
__device__ __forceinline__ void nm_ll_store_mode(unsigned long long* slot, double v, unsigned tag, int mode) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (u & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (u >> 32) | ((unsigned long long)tag << 32);
  if (mode == 1) asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
  else if (mode == 2) asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
  else asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool nm_ll_load(const unsigned long long* slot, unsigned tag, double* v) {
  unsigned a, fa, b, fb;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(slot) : "memory");
  if (fa != tag || fb != tag) return false;
  *v = __longlong_as_double((long long)(((unsigned long long)b << 32) | a));
  return true;
}

// ================================================================ streaming kernel
__device__ __forceinline__ uint32_t nm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nm_mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nm_smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void nm_mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nm_smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nm_mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done;
  const uint32_t a = nm_smem_u32(b);
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA bulk copy global -> shared (16-byte aligned src/dst/size), completion on an mbarrier, L2 evict-first.
__device__ __forceinline__ void nm_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          nm_smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(nm_smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t nm_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ void nm_cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(nm_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void nm_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---------------------------------------------------------------- product kernels (global-memory subwarp per row)
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
  if (row < nrow && lane == 0) epi.apply(row, acc, epi.load(row));
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
  if (br < nbrow && lane < 3) {
    const int row = 3 * br + lane;
    epi.apply(row, lane == 0 ? a0 : (lane == 1 ? a1 : a2), epi.load(row));
  }
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
  if (br < nbrow && lane < 3) {
    const int row = 3 * br + lane;
    epi.apply(row, lane == 0 ? a0 : (lane == 1 ? a1 : a2), epi.load(row));
  }
}

// ---------------------------------------------------------------- host-side dispatch
template <int W, class Epi>
static inline void nm_spmv_launch_w(NmParcsr& M, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  const double* xg = M.halo.xg_cur ? M.halo.xg_cur : x;
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
  nm_halo_exchange(M, x);
  const double r = M.avg_row;              // entries one subwarp walks through
  if (r <= 6.0) nm_spmv_launch_w<4, Epi>(M, x, epi);
  else if (r <= 12.0) nm_spmv_launch_w<8, Epi>(M, x, epi);
  else if (r <= 48.0) nm_spmv_launch_w<16, Epi>(M, x, epi);
  else nm_spmv_launch_w<32, Epi>(M, x, epi);
}
