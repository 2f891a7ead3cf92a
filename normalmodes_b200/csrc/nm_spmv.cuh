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
// Main kernel (k_stream): HBM-bound streaming work, no tensor cores.  The (block-)rows are cut on the
// host into ROW BLOCKS ("chunks") of contiguous rows whose values / column ids / row pointers fit one
// shared-memory stage.  A persistent CTA walks a contiguous range of chunks; one elected thread moves
// each chunk's three arrays into shared memory with TMA bulk copies (cp.async.bulk + mbarrier
// complete_tx, L2 evict-first so the vectors stay L2-resident) through an NSTAGE-deep ring, so the matrix
// stream is decoupled from the row structure and always has several chunks in flight.  Each scalar row of
// the staged chunk is owned by L adjacent lanes (L = 1,2,..,32, chosen per chunk on the host so that
// rows x L fills the CTA: L = 1 for the short edge-node rows, 4-16 for the long vertex-node rows); a lane
// walks every L-th entry of its row with 4 independent gathers in flight, so consecutive lanes gather the
// x triple of one node (one L1 line look-up per node) and the sum order is fixed (deterministic; for
// L = 1 it is the sequential order of the CPU loop).  The lane that owns the row then runs the epilogue
// with its vector operands prefetched before the gather phase; neighbouring rows sit in neighbouring
// lanes, so the epilogue loads/stores are coalesced and no shared-memory hand-off is needed.
//
// Fallback kernels (k_spmv_*): one subwarp per (block-)row straight from global memory; used for
// matrices whose rows do not fit a stage and for tiny matrices.
#pragma once
#include "nm_internal.h"

#define NM_SPMV_THREADS 256
#define NM_STREAM_WARPS (NM_SPMV_THREADS / 32)
#define NM_STREAM_MAXDESC 256          // chunk descriptors a CTA keeps in shared memory

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

struct NmStreamArgs {
  const NmChunk* chunks;
  int nchunk;
  int chunks_per_cta;
  const int* rp;            // row pointers: bia (ROW3/KRON3) or ia (CSR)
  const int* idx;           // bja or ja
  const double* val;        // mval (KRON3) or a
  const double* x;
  const double* xg;
  int ncol;
  int vcap, icap, rcap;     // bytes per stage region (each a multiple of 16)
  int nstage;
};

// FMT: NM_FMT_CSR / ROW3 / KRON3.  W: unused by the streaming kernel (kept for the dispatch signature).
template <int FMT, int W, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS, 4) k_stream(NmStreamArgs A, Epi epi) {
  constexpr int R = (FMT == NM_FMT_CSR) ? 1 : 3;          // scalar rows per (block-)row
  constexpr int VPE = (FMT == NM_FMT_ROW3) ? 9 : 1;       // values per index entry
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * A.chunks_per_cta;
  const int nmine = min(A.chunks_per_cta, A.nchunk - c0);
  if (nmine <= 0) return;
  // shared layout: [descs][barriers][stages...]
  NmChunk* sdesc = (NmChunk*)smem;
  uint64_t* bars = (uint64_t*)(smem + NM_STREAM_MAXDESC * sizeof(NmChunk));
  unsigned char* stage0 = (unsigned char*)(bars + 8);
  const int stage_bytes = A.vcap + A.icap + A.rcap;
  if (tid < nmine) sdesc[tid] = A.chunks[c0 + tid];
  if (tid == 0) {
    for (int s = 0; s < A.nstage; ++s) nm_mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t policy = 0;
  auto issue = [&](int it) {
    const NmChunk d = sdesc[it];
    const int s = it % A.nstage;
    unsigned char* st = stage0 + (size_t)s * stage_bytes;
    const char* gv = (const char*)(A.val + (size_t)d.e0 * VPE);
    const char* gi = (const char*)(A.idx + d.e0);
    const char* gr = (const char*)(A.rp + d.rb0);
    const uint32_t mv = (uint32_t)((uintptr_t)gv & 15), mi = (uint32_t)((uintptr_t)gi & 15), mr = (uint32_t)((uintptr_t)gr & 15);
    const uint32_t bv = (mv + (uint32_t)d.ne * VPE * 8 + 15) & ~15u;
    const uint32_t bi = (mi + (uint32_t)d.ne * 4 + 15) & ~15u;
    const uint32_t br = (mr + (uint32_t)((d.nr_l & 0xffff) + 1) * 4 + 15) & ~15u;
    nm_mbar_expect_tx(bars + s, bv + bi + br);
    nm_bulk_g2s(st, gv - mv, bv, bars + s, policy);
    nm_bulk_g2s(st + A.vcap, gi - mi, bi, bars + s, policy);
    nm_bulk_g2s(st + A.vcap + A.icap, gr - mr, br, bars + s, policy);
  };
  if (tid == 0) {
    policy = nm_policy_evict_first();
    for (int it = 0; it < min(A.nstage, nmine); ++it) issue(it);
  }
  const double* __restrict__ x = A.x;
  const double* __restrict__ xg = A.xg;
  const int ncol = A.ncol;
  for (int it = 0; it < nmine; ++it) {
    const NmChunk d = sdesc[it];
    const int s = it % A.nstage;
    const int nr = d.nr_l & 0xffff, logL = d.nr_l >> 16, L = 1 << logL;
    // thread = (scalar row of the chunk, lane l of the L lanes sharing that row)
    const int slot = tid >> logL, l = tid & (L - 1);
    const bool active = slot < R * nr;
    const int row = R * d.rb0 + slot;
    typename Epi::In in;
    if (active && l == 0) in = epi.load(row);                    // operands in flight during the gather phase
    unsigned char* st = stage0 + (size_t)s * stage_bytes;
    const double* sv = (const double*)(st + ((uintptr_t)(A.val + (size_t)d.e0 * VPE) & 15));
    const int* si = (const int*)(st + A.vcap + ((uintptr_t)(A.idx + d.e0) & 15));
    const int* sr = (const int*)(st + A.vcap + A.icap + ((uintptr_t)(A.rp + d.rb0) & 15));
    nm_mbar_wait(bars + s, (uint32_t)((it / A.nstage) & 1));
    double acc = 0.0;
    if (active) {
      const int r = slot / R, comp = slot - R * r;
      const int s0 = sr[r] - d.e0, e0 = sr[r + 1] - d.e0;
      if (FMT == NM_FMT_ROW3) {
        // scalar row 3r+comp: its value stream holds 3 consecutive doubles per block column
        const int nb = e0 - s0;
        const double* v = sv + 9 * (size_t)s0 + (size_t)comp * 3 * nb;
        double b0 = 0.0, b1 = 0.0;
        int q = l;
        for (; q + L < nb; q += 2 * L) {
          const int ca = 3 * si[s0 + q], cb = 3 * si[s0 + q + L];
          const double xa0 = nm_ldx(x, xg, ncol, ca), xa1 = nm_ldx(x, xg, ncol, ca + 1), xa2 = nm_ldx(x, xg, ncol, ca + 2);
          const double xb0 = nm_ldx(x, xg, ncol, cb), xb1 = nm_ldx(x, xg, ncol, cb + 1), xb2 = nm_ldx(x, xg, ncol, cb + 2);
          const double* va = v + 3 * q;
          const double* vb = v + 3 * (q + L);
          b0 += va[0] * xa0; b0 += va[1] * xa1; b0 += va[2] * xa2;
          b1 += vb[0] * xb0; b1 += vb[1] * xb1; b1 += vb[2] * xb2;
        }
        if (q < nb) {
          const int ca = 3 * si[s0 + q];
          const double* va = v + 3 * q;
          b0 += va[0] * nm_ldx(x, xg, ncol, ca); b0 += va[1] * nm_ldx(x, xg, ncol, ca + 1); b0 += va[2] * nm_ldx(x, xg, ncol, ca + 2);
        }
        acc = b0 + b1;
      } else {
        // KRON3: column 3*block + comp of the scalar value; CSR: the column itself
        double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int p = s0 + l;
        for (; p + 3 * L < e0; p += 4 * L) {
          int c0 = si[p], c1 = si[p + L], c2 = si[p + 2 * L], c3 = si[p + 3 * L];
          if (FMT == NM_FMT_KRON3) { c0 = 3 * c0 + comp; c1 = 3 * c1 + comp; c2 = 3 * c2 + comp; c3 = 3 * c3 + comp; }
          const double x0 = nm_ldx(x, xg, ncol, c0), x1 = nm_ldx(x, xg, ncol, c1), x2 = nm_ldx(x, xg, ncol, c2),
                       x3 = nm_ldx(x, xg, ncol, c3);
          b0 += sv[p] * x0; b1 += sv[p + L] * x1; b2 += sv[p + 2 * L] * x2; b3 += sv[p + 3 * L] * x3;
        }
        for (; p < e0; p += L) {
          int c0 = si[p];
          if (FMT == NM_FMT_KRON3) c0 = 3 * c0 + comp;
          b0 += sv[p] * nm_ldx(x, xg, ncol, c0);
        }
        acc = (b0 + b1) + (b2 + b3);
      }
    }
    // the L lanes of a row are adjacent and L divides 32: fixed-order butterfly (deterministic)
    for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();                                             // stage s fully consumed by every thread
    if (tid == 0 && it + A.nstage < nmine) issue(it + A.nstage);
    if (active && l == 0) epi.apply(row, acc, in);
  }
}

// ---------------------------------------------------------------- fallback kernels (global-memory subwarp per row)
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

template <int FMT, int W, class Epi>
static inline void nm_stream_launch(NmParcsr& M, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  NmStreamPlan& P = M.plan;
  NmStreamArgs A;
  A.chunks = P.chunks.p; A.nchunk = P.nchunk; A.chunks_per_cta = P.chunks_per_cta;
  A.rp = FMT == NM_FMT_CSR ? M.ia.p : M.bia.p;
  A.idx = FMT == NM_FMT_CSR ? M.ja.p : M.bja.p;
  A.val = FMT == NM_FMT_KRON3 ? M.mval.p : M.a.p;
  A.x = x; A.xg = M.halo.xg.p ? M.halo.xg.p : x; A.ncol = M.ncol;
  A.vcap = P.vcap; A.icap = P.icap; A.rcap = P.rcap; A.nstage = P.nstage;
  static bool attr_set = false;                                  // per template instantiation
  if (!attr_set) {
    NM_CUDA(cudaFuncSetAttribute(k_stream<FMT, W, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  k_stream<FMT, W, Epi><<<P.grid, NM_SPMV_THREADS, P.smem_bytes, c.stream>>>(A, epi);
  c.launches++;
}

// Halo exchange (if any) + SpMV with the given epilogue.  x: device, owned part only.
template <class Epi>
static inline void nm_spmv_epi(NmParcsr& M, const double* x, const Epi& epi) {
  if (M.halo.nghost > 0) nm_halo_exchange(M, x);
  if (M.plan.nchunk > 0) {
    if (M.format == NM_FMT_KRON3) nm_stream_launch<NM_FMT_KRON3, 32, Epi>(M, x, epi);
    else if (M.format == NM_FMT_ROW3) nm_stream_launch<NM_FMT_ROW3, 32, Epi>(M, x, epi);
    else nm_stream_launch<NM_FMT_CSR, 32, Epi>(M, x, epi);
    return;
  }
  const double r = M.avg_row;              // entries one subwarp walks through
  if (r <= 6.0) nm_spmv_launch_w<4, Epi>(M, x, epi);
  else if (r <= 12.0) nm_spmv_launch_w<8, Epi>(M, x, epi);
  else if (r <= 48.0) nm_spmv_launch_w<16, Epi>(M, x, epi);
  else nm_spmv_launch_w<32, Epi>(M, x, epi);
}
