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
// The Chebyshev iterations (the hot loop) run on k_slabws / k_slab of nm_slab.cuh; the kernels in this file are the
// general products (A, Ad, E, ET with the fused filter epilogue), the round-1a iteration kernels kept as fallbacks
// and regression references (k_pack, k_sell), and the plain subwarp-per-row kernels.
//
// k_pack: HBM-bound streaming work, no tensor cores.  The matrix is stored once more in the
// packed row-block format of nm_pack.cu: locality-ordered rows cut into chunks, each chunk one contiguous blob
// (values in jagged-diagonal order, the chunk's row ids, its DISTINCT column ids, 16-bit chunk-local column
// indices).  A persistent CTA walks a contiguous range of chunks; one elected thread moves each blob into
// shared memory with a TMA bulk copy (cp.async.bulk + mbarrier complete_tx, L2 evict-first so the vectors stay
// L2-resident) through an NSTAGE-deep ring, so the matrix stream never waits on the row structure.  The x
// values a chunk needs are gathered from global memory ONCE per distinct column into shared memory (the L1
// data pipe, one wavefront per cycle per SM, is what bounds a gather-per-entry SpMV on this part: ncu showed
// 76% L1 throughput at 30% of HBM); the inner loop then runs entirely out of shared memory with conflict-free
// JDS reads.  A scalar row is owned by L lanes (L chosen per chunk so rows x L fills the CTA); partial sums are
// combined in a fixed order (deterministic) and the row's thread applies the fused epilogue with operands it
// prefetched before the walk.
//
// Fallback kernels (k_spmv_*): one subwarp per (block-)row straight from global memory; used for
// matrices whose rows do not fit a stage and for tiny matrices.
#pragma once
#include "nm_internal.h"

#define NM_SPMV_THREADS 256
#define NM_PACK_MAXDESC 256            // chunk descriptors a CTA keeps in shared memory

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

struct NmPackArgs {
  const unsigned char* blob;
  const NmPackDesc* desc;
  int nchunk;
  int chunks_per_cta;
  const double* x;
  const double* xg;
  int ncol;
  int stage_bytes, xs_doubles, nstage;
  int dbg;                  // diagnostics (NM_PACK_DBG): 1 skip the x gather, 2 skip the JDS walk, 4 skip the epilogue
};

__device__ __forceinline__ void nm_cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(nm_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void nm_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// decoded view of a chunk blob sitting in shared memory
struct NmPackView {
  int nr, nd, ne, maxlen, L;
  const double* sv;
  const int* srows;
  const int* scols;
  const unsigned short* soff;
  const unsigned short* slen;
  const unsigned short* slidx;
};
template <int VPE>
__device__ __forceinline__ NmPackView nm_pack_view(const unsigned char* st) {
  const NmPackHeader h = *(const NmPackHeader*)st;
  NmPackView v;
  v.nr = h.nr; v.nd = h.nd; v.ne = h.ne; v.maxlen = h.maxlen_L & 0xffff; v.L = h.maxlen_L >> 16;
  v.sv = (const double*)(st + 16);
  v.srows = (const int*)(v.sv + (size_t)VPE * v.ne);
  v.scols = v.srows + v.nr;
  v.soff = (const unsigned short*)(((uintptr_t)(v.scols + v.nd) + 7) & ~(uintptr_t)7);
  v.slen = v.soff + (v.maxlen + 1);
  v.slidx = v.slen + v.nr;
  return v;
}

// One persistent CTA = a contiguous range of chunks, software-pipelined so that ONE barrier per chunk remains:
//   iteration it:  wait for blob it+1 (TMA, mbarrier ring) and start the asynchronous gather (cp.async, 8 bytes
//                  per distinct column component) of the x values chunk it+1 needs into the other xs buffer;
//                  load the epilogue operands of chunk it;
//                  walk the JDS columns of chunk it out of shared memory (lane = index row x one of L lanes);
//                  cp.async.wait_all + __syncthreads;  re-arm the freed stage with chunk it+NSTAGE;
//                  fixed-order reduction of the L partial sums and fused epilogue, one thread per scalar row.
template <int FMT, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS, 3) k_pack(NmPackArgs A, Epi epi) {
  constexpr int R = (FMT == NM_FMT_CSR) ? 1 : 3;          // scalar rows (and columns) per index entry
  constexpr int VPE = (FMT == NM_FMT_ROW3) ? 9 : 1;
  constexpr int PS = 3 * NM_SPMV_THREADS;                 // doubles per partial-sum buffer
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * A.chunks_per_cta;
  const int nmine = min(A.chunks_per_cta, A.nchunk - c0);
  if (nmine <= 0) return;
  // shared layout: [descs][barriers][xs x2][partials x2][stages...]
  NmPackDesc* sdesc = (NmPackDesc*)smem;
  uint64_t* bars = (uint64_t*)(smem + NM_PACK_MAXDESC * sizeof(NmPackDesc));
  double* xs0 = (double*)(bars + 8);
  double* ps0 = xs0 + 2 * (size_t)A.xs_doubles;
  unsigned char* stage0 = smem + ((NM_PACK_MAXDESC * sizeof(NmPackDesc) + 64 + 16 * (size_t)A.xs_doubles + 16 * PS + 15) &
                                  ~(size_t)15);
  if (tid < nmine) sdesc[tid] = A.desc[c0 + tid];
  if (tid == 0) {
    for (int s = 0; s < A.nstage; ++s) nm_mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint64_t policy = 0;
  auto issue = [&](int it) {
    const NmPackDesc d = sdesc[it];
    const int s = it % A.nstage;
    nm_mbar_expect_tx(bars + s, d.bytes);
    nm_bulk_g2s(stage0 + (size_t)s * A.stage_bytes, A.blob + 16ull * d.off16, d.bytes, bars + s, policy);
  };
  if (tid == 0) {
    policy = nm_policy_evict_first();
    for (int it = 0; it < min(A.nstage, nmine); ++it) issue(it);
  }
  const double* __restrict__ x = A.x;
  const double* __restrict__ xg = A.xg;
  const int ncol = A.ncol;
  // wait for blob `it`, then start the asynchronous gather of its distinct x values
  auto gather = [&](int it) {
    const int s = it % A.nstage;
    nm_mbar_wait(bars + s, (uint32_t)((it / A.nstage) & 1));
    const NmPackView v = nm_pack_view<VPE>(stage0 + (size_t)s * A.stage_bytes);
    double* xs = xs0 + (size_t)(it & 1) * A.xs_doubles;
    if (A.dbg & 1) return;
    for (int j = tid; j < R * v.nd; j += NM_SPMV_THREADS) {
      const int node = j / R;
      const int c = R * v.scols[node] + (j - R * node);
      nm_cp_async8(xs + j, c < ncol ? x + c : xg + (c - ncol));
    }
  };
  gather(0);
  nm_cp_async_wait_all();
  __syncthreads();
  for (int it = 0; it < nmine; ++it) {
    const NmPackView v = nm_pack_view<VPE>(stage0 + (size_t)(it % A.nstage) * A.stage_bytes);
    const int nr = v.nr, L = v.L, S = R * nr;                    // S scalar rows (<= 256)
    const double* xs = xs0 + (size_t)(it & 1) * A.xs_doubles;
    double* ps = ps0 + (size_t)(it & 1) * PS;
    if (it + 1 < nmine) gather(it + 1);
    int row = 0;
    typename Epi::In in;
    if (tid < S && !(A.dbg & 4)) {
      const int r = tid / R;
      row = R * v.srows[r] + (tid - R * r);
      in = epi.load(row);
    }
    // ---- JDS walk out of shared memory
    if (FMT == NM_FMT_ROW3) {
      // lane = (l, row r, column component c): 3 partial sums, one per scalar row of the node
      if (tid < S * L && !(A.dbg & 2)) {
        const int l = tid / S, slot = tid - l * S, r = slot / 3, c = slot - 3 * r;
        const int len = v.slen[r];
        const double* p0 = v.sv + c;
        const double* p1 = p0 + 3 * (size_t)v.ne;
        const double* p2 = p1 + 3 * (size_t)v.ne;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 2
        for (int k = l; k < len; k += L) {
          const int p = v.soff[k] + r;
          const double xv = xs[3 * v.slidx[p] + c];
          a0 += p0[3 * p] * xv;
          a1 += p1[3 * p] * xv;
          a2 += p2[3 * p] * xv;
        }
        // partial of scalar row 3r+i from column component c, lane l: ps[i][l][r][c]
        ps[((0 * L + l) * nr + r) * 3 + c] = a0;
        ps[((1 * L + l) * nr + r) * 3 + c] = a1;
        ps[((2 * L + l) * nr + r) * 3 + c] = a2;
      }
    } else {
      // lane = (l, index row r): R accumulators, x gathered once per entry for the R components
      if (tid < nr * L && !(A.dbg & 2)) {
        const int l = tid / nr, r = tid - l * nr;
        const int len = v.slen[r];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 2
        for (int k = l; k < len; k += L) {
          const int p = v.soff[k] + r;
          const double m = v.sv[p];
          const double* xp = xs + R * v.slidx[p];
          a0 += m * xp[0];
          if (R == 3) { a1 += m * xp[1]; a2 += m * xp[2]; }
        }
        double* q = ps + (size_t)(l * nr + r) * R;                // partials [l][r][component]
        q[0] = a0;
        if (R == 3) { q[1] = a1; q[2] = a2; }
      }
    }
    nm_cp_async_wait_all();
    __syncthreads();             // partials of chunk it visible; blob it consumed; xs of chunk it+1 complete
    if (tid == 0 && it + A.nstage < nmine) issue(it + A.nstage);
    if (tid < S && !(A.dbg & 4)) {
      double acc = 0.0;
      if (FMT == NM_FMT_ROW3) {
        const int r = tid / 3, i = tid - 3 * r;
        for (int l = 0; l < L; ++l) {
          const double* q = ps + ((i * L + l) * nr + r) * 3;
          acc += (q[0] + q[1]) + q[2];
        }
      } else {
        for (int l = 0; l < L; ++l) acc += ps[l * S + tid];
      }
      epi.apply(row, acc, in);
    }
  }
}

// ================================================================ sliced-JDS kernel (no staging, no barriers)
struct NmSellArgs {
  const NmSellChunk* chunks;
  const double* val;
  const int* col;
  const int* off;
  const int* rows;
  const int* rowlen;
  const double* x;
  const double* xg;
  int ncol;
};

// One thread block per slice; thread = (scalar row of the slice, lane l of L).  Values and column ids are read
// straight from global memory: in JDS order the lanes of a warp (neighbouring rows, same step) read neighbouring
// words, so the matrix stream is coalesced without shared memory, the kernel runs at full occupancy and every lane
// keeps 4 independent x gathers in flight.  L adjacent lanes share a long row (shuffle butterfly, fixed order).
template <int FMT, class Epi>
__global__ void __launch_bounds__(NM_SPMV_THREADS) k_sell(NmSellArgs A, Epi epi) {
  constexpr int R = (FMT == NM_FMT_CSR) ? 1 : 3;
  const NmSellChunk d = A.chunks[blockIdx.x];
  const int L = d.L;
  const int slot = threadIdx.x / L, l = threadIdx.x - slot * L;
  const bool active = slot < R * d.nr;
  const int r = active ? slot / R : 0, comp = slot - R * (slot / R);
  const int len = active ? A.rowlen[d.r0 + r] : 0;
  const int row = R * A.rows[d.r0 + r] + comp;
  typename Epi::In in;
  if (active && l == 0) in = epi.load(row);
  const int* __restrict__ off = A.off + d.o0;
  const int* __restrict__ col = A.col + d.e0;
  const double* __restrict__ x = A.x;
  const double* __restrict__ xg = A.xg;
  const int ncol = A.ncol;
  double acc = 0.0;
  if (FMT == NM_FMT_ROW3) {
    // scalar row 3r+comp: 3 consecutive values per block column at ((e0 + p)*9 + comp*3)
    const double* __restrict__ val = A.val + 9 * d.e0 + 3 * comp;
    double b0 = 0.0, b1 = 0.0;
    int k = l;
    for (; k + L < len; k += 2 * L) {
      const int pa = off[k] + r, pb = off[k + L] + r;
      const int ca = 3 * col[pa], cb = 3 * col[pb];
      const double* va = val + 9 * (size_t)pa;
      const double* vb = val + 9 * (size_t)pb;
      const double xa0 = nm_ldx(x, xg, ncol, ca), xa1 = nm_ldx(x, xg, ncol, ca + 1), xa2 = nm_ldx(x, xg, ncol, ca + 2);
      const double xb0 = nm_ldx(x, xg, ncol, cb), xb1 = nm_ldx(x, xg, ncol, cb + 1), xb2 = nm_ldx(x, xg, ncol, cb + 2);
      b0 += va[0] * xa0; b0 += va[1] * xa1; b0 += va[2] * xa2;
      b1 += vb[0] * xb0; b1 += vb[1] * xb1; b1 += vb[2] * xb2;
    }
    if (k < len) {
      const int pa = off[k] + r;
      const int ca = 3 * col[pa];
      const double* va = val + 9 * (size_t)pa;
      b0 += va[0] * nm_ldx(x, xg, ncol, ca); b0 += va[1] * nm_ldx(x, xg, ncol, ca + 1); b0 += va[2] * nm_ldx(x, xg, ncol, ca + 2);
    }
    acc = b0 + b1;
  } else {
    const double* __restrict__ val = A.val + d.e0;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
    int k = l;
    for (; k + 3 * L < len; k += 4 * L) {
      const int p0 = off[k] + r, p1 = off[k + L] + r, p2 = off[k + 2 * L] + r, p3 = off[k + 3 * L] + r;
      int c0 = col[p0], c1 = col[p1], c2 = col[p2], c3 = col[p3];
      if (FMT == NM_FMT_KRON3) { c0 = 3 * c0 + comp; c1 = 3 * c1 + comp; c2 = 3 * c2 + comp; c3 = 3 * c3 + comp; }
      const double v0 = val[p0], v1 = val[p1], v2 = val[p2], v3 = val[p3];
      const double x0 = nm_ldx(x, xg, ncol, c0), x1 = nm_ldx(x, xg, ncol, c1), x2 = nm_ldx(x, xg, ncol, c2),
                   x3 = nm_ldx(x, xg, ncol, c3);
      b0 += v0 * x0; b1 += v1 * x1; b2 += v2 * x2; b3 += v3 * x3;
    }
    for (; k < len; k += L) {
      const int p0 = off[k] + r;
      int c0 = col[p0];
      if (FMT == NM_FMT_KRON3) c0 = 3 * c0 + comp;
      b0 += val[p0] * nm_ldx(x, xg, ncol, c0);
    }
    acc = (b0 + b1) + (b2 + b3);
  }
  for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (active && l == 0) epi.apply(row, acc, in);
}

template <int FMT, class Epi>
static inline void nm_sell_launch(NmParcsr& M, NmSell& S, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  NmSellArgs A;
  A.chunks = S.chunks.p; A.val = S.val.p; A.col = S.col.p; A.off = S.off.p; A.rows = S.rows.p; A.rowlen = S.rowlen.p;
  A.x = x; A.xg = M.halo.xg_cur ? M.halo.xg_cur : x; A.ncol = M.ncol;
  k_sell<FMT, Epi><<<S.nchunk, NM_SPMV_THREADS, 0, c.stream>>>(A, epi);
  c.launches++;
}
template <class Epi>
static inline void nm_sell_dispatch(NmParcsr& M, NmSell& S, const double* x, const Epi& epi) {
  if (M.format == NM_FMT_KRON3) nm_sell_launch<NM_FMT_KRON3, Epi>(M, S, x, epi);
  else if (M.format == NM_FMT_ROW3) nm_sell_launch<NM_FMT_ROW3, Epi>(M, S, x, epi);
  else nm_sell_launch<NM_FMT_CSR, Epi>(M, S, x, epi);
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

template <int FMT, class Epi>
static inline void nm_pack_launch(NmParcsr& M, NmPack& P, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  NmPackArgs A;
  A.blob = P.blob.p; A.desc = P.desc.p; A.nchunk = P.nchunk; A.chunks_per_cta = P.chunks_per_cta;
  A.x = x; A.xg = M.halo.xg_cur ? M.halo.xg_cur : x; A.ncol = M.ncol;
  A.stage_bytes = P.stage_bytes; A.xs_doubles = P.xs_doubles; A.nstage = P.nstage;
  static const int dbg = getenv("NM_PACK_DBG") ? atoi(getenv("NM_PACK_DBG")) : 0;
  A.dbg = dbg;
  static bool attr_set = false;                                  // per template instantiation
  if (!attr_set) {
    NM_CUDA(cudaFuncSetAttribute(k_pack<FMT, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  k_pack<FMT, Epi><<<P.grid, NM_SPMV_THREADS, P.smem_bytes, c.stream>>>(A, epi);
  c.launches++;
}

// Product through an explicit pack (NmChebIter's pack-order copy): x and the epilogue vectors are in P's order.
template <class Epi>
static inline void nm_spmv_sell_epi(NmParcsr& M, NmSell& S, const double* x, const Epi& epi, const int* send_idx) {
  nm_halo_exchange(M, x, send_idx);
  nm_sell_dispatch(M, S, x, epi);
}
template <class Epi>
static inline void nm_spmv_pack_epi(NmParcsr& M, NmPack& P, const double* x, const Epi& epi, const int* send_idx) {
  nm_halo_exchange(M, x, send_idx);
  if (M.format == NM_FMT_KRON3) nm_pack_launch<NM_FMT_KRON3, Epi>(M, P, x, epi);
  else if (M.format == NM_FMT_ROW3) nm_pack_launch<NM_FMT_ROW3, Epi>(M, P, x, epi);
  else nm_pack_launch<NM_FMT_CSR, Epi>(M, P, x, epi);
}

// Halo exchange (if any) + SpMV with the given epilogue.  x: device, owned part only.
template <class Epi>
static inline void nm_spmv_epi(NmParcsr& M, const double* x, const Epi& epi) {
  nm_halo_exchange(M, x);
  if (M.sell.nchunk > 0) { nm_sell_dispatch(M, M.sell, x, epi); return; }
  if (M.pack.nchunk > 0) {
    if (M.format == NM_FMT_KRON3) nm_pack_launch<NM_FMT_KRON3, Epi>(M, M.pack, x, epi);
    else if (M.format == NM_FMT_ROW3) nm_pack_launch<NM_FMT_ROW3, Epi>(M, M.pack, x, epi);
    else nm_pack_launch<NM_FMT_CSR, Epi>(M, M.pack, x, epi);
    return;
  }
  const double r = M.avg_row;              // entries one subwarp walks through
  if (r <= 6.0) nm_spmv_launch_w<4, Epi>(M, x, epi);
  else if (r <= 12.0) nm_spmv_launch_w<8, Epi>(M, x, epi);
  else if (r <= 48.0) nm_spmv_launch_w<16, Epi>(M, x, epi);
  else nm_spmv_launch_w<32, Epi>(M, x, epi);
}
