// k_slab: the fused Chebyshev-iteration step (SpMV + three-term vector update, EpiCheb) on the warp-sliced ELL
// slabs of nm_slab.cu.  Replaces pEVSL parcsrmatvec + the ChebIter AXPYs called through pevsl_chebiter_f90
// (src/mod_matvec.f90:480,512; B-solve registration src/mod_pevsl.f90:72-73).
//
// HBM-bound streaming work (0.2 flop/B): no tensor cores.  Why this shape (ncu of the predecessor k_pack,
// profiles/r1b_kpack_summary.md): with L lanes per row, a JDS offset table and a shared-memory partial-sum
// reduction the walk cost ~480 warp instructions per chunk and warp, only 3% of them DFMA, and the kernel sat
// at 52% issue utilisation / 28% of HBM.  Here a chunk holds at most T lanes and ONE THREAD walks ONE index row
// (R = 3 scalar rows of a node for B = M (x) I3, R = 1 for Ap~; rows longer than NM_SLAB_SPLIT entries are
// shared by an aligned group of 2, 4, ... lanes and combined with shuffles): the inner loop is
//     LDS.64 value, LDS.U16 local column, R x LDS.64 x, R x DFMA
// with stride-32 conflict-free value/index reads, the row sums stay in registers, and the fused epilogue follows
// the walk directly -- one __syncthreads per chunk.  The matrix is streamed by TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx, L2 evict-first so the vectors stay L2-resident) through an NSTAGE ring; the x values of
// the NEXT chunk are gathered once per distinct column with cp.async while the current chunk is walked.
#pragma once
#include "nm_spmv.cuh"

#define NM_SLAB_MAXDESC 64             // chunk descriptors a CTA keeps in shared memory

struct NmSlabArgs {
  const unsigned char* blob;
  const NmPackDesc* desc;
  const int* cta_first;
  const double* x;
  const double* xg;
  int ncol;                 // owned scalar columns (gather ids >= ncol come from xg)
  int stage_bytes, xs_doubles, nstage;
  long long* trace;         // NM_SLAB_TRACE: per CTA and chunk 8 clock64 stamps of thread 0 (diagnostic), else null
};
#define NM_SLAB_STAMP(p) do { if (A.trace && tid == 0) A.trace[((size_t)blockIdx.x * NM_SLAB_MAXDESC + it) * 8 + (p)] = clock64(); } while (0)

struct NmSlabView {
  NmSlabHeader h;
  const uint2* tbl;         // per slice: entry offset, width
  const double* sv;
  const int* scols;
  const unsigned short* sidx;
  const unsigned short* slane;   // per lane: bit 15 first lane of its row, bits 10..14 shuffle-tree adds, bits 0..9 local row
};
__device__ __forceinline__ NmSlabView nm_slab_view(const unsigned char* st) {
  NmSlabView v;
  const int4 a = *(const int4*)st;
  const int4 b = *(const int4*)(st + 16);
  v.h.nr = a.x; v.h.nd = a.y; v.h.nslice = a.z; v.h.first = a.w; v.h.nep = b.x; v.h.gmax = b.y;
  v.tbl = (const uint2*)(st + 32);
  v.sv = (const double*)(st + 32 + ((8 * v.h.nslice + 15) & ~15));
  v.scols = (const int*)(v.sv + v.h.nep);
  v.sidx = (const unsigned short*)(v.scols + v.h.nd);
  v.slane = v.sidx + v.h.nep;
  return v;
}

template <int R, int T, class Epi>
__global__ void __launch_bounds__(T) k_slab(NmSlabArgs A, Epi epi) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = A.cta_first[blockIdx.x];
  const int nmine = A.cta_first[blockIdx.x + 1] - c0;
  if (nmine <= 0) return;
  // shared layout: [descs][barriers][xs x2][stages...]
  NmPackDesc* sdesc = (NmPackDesc*)smem;
  uint64_t* bars = (uint64_t*)(smem + NM_SLAB_MAXDESC * sizeof(NmPackDesc));
  double* xs0 = (double*)(smem + NM_SLAB_MAXDESC * sizeof(NmPackDesc) + 64);
  unsigned char* stage0 =
      smem + ((NM_SLAB_MAXDESC * sizeof(NmPackDesc) + 64 + 16 * (size_t)A.xs_doubles + 15) & ~(size_t)15);
  for (int i = tid; i < nmine; i += T) sdesc[i] = A.desc[c0 + i];
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
  // wait for blob `it`, then start the asynchronous gather of its distinct x values (8 bytes per component)
  auto gather = [&](int it) {
    const int s = it % A.nstage;
    nm_mbar_wait(bars + s, (uint32_t)((it / A.nstage) & 1));
    if (A.trace && tid == 0 && it > 0) A.trace[((size_t)blockIdx.x * NM_SLAB_MAXDESC + it - 1) * 8 + 1] = clock64();
    const NmSlabView v = nm_slab_view(stage0 + (size_t)s * A.stage_bytes);
    double* xs = xs0 + (size_t)(it & 1) * A.xs_doubles;
    const int tot = R * v.h.nd;
    for (int j = tid; j < tot; j += T) {
      const int node = (R == 1) ? j : j / 3;
      const int c = R * v.scols[node] + (j - R * node);
      nm_cp_async8(xs + j, c < ncol ? x + c : xg + (c - ncol));
    }
  };
  gather(0);
  nm_cp_async_wait_all();
  __syncthreads();
  for (int it = 0; it < nmine; ++it) {
    const NmSlabView v = nm_slab_view(stage0 + (size_t)(it % A.nstage) * A.stage_bytes);
    const double* xs = xs0 + (size_t)(it & 1) * A.xs_doubles;
    NM_SLAB_STAMP(0);
    if (it + 1 < nmine) gather(it + 1);
    NM_SLAB_STAMP(2);
    const bool walk = warp < v.h.nslice;
    const unsigned lw = walk ? (unsigned)v.slane[tid] : 0u;
    const bool own = (lw & 0x8000u) != 0u;                       // first lane of a row: does its epilogue
    const int row0 = R * (v.h.first + (int)(lw & 0x3ffu));
    typename Epi::In in[R];
    if (own) {
#pragma unroll
      for (int c = 0; c < R; ++c) in[c] = epi.load(row0 + c);
    }
    double acc[R];
#pragma unroll
    for (int c = 0; c < R; ++c) acc[c] = 0.0;
    NM_SLAB_STAMP(3);
    if (walk) {
      const uint2 t = v.tbl[warp];
      const double* pv = v.sv + t.x + lane;
      const unsigned short* pi = v.sidx + t.x + lane;
      const int w = (int)t.y;
#pragma unroll 4
      for (int k = 0; k < w; ++k) {
        const double m = pv[32 * k];
        const double* xp = xs + R * (int)pi[32 * k];
#pragma unroll
        for (int c = 0; c < R; ++c) acc[c] += m * xp[c];
      }
    }
    NM_SLAB_STAMP(4);
    // rows shared by several adjacent lanes of the warp: fixed-order segmented tree sum into the row's first lane
    for (int s = 0; (1 << s) < v.h.gmax; ++s) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        const double other = __shfl_down_sync(0xffffffffu, acc[c], 1 << s);
        if (lw & (1u << (10 + s))) acc[c] += other;
      }
    }
    if (own) {
#pragma unroll
      for (int c = 0; c < R; ++c) epi.apply(row0 + c, acc[c], in[c]);
    }
    NM_SLAB_STAMP(5);
    nm_cp_async_wait_all();
    NM_SLAB_STAMP(6);
    __syncthreads();             // blob it and xs[it&1] consumed by every warp; xs of chunk it+1 complete
    NM_SLAB_STAMP(7);
    if (tid == 0 && it + A.nstage < nmine) issue(it + A.nstage);
  }
}

template <int R, int T, class Epi>
static inline void nm_slab_launch_t(NmParcsr& M, NmSlab& S, const double* x, const Epi& epi) {
  NmCtx& c = nm_ctx();
  NmSlabArgs A;
  A.blob = S.blob.p; A.desc = S.desc.p; A.cta_first = S.cta_first.p;
  A.x = x; A.xg = M.halo.xg_cur ? M.halo.xg_cur : x; A.ncol = M.ncol;
  A.stage_bytes = S.stage_bytes; A.xs_doubles = S.xs_doubles; A.nstage = S.nstage;
  A.trace = S.trace.p;
  static bool attr_set = false;                                  // per template instantiation
  if (!attr_set) {
    NM_CUDA(cudaFuncSetAttribute(k_slab<R, T, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  k_slab<R, T, Epi><<<S.grid, T, S.smem_bytes, c.stream>>>(A, epi);
  c.launches++;
}

// Product through the slabs: x and the epilogue vectors are in pack order (S.order).
template <class Epi>
static inline void nm_spmv_slab_epi(NmParcsr& M, NmSlab& S, const double* x, const Epi& epi, const int* send_idx) {
  nm_halo_exchange(M, x, send_idx);
  const bool blk = M.format == NM_FMT_KRON3;
  if (S.threads == 512) {
    if (blk) nm_slab_launch_t<3, 512, Epi>(M, S, x, epi); else nm_slab_launch_t<1, 512, Epi>(M, S, x, epi);
  } else if (S.threads == 256) {
    if (blk) nm_slab_launch_t<3, 256, Epi>(M, S, x, epi); else nm_slab_launch_t<1, 256, Epi>(M, S, x, epi);
  } else if (S.threads == 64) {
    if (blk) nm_slab_launch_t<3, 64, Epi>(M, S, x, epi); else nm_slab_launch_t<1, 64, Epi>(M, S, x, epi);
  } else {
    if (blk) nm_slab_launch_t<3, 128, Epi>(M, S, x, epi); else nm_slab_launch_t<1, 128, Epi>(M, S, x, epi);
  }
}
