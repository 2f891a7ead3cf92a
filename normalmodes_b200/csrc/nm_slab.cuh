// k_slabws / k_slab: the fused Chebyshev-iteration step (SpMV + three-term vector update, EpiCheb) on the warp-sliced
// ELL slabs of nm_slab.cu.  Replaces pEVSL parcsrmatvec + the ChebIter AXPYs called through pevsl_chebiter_f90
// (src/mod_matvec.f90:480,512; B-solve registration src/mod_pevsl.f90:72-73).
//
// HBM-bound streaming work (0.2 flop/B): no tensor cores.  Why this shape (ncu of the predecessor k_pack,
// profiles/r1b_kpack_summary.md): with L lanes per row, a JDS offset table and a shared-memory partial-sum
// reduction the walk cost ~480 warp instructions per chunk and warp, only 3% of them DFMA, and the kernel sat
// at 52% issue utilisation / 28% of HBM.  Here a chunk holds at most T lanes and ONE LANE walks ONE index row
// (R = 3 scalar rows of a node for B = M (x) I3, R = 1 for Ap~; rows longer than NM_SLAB_SPLIT entries are
// shared by adjacent lanes of one warp and combined with a segmented shuffle tree): the inner loop is
//     LDS.64 value, LDS.U16 local column, R x LDS.64 x, R x DFMA
// with stride-32 conflict-free value/index reads, the row sums stay in registers and the fused epilogue follows the
// walk directly.  The matrix is streamed by TMA bulk copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first so
// the vectors stay L2-resident) through a stage ring; the x values of the chunks AHEAD are gathered once per distinct
// column with cp.async.
//   k_slab     one __syncthreads per chunk (all warps gather, then walk; NM_SLAB_WS=0);
//   k_slabws   (default) warp-specialised: producer warps keep the rings full, consumer warps only walk; stages and x
//              buffers are handed over with full/empty mbarriers; consecutive steps are chained by programmatic
//              dependent launch; on several GPUs (FUSED) the epilogue stores the boundary rows into the peers'
//              flag-in-data slots and the producers poll the slots of the ghost columns they gather;
//   k_slabpers (NM_SLAB_PERS=1) the whole iteration in one cooperative launch: pinned + ring stages, a grid barrier or
//              per-chunk dependency flags between steps, the same in-kernel halo.
#pragma once
#include "nm_spmv.cuh"

#define NM_SLAB_MAXDESC 256            // chunk descriptors a CTA keeps in shared memory

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
  const int* deps;               // h.pad2 chunk ids: the owners of the chunk's (non-ghost) columns
};
__device__ __forceinline__ NmSlabView nm_slab_view(const unsigned char* st) {
  NmSlabView v;
  const int4 a = *(const int4*)st;
  const int4 b = *(const int4*)(st + 16);
  v.h.nr = a.x; v.h.nd = a.y; v.h.nslice = a.z; v.h.first = a.w; v.h.nep = b.x; v.h.gmax = b.y;
  v.h.has_ghost = b.z; v.h.pad2 = b.w;
  v.tbl = (const uint2*)(st + 32);
  v.sv = (const double*)(st + 32 + ((8 * v.h.nslice + 15) & ~15));
  v.scols = (const int*)(v.sv + v.h.nep);
  v.sidx = (const unsigned short*)(v.scols + v.h.nd);
  v.slane = v.sidx + v.h.nep;
  v.deps = (const int*)(v.slane + 32 * v.h.nslice);
  return v;
}

// Asynchronous gather of a chunk's distinct x values (8 bytes per scalar component) into xs, by `nthreads` threads;
// 4 independent column-id loads / address computations / cp.async per round so the issue latency chain overlaps.
template <int R>
__device__ __forceinline__ void nm_slab_gather(const NmSlabView& v, double* xs, const double* __restrict__ x,
                                               const double* __restrict__ xg, int ncol, int t, int nthreads) {
  const int tot = R * v.h.nd;
  int j = t;
  for (; j + 3 * nthreads < tot; j += 4 * nthreads) {
    int node[4], cid[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int jj = j + u * nthreads; node[u] = (R == 1) ? jj : jj / 3; }
#pragma unroll
    for (int u = 0; u < 4; ++u) cid[u] = v.scols[node[u]];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int jj = j + u * nthreads;
      const int c = R * cid[u] + (jj - R * node[u]);
      nm_cp_async8(xs + jj, c < ncol ? x + c : xg + (c - ncol));
    }
  }
  for (; j < tot; j += nthreads) {
    const int node = (R == 1) ? j : j / 3;
    const int c = R * v.scols[node] + (j - R * node);
    nm_cp_async8(xs + j, c < ncol ? x + c : xg + (c - ncol));
  }
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
    nm_slab_gather<R>(v, xs, x, xg, ncol, tid, T);
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

// ---------------------------------------------------------------- warp-specialised variant
// ncu + clock64 traces of k_slab (profiles/r1c_kslab_summary.md): the SM's load/store pipe is the busy unit (shared-
// memory wavefronts of the walk + the x gather + the epilogue), but it idles half of the time because a CTA's
// phases -- wait for the blob, issue the gather, walk, epilogue -- are serialised by the per-chunk __syncthreads.
// k_slabws splits the CTA: NP producer warps keep the TMA ring full and issue the cp.async gathers of the chunks
// AHEAD (completion tracked by mbarriers through cp.async.mbarrier.arrive), NC consumer warps only walk and run the
// epilogue; stages and x buffers are handed over with full/empty mbarriers, no CTA-wide barrier inside the loop.
__device__ __forceinline__ void nm_mbar_wait_bounded(uint64_t* b, uint32_t parity) {
  uint32_t done;
  const uint32_t a = nm_smem_u32(b);
  const long long t0 = clock64();
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (!done && clock64() - t0 > 40000000000ll) __trap();      // ~20 s (above the 10 s peer waits): a lost arrival must not hang the GPU
  } while (!done);
}
__device__ __forceinline__ void nm_mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nm_smem_u32(b)) : "memory");
}
__device__ __forceinline__ void nm_cp_async_mbar_arrive_noinc(uint64_t* b) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(nm_smem_u32(b)) : "memory");
}

struct NmSlabWsArgs {
  NmSlabArgs a;
  int nxs;                  // x buffers in the ring
  int np;                   // producer warps
  // several GPUs: arrival flags of the ghost values this step gathers (raised by the peers' k_halo_push); polled by
  // the producers before the first chunk that has ghost columns (those come last in every CTA's range)
  const unsigned long long* hflags;
  unsigned hmask;
  unsigned long long hepoch;
  int* hstatus;
  int ghost_cg;             // 1: read ghost values with ld.global.cg (synchronous); 0: cp.async like the owned ones
};

// Fused multi-GPU step (default on several GPUs): no kernel, fence, flag or atomic between two steps.  The first lane of a
// boundary row stores its new direction straight into the peers' flag-in-data ghost slots (nm_ll_store: NVLink peer
// window, 16 bytes per value, tag = epoch of the step) from the epilogue; the next step's producers poll the slots they
// gather (nm_ll_load) before the chunks that have ghost columns, which come last in every CTA's range -- the values were
// sent a whole step earlier.  Three rotating slot buffers: the kernel boundary keeps a GPU's CTAs within one step,
// and a peer can only enter step k+2 after ALL of this rank's step k+1 (it gathers from it).  Measured against the
// round-1 scheme (peer stores + system fences + arrival flags: two serialised fence.sys round trips and a flag flight
// per step): 200k-tet bench workload on 2 GPUs 14.5 s -> 9.8 s per filter application (profiles/r2*).
struct NmSlabFusedArgs {
  const int* push_off;          // per pack-order index row: [off, off') into push_ent; null = nothing to push
  const NmPushEnt* push_ent;    // peer, scalar component, position in that peer's ghost buffer
  unsigned long long* ll_out[8];        // this rank's block of slots in each peer's buffer of tag_out
  const unsigned long long* ll_in;      // this rank's ghost slots of tag_in (null: no ghosts gathered through slots)
  unsigned tag_in, tag_out;
  const int* gslot;             // ghost column -> slot
  int* status;
  int debug;                    // NM_DEBUG_LL bit 0: no tag wait (timing diagnostics only)
};

// ghost values were written by a peer GPU during this kernel's lifetime: read them through L2 (no L1 allocation)
__device__ __forceinline__ double nm_ld_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
template <int R>
__device__ __forceinline__ void nm_slab_gather_ghost(const NmSlabView& v, double* xs, const double* __restrict__ x,
                                                     const double* xg, int ncol, int t, int nthreads) {
  const int tot = R * v.h.nd;
  for (int j = t; j < tot; j += nthreads) {
    const int node = (R == 1) ? j : j / 3;
    const int c = R * v.scols[node] + (j - R * node);
    if (c < ncol) nm_cp_async8(xs + j, x + c);
    else xs[j] = nm_ld_cg(xg + (c - ncol));
  }
  __threadfence_block();                                         // the plain stores above precede the barrier arrival
}

template <int R>
__device__ __forceinline__ void nm_slab_gather_ll(const NmSlabView& v, double* xs, const double* __restrict__ x,
                                                  const unsigned long long* ll, const int* __restrict__ gslot, unsigned tag,
                                                  int* status, int ncol, int t, int nthreads, bool nowait = false,
                                                  int lmode = 0) {
  const int tot = R * v.h.nd;
  // LLB independent column ids / slot loads in flight per thread: a ghost value is a dependent 16-byte L2 load, and a
  // boundary chunk has hundreds of them.  NM_DEBUG_LL (diagnostics, bit 0): do not wait for the tags (timing of the
  // kernel without the exchange latency; results are then wrong)
  constexpr int LLB = 2;
  for (int j = t; j < tot; j += LLB * nthreads) {
    int cc[LLB];
    unsigned a[LLB], fa[LLB], b[LLB], fb[LLB];
#pragma unroll
    for (int u = 0; u < LLB; ++u) {
      const int jj = j + u * nthreads;
      cc[u] = -1;
      if (jj < tot) {
        const int node = (R == 1) ? jj : jj / 3;
        cc[u] = R * v.scols[node] + (jj - R * node);
      }
    }
    if (gslot) {                                                 // ghost column -> its slot (sender's pack order layout)
#pragma unroll
      for (int u = 0; u < LLB; ++u) if (cc[u] >= ncol) cc[u] = ncol + gslot[cc[u] - ncol];
    }
#pragma unroll
    for (int u = 0; u < LLB; ++u) {
      const int jj = j + u * nthreads;
      fa[u] = fb[u] = tag; a[u] = b[u] = 0u;
      if (cc[u] < 0) continue;
      if (cc[u] < ncol) { nm_cp_async8(xs + jj, x + cc[u]); continue; }
      nm_ll_load_raw(ll + 2 * (size_t)(cc[u] - ncol), lmode, a[u], fa[u], b[u], fb[u]);
    }
#pragma unroll
    for (int u = 0; u < LLB; ++u) {
      if (cc[u] < ncol) continue;                                // nothing (-1) or an owned column (cp.async above)
      double val = __longlong_as_double((long long)(((unsigned long long)b[u] << 32) | a[u]));
      if ((fa[u] != tag || fb[u] != tag) && !nowait) {
        val = 0.0;
        const unsigned long long* slot = ll + 2 * (size_t)(cc[u] - ncol);
        const long long t0 = clock64();
        while (!nm_ll_load(slot, tag, &val)) {
          if (*(volatile int*)status != 0) break;                                  // an earlier wait already failed
          if (clock64() - t0 > 20000000000ll) { atomicOr(status, 1); break; }      // ~10 s: peer stalled or died
        }
      }
      xs[j + u * nthreads] = val;
    }
  }
  __threadfence_block();                                         // the plain stores above precede the barrier arrival
}

template <int R, int NC, bool FUSED, class Epi>
__global__ void __launch_bounds__(32 * (NC + 8)) k_slabws(NmSlabWsArgs W, Epi epi, NmSlabFusedArgs F) {
  const NmSlabArgs& A = W.a;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = A.cta_first[blockIdx.x];
  const int nmine = A.cta_first[blockIdx.x + 1] - c0;
  if (nmine <= 0) return;
  const int S = A.nstage, X = W.nxs, NP = W.np;
  // shared layout: [descs][full_blob 8][empty_blob 8][full_xs 8][empty_xs 8][xs x X][stages x S]
  NmPackDesc* sdesc = (NmPackDesc*)smem;
  uint64_t* full_blob = (uint64_t*)(smem + NM_SLAB_MAXDESC * sizeof(NmPackDesc));
  uint64_t* empty_blob = full_blob + 8;
  uint64_t* full_xs = full_blob + 16;
  uint64_t* empty_xs = full_blob + 24;
  double* xs0 = (double*)(full_blob + 32);
  unsigned char* stage0 =
      smem + ((NM_SLAB_MAXDESC * sizeof(NmPackDesc) + 256 + 8 * (size_t)X * A.xs_doubles + 15) & ~(size_t)15);
  // programmatic dependent launch: the NEXT step's kernel may start as SM resources free up; everything it does before
  // its griddepcontrol.wait (descriptors, barrier init, the first TMA blob copies) touches read-only matrix data only
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = tid; i < nmine; i += blockDim.x) sdesc[i] = A.desc[c0 + i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { nm_mbar_init(full_blob + s, 1); nm_mbar_init(empty_blob + s, NC); }
    for (int x = 0; x < X; ++x) { nm_mbar_init(full_xs + x, 32 * NP); nm_mbar_init(empty_xs + x, NC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp >= NC) {
    // ================================================= producers
    const int pw = warp - NC;                                   // producer warp index
    const int ptid = pw * 32 + lane, pthreads = 32 * NP;
    uint64_t policy = 0;
    auto issue = [&](int j) {
      const NmPackDesc d = sdesc[j];
      const int s = j % S;
      nm_mbar_expect_tx(full_blob + s, d.bytes);
      nm_bulk_g2s(stage0 + (size_t)s * A.stage_bytes, A.blob + 16ull * d.off16, d.bytes, full_blob + s, policy);
    };
    if (ptid == 0) {
      policy = nm_policy_evict_first();
      for (int j = 0; j < min(S, nmine); ++j) issue(j);
    }
    const double* __restrict__ x = A.x;
    const double* __restrict__ xg = A.xg;
    const int ncol = A.ncol;
    asm volatile("griddepcontrol.wait;" ::: "memory");           // the previous step's vectors are complete and visible
    bool flags_seen = false;
    for (int it = 0; it < nmine; ++it) {
      // a. x values of chunk it -> xs[it % X] (free once the consumers are done with chunk it - X)
      const int s = it % S, xb = it % X;
      nm_mbar_wait_bounded(full_blob + s, (uint32_t)((it / S) & 1));
      if (it >= X) nm_mbar_wait_bounded(empty_xs + xb, (uint32_t)(((it / X) - 1) & 1));
      const NmSlabView v = nm_slab_view(stage0 + (size_t)s * A.stage_bytes);
      double* xs = xs0 + (size_t)xb * A.xs_doubles;
      if (FUSED && v.h.has_ghost && F.ll_in) {
        nm_slab_gather_ll<R>(v, xs, x, F.ll_in, F.gslot, F.tag_in, F.status, ncol, ptid, pthreads, (F.debug & 1) != 0, (F.debug >> 2) & 3);
      } else if (v.h.has_ghost && W.hmask) {
        if (!flags_seen) {
          if (lane == 0) {
            for (int r = 0; r < 8; ++r) {
              if (!(W.hmask & (1u << r))) continue;
              unsigned long long f;
              const long long t0 = clock64();
              for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(W.hflags + r) : "memory");
                if (f >= W.hepoch) break;
                // ~10 s: peer stalled or died; once the status word is set every later wait gives up at once, so a
                // failed run ends at the host's next status check instead of stalling per kernel
                if (*(volatile int*)W.hstatus != 0) break;
                if (clock64() - t0 > 20000000000ll) { atomicOr(W.hstatus, 1); break; }
              }
            }
          }
          __syncwarp();
          flags_seen = true;
        }
        // the ghost buffer of this parity is only read after the flags, and L1 does not outlive a grid (the d buffers
        // alternate the same way): the asynchronous gather is as safe for ghosts as for owned values
        if (W.ghost_cg) nm_slab_gather_ghost<R>(v, xs, x, xg, ncol, ptid, pthreads);
        else nm_slab_gather<R>(v, xs, x, xg, ncol, ptid, pthreads);
      } else {
        nm_slab_gather<R>(v, xs, x, xg, ncol, ptid, pthreads);
      }
      nm_cp_async_mbar_arrive_noinc(full_xs + xb);
      // b. refill the stage of chunk it-1 (consumed once all NC warps released it) with chunk it-1+S
      if (ptid == 0 && it >= 1) {
        const int j = it - 1 + S;
        if (j < nmine) {
          nm_mbar_wait_bounded(empty_blob + (j % S), (uint32_t)(((j / S) - 1) & 1));
          issue(j);
        }
      }
      __syncwarp();
    }
    return;
  }
  // =================================================== consumers
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int it = 0; it < nmine; ++it) {
    const int s = it % S, xb = it % X;
    nm_mbar_wait_bounded(full_blob + s, (uint32_t)((it / S) & 1));
    const NmSlabView v = nm_slab_view(stage0 + (size_t)s * A.stage_bytes);
    const double* xs = xs0 + (size_t)xb * A.xs_doubles;
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
    uint2 t = make_uint2(0u, 0u);
    if (walk) t = v.tbl[warp];
    nm_mbar_wait_bounded(full_xs + xb, (uint32_t)((it / X) & 1));
    if (walk) {
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
    const int gmax = v.h.gmax;
    __syncwarp();
    if (lane == 0) { nm_mbar_arrive(empty_blob + s); nm_mbar_arrive(empty_xs + xb); }   // blob and x buffer consumed
    for (int q = 0; (1 << q) < gmax; ++q) {
#pragma unroll
      for (int c = 0; c < R; ++c) {
        const double other = __shfl_down_sync(0xffffffffu, acc[c], 1 << q);
        if (lw & (1u << (10 + q))) acc[c] += other;
      }
    }
    if (own) {
      if constexpr (FUSED) {
        double dn[R];
#pragma unroll
        for (int c = 0; c < R; ++c) dn[c] = epi.apply_dn(row0 + c, acc[c], in[c]);
        if (F.push_off && v.h.has_ghost) {                        // symmetric pattern: sent rows gather ghost columns
          const int prow = v.h.first + (int)(lw & 0x3ffu);
          for (int e = F.push_off[prow]; e < F.push_off[prow + 1]; ++e) {
            const NmPushEnt pe = F.push_ent[e];
            double val = dn[0];
#pragma unroll
            for (int c = 1; c < R; ++c) if (pe.comp == c) val = dn[c];
            if (!(F.debug & 2)) nm_ll_store_mode(F.ll_out[pe.peer] + 2 * (size_t)pe.dst, val, F.tag_out, (F.debug >> 4) & 3);   // bit 1: timing without the stores
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < R; ++c) epi.apply(row0 + c, acc[c], in[c]);
      }
    }
  }
}

// ---------------------------------------------------------------- persistent variant: the WHOLE iteration in one launch
// "one persistent kernel per solve" (opt-in, NM_SLAB_PERS=1).  Motivation: the per-step launches of
// k_slabws cost ~5 us of drain + launch + first-gather latency per step even with dependent launch (an Ap~ step on the
// 200k-tet workload is 9 us whatever the tile shape, profiles/r1d_sweep_slab.json).  Measured outcome (one B200,
// profiles/r2b_kernel_times_pers_barrier.json, gpurun_out/r2c_*): the software grid barrier costs MORE than the
// hardware's grid completion + dependent launch -- B~ step 32.8 us against 28.4, Ap~ 10.4 against 8.7 -- and the
// dataflow variant more still (42.4 / 17.5: one poll + fence + named barrier per chunk visit serialises the
// producers); on 2 GPUs the barrier variant is ~5% ahead of the per-step kernels (profiles/r2l_*).  Kept, tested, not
// the default.  k_slabpers is launched cooperatively ONCE per solve: every CTA keeps its chunk range for all deg steps;
//   * matrix: the first P chunks of a CTA stay pinned in their stages (all of them when the range fits the ring: the
//     slab is then read from HBM once per SOLVE); the others stream through 2 ring stages, the copies of step k+1
//     running ahead across the step boundary (the matrix is read-only);
//   * vectors: a row's r, x, d are read and written by the same thread in every step (same-thread RAW through
//     global memory); only the GATHERED direction crosses CTAs -> one grid barrier per step, waited for by the
//     producer warps only (arrive: one fence + atomic per CTA after its consumers' named barrier; wait: one polling
//     thread, then a gpu-scope fence -- which also drops the stale L1 lines of the alternating d buffers -- and the
//     producers' named barrier).  Consumer warps never wait for the grid: they block on full_xs as before;
//   * halo (several GPUs): boundary rows go straight from the epilogue into the peers' flag-in-data slots
//     (nm_ll_store, three rotating buffers, tag = step epoch); the producers poll the slots they gather
//     (nm_ll_load).  No fence, flag or atomic between GPUs; chunks with ghost columns come last in every CTA's range
//     and their rows are sent a whole step before they are needed.  Why three buffers are enough: the local grid
//     barrier keeps a GPU's CTAs within one step of each other, and a peer's step k+2 needs ALL of our step k+1.
struct NmSlabPersArgs {
  NmSlabArgs a;             // blob, desc, cta_first, ncol, stage_bytes, xs_doubles, nstage (x, xg unused)
  int nxs, np, deg;
  const double* b;          // right-hand side, pack order
  double* r; double* d0; double* d1; double* xout;
  const double* ak; const double* bk;
  double inv_theta;
  unsigned long long* gbar; // grid-barrier counter (monotonic over the solves of this ChebIter)
  unsigned long long gbase; // its value when this launch starts
  // dataflow execution (null: grid barrier per step instead): per-chunk completion flags, chunk id per descriptor
  unsigned* cflag; const int* desc_cid;
  unsigned ftag0;           // a chunk that has finished step k carries ftag0 + k + 1
  // several GPUs (null / 0 on one)
  const unsigned long long* ll_in[3];   // this rank's ghost slots, by tag % 3
  unsigned long long* ll_out[3][8];     // this rank's block of slots in each peer's buffers
  const int* push_off; const NmPushEnt* push_ent;
  const int* gslot;         // ghost column -> slot
  unsigned tag0;            // tag of the values step 0 gathers (pushed by nm_halo_push_ll)
  int* hstatus;
};

// shared memory in front of the x buffers: descriptors, chunk ids, 5 x 8 mbarriers
#define NM_SLABPERS_FIXED (NM_SLAB_MAXDESC * (sizeof(NmPackDesc) + sizeof(int)) + 384)
template <int R, int NC>
__global__ void __launch_bounds__(32 * (NC + 9)) k_slabpers(NmSlabPersArgs W) {
  const NmSlabArgs& A = W.a;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = A.cta_first[blockIdx.x];
  const int n = A.cta_first[blockIdx.x + 1] - c0;                // >= 1 for every CTA (packer)
  const int S = A.nstage, X = W.nxs, NP = W.np, deg = W.deg;
  const bool flow = W.cflag != nullptr;                          // dataflow: per-chunk flags instead of the grid barrier
  NmPackDesc* sdesc = (NmPackDesc*)smem;
  int* scid = (int*)(smem + NM_SLAB_MAXDESC * sizeof(NmPackDesc));
  uint64_t* full_blob = (uint64_t*)(smem + NM_SLAB_MAXDESC * (sizeof(NmPackDesc) + sizeof(int)));
  uint64_t* empty_blob = full_blob + 8;
  uint64_t* full_xs = full_blob + 16;
  uint64_t* empty_xs = full_blob + 24;
  uint64_t* done_xs = full_blob + 32;                            // dataflow: all consumer warps have STORED a visit's rows
  double* xs0 = (double*)(full_blob + 48);
  unsigned char* stage0 = smem + ((NM_SLABPERS_FIXED + 8 * (size_t)X * A.xs_doubles + 15) & ~(size_t)15);
  for (int i = tid; i < n; i += blockDim.x) { sdesc[i] = A.desc[c0 + i]; if (flow) scid[i] = W.desc_cid[c0 + i]; }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { nm_mbar_init(full_blob + s, 1); nm_mbar_init(empty_blob + s, NC); }
    // dataflow: an x buffer is reused only after the publisher warp has raised the visit's flag as well
    for (int x = 0; x < X; ++x) { nm_mbar_init(full_xs + x, 32 * NP); nm_mbar_init(empty_xs + x, flow ? NC + 1 : NC); nm_mbar_init(done_xs + x, NC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // pinned chunks [0, P) sit in stages [0, P); the others cycle through the RG ring stages behind them in the order
  // q = k (n - P) + (it - P) over the whole solve
  const int P = n <= S ? n : S - 2, RG = n <= S ? 0 : 2, NR = n - P;
  const bool multi = W.push_off != nullptr;
  if (warp >= NC + NP) {
    // ================================================= publisher (dataflow only): one thread
    // When every consumer warp has stored its rows of a visit (done_xs), ONE gpu-scope release makes them visible and
    // raises the chunk's flag; only then may the visit's x buffer be reused, so the publisher is never lapped.
    if (!flow || lane != 0) return;
    for (int k = 0; k < deg; ++k)
      for (int it = 0; it < n; ++it) {
        const long long gx = (long long)k * n + it;
        const int xb = (int)(gx % X);
        nm_mbar_wait_bounded(done_xs + xb, (uint32_t)((gx / X) & 1));
        if (k < deg - 1)
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(W.cflag + scid[it]), "r"(W.ftag0 + (unsigned)k + 1u) : "memory");
        nm_mbar_arrive(empty_xs + xb);
      }
    return;
  }
  if (warp >= NC) {
    // ================================================= producers
    const int pw = warp - NC;
    const int ptid = pw * 32 + lane, pthreads = 32 * NP;
    uint64_t policy = 0;
    const long long Q = (long long)deg * NR;                     // ring copies of the whole solve
    long long issued = 0;
    auto issue = [&](int chunk, int s) {
      const NmPackDesc d = sdesc[chunk];
      nm_mbar_expect_tx(full_blob + s, d.bytes);
      nm_bulk_g2s(stage0 + (size_t)s * A.stage_bytes, A.blob + 16ull * d.off16, d.bytes, full_blob + s, policy);
    };
    // ring copy `issued` goes into stage P + issued % RG once copy issued - RG has been released by the consumers;
    // need >= 0: block until copy `need` is on its way, otherwise only take what is free
    auto pump = [&](long long need) {
      while (issued < Q) {
        const long long rel = issued - RG;
        if (rel >= 0) {
          uint64_t* eb = empty_blob + P + (int)(rel % RG);
          const uint32_t par = (uint32_t)((rel / RG) & 1);
          if (issued <= need) nm_mbar_wait_bounded(eb, par);
          else {
            uint32_t done;
            asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(nm_smem_u32(eb)), "r"(par) : "memory");
            if (!done) break;
          }
        }
        issue(P + (int)(issued % NR), P + (int)(issued % RG));
        ++issued;
      }
    };
    if (ptid == 0) {
      policy = nm_policy_evict_first();
      for (int j = 0; j < P; ++j) issue(j, j);
      pump(-1);
    }
    const int ncol = A.ncol;
    for (int k = 0; k < deg; ++k) {
      const double* __restrict__ x = k == 0 ? W.b : ((k & 1) ? W.d0 : W.d1);      // direction written by step k-1
      if (k > 0 && !flow) {
        // grid barrier: every CTA has stored (and fenced) its step k-1 directions
        if (ptid == 0) {
          const unsigned long long target = W.gbase + (unsigned long long)k * gridDim.x;
          unsigned long long v;
          const long long t0 = clock64();
          for (;;) {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(W.gbar) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > 60000000000ll) __trap();                         // ~30 s (above the peer and mbarrier limits): never hang the GPU
          }
          __threadfence();                                                       // also drops this SM's stale L1 lines of d
        }
        asm volatile("bar.sync 1, %0;" ::"r"(pthreads) : "memory");
      }
      const unsigned tag = W.tag0 + (unsigned)k;
      const unsigned long long* ll = multi ? W.ll_in[tag % 3u] : nullptr;
      for (int it = 0; it < n; ++it) {
        const long long gx = (long long)k * n + it;
        const int xb = (int)(gx % X);
        int s; uint32_t par;
        if (it < P) { s = it; par = 0; }
        else {
          const long long q = (long long)k * NR + (it - P);
          s = P + (int)(q % RG); par = (uint32_t)((q / RG) & 1);
          if (ptid == 0) pump(q);
        }
        nm_mbar_wait_bounded(full_blob + s, par);
        const NmSlabView v = nm_slab_view(stage0 + (size_t)s * A.stage_bytes);
        if (flow && k > 0) {
          // dataflow: the chunks owning this chunk's columns have finished step k-1 (relaxed polls, one acquire fence --
          // which also drops the SM's stale L1 lines of the alternating d buffers -- then the producers' barrier)
          const unsigned target = W.ftag0 + (unsigned)k;
          for (int j = ptid; j < v.h.pad2; j += pthreads) {
            const unsigned* f = W.cflag + v.deps[j];
            unsigned fv;
            const long long t0 = clock64();
            for (;;) {
              asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(fv) : "l"(f) : "memory");
              if ((int)(fv - target) >= 0) break;
              if (clock64() - t0 > 60000000000ll) __trap();                       // ~30 s: never hang the GPU
            }
          }
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("bar.sync 1, %0;" ::"r"(pthreads) : "memory");
        }
        if (gx >= X) nm_mbar_wait_bounded(empty_xs + xb, (uint32_t)(((gx / X) - 1) & 1));
        double* xs = xs0 + (size_t)xb * A.xs_doubles;
        if (multi && v.h.has_ghost) nm_slab_gather_ll<R>(v, xs, x, ll, W.gslot, tag, W.hstatus, ncol, ptid, pthreads);
        else nm_slab_gather<R>(v, xs, x, x, ncol, ptid, pthreads);
        nm_cp_async_mbar_arrive_noinc(full_xs + xb);
        if (ptid == 0 && RG) pump(-1);
        __syncwarp();
      }
    }
    return;
  }
  // =================================================== consumers
  const unsigned cthreads = 32 * NC;
  for (int k = 0; k < deg; ++k) {
    EpiCheb epi;
    epi.first = (k == 0); epi.last = (k == deg - 1);
    epi.r_in = epi.first ? W.b : W.r;
    epi.d_in = k == 0 ? W.b : ((k & 1) ? W.d0 : W.d1);
    epi.r_out = W.r;
    epi.d_out = (k & 1) ? W.d1 : W.d0;
    epi.x = W.xout;
    epi.inv_theta = W.inv_theta; epi.ak = W.ak[k]; epi.bk = W.bk[k];
    const unsigned otag = W.tag0 + (unsigned)k + 1u;             // tag of this step's directions
    const bool push = multi && !epi.last;
    for (int it = 0; it < n; ++it) {
      const long long gx = (long long)k * n + it;
      const int xb = (int)(gx % X);
      int s; uint32_t par;
      if (it < P) { s = it; par = 0; }
      else {
        const long long q = (long long)k * NR + (it - P);
        s = P + (int)(q % RG); par = (uint32_t)((q / RG) & 1);
      }
      nm_mbar_wait_bounded(full_blob + s, par);
      const NmSlabView v = nm_slab_view(stage0 + (size_t)s * A.stage_bytes);
      const double* xs = xs0 + (size_t)xb * A.xs_doubles;
      const bool walk = warp < v.h.nslice;
      const unsigned lw = walk ? (unsigned)v.slane[tid] : 0u;
      const bool own = (lw & 0x8000u) != 0u;
      const int prow = v.h.first + (int)(lw & 0x3ffu);
      const int row0 = R * prow;
      typename EpiCheb::In in[R];
      if (own) {
#pragma unroll
        for (int c = 0; c < R; ++c) in[c] = epi.load(row0 + c);
      }
      double acc[R];
#pragma unroll
      for (int c = 0; c < R; ++c) acc[c] = 0.0;
      uint2 t = make_uint2(0u, 0u);
      if (walk) t = v.tbl[warp];
      nm_mbar_wait_bounded(full_xs + xb, (uint32_t)((gx / X) & 1));
      if (walk) {
        const double* pv = v.sv + t.x + lane;
        const unsigned short* pi = v.sidx + t.x + lane;
        const int w = (int)t.y;
#pragma unroll 4
        for (int kk = 0; kk < w; ++kk) {
          const double m = pv[32 * kk];
          const double* xp = xs + R * (int)pi[32 * kk];
#pragma unroll
          for (int c = 0; c < R; ++c) acc[c] += m * xp[c];
        }
      }
      const int gmax = v.h.gmax;
      __syncwarp();
      if (lane == 0) {
        if (it >= P) nm_mbar_arrive(empty_blob + s);             // pinned stages are never refilled
        nm_mbar_arrive(empty_xs + xb);
      }
      for (int q = 0; (1 << q) < gmax; ++q) {
#pragma unroll
        for (int c = 0; c < R; ++c) {
          const double other = __shfl_down_sync(0xffffffffu, acc[c], 1 << q);
          if (lw & (1u << (10 + q))) acc[c] += other;
        }
      }
      if (own) {
        double dn[R];
#pragma unroll
        for (int c = 0; c < R; ++c) dn[c] = epi.apply_dn(row0 + c, acc[c], in[c]);
        if (push && v.h.has_ghost) {                              // symmetric pattern: sent rows gather ghost columns
          for (int e = W.push_off[prow]; e < W.push_off[prow + 1]; ++e) {
            const NmPushEnt pe = W.push_ent[e];
            double val = dn[0];
#pragma unroll
            for (int c = 1; c < R; ++c) if (pe.comp == c) val = dn[c];
            nm_ll_store(W.ll_out[otag % 3u][pe.peer] + 2 * (size_t)pe.dst, val, otag);
          }
        }
      }
      if (flow) {
        __syncwarp();                                            // the warp's stores of this visit precede lane 0's arrival
        if (lane == 0) nm_mbar_arrive(done_xs + xb);
      }
    }
    if (k < deg - 1 && !flow) {
      // this CTA's directions of step k are stored: count it in at the grid barrier (the producers wait)
      asm volatile("bar.sync 2, %0;" ::"r"(cthreads) : "memory");
      if (tid == 0) {
        __threadfence();
        atomicAdd(W.gbar, 1ull);
      }
    }
  }
}

template <int R, int NC>
static inline void nm_slabpers_launch_t(NmSlab& S, const NmSlabPersArgs& W) {
  NmCtx& c = nm_ctx();
  static bool attr_set = false;                                  // per template instantiation
  if (!attr_set) {
    NM_CUDA(cudaFuncSetAttribute(k_slabpers<R, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(S.grid); cfg.blockDim = dim3(32 * (NC + S.nprod + 1)); cfg.dynamicSmemBytes = S.pers_smem; cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;                   // every CTA resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  NM_CUDA(cudaLaunchKernelEx(&cfg, k_slabpers<R, NC>, W));
  c.launches++;
}
static inline void nm_slabpers_dispatch(NmParcsr& M, NmSlab& S, const NmSlabPersArgs& W) {
  const bool blk = M.format == NM_FMT_KRON3;
  if (S.threads == 512) { if (blk) nm_slabpers_launch_t<3, 16>(S, W); else nm_slabpers_launch_t<1, 16>(S, W); }
  else if (S.threads == 256) { if (blk) nm_slabpers_launch_t<3, 8>(S, W); else nm_slabpers_launch_t<1, 8>(S, W); }
  else if (S.threads == 64) { if (blk) nm_slabpers_launch_t<3, 2>(S, W); else nm_slabpers_launch_t<1, 2>(S, W); }
  else { if (blk) nm_slabpers_launch_t<3, 4>(S, W); else nm_slabpers_launch_t<1, 4>(S, W); }
}

template <int R, int NC, bool FUSED, class Epi>
static inline void nm_slabws_launch_t(NmParcsr& M, NmSlab& S, const double* x, const Epi& epi, const NmHaloWait& hw,
                                      const NmSlabFusedArgs& F) {
  NmCtx& c = nm_ctx();
  NmSlabWsArgs W;
  NmSlabArgs& A = W.a;
  A.blob = S.blob.p; A.desc = S.desc.p; A.cta_first = S.cta_first.p;
  A.x = x; A.xg = M.halo.xg_cur ? M.halo.xg_cur : x; A.ncol = M.ncol;
  A.stage_bytes = S.stage_bytes; A.xs_doubles = S.xs_doubles; A.nstage = S.nstage;
  A.trace = nullptr;
  W.nxs = S.nxs; W.np = S.nprod;
  W.hflags = hw.flags; W.hmask = hw.mask; W.hepoch = hw.epoch; W.hstatus = hw.status;
  static const int ghost_cg = nm_env_int("NM_HALO_GHOST_CG", 0);
  W.ghost_cg = ghost_cg;
  static bool attr_set = false;                                  // per template instantiation
  if (!attr_set) {
    NM_CUDA(cudaFuncSetAttribute(k_slabws<R, NC, FUSED, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(S.grid); cfg.blockDim = dim3(32 * (NC + S.nprod)); cfg.dynamicSmemBytes = S.smem_bytes; cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = S.pdl ? 1 : 0;
  NM_CUDA(cudaLaunchKernelEx(&cfg, k_slabws<R, NC, FUSED, Epi>, W, epi, F));
  c.launches++;
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

template <bool FUSED, class Epi>
static inline void nm_slabws_dispatch(NmParcsr& M, NmSlab& S, const double* x, const Epi& epi, const NmHaloWait& hw,
                                      const NmSlabFusedArgs& F) {
  const bool blk = M.format == NM_FMT_KRON3;
  if (S.threads == 512) {
    if (blk) nm_slabws_launch_t<3, 16, FUSED, Epi>(M, S, x, epi, hw, F); else nm_slabws_launch_t<1, 16, FUSED, Epi>(M, S, x, epi, hw, F);
  } else if (S.threads == 256) {
    if (blk) nm_slabws_launch_t<3, 8, FUSED, Epi>(M, S, x, epi, hw, F); else nm_slabws_launch_t<1, 8, FUSED, Epi>(M, S, x, epi, hw, F);
  } else if (S.threads == 64) {
    if (blk) nm_slabws_launch_t<3, 2, FUSED, Epi>(M, S, x, epi, hw, F); else nm_slabws_launch_t<1, 2, FUSED, Epi>(M, S, x, epi, hw, F);
  } else {
    if (blk) nm_slabws_launch_t<3, 4, FUSED, Epi>(M, S, x, epi, hw, F); else nm_slabws_launch_t<1, 4, FUSED, Epi>(M, S, x, epi, hw, F);
  }
}

// Product through the slabs: x and the epilogue vectors are in pack order (S.order).
template <class Epi>
static inline void nm_spmv_slab_epi(NmParcsr& M, NmSlab& S, const double* x, const Epi& epi, const int* send_idx) {
  const bool blk = M.format == NM_FMT_KRON3;
  if (S.ws) {
    // several GPUs: push without waiting -- the kernel's producers poll the arrival flags before the chunks that have
    // ghost columns, so the exchange overlaps the launch, the matrix prefetch and the interior chunks
    NmHaloWait hw;
    static const bool overlap = nm_env_int("NM_HALO_OVERLAP", 1) != 0;
    if (!overlap || !nm_halo_push_nowait(M, x, send_idx, &hw)) {
      nm_halo_exchange(M, x, send_idx);
      hw.flags = nullptr; hw.mask = 0; hw.epoch = 0; hw.status = nullptr;
    }
    NmSlabFusedArgs F;
    memset(&F, 0, sizeof(F));
    nm_slabws_dispatch<false>(M, S, x, epi, hw, F);
    return;
  }
  nm_halo_exchange(M, x, send_idx);
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
