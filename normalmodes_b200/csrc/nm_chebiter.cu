// Fixed-degree Chebyshev iteration x = q_deg(M) b ~ M^-1 b: replaces pEVSL's
// EXTERNAL/ITERSOL/chebiter.c as used through pevsl_setup_chebiter_f90 / pevsl_chebiter_f90
// (src/mod_matvec.f90:88-93,167-174,480,512) and as the B-solve registered by
// pevsl_setbsol_chebiter_f90 (src/mod_pevsl.f90:72-73).  Saad, Iterative Methods, Alg. 12.1 with a
// zero initial guess; no inner products.  Each step is ONE kernel: the SpMV r -= M d with the
// x += d and d = a d + b r updates in its epilogue (EpiCheb), so a step streams M once and touches
// each vector once.
#include "nm_slab.cuh"
#include <algorithm>

// The iteration keeps its vectors in the PACK ORDER of a second, permuted copy of M (the slabs of nm_slab.cu): every
// step's epilogue then reads and writes contiguous ranges and the x values a chunk gathers sit in a few
// contiguous runs.  b is permuted in once and x out once per solve (2 of the deg+2 vector passes).
__global__ void k_perm_gather(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ order,
                              int n, int R) {                    // dst[R*i+c] = src[R*order[i]+c]
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * R) return;
  const int i = t / R;
  dst[t] = src[R * order[i] + (t - R * i)];
}
__global__ void k_perm_scatter(double* __restrict__ dst, const double* __restrict__ src, const int* __restrict__ order,
                               int n, int R) {                   // dst[R*order[i]+c] = src[R*i+c]
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * R) return;
  const int i = t / R;
  dst[R * order[i] + (t - R * i)] = src[t];
}

static void cheb_build_ppack(NmChebIter& C) {
  NmParcsr& M = *C.M;
  const char* off = getenv("NM_CHEB_PERMUTED");
  if (M.format == NM_FMT_ROW3 || M.nrow == 0 || (off && off[0] == '0')) return;
  const bool blk = M.format == NM_FMT_KRON3;
  const int n = blk ? M.nbrow : M.nrow, R = blk ? 3 : 1;
  std::vector<int> rp(n + 1), idx((size_t)(blk ? M.bja.n : M.ja.n));
  (blk ? M.bia : M.ia).download(rp.data(), rp.size());
  idx.resize(rp[n]);
  (blk ? M.bja : M.ja).download(idx.data(), idx.size());
  // NM_CHEB_KERNEL=plain: the subwarp-per-row kernels on the caller's numbering (also the fallback of matrices the
  // slab packer refuses: a row longer than a warp can share, too many distinct columns for a stage)
  const char* kk = getenv("NM_CHEB_KERNEL");
  if (kk && strcmp(kk, "plain") == 0) return;
  nm_slab_build_into(M, C.pslab, rp, idx, n);
  if (C.pslab.nchunk == 0) return;
  const int* order_dev = C.pslab.order.p;
  C.ppack_version = M.values_version;
  C.bp.alloc(std::max(M.nrow, 1)); C.xp.alloc(std::max(M.nrow, 1));
  if (M.halo.nsend > 0) {
    std::vector<int> order(n), newid(n), sidx(M.halo.nsend);
    NM_CUDA(cudaMemcpy(order.data(), order_dev, n * sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) newid[order[i]] = i;
    M.halo.send_idx.download(sidx.data(), sidx.size());
    // fused multi-GPU step (opt-in): every sent row must sit in a chunk that polls the arrival flags before it is
    // walked, i.e. the row itself must reference a ghost column (true for the symmetric B~ / Ap~)
    // in-kernel halo (boundary rows stored to the peers from the step's epilogue): the per-step fused kernel
    // (NM_HALO_FUSED=1) and the persistent kernel.  Both need what a symmetric pattern gives: a row that gathers a
    // ghost column owned by rank r is itself sent to r (its chunk then produces what r waits for before r can run
    // ahead and overwrite the slots the chunk still reads), and every sent row gathers a ghost column.
    bool sym = C.pslab.nchunk > 0 && C.pslab.ws && M.halo.p2p;
    if (sym) {
      const int P = nm_ctx().nranks;
      std::vector<unsigned char> sent((size_t)n, 0), needs((size_t)n, 0);       // bit r: sent to / gathers from rank r
      for (int r = 0; r < P; ++r)
        for (int i = M.halo.send_off[r]; i < M.halo.send_off[r + 1]; ++i) sent[sidx[i] / R] |= (unsigned char)(1u << r);
      const int ncolb = M.ncol / R;
      for (int row = 0; row < n; ++row)
        for (int p = rp[row]; p < rp[row + 1]; ++p)
          if (idx[p] >= ncolb) {
            const int g = (idx[p] - ncolb) * R;                                  // position in the ghost tail
            int r = 0;
            while (r + 1 < P && g >= M.halo.recv_off[r + 1]) ++r;
            needs[row] |= (unsigned char)(1u << r);
          }
      for (int row = 0; row < n && sym; ++row) sym = sent[row] == needs[row];
    }
    C.pers_multi_ok = sym;
    for (int& v : sidx) v = R * newid[v / R] + v % R;
    C.send_idx_p.from_host(sidx);
    C.h_sidx_p = sidx;                                           // the push tables follow once every rank agrees
  }
}

// Slot numbering + push tables of the in-kernel exchange (collective: every rank, after the common decision).
// WHERE a boundary value lands in the receiver's slot buffer is free as long as both sides agree, and it decides how
// many NVLink write transactions a step costs: with the receiver's natural order (sorted by global id) the 16-byte
// stores of a warp's epilogue scatter -- one transaction per value, and the SM's queue of outstanding peer writes
// (not the link) bounds the step: 200k-tet workload on 2 GPUs, 30 k values per step, B~ step 32 us whatever the
// matrix size (profiles/r2e_*).  Here the block of sender s in the receiver's buffer is laid out in the SENDER's pack
// order, component-major: the k-th boundary node sent to a peer (pack order) owns slots {k, K + k, 2K + k}, so the
// lanes of a warp that send component c of consecutive boundary rows write CONSECUTIVE slots (coalesced into full
// lines).  The receiver learns the layout once (its ghost column -> slot table, gathered through with one indirection).
static void cheb_build_push_tables(NmChebIter& C) {
  NmParcsr& M = *C.M;
  NmHalo& h = M.halo;
  NmCtx& c = nm_ctx();
  const int P = c.nranks, R = M.format == NM_FMT_KRON3 ? 3 : 1;
  const int n = M.format == NM_FMT_KRON3 ? M.nbrow : M.nrow;
  const std::vector<int>& sidx = C.h_sidx_p;
  std::vector<int> newslot(std::max(h.nsend, 1), 0);
  // NM_LL_PACKORDER=1: the layout described above.  Default 0 = the receiver's natural order (no indirection in the
  // gather): measured on the 200k-tet workload on 2 GPUs the coalesced stores bought nothing and the extra dependent
  // load in the boundary chunks' gather cost 6 us per B~ step (profiles/r2h_*) -- the stores were not the limiter.
  const bool packorder = nm_env_int("NM_LL_PACKORDER", 0) != 0;
  for (int r = 0; r < P; ++r) {
    const int o0 = h.send_off[r], o1 = h.send_off[r + 1];
    if (o1 == o0) continue;
    if (!packorder) { for (int i = o0; i < o1; ++i) newslot[i] = i - o0; continue; }
    std::vector<int> rows;
    for (int i = o0; i < o1; ++i) rows.push_back(sidx[i] / R);
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    const int K = (int)rows.size();
    NM_REQUIRE(o1 - o0 == R * K, "in-kernel halo: a boundary node is sent with %d of its %d components", (o1 - o0), R * K);
    for (int i = o0; i < o1; ++i) {
      const int k = (int)(std::lower_bound(rows.begin(), rows.end(), sidx[i] / R) - rows.begin());
      newslot[i] = (sidx[i] % R) * K + k;
    }
  }
  // tell every receiver where its ghosts from this rank sit inside this rank's block
  DBuf<int> d_send(std::max(h.nsend, 1)), d_recv(std::max(h.nghost, 1));
  if (h.nsend) d_send.upload(newslot.data(), h.nsend);
  NM_NCCL(ncclGroupStart());
  for (int r = 0; r < P; ++r) {
    if (r == c.rank) continue;
    if (h.send_cnt[r] > 0) NM_NCCL(ncclSend(d_send.p + h.send_off[r], h.send_cnt[r], ncclInt, r, c.nccl, c.stream));
    if (h.recv_cnt[r] > 0) NM_NCCL(ncclRecv(d_recv.p + h.recv_off[r], h.recv_cnt[r], ncclInt, r, c.nccl, c.stream));
  }
  NM_NCCL(ncclGroupEnd());
  std::vector<int> gslot(std::max(h.nghost, 1), 0);
  if (h.nghost) d_recv.download(gslot.data(), h.nghost);
  else NM_CUDA(cudaStreamSynchronize(c.stream));
  for (int r = 0; r < P; ++r) {
    std::vector<char> seen(h.recv_cnt[r], 0);
    for (int g = h.recv_off[r]; g < h.recv_off[r + 1]; ++g) {
      NM_REQUIRE(gslot[g] >= 0 && gslot[g] < h.recv_cnt[r] && !seen[gslot[g]], "in-kernel halo: rank %d sent a bad slot layout", r);
      seen[gslot[g]] = 1;
      gslot[g] += h.recv_off[r];
    }
  }
  if (packorder) C.ghost_slot.from_host(gslot);
  if (h.nsend == 0) return;
  std::vector<int> sslot(h.nsend);
  std::vector<int> cnt(n + 1, 0);
  for (int v : sidx) cnt[v / R + 1]++;
  for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
  std::vector<NmPushEnt> ent(sidx.size());
  std::vector<int> fill(cnt.begin(), cnt.end() - 1);
  for (int r = 0; r < P; ++r)
    for (int i = h.send_off[r]; i < h.send_off[r + 1]; ++i) {
      const int v = sidx[i];
      sslot[i] = h.peer_base[r] + newslot[i];
      NmPushEnt e;
      e.dst = (unsigned)sslot[i];
      e.peer = (unsigned short)r; e.comp = (unsigned short)(v % R);
      ent[fill[v / R]++] = e;
    }
  C.send_slot.from_host(sslot);
  C.push_off.from_host(cnt);
  C.push_ent.alloc(std::max<size_t>(ent.size(), 1)); C.push_ent.upload(ent.data(), ent.size());
}

// In-kernel halo on several GPUs (collective decision, every rank takes the same branch) and the opt-in persistent
// kernel (k_slabpers, NM_SLAB_PERS=1): stages per CTA from the shared memory left beside the co-resident CTAs of an
// SM, Chebyshev coefficients and the grid-barrier counter on the device.
static void cheb_setup_multi_pers(NmChebIter& C) {
  NmParcsr& M = *C.M;
  NmCtx& c = nm_ctx();
  NmSlab& S = C.pslab;
  C.pers = false; C.fused = false;
  const bool slab_ok = S.nchunk > 0 && S.ws;
  bool ll_ok = true;
  if (c.nranks > 1) {
    double v = (slab_ok && (M.halo.nsend == 0 || C.pers_multi_ok) && (M.halo.nghost == 0) == (M.halo.nsend == 0)) ? 1.0 : 0.0;
    if (!c.p2p || nm_env_int("NM_HALO_FUSED", 1) == 0) v = 0.0;
    DBuf<double> d(1);
    d.upload(&v, 1);
    NM_NCCL(ncclAllReduce(d.p, d.p, 1, ncclDouble, ncclMin, c.nccl, c.stream));
    d.download(&v, 1);
    ll_ok = v > 0.5;
    if (ll_ok) ll_ok = nm_halo_ll_setup(M.halo);
    if (ll_ok) cheb_build_push_tables(C);
    C.fused = ll_ok;
  }
  // NM_SLAB_PERS: 0 (default) one launch per step chained by programmatic dependent launch; 1 the whole iteration in one
  // cooperative launch.  Measured on the bench workload (profiles/r2c_*): the software grid barrier (fence + atomic,
  // poll + fence) costs 2-4 us per step more than the hardware's grid completion + dependent launch, and the dataflow
  // variant (NM_SLAB_FLOW=1: per-chunk flags instead of the barrier) another 7-10 us (a serial poll + fence per chunk
  // visit in the producers) -- kept as tested options, not the default.
  const bool want = nm_env_int("NM_SLAB_PERS", 0) != 0 && slab_ok && S.max_chunks_per_cta <= NM_SLAB_MAXDESC;
  if (!want || !ll_ok) return;
  const int per_sm = std::max(1, nm_div_up(S.grid, c.sm_count));
  const int budget = std::min(227 * 1024, (228 * 1024) / per_sm - 1024);
  const int fixed = (int)((NM_SLABPERS_FIXED + 8 * (size_t)S.nxs * S.xs_doubles + 15) & ~(size_t)15);
  int nst = (budget - fixed) / std::max(S.stage_bytes, 16);
  nst = std::max(0, std::min(8, std::min(nst, nm_env_int("NM_SLAB_PERS_STAGES", 8))));
  if (nst < 2) return;
  S.pers_nstage = nst; S.pers_smem = fixed + nst * S.stage_bytes;
  C.ak_dev.from_host(C.ak); C.bk_dev.from_host(C.bk);
  C.gbar.alloc(2); C.gbar.zero();
  C.gbar_base = 0;
  C.pers = true;
}

// One step of the fused multi-GPU iteration: kernel k polls the flag-in-data slots of the ghost values it gathers
// (written by the peers' step k-1, or by nm_halo_push_ll of b before step 0) and stores its own new direction into the
// peers' slots of the next tag.
static void cheb_step_fused(NmChebIter& C, const double* din, const EpiCheb& e, int k, unsigned tag0) {
  NmParcsr& M = *C.M;
  NmHalo& h = M.halo;
  NmCtx& c = nm_ctx();
  NmSlabFusedArgs F;
  memset(&F, 0, sizeof(F));
  if (h.nsend > 0) {
    const unsigned tin = tag0 + (unsigned)k;
    F.ll_in = (const unsigned long long*)(c.win + h.win_ll[tin % 3u]);
    F.tag_in = tin; F.status = c.dev_status; F.gslot = C.ghost_slot.p;
    static const int dbg = nm_env_int("NM_DEBUG_LL", 0);
    F.debug = dbg;
    if (k < C.deg - 1) {
      F.tag_out = tin + 1u;
      F.push_off = C.push_off.p; F.push_ent = C.push_ent.p;
      for (int r = 0; r < c.nranks; ++r)
        if (r != c.rank && h.send_cnt[r] > 0) F.ll_out[r] = (unsigned long long*)(c.peer_win[r] + h.peer_ll[F.tag_out % 3u][r]);
    }
  }
  NmHaloWait none;
  none.flags = nullptr; none.mask = 0; none.epoch = 0; none.status = nullptr;
  nm_slabws_dispatch<true>(M, C.pslab, din, e, none, F);
}

NmChebIter* nm_chebiter_build(double lb, double ub, int deg, NmParcsr* M) {
  NM_REQUIRE(M && M->nrow == M->ncol, "setup_chebiter: square matrix required");
  NM_REQUIRE(deg >= 1, "setup_chebiter: degree %d < 1", deg);
  NM_REQUIRE(ub > lb && lb > 0.0, "setup_chebiter: need 0 < lmin < lmax (got %g, %g)", lb, ub);
  std::unique_ptr<NmChebIter> C(new NmChebIter());
  C->M = M; C->lb = lb; C->ub = ub; C->deg = deg;
  const double theta = (ub + lb) / 2.0, delta = (ub - lb) / 2.0;
  const double sigma1 = theta / delta;
  double rho = 1.0 / sigma1;
  C->theta = theta;
  for (int k = 0; k < deg; ++k) {
    const double rho1 = 1.0 / (2.0 * sigma1 - rho);
    C->ak.push_back(rho1 * rho);
    C->bk.push_back(2.0 * rho1 / delta);
    rho = rho1;
  }
  const size_t n = M->nrow > 0 ? M->nrow : 1;
  C->r.alloc(n); C->d0.alloc(n); C->d1.alloc(n);
  cheb_build_ppack(*C);
  cheb_setup_multi_pers(*C);
  return C.release();
}

void nm_chebiter_solve(NmChebIter& C, const double* b, double* x) {
  NM_REQUIRE(b != x, "chebiter: b and x must not alias");
  NmParcsr& M = *C.M;
  NmCtx& c = nm_ctx();
  double* dbuf[2] = {C.d0.p, C.d1.p};
  const bool perm = C.pslab.nchunk > 0;
  const int* order_dev = C.pslab.order.p;
  const int nblk = M.format == NM_FMT_KRON3 ? M.nbrow : M.nrow, R = M.format == NM_FMT_KRON3 ? 3 : 1;
  double* xout = x;
  if (perm) {
    if (C.ppack_version != M.values_version) {
      nm_slab_fill_from(M, C.pslab);
      C.ppack_version = M.values_version;
    }
    k_perm_gather<<<nm_div_up(M.nrow, 256), 256, 0, c.stream>>>(C.bp.p, b, order_dev, nblk, R);
    c.launches++;
    b = C.bp.p;
    x = C.xp.p;
  }
  if (C.pers) {
    // the whole iteration in ONE cooperative launch (k_slabpers); several GPUs: the boundary values of b go into the
    // peers' flag-in-data slots first (what step 0 gathers), everything else is exchanged inside the kernel
    NmSlabPersArgs W;
    memset(&W, 0, sizeof(W));
    NmSlab& S = C.pslab;
    W.a.blob = S.blob.p; W.a.desc = S.desc.p; W.a.cta_first = S.cta_first.p; W.a.ncol = M.ncol;
    W.a.stage_bytes = S.stage_bytes; W.a.xs_doubles = S.xs_doubles; W.a.nstage = S.pers_nstage;
    W.nxs = S.nxs; W.np = S.nprod; W.deg = C.deg;
    W.b = b; W.r = C.r.p; W.d0 = C.d0.p; W.d1 = C.d1.p; W.xout = x;
    W.ak = C.ak_dev.p; W.bk = C.bk_dev.p; W.inv_theta = 1.0 / C.theta;
    W.gbar = C.gbar.p; W.gbase = C.gbar_base;
    static const bool want_flow = nm_env_int("NM_SLAB_FLOW", 1) != 0;
    const bool flow = want_flow && S.deps_ok;
    if (flow) { W.cflag = S.cflag.p; W.desc_cid = S.desc_cid.p; W.ftag0 = C.ftag; }
    NmHalo& h = M.halo;
    if (c.nranks > 1 && h.nsend > 0) {
      const unsigned long long e0 = ++h.epoch;
      W.tag0 = (unsigned)e0;
      nm_halo_push_ll(M, b, C.send_idx_p.p, C.send_slot.p, W.tag0, (int)(W.tag0 % 3u));
      for (int q = 0; q < 3; ++q) {
        W.ll_in[q] = (const unsigned long long*)(c.win + h.win_ll[q]);
        for (int r = 0; r < c.nranks; ++r)
          if (r != c.rank && h.send_cnt[r] > 0) W.ll_out[q][r] = (unsigned long long*)(c.peer_win[r] + h.peer_ll[q][r]);
      }
      W.push_off = C.push_off.p; W.push_ent = C.push_ent.p; W.hstatus = c.dev_status; W.gslot = C.ghost_slot.p;
      h.epoch = e0 + (unsigned long long)(C.deg - 1);
    }
    bool launched = true;
    try {
      nm_slabpers_dispatch(M, S, W);
    } catch (const NmError&) {
      // not co-resident on this device (cooperative launch refused): per-step launches from now on.  Only legal on one
      // rank, where nothing has been exchanged yet; ranks must not diverge.
      cudaGetLastError();
      NM_REQUIRE(c.nranks == 1, "persistent ChebIter kernel could not be launched on rank %d", c.rank);
      C.pers = false;
      launched = false;
    }
    if (launched) {
      if (flow) C.ftag += (unsigned)C.deg;
      else C.gbar_base += (unsigned long long)(C.deg - 1) * (unsigned long long)S.grid;
      k_perm_scatter<<<nm_div_up(M.nrow, 256), 256, 0, c.stream>>>(xout, C.xp.p, order_dev, nblk, R);
      c.launches++;
      C.nsolve++;
      C.nmatvec += C.deg;
      return;
    }
  }
  const double* din = b;
  unsigned fused_tag0 = 0;
  if (C.fused && M.halo.nsend > 0) {
    // boundary values of b -> the peers' slots (what step 0 gathers); deg epochs for the solve
    const unsigned long long e0 = ++M.halo.epoch;
    fused_tag0 = (unsigned)e0;
    nm_halo_push_ll(M, b, C.send_idx_p.p, C.send_slot.p, fused_tag0, (int)(fused_tag0 % 3u));
    M.halo.epoch = e0 + (unsigned long long)(C.deg - 1);
  }
  for (int k = 0; k < C.deg; ++k) {
    EpiCheb e;
    e.first = (k == 0); e.last = (k == C.deg - 1);
    e.r_in = e.first ? b : C.r.p;
    e.d_in = din;
    e.r_out = C.r.p;
    e.d_out = dbuf[k & 1];
    e.x = x;
    e.inv_theta = 1.0 / C.theta; e.ak = C.ak[k]; e.bk = C.bk[k];
    if (C.fused) cheb_step_fused(C, din, e, k, fused_tag0);
    else if (C.pslab.nchunk > 0) nm_spmv_slab_epi(M, C.pslab, din, e, C.send_idx_p.p);
    else nm_spmv_epi(M, din, e);
    din = dbuf[k & 1];
  }
  if (perm) {
    k_perm_scatter<<<nm_div_up(M.nrow, 256), 256, 0, c.stream>>>(xout, C.xp.p, order_dev, nblk, R);
    c.launches++;
  }
  C.nsolve++;
  C.nmatvec += C.deg;
}

// ---------------------------------------------------------------- C ABI
extern "C" int nm_chebiter_create(double lmin, double lmax, int deg, void* mat, void** out) {
  NM_API_BEGIN
  *out = nm_chebiter_build(lmin, lmax, deg, (NmParcsr*)mat);
  NM_API_END
}
extern "C" int nm_chebiter_free(void* h) {
  NM_API_BEGIN
  if (h) { NM_CUDA(cudaStreamSynchronize(nm_ctx().stream)); delete (NmChebIter*)h; }
  NM_API_END
}
// x = q(M) b with HOST vectors (pevsl_chebiter_f90 contract).
extern "C" int nm_chebiter_solve_host(void* h, const double* b, double* x) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  NmCtx& c = nm_ctx();
  const int n = C.M->nrow;
  DBuf<double> db(n > 0 ? n : 1), dx(n > 0 ? n : 1);
  if (n) NM_CUDA(cudaMemcpyAsync(db.p, b, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  nm_chebiter_solve(C, db.p, dx.p);
  if (n) NM_CUDA(cudaMemcpyAsync(x, dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_API_END
}
extern "C" int nm_chebiter_solve_dev(void* h, const double* b_dev, double* x_dev) {
  NM_API_BEGIN
  nm_chebiter_solve(*(NmChebIter*)h, b_dev, x_dev);
  NM_API_END
}
// kind: 0 = plain kernels on the caller's numbering (1, 2: the retired round-1a kernels),
// 3 = TMA-staged warp-sliced ELL slabs (k_slab), 4 = the same, warp-specialised (k_slabws), 5 = the whole iteration in one
// persistent cooperative launch (k_slabpers, NM_SLAB_PERS=1); all on vectors kept in pack order; bytes = matrix bytes one iteration step streams
extern "C" int nm_chebiter_pack_info(void* h, int* kind, long long* bytes) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  const int k = C.pslab.nchunk > 0 ? (C.pers ? 5 : (C.pslab.ws ? 4 : 3)) : 0;
  if (kind) *kind = k;
  if (bytes) *bytes = k >= 3 ? C.pslab.bytes : C.M->fmt_bytes;
  NM_API_END
}
// diagnostic (NM_SLAB_TRACE=1): clock64 stamps of the last k_slab launch, [grid][64 chunks][8 stamps]; returns grid
extern "C" int nm_chebiter_trace_dump(void* h, long long* out, long long cap, int* grid, int* cta_first) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  NM_REQUIRE(C.pslab.trace.n > 0, "no trace (set NM_SLAB_TRACE=1 before creating the ChebIter)");
  NM_REQUIRE(cap >= (long long)C.pslab.trace.n, "trace buffer too small");
  C.pslab.trace.download(out, C.pslab.trace.n);
  if (grid) *grid = C.pslab.grid;
  if (cta_first) C.pslab.cta_first.download(cta_first, C.pslab.grid + 1);
  NM_API_END
}
extern "C" int nm_chebiter_stats(void* h, long long* nsolve, long long* nmatvec, int* deg, double* lmin, double* lmax) {
  NM_API_BEGIN
  NmChebIter& C = *(NmChebIter*)h;
  if (nsolve) *nsolve = C.nsolve;
  if (nmatvec) *nmatvec = C.nmatvec;
  if (deg) *deg = C.deg;
  if (lmin) *lmin = C.lb;
  if (lmax) *lmax = C.ub;
  NM_API_END
}
