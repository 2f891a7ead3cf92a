// Distributed CSR container: replaces pEVSL's parcsr object created by
// pevsl_parcsrcreate_f90 (src/mod_matvec.f90:69-71,117-119,146-148,196-198,218-220,242-244) from
// the reference's COOmat arrays (src/mod_cg_datatype.f90:35-49): global row/column offsets per rank,
// 0-based local row pointers, 0-based GLOBAL column ids, fp64 values.
#include "nm_spmv.cuh"
#include <algorithm>

// ---------------------------------------------------------------- halo (ghost-DOF) exchange
__global__ void k_halo_gather(double* __restrict__ buf, const double* __restrict__ x, const int* __restrict__ idx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = x[idx[i]];
}

// Peer-store exchange (NVLink, CUDA IPC window): every rank writes the x values its peers need straight into THEIR
// ghost buffers, then raises an arrival flag per destination; the last block of the same kernel waits for the
// flags of the ranks that write to us, so the next kernel on the stream may gather from the ghost buffer.  Two
// ghost buffers alternate (parity of the exchange counter): a peer can be at most one exchange ahead, because its
// next push follows its wait for OUR flag of the current one, which we raise only after our previous product.
struct NmPushArgs {
  int nranks, me, nsend;
  int send_off[9];
  double* peer_xg[8];                       // destination of this rank's block in each peer's ghost buffer (parity applied)
  unsigned long long* peer_flag[8];         // this rank's slot in each peer's flag array
  const unsigned long long* my_flag;        // flag array of this rank (slot s written by rank s)
  unsigned recv_mask, send_mask;
  unsigned long long epoch;
  unsigned* ctr;
  int* status;
  int wait;                                 // 1: the last block waits for the peers' flags; 0: the consumer kernel polls
};

__global__ void k_halo_push(NmPushArgs A, const double* __restrict__ x, const int* __restrict__ idx) {
  // programmatic dependent launch (no-ops for plain launches): let the consumer kernel start its matrix prefetch,
  // and make sure the producer of x has completed before x is read
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.nsend; i += gridDim.x * blockDim.x) {
    int r = 0;
    while (i >= A.send_off[r + 1]) ++r;
    A.peer_xg[r][i - A.send_off[r]] = x[idx[i]];
  }
  __shared__ int last;
  __syncthreads();                           // the block's stores happen-before thread 0's fence (cumulative)
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(A.ctr, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence_system();                    // every block's stores are ordered before the flags below
  const int r = threadIdx.x;
  if (r < A.nranks && r != A.me) {
    if (A.send_mask & (1u << r)) {
      asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(A.peer_flag[r]), "l"(A.epoch) : "memory");
    }
    if (A.wait && (A.recv_mask & (1u << r))) {
      const unsigned long long* f = A.my_flag + r;
      unsigned long long v;
      const long long t0 = clock64();
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= A.epoch) break;
        if (*(volatile int*)A.status != 0) break;                                 // an earlier wait already failed
        if (clock64() - t0 > 20000000000ll) { atomicOr(A.status, 1); break; }     // ~10 s: peer stalled or died
      }
    }
  }
  if (threadIdx.x == 0) *A.ctr = 0;
}

static void halo_push_launch(NmParcsr& M, const double* x, const int* send_idx, int wait, NmHaloWait* w) {
  NmCtx& c = nm_ctx();
  NmHalo& h = M.halo;
  const unsigned long long epoch = ++h.epoch;
  const int par = (int)(epoch & 1);
  NmPushArgs A;
  A.nranks = c.nranks; A.me = c.rank; A.nsend = h.nsend;
  A.recv_mask = A.send_mask = 0;
  for (int r = 0; r < 8; ++r) { A.peer_xg[r] = nullptr; A.peer_flag[r] = nullptr; }
  for (int r = 0; r <= 8; ++r) A.send_off[r] = r <= c.nranks ? h.send_off[std::min(r, c.nranks)] : h.nsend;
  // Flags travel in BOTH directions of every link, also where values travel in one only (rectangular E / ET: a rank
  // may send to a peer it receives nothing from): the receiver's flag is then a zero-length acknowledgement that its
  // previous product is done, which is what bounds a sender to one exchange ahead of the slowest reader of the
  // parity buffer it is about to overwrite.
  for (int r = 0; r < c.nranks; ++r) {
    if (r == c.rank) continue;
    if (h.send_cnt[r] > 0) A.peer_xg[r] = (double*)(c.peer_win[r] + h.peer_xg[par][r]) + h.peer_base[r];
    if (h.send_cnt[r] > 0 || h.recv_cnt[r] > 0) {
      A.send_mask |= 1u << r;
      A.recv_mask |= 1u << r;
      A.peer_flag[r] = (unsigned long long*)(c.peer_win[r] + h.peer_flag[r]) + c.rank;
    }
  }
  A.my_flag = (const unsigned long long*)(c.win + h.win_flag);
  A.epoch = epoch; A.ctr = c.push_ctr; A.status = c.dev_status; A.wait = wait;
  const int blocks = std::max(1, std::min(16, nm_div_up(h.nsend, 2048)));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = c.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = wait ? 0 : 1;               // chained into the ChebIter steps' dependent launches
  NM_CUDA(cudaLaunchKernelEx(&cfg, k_halo_push, A, x, send_idx ? send_idx : (const int*)h.send_idx.p));
  c.launches++;
  h.xg_cur = (double*)(c.win + h.win_xg[par]);
  if (w) { w->flags = A.my_flag; w->mask = A.recv_mask; w->epoch = epoch; w->status = c.dev_status; }
}

bool nm_halo_push_nowait(NmParcsr& M, const double* x, const int* send_idx, NmHaloWait* w) {
  NmCtx& c = nm_ctx();
  NmHalo& h = M.halo;
  w->flags = nullptr; w->mask = 0; w->epoch = 0; w->status = nullptr;
  if (c.nranks == 1 || (h.nghost == 0 && h.nsend == 0)) return true;       // nothing to exchange, nothing to wait for
  if (!h.p2p) return false;
  halo_push_launch(M, x, send_idx, 0, w);
  return true;
}

void nm_halo_exchange(NmParcsr& M, const double* x, const int* send_idx) {
  NmCtx& c = nm_ctx();
  NmHalo& h = M.halo;
  if (c.nranks == 1 || (h.nghost == 0 && h.nsend == 0)) return;
  if (h.p2p) {
    halo_push_launch(M, x, send_idx, 1, nullptr);
    return;
  }
  if (h.nsend > 0) {
    k_halo_gather<<<nm_div_up(h.nsend, 256), 256, 0, c.stream>>>(h.sendbuf.p, x, send_idx ? send_idx : h.send_idx.p, h.nsend);
    c.launches++;
  }
  NM_NCCL(ncclGroupStart());
  for (int r = 0; r < c.nranks; ++r) {
    if (r == c.rank) continue;
    if (h.send_cnt[r] > 0)
      NM_NCCL(ncclSend(h.sendbuf.p + h.send_off[r], h.send_cnt[r], ncclDouble, r, c.nccl, c.stream));
    if (h.recv_cnt[r] > 0)
      NM_NCCL(ncclRecv(h.xg.p + h.recv_off[r], h.recv_cnt[r], ncclDouble, r, c.nccl, c.stream));
  }
  NM_NCCL(ncclGroupEnd());
  h.xg_cur = h.xg.p;
}

// ---- flag-in-data exchange of the persistent ChebIter kernel: slots + the push of the right-hand side (step 0's gather)
struct NmPushLLArgs {
  int nsend;
  int send_off[9];
  unsigned long long* peer_ll[8];           // this rank's block of slots in each peer's buffer
  unsigned tag;
};
__global__ void k_halo_push_ll(NmPushLLArgs A, const double* __restrict__ x, const int* __restrict__ idx,
                               const int* __restrict__ slot) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.nsend; i += gridDim.x * blockDim.x) {
    int r = 0;
    while (i >= A.send_off[r + 1]) ++r;
    nm_ll_store(A.peer_ll[r] + 2 * (size_t)slot[i], x[idx[i]], A.tag);
  }
}
void nm_halo_push_ll(NmParcsr& M, const double* x, const int* send_idx, const int* send_slot, unsigned tag, int buf) {
  NmCtx& c = nm_ctx();
  NmHalo& h = M.halo;
  if (h.nsend == 0) return;
  NM_REQUIRE(h.ll, "nm_halo_push_ll without LL slots");
  NmPushLLArgs A;
  A.nsend = h.nsend; A.tag = tag;
  for (int r = 0; r < 8; ++r) A.peer_ll[r] = nullptr;
  for (int r = 0; r <= 8; ++r) A.send_off[r] = h.send_off[std::min(r, c.nranks)];
  for (int r = 0; r < c.nranks; ++r)
    if (r != c.rank && h.send_cnt[r] > 0)
      A.peer_ll[r] = (unsigned long long*)(c.peer_win[r] + h.peer_ll[buf][r]);     // send_slot includes the block's base
  const int blocks = std::max(1, std::min(32, nm_div_up(h.nsend, 1024)));
  k_halo_push_ll<<<blocks, 256, 0, c.stream>>>(A, x, send_idx ? send_idx : (const int*)h.send_idx.p, send_slot);
  c.launches++;
}
bool nm_halo_ll_setup(NmHalo& h) {
  NmCtx& c = nm_ctx();
  if (h.ll) return true;
  if (!c.p2p || !h.p2p) return false;
  const int P = c.nranks;
  NM_REQUIRE(P <= 8, "peer-window halo: at most 8 ranks (one NVSwitch box)");
  long long mine[3];
  for (int b = 0; b < 3; ++b) { h.win_ll[b] = nm_win_alloc(16 * (size_t)std::max(h.nghost, 1)); mine[b] = (long long)h.win_ll[b]; }
  DBuf<long long> d_mine(3), d_all((size_t)3 * P);
  d_mine.upload(mine, 3);
  NM_NCCL(ncclAllGather(d_mine.p, d_all.p, 3, ncclInt64, c.nccl, c.stream));
  std::vector<long long> all((size_t)3 * P);
  d_all.download(all.data(), all.size());
  for (int b = 0; b < 3; ++b) {
    h.peer_ll[b].assign(P, 0);
    for (int r = 0; r < P; ++r) h.peer_ll[b][r] = (size_t)all[3 * r + b];
  }
  h.ll = true;
  return true;
}

// Window slots of one matrix' halo (collective: every rank calls this for the same matrices in the same order).
// cnt: P x P matrix, cnt[r*P + s] = ghosts rank r receives from rank s.
static void halo_p2p_setup(NmHalo& h, const std::vector<int>& cnt) {
  NmCtx& c = nm_ctx();
  h.p2p = false;
  if (!c.p2p) return;
  const int P = c.nranks;
  const size_t gb = (size_t)std::max(h.nghost, 1) * sizeof(double);
  long long mine[3];
  h.win_xg[0] = nm_win_alloc(gb); h.win_xg[1] = nm_win_alloc(gb);
  h.win_flag = nm_win_alloc(sizeof(unsigned long long) * 8);
  mine[0] = (long long)h.win_xg[0]; mine[1] = (long long)h.win_xg[1]; mine[2] = (long long)h.win_flag;
  DBuf<long long> d_mine(3), d_all((size_t)3 * P);
  d_mine.upload(mine, 3);
  NM_NCCL(ncclAllGather(d_mine.p, d_all.p, 3, ncclInt64, c.nccl, c.stream));
  std::vector<long long> all((size_t)3 * P);
  d_all.download(all.data(), all.size());
  h.peer_xg[0].assign(P, 0); h.peer_xg[1].assign(P, 0); h.peer_flag.assign(P, 0); h.peer_base.assign(P, 0);
  for (int r = 0; r < P; ++r) {
    h.peer_xg[0][r] = (size_t)all[3 * r]; h.peer_xg[1][r] = (size_t)all[3 * r + 1]; h.peer_flag[r] = (size_t)all[3 * r + 2];
    int base = 0;                                   // ghosts of rank r are grouped by owner: blocks of ranks < me first
    for (int s2 = 0; s2 < c.rank; ++s2) base += cnt[(size_t)r * P + s2];
    h.peer_base[r] = base;
  }
  h.epoch = 0;
  h.p2p = true;
  h.xg_cur = (double*)(c.win + h.win_xg[0]);
}

// Host part of the plan (no device, no communicator): the sorted unique global ids of the columns a rank's block
// references outside its own range [lo, hi) -- sorted by id = grouped by owner rank.
static void halo_ghosts_host(int ncol_glob, int lo, int hi, long long nnz, const int* ja, std::vector<int>& gl) {
  gl.clear();
  for (long long p = 0; p < nnz; ++p) {
    const int g = ja[p];
    NM_REQUIRE(g >= 0 && g < ncol_glob, "parcsrcreate: column id %d out of range (0-based global ids expected)", g);
    if (g < lo || g >= hi) gl.push_back(g);
  }
  std::sort(gl.begin(), gl.end());
  gl.erase(std::unique(gl.begin(), gl.end()), gl.end());
}
static int halo_owner(const int* col_starts, int P, int g) {
  return (int)(std::upper_bound(col_starts, col_starts + P + 1, g) - col_starts) - 1;
}

// Host-only view of the receive side of a rank's halo plan, for hosts and tests without a GPU: ghost ids (sorted)
// and how many come from each owner.  ghost_glob may be null (count only) or hold up to nnz ids.
extern "C" int nm_halo_plan_host(int nranks, int rank, const int* col_starts, long long nnz, const int* ja, int* nghost,
                                 int* ghost_glob, int* recv_cnt) {
  NM_API_BEGIN
  NM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "halo_plan_host: bad rank %d of %d", rank, nranks);
  std::vector<int> gl;
  halo_ghosts_host(col_starts[nranks], col_starts[rank], col_starts[rank + 1], nnz, ja, gl);
  for (int r = 0; r < nranks; ++r) recv_cnt[r] = 0;
  for (int g : gl) {
    const int owner = halo_owner(col_starts, nranks, g);
    NM_REQUIRE(owner >= 0 && owner < nranks && owner != rank, "ghost column %d has no remote owner", g);
    recv_cnt[owner]++;
  }
  *nghost = (int)gl.size();
  if (ghost_glob) std::copy(gl.begin(), gl.end(), ghost_glob);
  NM_API_END
}

// Build the send side of the plan: every rank tells every owner which of its columns it needs.
static void build_halo_plan(NmParcsr& M, const int* col_starts) {
  NmCtx& c = nm_ctx();
  NmHalo& h = M.halo;
  const int P = c.nranks;
  h.recv_cnt.assign(P, 0); h.recv_off.assign(P + 1, 0);
  h.send_cnt.assign(P, 0); h.send_off.assign(P + 1, 0);
  for (int g : h.ghost_glob) {
    const int owner = halo_owner(col_starts, P, g);
    NM_REQUIRE(owner >= 0 && owner < P && owner != c.rank, "ghost column %d has no remote owner", g);
    h.recv_cnt[owner]++;
  }
  for (int r = 0; r < P; ++r) h.recv_off[r + 1] = h.recv_off[r] + h.recv_cnt[r];
  if (P == 1) return;
  // counts: allgather of the P x P matrix
  DBuf<int> d_cnt((size_t)P * P);
  DBuf<int> d_mine(P);
  d_mine.upload(h.recv_cnt.data(), P);
  NM_NCCL(ncclAllGather(d_mine.p, d_cnt.p, P, ncclInt, c.nccl, c.stream));
  std::vector<int> cnt((size_t)P * P);
  d_cnt.download(cnt.data(), cnt.size());
  h.cnt_all = cnt;
  for (int r = 0; r < P; ++r) h.send_cnt[r] = cnt[(size_t)r * P + c.rank];      // what rank r needs from me
  for (int r = 0; r < P; ++r) h.send_off[r + 1] = h.send_off[r] + h.send_cnt[r];
  h.nsend = h.send_off[P];
  DBuf<int> d_need(std::max<size_t>(h.ghost_glob.size(), 1));
  if (!h.ghost_glob.empty()) d_need.upload(h.ghost_glob.data(), h.ghost_glob.size());
  DBuf<int> d_send(std::max(h.nsend, 1));
  NM_NCCL(ncclGroupStart());
  for (int r = 0; r < P; ++r) {
    if (r == c.rank) continue;
    if (h.recv_cnt[r] > 0) NM_NCCL(ncclSend(d_need.p + h.recv_off[r], h.recv_cnt[r], ncclInt, r, c.nccl, c.stream));
    if (h.send_cnt[r] > 0) NM_NCCL(ncclRecv(d_send.p + h.send_off[r], h.send_cnt[r], ncclInt, r, c.nccl, c.stream));
  }
  NM_NCCL(ncclGroupEnd());
  std::vector<int> sidx(std::max(h.nsend, 1));
  d_send.download(sidx.data(), h.nsend);
  for (int i = 0; i < h.nsend; ++i) {
    sidx[i] -= M.col0;
    NM_REQUIRE(sidx[i] >= 0 && sidx[i] < M.ncol, "halo plan: peer asked for a column this rank does not own");
  }
  if (h.nsend) { h.send_idx.alloc(h.nsend); h.send_idx.upload(sidx.data(), h.nsend); h.sendbuf.alloc(h.nsend); }
  halo_p2p_setup(h, cnt);
}

// ---------------------------------------------------------------- format detection (host)
// ROW3: rows 3b..3b+2 have identical column lists consisting of aligned triples (c, c+1, c+2), c%3==0.
static bool detect_row3(int nrow, const std::vector<int>& ia, const std::vector<int>& ja, std::vector<int>& bia,
                        std::vector<int>& bja) {
  if (nrow == 0 || nrow % 3) return false;
  const int nb = nrow / 3;
  bia.assign(nb + 1, 0);
  bja.clear();
  bja.reserve(ja.size() / 9 + 1);
  for (int b = 0; b < nb; ++b) {
    const int s0 = ia[3 * b], s1 = ia[3 * b + 1], s2 = ia[3 * b + 2], s3 = ia[3 * b + 3];
    const int len = s1 - s0;
    if (s2 - s1 != len || s3 - s2 != len || len % 3) return false;
    if (s0 != 9 * bia[b]) return false;
    for (int u = 0; u < len; u += 3) {
      const int c = ja[s0 + u];
      if (c % 3 || ja[s0 + u + 1] != c + 1 || ja[s0 + u + 2] != c + 2) return false;
      bja.push_back(c / 3);
    }
    for (int u = 0; u < len; ++u)
      if (ja[s1 + u] != ja[s0 + u] || ja[s2 + u] != ja[s0 + u]) return false;
    bia[b + 1] = bia[b] + len / 3;
  }
  return true;
}
// KRON3: row 3b+p holds the same values at columns c_j + p, c_j % 3 == 0.
static bool detect_kron3(int nrow, const std::vector<int>& ia, const std::vector<int>& ja, const double* a,
                         std::vector<int>& bia, std::vector<int>& bja, std::vector<double>& mval) {
  if (nrow == 0 || nrow % 3) return false;
  const int nb = nrow / 3;
  bia.assign(nb + 1, 0);
  bja.clear(); mval.clear();
  bja.reserve(ja.size() / 3 + 1); mval.reserve(ja.size() / 3 + 1);
  for (int b = 0; b < nb; ++b) {
    const int s0 = ia[3 * b], s1 = ia[3 * b + 1], s2 = ia[3 * b + 2], s3 = ia[3 * b + 3];
    const int len = s1 - s0;
    if (s2 - s1 != len || s3 - s2 != len) return false;
    for (int u = 0; u < len; ++u) {
      const int c = ja[s0 + u];
      if (c % 3 || ja[s1 + u] != c + 1 || ja[s2 + u] != c + 2) return false;
      const double v = a[s0 + u];
      if (a[s1 + u] != v || a[s2 + u] != v) return false;
      bja.push_back(c / 3);
      mval.push_back(v);
    }
    bia[b + 1] = bia[b] + len;
  }
  return true;
}

static void choose_format(NmParcsr& M, const std::vector<int>& ia, const std::vector<int>& ja, const double* a) {
  std::vector<int> bia, bja;
  std::vector<double> mval;
  const bool aligned = (M.ncol % 3 == 0);
  const char* force = getenv("NM_FORCE_CSR");
  M.format = NM_FMT_CSR;
  M.avg_row = M.nrow ? (double)M.nnz / M.nrow : 0.0;
  M.fmt_bytes = 12ll * M.nnz + 4ll * (M.nrow + 1);
  if (force && force[0] == '1') return;
  if (aligned && M.nnz > 0 && detect_kron3(M.nrow, ia, ja, a, bia, bja, mval)) {
    M.format = NM_FMT_KRON3;
    M.nbrow = M.nrow / 3;
    M.bia.from_host(bia); M.bja.from_host(bja); M.mval.from_host(mval);
    M.avg_row = (double)bja.size() / M.nbrow;
    M.fmt_bytes = 12ll * (long long)bja.size() + 4ll * (M.nbrow + 1);
  } else if (aligned && M.nnz > 0 && detect_row3(M.nrow, ia, ja, bia, bja)) {
    M.format = NM_FMT_ROW3;
    M.nbrow = M.nrow / 3;
    M.bia.from_host(bia); M.bja.from_host(bja);
    M.avg_row = 3.0 * (double)bja.size() / M.nbrow;
    M.fmt_bytes = 8ll * M.nnz + 4ll * (long long)bja.size() + 4ll * (M.nbrow + 1);
  }
}

// ---------------------------------------------------------------- build
NmParcsr* nm_parcsr_build(int nrow_glob, int ncol_glob, const int* row_starts, const int* col_starts,
                          const int* ia, const int* ja, const double* a) {
  nm_ensure_init();
  NmCtx& c = nm_ctx();
  std::unique_ptr<NmParcsr> M(new NmParcsr());
  M->nrow_glob = nrow_glob; M->ncol_glob = ncol_glob;
  M->row0 = row_starts[c.rank]; M->nrow = row_starts[c.rank + 1] - row_starts[c.rank];
  M->col0 = col_starts[c.rank]; M->ncol = col_starts[c.rank + 1] - col_starts[c.rank];
  NM_REQUIRE(row_starts[c.nranks] == nrow_glob && col_starts[c.nranks] == ncol_glob,
             "parcsrcreate: row/col_starts[%d] do not match the global sizes", c.nranks);
  NM_REQUIRE(ia[0] == 0, "parcsrcreate: ia must be 0-based (src/mod_matvec.f90:59,112)");
  const int n = M->nrow;
  M->nnz = ia[n];
  const long long nnz = M->nnz;
  // ghost columns: sorted unique global ids outside [col0, col0+ncol)
  std::vector<int>& gl = M->halo.ghost_glob;
  const int lo = M->col0, hi = M->col0 + M->ncol;
  halo_ghosts_host(ncol_glob, lo, hi, nnz, ja, gl);
  M->halo.nghost = (int)gl.size();
  NM_REQUIRE(c.nranks > 1 || gl.empty(), "parcsrcreate: ghost columns on a single rank");
  std::vector<int> hia(ia, ia + n + 1), hja((size_t)nnz);
  for (long long p = 0; p < nnz; ++p) {
    const int g = ja[p];
    if (g >= lo && g < hi) hja[p] = g - lo;
    else hja[p] = M->ncol + (int)(std::lower_bound(gl.begin(), gl.end(), g) - gl.begin());
  }
  M->ia.from_host(hia);
  M->ja.alloc(std::max<size_t>(hja.size(), 1)); M->ja.upload(hja.data(), hja.size());
  M->a.alloc(std::max<size_t>((size_t)nnz, 1)); M->a.upload(a, (size_t)nnz);
  choose_format(*M, hia, hja, a);
  if (M->halo.nghost) M->halo.xg.alloc(M->halo.nghost);
  M->halo.xg_cur = M->halo.xg.p;
  build_halo_plan(*M, col_starts);
  return M.release();
}

// ---------------------------------------------------------------- scaled copy  S = diag(dr) M diag(dc)
// Folds the reference's per-call vector scalings (src/mod_matvec.f90:451-456, 505-519) into the
// matrix once.  The multiplication order (a*dc)*dr is the one Bdiagscaling uses (:336).
__global__ void k_scale_csr(int nrow, int ncol, const int* __restrict__ ia, const int* __restrict__ ja,
                            const double* __restrict__ a, double* __restrict__ out, const double* __restrict__ dr,
                            const double* __restrict__ dc, const double* __restrict__ dcg) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrow) return;
  const double r = dr[row];
  for (int p = ia[row]; p < ia[row + 1]; ++p) {
    const int c = ja[p];
    const double cv = c < ncol ? dc[c] : dcg[c - ncol];
    out[p] = (a[p] * cv) * r;
  }
}

NmParcsr* nm_parcsr_scaled_copy(const NmParcsr& M, const double* dr, const double* dc) {
  NmCtx& c = nm_ctx();
  std::unique_ptr<NmParcsr> S(new NmParcsr());
  S->nrow_glob = M.nrow_glob; S->ncol_glob = M.ncol_glob; S->nrow = M.nrow; S->ncol = M.ncol;
  S->row0 = M.row0; S->col0 = M.col0; S->nnz = M.nnz;
  S->format = M.format == NM_FMT_KRON3 ? NM_FMT_CSR : M.format;   // scaling keeps ROW3 structure, not KRON3 values
  S->nbrow = M.nbrow; S->avg_row = M.avg_row; S->fmt_bytes = M.fmt_bytes;
  auto clone_i = [&](const DBuf<int>& src, DBuf<int>& dst) {
    if (!src.n) return;
    dst.alloc(src.n);
    NM_CUDA(cudaMemcpyAsync(dst.p, src.p, src.n * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
  };
  clone_i(M.ia, S->ia); clone_i(M.ja, S->ja);
  if (S->format == NM_FMT_ROW3) { clone_i(M.bia, S->bia); clone_i(M.bja, S->bja); }
  if (S->format == NM_FMT_CSR) { S->avg_row = M.nrow ? (double)M.nnz / M.nrow : 0; S->fmt_bytes = 12ll * M.nnz + 4ll * (M.nrow + 1); }
  S->a.alloc(std::max<size_t>((size_t)M.nnz, 1));
  // halo plan is shared by value (index lists cloned)
  S->halo.nghost = M.halo.nghost; S->halo.ghost_glob = M.halo.ghost_glob;
  S->halo.recv_cnt = M.halo.recv_cnt; S->halo.recv_off = M.halo.recv_off;
  S->halo.send_cnt = M.halo.send_cnt; S->halo.send_off = M.halo.send_off; S->halo.nsend = M.halo.nsend;
  clone_i(M.halo.send_idx, S->halo.send_idx);
  if (S->halo.nsend) S->halo.sendbuf.alloc(S->halo.nsend);
  if (S->halo.nghost) S->halo.xg.alloc(S->halo.nghost);
  S->halo.xg_cur = S->halo.xg.p;
  S->halo.cnt_all = M.halo.cnt_all;
  if (c.nranks > 1) halo_p2p_setup(S->halo, S->halo.cnt_all);
  // ghost part of the column scaling through the halo plan of the copy
  if (S->halo.nghost > 0 || S->halo.nsend > 0) nm_halo_exchange(*S, dc);
  if (M.nrow) {
    k_scale_csr<<<nm_div_up(M.nrow, 128), 128, 0, c.stream>>>(M.nrow, M.ncol, M.ia.p, M.ja.p, M.a.p, S->a.p, dr, dc,
                                                              S->halo.xg_cur);
    c.launches++;
  }
  NM_CUDA(cudaStreamSynchronize(c.stream));
  return S.release();
}

// ---------------------------------------------------------------- Jacobi scaling (K10)
// Bdiagscaling / Apdiagscaling (src/mod_matvec.f90:252-342, 345-441): d_i = 1/sqrt(sign*M_ii), ghost d_j
// through the halo plan (the reference's alltoallv, :303-325), M~_ij = (sign*M_ij*d_j)*d_i in place.
__global__ void k_diag_rsqrt(int nrow, const int* __restrict__ ia, const int* __restrict__ ja,
                             const double* __restrict__ a, double sign, double* __restrict__ d, int* __restrict__ bad) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrow) return;
  double diag = 0.0;
  int found = 0;
  for (int p = ia[row]; p < ia[row + 1]; ++p)
    if (ja[p] == row) { diag = a[p]; found = 1; }
  diag *= sign;
  if (!found || !(diag > 0.0)) { atomicAdd(bad, 1); d[row] = 1.0; return; }
  d[row] = 1.0 / sqrt(diag);
}
__global__ void k_scale_inplace(int nrow, int ncol, const int* __restrict__ ia, const int* __restrict__ ja,
                                double* __restrict__ a, double sign, const double* __restrict__ d,
                                const double* __restrict__ dg) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrow) return;
  const double r = d[row];
  for (int p = ia[row]; p < ia[row + 1]; ++p) {
    const int c = ja[p];
    const double cv = c < ncol ? d[c] : dg[c - ncol];
    a[p] = (sign * a[p] * cv) * r;
  }
}
__global__ void k_kron_refresh(int nbrow, const int* __restrict__ bia, const int* __restrict__ ia,
                               const double* __restrict__ a, double* __restrict__ mval) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbrow) return;
  const int s = bia[b], len = bia[b + 1] - s, r0 = ia[3 * b];
  for (int u = 0; u < len; ++u) mval[s + u] = a[r0 + u];
}

extern "C" int nm_parcsr_jacobi_scale(void* h, double sign, double* d_host) {
  NM_API_BEGIN
  NmParcsr& M = *(NmParcsr*)h;
  NmCtx& c = nm_ctx();
  NM_REQUIRE(M.nrow == M.ncol && M.row0 == M.col0, "jacobi_scale: square matrix with matching row/col distribution required");
  const int n = M.nrow;
  DBuf<double> d(std::max(n, 1));
  DBuf<int> bad(1);
  bad.zero();
  if (n) {
    k_diag_rsqrt<<<nm_div_up(n, 128), 128, 0, c.stream>>>(n, M.ia.p, M.ja.p, M.a.p, sign, d.p, bad.p);
    c.launches++;
  }
  int hbad = 0;
  bad.download(&hbad, 1);
  NM_REQUIRE(hbad == 0, "jacobi_scale: %d rows without a positive diagonal (sign %g)", hbad, sign);
  nm_halo_exchange(M, d.p);
  if (n) {
    k_scale_inplace<<<nm_div_up(n, 128), 128, 0, c.stream>>>(n, M.ncol, M.ia.p, M.ja.p, M.a.p, sign, d.p, M.halo.xg_cur);
    c.launches++;
    if (M.format == NM_FMT_KRON3) {
      k_kron_refresh<<<nm_div_up(M.nbrow, 128), 128, 0, c.stream>>>(M.nbrow, M.bia.p, M.ia.p, M.a.p, M.mval.p);
      c.launches++;
    }
  }
  if (d_host) d.download(d_host, n);
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_API_END
}

// copy of the (local) values back to the host, e.g. to inspect B~ after nm_parcsr_jacobi_scale
extern "C" int nm_parcsr_get_values(void* h, double* a_host) {
  NM_API_BEGIN
  NmParcsr& M = *(NmParcsr*)h;
  M.a.download(a_host, (size_t)M.nnz);
  NM_API_END
}

// ---------------------------------------------------------------- plain products
void nm_spmv(NmParcsr& M, const double* x, double* y) { nm_spmv_epi(M, x, EpiStore{y}); }
void nm_spmv_add(NmParcsr& M, const double* x, double* y) { nm_spmv_epi(M, x, EpiAdd{y}); }

// ---------------------------------------------------------------- C ABI
extern "C" int nm_parcsr_create(int nrow_glob, int ncol_glob, const int* row_starts, const int* col_starts,
                                const int* ia, const int* ja, const double* a, void** out) {
  NM_API_BEGIN
  NM_REQUIRE(out, "nm_parcsr_create: null output handle");
  *out = nm_parcsr_build(nrow_glob, ncol_glob, row_starts, col_starts, ia, ja, a);
  NM_API_END
}

extern "C" int nm_parcsr_free(void* h) {
  NM_API_BEGIN
  if (h) { NM_CUDA(cudaStreamSynchronize(nm_ctx().stream)); delete (NmParcsr*)h; }
  NM_API_END
}

// y = M x with HOST vectors (the reference's pevsl_parcsrmatvec_f90 contract, src/mod_matvec.f90:453).
extern "C" int nm_parcsr_matvec(void* h, const double* x, double* y) {
  NM_API_BEGIN
  NmParcsr& M = *(NmParcsr*)h;
  NmCtx& c = nm_ctx();
  DBuf<double> dx(std::max(M.ncol, 1)), dy(std::max(M.nrow, 1));
  if (M.ncol) NM_CUDA(cudaMemcpyAsync(dx.p, x, M.ncol * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  nm_spmv(M, dx.p, dy.p);
  if (M.nrow) NM_CUDA(cudaMemcpyAsync(y, dy.p, M.nrow * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_API_END
}

// y = M x with DEVICE vectors (owned parts), asynchronous on nm_stream().
extern "C" int nm_parcsr_matvec_dev(void* h, const double* x_dev, double* y_dev) {
  NM_API_BEGIN
  nm_spmv(*(NmParcsr*)h, x_dev, y_dev);
  NM_API_END
}

// halo exchange alone (device vector of the owned columns), asynchronous on nm_stream(): timing / diagnostics
extern "C" int nm_parcsr_halo_exchange_dev(void* h, const double* x_dev) {
  NM_API_BEGIN
  nm_halo_exchange(*(NmParcsr*)h, x_dev);
  NM_API_END
}
// mode: 0 single rank / no halo, 1 NCCL send/recv, 2 NVLink peer window (direct peer stores + arrival flags)
extern "C" int nm_parcsr_halo_info(void* h, int* mode, int* nghost, int* nsend) {
  NM_API_BEGIN
  NmParcsr& M = *(NmParcsr*)h;
  if (mode) *mode = nm_ctx().nranks == 1 ? 0 : (M.halo.p2p ? 2 : 1);
  if (nghost) *nghost = M.halo.nghost;
  if (nsend) *nsend = M.halo.nsend;
  NM_API_END
}
extern "C" int nm_parcsr_info(void* h, int* nrow, int* ncol, long long* nnz, int* format, int* nghost,
                              long long* fmt_bytes) {
  NM_API_BEGIN
  NmParcsr& M = *(NmParcsr*)h;
  if (nrow) *nrow = M.nrow;
  if (ncol) *ncol = M.ncol;
  if (nnz) *nnz = M.nnz;
  if (format) *format = M.format;
  if (nghost) *nghost = M.halo.nghost;
  if (fmt_bytes) *fmt_bytes = M.fmt_bytes;
  NM_API_END
}
