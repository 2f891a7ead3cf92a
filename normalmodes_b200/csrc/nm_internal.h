// normalmodes_b200 -- internal declarations shared by the CUDA translation units.
// Nothing in here is part of the C ABI (see include/nm_b200.h, include/pevsl_f90.h).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>
#include <memory>

#define NM_OK 0
#define NM_ERR 1

// ---------------------------------------------------------------- errors
struct NmError : public std::runtime_error {
  explicit NmError(const std::string& s) : std::runtime_error(s) {}
};
void nm_fail(const char* file, int line, const char* fmt, ...);
void nm_record_error(const char* msg);
#define NM_CUDA(x)                                                                         \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) nm_fail(__FILE__, __LINE__, "CUDA: %s (%s)", cudaGetErrorString(e_), #x); \
  } while (0)
#define NM_NCCL(x)                                                                         \
  do {                                                                                     \
    ncclResult_t e_ = (x);                                                                 \
    if (e_ != ncclSuccess) nm_fail(__FILE__, __LINE__, "NCCL: %s (%s)", ncclGetErrorString(e_), #x); \
  } while (0)
#define NM_REQUIRE(c, ...)                                  \
  do {                                                      \
    if (!(c)) nm_fail(__FILE__, __LINE__, __VA_ARGS__);     \
  } while (0)
// C-ABI wrappers: run body, convert exceptions into an error code + message.
#define NM_API_BEGIN try {
#define NM_API_END                               \
  }                                              \
  catch (const std::exception& e) {              \
    nm_record_error(e.what());                   \
    return NM_ERR;                               \
  }                                              \
  return NM_OK;

// ---------------------------------------------------------------- runtime context
struct NmCtx {
  bool ready = false;
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  int rank = 0, nranks = 1;
  ncclComm_t nccl = nullptr;
  double* d_red = nullptr;        // reduction scratch (partials), device
  size_t d_red_cap = 0;
  double* h_pin = nullptr;        // small pinned host scratch
  size_t h_pin_cap = 0;
  long launches = 0;              // kernels launched by this library (bench 'gpu_launches')
  // NVLink peer window (one process per GPU, CUDA IPC): ghost buffers and arrival flags of every matrix live in
  // one device allocation per rank that all peers map, so the halo exchange is direct peer stores + a flag
  // (k_halo_push, nm_parcsr.cu) instead of NCCL send/recv.  p2p == false: NCCL fallback.
  bool p2p = false;
  unsigned char* win = nullptr;   // this rank's window
  size_t win_bytes = 0, win_used = 0;
  std::vector<unsigned char*> peer_win;   // peer_win[r]: rank r's window mapped here (peer_win[rank] = win)
  unsigned* push_ctr = nullptr;   // last-block detection counter of k_halo_push (device)
  int* dev_status = nullptr;      // device-side error word (bit 0: halo flag wait timed out)
};
NmCtx& nm_ctx();
void nm_ensure_init();
double* nm_red_scratch(size_t n);
size_t nm_win_alloc(size_t bytes);          // slot in the peer window (256-byte granules, zeroed), collective order
void nm_win_free(size_t off, size_t bytes);
void nm_check_device_status();              // raises if a device-side wait timed out
double* nm_pinned(size_t n);

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  explicit DBuf(size_t n_) { alloc(n_); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t n_) {
    release();
    n = n_;
    // +16 bytes: the TMA bulk copies of the streaming SpMV round slice ends up to the next 16-byte boundary
    if (n) NM_CUDA(cudaMalloc((void**)&p, n * sizeof(T) + 16));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void upload(const T* h, size_t cnt) {
    if (cnt) NM_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, nm_ctx().stream));
    NM_CUDA(cudaStreamSynchronize(nm_ctx().stream));
  }
  void from_host(const std::vector<T>& h) { alloc(h.size()); upload(h.data(), h.size()); }
  void download(T* h, size_t cnt) const {
    if (cnt) NM_CUDA(cudaMemcpyAsync(h, p, cnt * sizeof(T), cudaMemcpyDeviceToHost, nm_ctx().stream));
    NM_CUDA(cudaStreamSynchronize(nm_ctx().stream));
  }
  void zero() { if (n) NM_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), nm_ctx().stream)); }
};

// ---------------------------------------------------------------- distributed CSR
enum NmFormat { NM_FMT_CSR = 0, NM_FMT_ROW3 = 1, NM_FMT_KRON3 = 2 };

struct NmHalo {
  int nghost = 0;                         // ghost columns, sorted by global id (= grouped by owner rank)
  std::vector<int> ghost_glob;            // global column ids of the ghosts
  std::vector<int> recv_cnt, recv_off;    // per peer, into the ghost tail
  std::vector<int> send_cnt, send_off;    // per peer, into sendbuf
  int nsend = 0;
  DBuf<int> send_idx;                     // local (owned) column ids to pack, device
  DBuf<double> sendbuf;                   // device
  DBuf<double> xg;                        // ghost values, device (gather target c >= ncol -> xg[c-ncol])
  double* xg_cur = nullptr;               // ghost buffer the kernels read after the last exchange (xg.p or a window half)
  // peer-store exchange (NmCtx::p2p): two ghost buffers (parity of the exchange counter) + one arrival flag per
  // source rank in this rank's window; where this rank's values go in every peer's window
  size_t win_xg[2] = {0, 0}, win_flag = 0;             // byte offsets in this rank's window
  std::vector<size_t> peer_xg[2], peer_flag;           // per peer: byte offsets in THAT peer's window
  std::vector<int> peer_base;                           // per peer: where this rank's block starts in its ghost tail
  std::vector<int> cnt_all;                             // P x P: cnt_all[r*P+s] = ghosts rank r receives from rank s
  unsigned long long epoch = 0;                         // exchanges done (flag value of the next one = epoch + 1)
  bool p2p = false;
  // flag-in-data ("LL") ghost slots of the persistent ChebIter kernel (k_slabpers): 16 bytes per ghost value =
  // {lo32, tag, hi32, tag}, three rotating buffers in the window; no fences, flags or atomics on the exchange path
  bool ll = false;
  size_t win_ll[3] = {0, 0, 0};
  std::vector<size_t> peer_ll[3];
  NmHalo() {}
  NmHalo(const NmHalo&) = delete;
  NmHalo& operator=(const NmHalo&) = delete;
  ~NmHalo() {                               // hand the window slots back (nm_parcsr_free / nm_op_free)
    if (p2p) {
      const size_t gb = (size_t)(nghost > 0 ? nghost : 1) * sizeof(double);
      nm_win_free(win_xg[0], gb); nm_win_free(win_xg[1], gb); nm_win_free(win_flag, sizeof(unsigned long long) * 8);
    }
    if (ll) for (int b = 0; b < 3; ++b) nm_win_free(win_ll[b], 16 * (size_t)(nghost > 0 ? nghost : 1));
  }
};

struct NmPackDesc { unsigned off16, bytes; };   // blob offset in 16-byte units, blob size in bytes

// Warp-sliced ELL slabs in pack order (k_slab, nm_slab.cuh / nm_slab.cu): index rows grouped breadth-first into
// compact chunks of at most T lanes (T = threads per CTA); a row of up to NM_SLAB_SPLIT entries is walked by ONE
// thread, a longer one by an aligned group of 2, 4, ... lanes (combined by shuffles); lanes are grouped into
// slices of 32 (one warp), each padded to its longest lane, entries step-major -> stride-32 conflict-free
// shared-memory reads, row sums in registers, no partial sums in memory, no offsets table.
// A chunk is one contiguous 16-byte aligned blob (header, slice table, values, distinct column ids, 16-bit
// chunk-local column indices) moved by one TMA bulk copy.  Vectors live in pack order (row j of chunk c is
// element first(c)+j).
struct NmSlabHeader { int nr, nd, nslice, first, nep, gmax, has_ghost, pad2; };   // 32 bytes; gmax: most lanes per row
struct NmSlab {
  // persistent kernel (k_slabpers): stages per CTA when the whole iteration runs in one cooperative launch
  int pers_nstage = 0, pers_smem = 0;
  // dataflow execution: chunk id of every descriptor, per-chunk completion flags (epoch of the last finished step);
  // the ids of the chunks a chunk depends on sit at the end of its blob (header word 7 = their number)
  DBuf<int> desc_cid;
  DBuf<unsigned> cflag;
  bool deps_ok = false;
  DBuf<unsigned char> blob;
  DBuf<NmPackDesc> desc;
  DBuf<int> cta_first;                    // grid+1: first chunk of each CTA (balanced by bytes)
  DBuf<unsigned> slot_off8;
  DBuf<int> slot_src;
  long long nslot = 0;
  int nchunk = 0;                         // 0: not built
  int grid = 0, threads = 0, max_chunks_per_cta = 0;
  int stage_bytes = 0, xs_doubles = 0, nstage = 0, smem_bytes = 0;
  bool ws = false;                        // warp-specialised kernel (k_slabws): nprod producer warps, nxs x buffers
  int nxs = 2, nprod = 0;
  bool pdl = true;                        // programmatic dependent launch of consecutive steps (NM_SLAB_PDL)
  long long bytes = 0;                    // blob bytes = what one product streams
  long long entries = 0, padded_entries = 0;
  DBuf<int> order;                        // pack position -> caller's index row
  DBuf<long long> trace;                  // NM_SLAB_TRACE=1: grid x 64 x 8 clock stamps of the last launch
};

struct NmParcsr {
  int nrow_glob = 0, ncol_glob = 0;
  int nrow = 0, ncol = 0;                 // local (owned) rows / columns
  int row0 = 0, col0 = 0;                 // first global row / column owned
  long long nnz = 0;
  DBuf<int> ia, ja;                       // local CSR; ja in local column space (owned first, ghosts after)
  DBuf<double> a;
  int format = NM_FMT_CSR;
  // ROW3: rows 3b..3b+2 share one column list made of aligned triples -> one block-column id per
  //       triple (bja), values stay in CSR order.  KRON3: additionally M (x) I3 -> scalar values mval.
  int nbrow = 0;
  DBuf<int> bia, bja;
  DBuf<double> mval;
  NmHalo halo;
  long long values_version = 0;           // bumped whenever the values change (dependants refill their packs)
  double avg_row = 0.0;                   // mean entries per (block-)row processed by one subwarp
  long long fmt_bytes = 0;                // bytes one SpMV streams in the chosen format (matrix part)
};

// fused multi-GPU step (nm_slab.cuh): where the new direction of a boundary row goes
struct NmPushEnt { unsigned dst; unsigned short peer, comp; };   // position in that peer's ghost buffer, peer rank, scalar component

// ---------------------------------------------------------------- ChebIter (B^-1, Ap^-1)
struct NmChebIter {
  NmParcsr* M = nullptr;
  double lb = 0, ub = 0, theta = 0;
  int deg = 0;
  std::vector<double> ak, bk;             // d_{k+1} = ak d_k + bk r_{k+1}
  DBuf<double> r, d0, d1;
  // pack-order slab copy of M (nm_slab_build_into): the iteration runs on vectors kept in that order
  NmSlab pslab;
  long long ppack_version = -1;
  DBuf<double> bp, xp;                    // b and x in pack order
  DBuf<int> send_idx_p;                   // halo send list in pack order
  // per pack-order index row the peer stores of its new direction: in-kernel halo of the persistent kernel
  // (k_slabpers, NM_SLAB_PERS=1) and of the per-step fused kernel (default on several GPUs)
  bool fused = false;
  bool pers = false;                      // whole iteration in ONE cooperative launch (grid barrier between steps)
  bool pers_multi_ok = false;             // the halo pattern allows the in-kernel exchange (symmetric)
  DBuf<int> push_off;
  DBuf<NmPushEnt> push_ent;
  std::vector<int> h_sidx_p;              // host copy of send_idx_p
  DBuf<int> ghost_slot;                   // ghost column (position in the ghost tail) -> slot in this rank's buffers
  DBuf<int> send_slot;                    // send-list entry -> slot in the receiving peer's buffers
  DBuf<double> ak_dev, bk_dev;
  DBuf<unsigned long long> gbar;          // grid-barrier counter of the persistent kernel (monotonic)
  unsigned long long gbar_base = 0;
  unsigned ftag = 0;                      // dataflow: chunk-flag epoch consumed so far (a solve uses deg of them)
  long long nsolve = 0, nmatvec = 0;
  double t_total = 0;
};

// ---------------------------------------------------------------- operators (the reference's callbacks)
typedef void (*nm_matvec_fn)(double* x, double* y, void* data);
enum NmOpKind { NM_OP_CSR = 0, NM_OP_SOLID = 1, NM_OP_FLUIDSOLID = 2, NM_OP_CALLBACK = 3 };

struct NmOp {
  int kind = NM_OP_CSR;
  int n = 0;                               // local length of x and y
  NmParcsr* M = nullptr;                   // NM_OP_CSR: borrowed
  std::unique_ptr<NmParcsr> As, Es, ETs;   // scaled copies D A D, D E Dp, Dp ET D (owned)
  NmChebIter* chebAp = nullptr;            // borrowed
  DBuf<double> x1, y0, w1;                 // fluid work vectors (pressure, pressure, displacement)
  nm_matvec_fn fn = nullptr;               // NM_OP_CALLBACK
  void* fn_data = nullptr;
  std::vector<double> hx, hy;
  long long napply = 0;
};

// ---------------------------------------------------------------- filter polynomial
struct NmPol {
  int deg = 0, type = 0;
  std::vector<double> mu;
  double cc = 0, dd = 0, gam = 0, bar = 0;
  double intv[4] = {0, 0, 0, 0};
};

// ---------------------------------------------------------------- solver context (pevsl handle)
struct NmPevsl {
  int N = 0, n = 0, nfirst = -1;
  NmOp* A = nullptr;
  NmOp* B = nullptr;
  NmChebIter* bsol = nullptr;
  bool geneig = false;
  std::vector<std::unique_ptr<NmOp>> owned_ops;
  // results of the last cheblannr
  int nev = 0;
  std::vector<double> lam, res;
  DBuf<double> Y;                          // n x nev, column-major, B-orthonormal (scaled coordinates)
  // statistics
  int last_steps = 0, last_deg = 0;
  double t_filter = 0, t_reorth = 0, t_total = 0, t_ritz = 0;
  long long n_filter_apply = 0;
  unsigned long long seed = 4321;
  double ritz_tol = 0.0;                   // > 0: per-pair residual-estimate gate on top of the trace test
  DBuf<double> fwork;                      // device scratch of the host-vector filter entry points (5n)
};

// ---------------------------------------------------------------- internal entry points
// parcsr
NmParcsr* nm_parcsr_build(int nrow_glob, int ncol_glob, const int* row_starts, const int* col_starts,
                          const int* ia, const int* ja, const double* a);
NmParcsr* nm_parcsr_scaled_copy(const NmParcsr& M, const double* drow_dev, const double* dcol_dev);
void nm_halo_exchange(NmParcsr& M, const double* x, const int* send_idx = nullptr);   // send_idx: override of the pack list
// Peer-window exchange WITHOUT the arrival wait: pushes x to the peers and raises the flags; the consumer kernel polls
// (out: this rank's flag array, the sources to wait for, the flag value).  Returns false when the matrix uses NCCL or
// has no halo (then nothing was done and the caller uses nm_halo_exchange).
struct NmHaloWait { const unsigned long long* flags; unsigned mask; unsigned long long epoch; int* status; };
bool nm_halo_push_nowait(NmParcsr& M, const double* x, const int* send_idx, NmHaloWait* w);
bool nm_halo_ll_setup(NmHalo& h);            // collective: LL ghost slots in the peer window; false without a window
void nm_halo_push_ll(NmParcsr& M, const double* x, const int* send_idx, const int* send_slot, unsigned tag, int buf);
void nm_spmv(NmParcsr& M, const double* x, double* y);                 // y = M x   (device pointers)
void nm_spmv_add(NmParcsr& M, const double* x, double* y);             // y += M x
void nm_slab_build_into(NmParcsr& M, NmSlab& S, const std::vector<int>& rp, const std::vector<int>& idx, int n);
void nm_slab_fill_from(NmParcsr& M, NmSlab& S);
int nm_env_int(const char* name, int dflt);
// chebiter
NmChebIter* nm_chebiter_build(double lb, double ub, int deg, NmParcsr* M);
void nm_chebiter_solve(NmChebIter& C, const double* b, double* x);
// operators
void nm_op_apply(NmOp& op, const double* x, double* y);
// y-update of ChebAv fused into the operator: vout = t*(Op(w) - cc*vk) - vkm1 ; y (+)= mu*vout
void nm_op_apply_filter(NmOp& op, const double* w, const double* vk, const double* vkm1, double* vout, double* y,
                        double t, double cc, double mu, double mu0, int first);
// vector kernels (all on nm_ctx().stream)
void nm_vec_copy(double* dst, const double* src, size_t n);
void nm_vec_set(double* dst, double v, size_t n);
void nm_vec_scale(double* x, double s, size_t n);
void nm_vec_mul(double* y, const double* x, const double* d, size_t n);      // y = x .* d
void nm_vec_axpy(double* y, double a, const double* x, size_t n);
void nm_vec_axpy_dev(double* y, const double* a_dev, double sign, const double* x, size_t n);  // y += sign*(*a)*x
void nm_vec_dot_dev(const double* x, const double* y, size_t n, double* out_dev);              // allreduced
double nm_vec_dot(const double* x, const double* y, size_t n);
void nm_vec_random(double* x, size_t n, unsigned long long seed, unsigned long long offset);
void nm_allreduce_sum(double* buf_dev, size_t n);
void nm_filter_update(const double* vkm1, double* vout, const double* vk, const double* u, double* y, double t,
                      double cc, double mu, double mu0, int first, size_t n);
// host maths
int nm_tridiag_eig(int k, const double* d, const double* e, double* w, double* Z /* k*k col-major or null */);
int nm_tridiag_eig_ex(int k, const double* d, const double* e, double* w, double* Z, double* lastrow);
int nm_sym_geneig(int m, const double* H, const double* G, double* w, double* C);   // dense H c = w G c (Cholesky + Jacobi)
void nm_findpol(const double xintv[4], double thresh_int, double thresh_ext, NmPol& pol);
// solver
void nm_lanbounds(NmPevsl& P, int mlan, int lanstep, double tol, double* lmin, double* lmax);
void nm_cheblannr(NmPevsl& P, const double xintv[4], int maxit, double tol, const NmPol& pol);
void nm_filter_apply(NmPevsl& P, const NmPol& pol, const double* z, double* y, double* work /* 3n */, int kmax = 0);

static inline int nm_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
