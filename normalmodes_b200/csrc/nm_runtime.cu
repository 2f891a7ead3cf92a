// Runtime context, error reporting, NCCL plumbing and the small vector kernels of the
// Lanczos driver (axpy / dot / scale).  All work is queued on one stream (nm_ctx().stream).
#include "nm_internal.h"
#include <algorithm>
#include <cstdarg>
#include <mutex>

// ---------------------------------------------------------------- errors
static thread_local std::string g_last_error;
static int g_last_code = 0;

void nm_record_error(const char* msg) {
  g_last_error = msg ? msg : "unknown error";
  g_last_code = NM_ERR;
  fprintf(stderr, "[nm_b200] Error: %s\n", g_last_error.c_str());
}

void nm_fail(const char* file, int line, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  char out[1200];
  snprintf(out, sizeof(out), "%s [%s:%d]", buf, file, line);
  throw NmError(out);
}

extern "C" int nm_last_error(void) { return g_last_code; }
extern "C" const char* nm_last_error_message(void) { return g_last_error.c_str(); }
extern "C" void nm_clear_error(void) { g_last_code = 0; g_last_error.clear(); }

// ---------------------------------------------------------------- context
static NmCtx g_ctx;
NmCtx& nm_ctx() { return g_ctx; }

void nm_ensure_init() {
  NmCtx& c = g_ctx;
  if (c.ready) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    nm_fail(__FILE__, __LINE__, "no CUDA device visible (%s): normalmodes_b200 has no CPU fallback",
            cudaGetErrorString(e));
  NM_CUDA(cudaGetDevice(&c.device));
  cudaDeviceProp prop;
  NM_CUDA(cudaGetDeviceProperties(&prop, c.device));
  c.sm_count = prop.multiProcessorCount;
  NM_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.ready = true;
}

double* nm_red_scratch(size_t n) {
  NmCtx& c = g_ctx;
  if (n > c.d_red_cap) {
    if (c.d_red) { NM_CUDA(cudaStreamSynchronize(c.stream)); cudaFree(c.d_red); }
    size_t cap = n < 65536 ? 65536 : n;
    NM_CUDA(cudaMalloc((void**)&c.d_red, cap * sizeof(double)));
    c.d_red_cap = cap;
  }
  return c.d_red;
}

double* nm_pinned(size_t n) {
  NmCtx& c = g_ctx;
  if (n > c.h_pin_cap) {
    if (c.h_pin) cudaFreeHost(c.h_pin);
    size_t cap = n < 4096 ? 4096 : n;
    NM_CUDA(cudaMallocHost((void**)&c.h_pin, cap * sizeof(double)));
    c.h_pin_cap = cap;
  }
  return c.h_pin;
}

extern "C" int nm_init(int device) {
  NM_API_BEGIN
  if (g_ctx.ready) {
    NM_REQUIRE(device < 0 || device == g_ctx.device, "nm_init: already initialised on device %d", g_ctx.device);
  } else {
    if (device >= 0) NM_CUDA(cudaSetDevice(device));
    nm_ensure_init();
  }
  NM_API_END
}

extern "C" int nm_device_info(int* device, int* sm_count, int* rank, int* nranks) {
  NM_API_BEGIN
  nm_ensure_init();
  if (device) *device = g_ctx.device;
  if (sm_count) *sm_count = g_ctx.sm_count;
  if (rank) *rank = g_ctx.rank;
  if (nranks) *nranks = g_ctx.nranks;
  NM_API_END
}

extern "C" long long nm_launch_count(void) { return g_ctx.launches; }
extern "C" void* nm_stream(void) { return (void*)g_ctx.stream; }
extern "C" int nm_sync(void) {
  NM_API_BEGIN
  nm_ensure_init();
  NM_CUDA(cudaStreamSynchronize(g_ctx.stream));
  nm_check_device_status();
  NM_API_END
}

// ---------------------------------------------------------------- NCCL communicator (one rank per GPU)
extern "C" int nm_comm_unique_id(char* id128) {
  NM_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  NM_NCCL(ncclGetUniqueId(&id));
  memcpy(id128, &id, 128);
  NM_API_END
}

// ---------------------------------------------------------------- NVLink peer window (CUDA IPC)
// Every rank allocates one window, publishes its IPC handle through the NCCL communicator and maps the windows
// of all peers.  All-or-nothing: if any rank fails to map any peer, every rank falls back to NCCL send/recv.
static void p2p_setup() {
  NmCtx& c = g_ctx;
  const char* off = getenv("NM_P2P");
  int want = !(off && off[0] == '0') && c.nranks <= 8;
  const size_t bytes = (size_t)std::max(16, nm_env_int("NM_P2P_WINDOW_MB", 256)) << 20;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  int ok = want;
  if (ok && cudaMalloc((void**)&c.win, bytes) != cudaSuccess) { cudaGetLastError(); c.win = nullptr; ok = 0; }
  if (ok) {
    NM_CUDA(cudaMemset(c.win, 0, bytes));
    if (cudaIpcGetMemHandle(&mine, c.win) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // handles + per-rank ok flag in one all-gather (sizeof handle = 64)
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad[3]; };
  static_assert(sizeof(Rec) == 80, "Rec layout");
  Rec rec; rec.h = mine; rec.ok = ok; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
  DBuf<unsigned char> d_mine(sizeof(Rec)), d_all(sizeof(Rec) * c.nranks);
  d_mine.upload((const unsigned char*)&rec, sizeof(Rec));
  NM_NCCL(ncclAllGather(d_mine.p, d_all.p, sizeof(Rec), ncclChar, c.nccl, c.stream));
  std::vector<Rec> all(c.nranks);
  d_all.download((unsigned char*)all.data(), sizeof(Rec) * c.nranks);
  for (const Rec& r : all) ok = ok && r.ok;
  c.peer_win.assign(c.nranks, nullptr);
  if (ok) {
    for (int r = 0; r < c.nranks; ++r) {
      if (r == c.rank) { c.peer_win[r] = c.win; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
      c.peer_win[r] = (unsigned char*)p;
    }
  }
  // agree: every rank must have mapped every peer
  DBuf<double> d_ok(1);
  double v = ok ? 1.0 : 0.0;
  d_ok.upload(&v, 1);
  NM_NCCL(ncclAllReduce(d_ok.p, d_ok.p, 1, ncclDouble, ncclMin, c.nccl, c.stream));
  d_ok.download(&v, 1);
  if (v < 0.5) {
    for (int r = 0; r < c.nranks; ++r)
      if (r != c.rank && c.peer_win[r]) cudaIpcCloseMemHandle(c.peer_win[r]);
    c.peer_win.clear();
    if (c.win) cudaFree(c.win);
    c.win = nullptr; c.p2p = false;
    if (want && c.rank == 0) fprintf(stderr, "[nm_b200] peer window unavailable: halo exchange falls back to NCCL send/recv\n");
    return;
  }
  c.win_bytes = bytes; c.win_used = 0; c.p2p = true;
  NM_CUDA(cudaMalloc((void**)&c.push_ctr, 256));
  NM_CUDA(cudaMemset(c.push_ctr, 0, 256));
  c.dev_status = (int*)(c.push_ctr + 32);
}

// Window slots: first-fit from the free list of released slots, else bump allocation; a slot is zeroed before it is
// handed out (arrival flags and flag-in-data tags of its previous owner must not satisfy a new matrix' first wait).
// Allocation order is collective (every rank creates / frees the same matrices in the same order), and the offsets
// are exchanged through the communicator afterwards, which also orders the zeroing before any peer's first store.
static std::vector<std::pair<size_t, size_t>> g_win_free;       // (offset, bytes), 256-byte granules
int nm_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}

size_t nm_win_alloc(size_t bytes) {
  NmCtx& c = g_ctx;
  NM_REQUIRE(c.p2p, "nm_win_alloc without a peer window");
  bytes = (bytes + 255) & ~(size_t)255;
  size_t off = (size_t)-1;
  for (size_t i = 0; i < g_win_free.size(); ++i)
    if (g_win_free[i].second >= bytes) {
      off = g_win_free[i].first;
      if (g_win_free[i].second == bytes) g_win_free.erase(g_win_free.begin() + i);
      else { g_win_free[i].first += bytes; g_win_free[i].second -= bytes; }
      break;
    }
  if (off == (size_t)-1) {
    off = (c.win_used + 255) & ~(size_t)255;
    NM_REQUIRE(off + bytes <= c.win_bytes, "peer window exhausted (%zu + %zu > %zu bytes): raise NM_P2P_WINDOW_MB", off, bytes,
               c.win_bytes);
    c.win_used = off + bytes;
  }
  NM_CUDA(cudaMemsetAsync(c.win + off, 0, bytes, c.stream));
  return off;
}
void nm_win_free(size_t off, size_t bytes) {
  NmCtx& c = g_ctx;
  if (!c.p2p || !c.win) return;
  bytes = (bytes + 255) & ~(size_t)255;
  g_win_free.emplace_back(off, bytes);
  // merge neighbours so that a sequence of create / free of the same matrices does not fragment the window
  std::sort(g_win_free.begin(), g_win_free.end());
  for (size_t i = 0; i + 1 < g_win_free.size();)
    if (g_win_free[i].first + g_win_free[i].second == g_win_free[i + 1].first) {
      g_win_free[i].second += g_win_free[i + 1].second;
      g_win_free.erase(g_win_free.begin() + i + 1);
    } else ++i;
  if (!g_win_free.empty() && g_win_free.back().first + g_win_free.back().second == c.win_used) {
    c.win_used = g_win_free.back().first;
    g_win_free.pop_back();
  }
}

void nm_check_device_status() {
  NmCtx& c = g_ctx;
  if (!c.dev_status) return;
  int st = 0;
  NM_CUDA(cudaMemcpyAsync(&st, c.dev_status, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  NM_CUDA(cudaStreamSynchronize(c.stream));
  NM_REQUIRE(st == 0, "device status %d: a halo arrival flag was not seen within the time limit (peer rank stalled or died)", st);
}

extern "C" int nm_comm_init(int rank, int nranks, const char* id128) {
  NM_API_BEGIN
  nm_ensure_init();
  NM_REQUIRE(g_ctx.nccl == nullptr, "nm_comm_init: communicator already initialised");
  NM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "nm_comm_init: bad rank %d of %d", rank, nranks);
  g_ctx.rank = rank;
  g_ctx.nranks = nranks;
  if (nranks > 1) {
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NM_NCCL(ncclCommInitRank(&g_ctx.nccl, nranks, id, rank));
    p2p_setup();
  }
  NM_API_END
}

extern "C" int nm_comm_finalize(void) {
  NM_API_BEGIN
  if (g_ctx.nccl) {
    NM_CUDA(cudaStreamSynchronize(g_ctx.stream));
    if (g_ctx.p2p) {
      for (int r = 0; r < g_ctx.nranks; ++r)
        if (r != g_ctx.rank && g_ctx.peer_win[r]) cudaIpcCloseMemHandle(g_ctx.peer_win[r]);
      g_ctx.peer_win.clear();
      cudaFree(g_ctx.win); cudaFree(g_ctx.push_ctr);
      g_ctx.win = nullptr; g_ctx.push_ctr = nullptr; g_ctx.dev_status = nullptr; g_ctx.p2p = false;
      g_win_free.clear(); g_ctx.win_used = 0;
    }
    ncclCommDestroy(g_ctx.nccl);
    g_ctx.nccl = nullptr;
  }
  g_ctx.rank = 0;
  g_ctx.nranks = 1;
  NM_API_END
}

void nm_allreduce_sum(double* buf, size_t n) {
  NmCtx& c = g_ctx;
  if (c.nranks > 1) NM_NCCL(ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, c.nccl, c.stream));
}

// ---------------------------------------------------------------- vector kernels
#define NM_VEC_THREADS 256
static inline int vec_grid(size_t n) {
  long long g = (long long)((n + NM_VEC_THREADS - 1) / NM_VEC_THREADS);
  long long cap = (long long)g_ctx.sm_count * 8;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

__global__ void k_set(double* x, double v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void k_scale(double* x, double s, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}
__global__ void k_mul(double* y, const double* x, const double* d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = x[i] * d[i];
}
__global__ void k_axpy(double* y, double a, const double* x, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}
__global__ void k_axpy_dev(double* y, const double* a, double sign, const double* x, size_t n) {
  const double s = sign * (*a);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] += s * x[i];
}
__global__ void k_filter_update(const double* vkm1, double* vout, const double* vk, const double* u, double* y,
                                double t, double cc, double mu, double mu0, int first, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = vk[i];
    double vn = t * (u[i] - cc * v);
    if (!first) vn -= vkm1[i];
    vout[i] = vn;
    y[i] = first ? (mu0 * v + mu * vn) : (y[i] + mu * vn);
  }
}

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
// Counter-based normal deviates keyed by the GLOBAL row id, so a start vector does not depend
// on how rows are distributed over GPUs.
__global__ void k_random(double* x, size_t n, unsigned long long seed, unsigned long long offset) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long g = offset + i;
    unsigned long long a = splitmix64(seed * 0x100000001B3ULL + 2 * g);
    unsigned long long b = splitmix64(seed * 0x100000001B3ULL + 2 * g + 1);
    double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    double u2 = (b >> 11) * (1.0 / 9007199254740992.0);
    x[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
}

// dot product: per-block partials in a fixed order, summed by the last block to finish
// (deterministic for a given grid).
__device__ unsigned int g_dot_ticket = 0;
__global__ void k_dot(const double* x, const double* y, size_t n, double* partial, double* out) {
  __shared__ double sm[NM_VEC_THREADS / 32];
  __shared__ bool last;
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += x[i] * y[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < NM_VEC_THREADS / 32; ++w) s += sm[w];
    partial[blockIdx.x] = s;
    __threadfence();
    unsigned int t = atomicAdd(&g_dot_ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    double s = 0.0;
    for (int i = threadIdx.x; i < gridDim.x; i += blockDim.x) s += ((volatile double*)partial)[i];
    // fixed-shape tree
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < NM_VEC_THREADS / 32; ++w) tot += sm[w];
      *out = tot;
      g_dot_ticket = 0;
    }
  }
}

#define LAUNCH_VEC(kern, n, ...)                                                     \
  do {                                                                               \
    nm_ensure_init();                                                                \
    if ((n) > 0) {                                                                   \
      kern<<<vec_grid(n), NM_VEC_THREADS, 0, g_ctx.stream>>>(__VA_ARGS__);           \
      g_ctx.launches++;                                                              \
    }                                                                                \
  } while (0)

void nm_vec_copy(double* dst, const double* src, size_t n) {
  if (n && dst != src)
    NM_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToDevice, g_ctx.stream));
}
void nm_vec_set(double* dst, double v, size_t n) { LAUNCH_VEC(k_set, n, dst, v, n); }
void nm_vec_scale(double* x, double s, size_t n) { LAUNCH_VEC(k_scale, n, x, s, n); }
void nm_vec_mul(double* y, const double* x, const double* d, size_t n) { LAUNCH_VEC(k_mul, n, y, x, d, n); }
void nm_vec_axpy(double* y, double a, const double* x, size_t n) { LAUNCH_VEC(k_axpy, n, y, a, x, n); }
void nm_vec_axpy_dev(double* y, const double* a_dev, double sign, const double* x, size_t n) {
  LAUNCH_VEC(k_axpy_dev, n, y, a_dev, sign, x, n);
}
void nm_filter_update(const double* vkm1, double* vout, const double* vk, const double* u, double* y, double t,
                      double cc, double mu, double mu0, int first, size_t n) {
  LAUNCH_VEC(k_filter_update, n, vkm1, vout, vk, u, y, t, cc, mu, mu0, first, n);
}
void nm_vec_random(double* x, size_t n, unsigned long long seed, unsigned long long offset) {
  LAUNCH_VEC(k_random, n, x, n, seed, offset);
}

void nm_vec_dot_dev(const double* x, const double* y, size_t n, double* out_dev) {
  nm_ensure_init();
  int grid = vec_grid(n);
  double* part = nm_red_scratch(grid);
  k_dot<<<grid, NM_VEC_THREADS, 0, g_ctx.stream>>>(x, y, n, part, out_dev);
  g_ctx.launches++;
  nm_allreduce_sum(out_dev, 1);
}

double nm_vec_dot(const double* x, const double* y, size_t n) {
  nm_ensure_init();
  double* scratch = nm_red_scratch(4096);
  double* out = scratch + 2048;           // partials use the front of the scratch
  int grid = vec_grid(n);
  if (grid > 2048) grid = 2048;
  k_dot<<<grid, NM_VEC_THREADS, 0, g_ctx.stream>>>(x, y, n, scratch, out);
  g_ctx.launches++;
  nm_allreduce_sum(out, 1);
  double h;
  NM_CUDA(cudaMemcpyAsync(&h, out, sizeof(double), cudaMemcpyDeviceToHost, g_ctx.stream));
  NM_CUDA(cudaStreamSynchronize(g_ctx.stream));
  return h;
}
