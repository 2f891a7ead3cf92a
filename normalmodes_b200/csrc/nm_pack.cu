// Host-side construction of the packed row-block format streamed by k_pack (nm_spmv.cuh).
//
// What it stores is exactly the matrix handed to pevsl_parcsrcreate_f90 (src/mod_matvec.f90:69-71): nothing
// is dropped or approximated, rows and columns keep the reference's numbering (vectors are never permuted).
// What changes is the ORDER in which rows are visited and the way a chunk of rows is laid out:
//   1. rows are ordered by a Cuthill-McKee sweep over the (block) pattern so that consecutive rows share
//      columns (finite-element neighbours); rectangular matrices keep their natural order; the order is then
//      stably partitioned into length classes so that every chunk holds rows of similar length;
//   2. the order is cut into chunks bounded by entries, rows (one epilogue thread per scalar row) and
//      DISTINCT columns (the x values a chunk needs are staged once in shared memory);
//   3. inside a chunk rows are sorted by length (stable) and entries stored in jagged-diagonal order:
//      entry k of the j-th longest row sits at off[k] + j, so lanes that walk neighbouring rows read
//      neighbouring shared-memory words, with no padding;
//   4. column ids become 16-bit indices into the chunk's distinct-column list.
// Per entry the stream is 8 (value) + 2 (index) bytes plus ~1 byte of lists, against 12 for CSR.
#include "nm_spmv.cuh"
#include <algorithm>
#include <numeric>
#include <queue>

int nm_env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && v[0]) ? atoi(v) : dflt;
}
static int env_int_(const char* name, int dflt) { return nm_env_int(name, dflt); }

// Cuthill-McKee order of the rows of a square pattern (columns >= n, i.e. ghosts, are ignored).
static void cm_order(int n, const std::vector<int>& rp, const std::vector<int>& idx, std::vector<int>& order) {
  order.clear();
  order.reserve(n);
  std::vector<char> seen(n, 0);
  std::vector<int> nb;
  // start each component from its lowest-degree unvisited row (scan in degree order)
  std::vector<int> bydeg(n);
  std::iota(bydeg.begin(), bydeg.end(), 0);
  std::stable_sort(bydeg.begin(), bydeg.end(), [&](int a, int b) { return rp[a + 1] - rp[a] < rp[b + 1] - rp[b]; });
  size_t head = 0;
  for (int seed : bydeg) {
    if (seen[seed]) continue;
    seen[seed] = 1;
    order.push_back(seed);
    while (head < order.size()) {
      const int u = order[head++];
      nb.clear();
      for (int p = rp[u]; p < rp[u + 1]; ++p) {
        const int v = idx[p];
        if (v < n && !seen[v]) { seen[v] = 1; nb.push_back(v); }
      }
      std::sort(nb.begin(), nb.end(), [&](int a, int b) {
        const int da = rp[a + 1] - rp[a], db = rp[b + 1] - rp[b];
        return da != db ? da < db : a < b;
      });
      order.insert(order.end(), nb.begin(), nb.end());
    }
  }
}

void nm_cm_order(int n, const std::vector<int>& rp, const std::vector<int>& idx, std::vector<int>& order) {
  cm_order(n, rp, idx, order);
}

bool nm_use_sell() { return env_int_("NM_KERNEL_SELL", 0) != 0; }

static inline size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// Pack of a matrix whose vectors stay in the caller's numbering.  Off by default: measured on B200 (tools/
// sweep_stream.py, profiles/) the plain global-memory kernels are as fast for one-off products in natural order
// (ROW3 A: 135 us vs 144-244 us; E / ET within 10%); the packed kernels pay off where the vectors can live in
// pack order, i.e. inside the Chebyshev iterations (NmChebIter builds its own permuted pack).
void nm_pack_build(NmParcsr& M, const std::vector<int>& rp, const std::vector<int>& idx, int n) {
  if (!env_int_("NM_NATURAL_PACK", 0)) return;
  if (nm_use_sell()) {
    nm_sell_build_into(M, M.sell, rp, idx, n, false);
    if (M.sell.nchunk > 0) M.fmt_bytes = M.sell.bytes;
    return;
  }
  nm_pack_build_into(M, M.pack, rp, idx, n, false);
  if (M.pack.nchunk > 0) M.fmt_bytes = M.pack.bytes;
}

// permuted = true (square matrices): the vectors the kernel touches live in PACK ORDER -- row j of chunk c is
// element first(c)+j and owned columns are renumbered the same way -- so epilogue accesses are contiguous and the
// x values a chunk gathers sit in a few contiguous runs.  P.order (pack position -> caller's index row) is kept on
// the device for the permute-in / permute-out kernels of the owner (NmChebIter keeps its work vectors that way).
void nm_pack_build_into(NmParcsr& M, NmPack& P, const std::vector<int>& rp, const std::vector<int>& idx, int n,
                        bool permuted) {
  P.nchunk = 0;
  P.permuted = false;
  if (n == 0 || env_int_("NM_NO_PACK", 0)) return;
  if (permuted && M.nrow != M.ncol) return;
  const int fmt = M.format;
  const int R = fmt == NM_FMT_CSR ? 1 : 3;
  const int VPE = fmt == NM_FMT_ROW3 ? 9 : 1;
  const int ecap = std::max(32, env_int_("NM_PACK_ENTRIES", fmt == NM_FMT_ROW3 ? 192 : 1536));
  const int rcap = std::min(NM_SPMV_THREADS / R, std::max(1, env_int_("NM_PACK_ROWS", NM_SPMV_THREADS)));
  const int dcap = std::max(16, std::min(4096, env_int_("NM_PACK_DISTINCT", fmt == NM_FMT_ROW3 ? 160 : 448)));
  const int ncolb = (M.ncol + M.halo.nghost + R - 1) / R;                   // column ids are < ncolb
  // ---- 1. row order
  std::vector<int> order;
  const bool square = (M.nrow == M.ncol);
  if (square && env_int_("NM_PACK_ORDER", 1)) cm_order(n, rp, idx, order);
  else { order.resize(n); std::iota(order.begin(), order.end(), 0); }
  // rows of similar length share chunks (length classes by powers of two, locality order kept inside a class): a
  // chunk's walk takes max(len)/L steps, so one long vertex-node row among short edge-node rows would leave all
  // but one warp idle; with classes the long rows get L = 4..16 lanes each and every chunk is balanced
  if (env_int_("NM_PACK_CLASSES", 1)) {
    auto cls = [&](int row) { int len = rp[row + 1] - rp[row], c = 0; while ((1 << c) <= len) ++c; return c; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cls(a) < cls(b); });
  }
  // ---- 2. chunks
  struct Chunk { int first, nr, ne, nd, maxlen; };
  std::vector<Chunk> chunks;
  std::vector<int> stamp(ncolb, -1);
  {
    int r = 0;
    while (r < n) {
      Chunk c{r, 0, 0, 0, 0};
      const int id = (int)chunks.size();
      while (r + c.nr < n && c.nr < rcap) {
        const int row = order[r + c.nr];
        const int len = rp[row + 1] - rp[row];
        if (c.ne + len > ecap) break;
        int fresh = 0;
        for (int p = rp[row]; p < rp[row + 1]; ++p) if (stamp[idx[p]] != id) ++fresh;
        if (c.nd + fresh > dcap) break;
        for (int p = rp[row]; p < rp[row + 1]; ++p) stamp[idx[p]] = id;
        c.nd += fresh; c.ne += len; c.nr++;
        c.maxlen = std::max(c.maxlen, len);
      }
      if (c.nr == 0) return;                      // one row alone exceeds a chunk: not packable, fallback kernels
      if (c.maxlen > 65535 || c.ne > 65535) return;
      chunks.push_back(c);
      r += c.nr;
    }
  }
  // ---- 2b. final row order: chunk after chunk, longest row first inside a chunk (stable)
  std::vector<int> final_order(n), newid;
  for (const Chunk& c : chunks) {
    std::copy(order.begin() + c.first, order.begin() + c.first + c.nr, final_order.begin() + c.first);
    std::stable_sort(final_order.begin() + c.first, final_order.begin() + c.first + c.nr,
                     [&](int a, int b) { return rp[a + 1] - rp[a] > rp[b + 1] - rp[b]; });
  }
  if (permuted) {
    newid.resize(n);
    for (int i = 0; i < n; ++i) newid[final_order[i]] = i;
  }
  auto colid = [&](int c) { return (permuted && c < n) ? newid[c] : c; };     // ghosts (>= n) keep their id
  // ---- 3. blobs
  std::vector<NmPackDesc> desc(chunks.size());
  size_t total = 0, max_blob = 0;
  int max_nd = 0;
  std::vector<size_t> start(chunks.size());
  for (size_t i = 0; i < chunks.size(); ++i) {
    const Chunk& c = chunks[i];
    size_t b = 16 + (size_t)8 * VPE * c.ne + 4 * (size_t)c.nr + 4 * (size_t)c.nd;
    b = (b + 7) & ~(size_t)7;
    b += 2 * (size_t)(c.maxlen + 1) + 2 * (size_t)c.nr + 2 * (size_t)c.ne;
    b = up16(b);
    start[i] = total;
    desc[i].off16 = (unsigned)(total / 16);
    desc[i].bytes = (unsigned)b;
    total += b;
    max_blob = std::max(max_blob, b);
    max_nd = std::max(max_nd, c.nd);
  }
  NM_REQUIRE(total / 16 < 0xffffffffull, "pack: matrix too large for 32-bit blob offsets");
  std::vector<unsigned char> blob(total, 0);
  std::vector<unsigned> slot_off8;
  std::vector<int> slot_src;
  slot_off8.reserve((size_t)rp[n] * VPE);
  slot_src.reserve((size_t)rp[n] * VPE);
  std::vector<int> rows_sorted, cols, lidx_of(ncolb, -1);
  std::vector<std::vector<int>> rem;
  const bool bank_aware = env_int_("NM_PACK_BANK_AWARE", 1) != 0;
  for (size_t i = 0; i < chunks.size(); ++i) {
    const Chunk& c = chunks[i];
    unsigned char* base = blob.data() + start[i];
    // lanes per row: fill the CTA.  ROW3: a lane per (row, column component); CSR/KRON3: a lane per index row
    const int L = std::max(1, NM_SPMV_THREADS / ((fmt == NM_FMT_ROW3 ? 3 : 1) * c.nr));
    NmPackHeader h{c.nr, c.nd, c.ne, c.maxlen | (L << 16)};
    memcpy(base, &h, sizeof(h));
    const size_t o_val = 16;
    const size_t o_rows = o_val + (size_t)8 * VPE * c.ne;
    const size_t o_cols = o_rows + 4 * (size_t)c.nr;
    const size_t o_off = (o_cols + 4 * (size_t)c.nd + 7) & ~(size_t)7;
    const size_t o_len = o_off + 2 * (size_t)(c.maxlen + 1);
    const size_t o_lidx = o_len + 2 * (size_t)c.nr;
    int* brows = (int*)(base + o_rows);
    int* bcols = (int*)(base + o_cols);
    unsigned short* boff = (unsigned short*)(base + o_off);
    unsigned short* blen = (unsigned short*)(base + o_len);
    unsigned short* blidx = (unsigned short*)(base + o_lidx);
    // rows of the chunk, longest first
    rows_sorted.assign(final_order.begin() + c.first, final_order.begin() + c.first + c.nr);
    // distinct columns, ascending in the numbering the vectors use (neighbouring ids share cache lines)
    cols.clear();
    for (int row : rows_sorted)
      for (int p = rp[row]; p < rp[row + 1]; ++p)
        if (lidx_of[idx[p]] < 0) { lidx_of[idx[p]] = 0; cols.push_back(idx[p]); }
    std::sort(cols.begin(), cols.end(), [&](int a, int b) { return colid(a) < colid(b); });
    NM_REQUIRE((int)cols.size() == c.nd, "pack: distinct-column count mismatch");
    for (int j = 0; j < c.nd; ++j) { lidx_of[cols[j]] = j; bcols[j] = colid(cols[j]); }
    for (int j = 0; j < c.nr; ++j) {
      brows[j] = permuted ? c.first + j : rows_sorted[j];
      blen[j] = (unsigned short)(rp[rows_sorted[j] + 1] - rp[rows_sorted[j]]);
    }
    // JDS offsets: column k holds the rows with len > k (a prefix of rows_sorted)
    int pos = 0;
    for (int k = 0; k <= c.maxlen; ++k) {
      boff[k] = (unsigned short)pos;
      int cnt = 0;
      while (cnt < c.nr && blen[cnt] > k) ++cnt;
      pos += cnt;
    }
    NM_REQUIRE(pos == c.ne, "pack: JDS offsets do not add up");
    const size_t val8 = (start[i] + o_val) / 8;                   // blob position of the value region in doubles
    if (VPE == 1) {
      // Entry order inside a row is free (it only reassociates the row sum).  Pick it so that the 16 lanes of a
      // half-warp, which at loop step i read entry k = l + i*L of their rows, hit 16 different shared-memory
      // bank pairs when they gather x (bank pair = chunk-local column index mod 16): greedy, deterministic.
      rem.resize(c.nr);
      for (int j = 0; j < c.nr; ++j) {
        const int row = rows_sorted[j];
        rem[j].clear();
        for (int p = rp[row + 1] - 1; p >= rp[row]; --p) rem[j].push_back(p);    // back() = first in CSR order
      }
      const int nthr = c.nr * L;
      const int steps = (c.maxlen + L - 1) / L;
      for (int st = 0; st < steps; ++st)
        for (int h0 = 0; h0 < nthr; h0 += 16) {
          unsigned taken = 0;
          for (int t = h0; t < std::min(h0 + 16, nthr); ++t) {
            const int l = t / c.nr, j = t - l * c.nr, k = l + st * L;
            if (k >= (int)blen[j]) continue;
            std::vector<int>& rj = rem[j];
            int pick = (int)rj.size() - 1;
            if (bank_aware)
              for (int q = (int)rj.size() - 1; q >= 0; --q)
                if (!(taken & (1u << (lidx_of[idx[rj[q]]] & 15)))) { pick = q; break; }
            const int src = rj[pick];
            rj.erase(rj.begin() + pick);
            const int li = lidx_of[idx[src]];
            taken |= 1u << (li & 15);
            const int p = boff[k] + j;
            blidx[p] = (unsigned short)li;
            slot_off8.push_back((unsigned)(val8 + p));
            slot_src.push_back(src);
          }
        }
    } else {
      for (int j = 0; j < c.nr; ++j) {
        const int row = rows_sorted[j];
        const int s = rp[row], len = rp[row + 1] - s;
        for (int k = 0; k < len; ++k) {
          const int p = boff[k] + j;
          blidx[p] = (unsigned short)lidx_of[idx[s + k]];
          // ROW3: plane i (row 3*row+i of the node) holds [p][c]; source a[9*s + i*3*len + 3*k + c]
          for (int pi = 0; pi < 3; ++pi)
            for (int cc = 0; cc < 3; ++cc) {
              slot_off8.push_back((unsigned)(val8 + ((size_t)pi * c.ne + p) * 3 + cc));
              slot_src.push_back(9 * s + pi * 3 * len + 3 * k + cc);
            }
        }
      }
    }
    for (int j = 0; j < c.nd; ++j) lidx_of[cols[j]] = -1;
  }
  NM_REQUIRE((total + 7) / 8 < 0xffffffffull, "pack: matrix too large for 32-bit value offsets");
  // ---- 4. launch geometry
  P.xs_doubles = R * max_nd;
  P.stage_bytes = (int)up16(max_blob);
  P.nstage = std::max(2, std::min(8, env_int_("NM_PACK_STAGES", 2)));
  const int fixed = NM_PACK_MAXDESC * (int)sizeof(NmPackDesc) + 64 + 16 * P.xs_doubles + 16 * 3 * NM_SPMV_THREADS;
  P.smem_bytes = (int)up16(fixed) + P.nstage * P.stage_bytes;
  if (P.smem_bytes > 200 * 1024) return;
  int ctas = nm_ctx().sm_count * std::max(1, env_int_("NM_PACK_CTAS_PER_SM", 3));
  ctas = std::max(1, std::min(ctas, env_int_("NM_PACK_MAXGRID", ctas)));     // tests: force several chunks per CTA
  const int nchunk = (int)chunks.size();
  P.grid = std::min(nchunk, ctas);
  P.chunks_per_cta = nm_div_up(nchunk, P.grid);
  if (P.chunks_per_cta > NM_PACK_MAXDESC) P.chunks_per_cta = NM_PACK_MAXDESC;
  P.grid = nm_div_up(nchunk, P.chunks_per_cta);
  P.bytes = (long long)total;
  P.nslot = (long long)slot_src.size();
  P.blob.alloc(total); P.blob.upload(blob.data(), total);
  P.desc.from_host(desc);
  P.slot_off8.from_host(slot_off8);
  P.slot_src.from_host(slot_src);
  P.nchunk = nchunk;
  P.permuted = permuted;
  if (permuted) P.order.from_host(final_order);
  nm_pack_fill_from(M, P);
}

__global__ void k_pack_fill(long long nslot, const unsigned* __restrict__ off8, const int* __restrict__ src,
                            const double* __restrict__ vals, double* __restrict__ blob) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nslot) blob[off8[i]] = vals[src[i]];
}

void nm_pack_fill(NmParcsr& M) {
  M.values_version++;
  nm_pack_fill_from(M, M.pack);
  nm_sell_fill_from(M, M.sell);
}

void nm_pack_fill_from(NmParcsr& M, NmPack& P) {
  if (P.nchunk == 0 || P.nslot == 0) return;
  NmCtx& c = nm_ctx();
  const double* vals = M.format == NM_FMT_KRON3 ? M.mval.p : M.a.p;
  k_pack_fill<<<nm_div_up(P.nslot, 256), 256, 0, c.stream>>>(P.nslot, P.slot_off8.p, P.slot_src.p, vals, (double*)P.blob.p);
  c.launches++;
}

void nm_pack_clone(const NmParcsr& src, NmParcsr& dst) {
  nm_sell_clone(src, dst);
  if (dst.sell.nchunk > 0) dst.fmt_bytes = dst.sell.bytes;
  const NmPack& S = src.pack;
  NmPack& D = dst.pack;
  D.nchunk = 0;
  if (S.nchunk == 0) return;
  NmCtx& c = nm_ctx();
  auto cp = [&](auto& d, const auto& s) {
    d.alloc(s.n);
    NM_CUDA(cudaMemcpyAsync(d.p, s.p, s.n * sizeof(*s.p), cudaMemcpyDeviceToDevice, c.stream));
  };
  cp(D.blob, S.blob); cp(D.desc, S.desc); cp(D.slot_off8, S.slot_off8); cp(D.slot_src, S.slot_src);
  D.nslot = S.nslot; D.chunks_per_cta = S.chunks_per_cta; D.grid = S.grid; D.stage_bytes = S.stage_bytes;
  D.xs_doubles = S.xs_doubles; D.nstage = S.nstage; D.smem_bytes = S.smem_bytes; D.bytes = S.bytes;
  D.nchunk = S.nchunk;
  dst.fmt_bytes = D.bytes;
  nm_pack_fill_from(dst, D);
}

// ================================================================ sliced JDS (k_sell)
void nm_sell_build_into(NmParcsr& M, NmSell& S, const std::vector<int>& rp, const std::vector<int>& idx, int n,
                        bool permuted) {
  S.nchunk = 0;
  S.permuted = false;
  if (n == 0 || env_int_("NM_NO_SELL", 0)) return;
  if (permuted && M.nrow != M.ncol) return;
  const int fmt = M.format;
  const int R = fmt == NM_FMT_CSR ? 1 : 3;
  const int VPE = fmt == NM_FMT_ROW3 ? 9 : 1;
  const int target = std::max(1, env_int_("NM_SELL_TARGET", fmt == NM_FMT_ROW3 ? 12 : 24));   // entries per lane
  std::vector<int> order;
  if (M.nrow == M.ncol && env_int_("NM_PACK_ORDER", 1)) cm_order(n, rp, idx, order);
  else { order.resize(n); std::iota(order.begin(), order.end(), 0); }
  auto cls = [&](int row) { int len = rp[row + 1] - rp[row], c = 0; while ((1 << c) <= len) ++c; return c; };
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cls(a) < cls(b); });
  // ---- slices: rows of one length class, L lanes per row
  std::vector<NmSellChunk> chunks;
  std::vector<int> final_order(n);
  long long e0 = 0;
  int o0 = 0;
  for (int r = 0; r < n;) {
    const int c0 = cls(order[r]);
    const int len0 = rp[order[r] + 1] - rp[order[r]];
    int L = 1;
    while (L < 32 && len0 >= 2 * L * target) L *= 2;
    const int cap = std::max(1, NM_SPMV_THREADS / (R * L));
    int k = 0;
    while (r + k < n && k < cap && cls(order[r + k]) == c0) ++k;
    std::copy(order.begin() + r, order.begin() + r + k, final_order.begin() + r);
    std::stable_sort(final_order.begin() + r, final_order.begin() + r + k,
                     [&](int a, int b) { return rp[a + 1] - rp[a] > rp[b + 1] - rp[b]; });
    NmSellChunk c;
    c.e0 = e0; c.o0 = o0; c.r0 = r; c.nr = k; c.L = L; c.pad = 0;
    c.maxlen = rp[final_order[r] + 1] - rp[final_order[r]];
    long long ne = 0;
    for (int j = 0; j < k; ++j) ne += rp[final_order[r + j] + 1] - rp[final_order[r + j]];
    e0 += ne; o0 += c.maxlen + 1;
    chunks.push_back(c);
    r += k;
  }
  NM_REQUIRE(e0 == rp[n], "sell: entry count mismatch");
  std::vector<int> newid;
  if (permuted) {
    newid.resize(n);
    for (int i = 0; i < n; ++i) newid[final_order[i]] = i;
  }
  auto colid = [&](int c) { return (permuted && c < n) ? newid[c] : c; };
  std::vector<int> col((size_t)e0), off((size_t)o0), rows(n), rowlen(n), src((size_t)e0 * VPE);
  for (const NmSellChunk& c : chunks) {
    int pos = 0;
    for (int k = 0; k <= c.maxlen; ++k) {
      off[c.o0 + k] = pos;
      int cnt = 0;
      while (cnt < c.nr && rp[final_order[c.r0 + cnt] + 1] - rp[final_order[c.r0 + cnt]] > k) ++cnt;
      pos += cnt;
    }
    for (int j = 0; j < c.nr; ++j) {
      const int row = final_order[c.r0 + j];
      const int s = rp[row], len = rp[row + 1] - s;
      rows[c.r0 + j] = permuted ? c.r0 + j : row;
      rowlen[c.r0 + j] = len;
      for (int k = 0; k < len; ++k) {
        const long long p = c.e0 + off[c.o0 + k] + j;
        col[p] = colid(idx[s + k]);
        if (VPE == 1) src[p] = s + k;
        else
          for (int pi = 0; pi < 3; ++pi)
            for (int cc = 0; cc < 3; ++cc) src[p * 9 + pi * 3 + cc] = 9 * s + pi * 3 * len + 3 * k + cc;
      }
    }
  }
  S.val.alloc((size_t)e0 * VPE);
  S.col.from_host(col); S.off.from_host(off); S.rows.from_host(rows); S.rowlen.from_host(rowlen);
  S.chunks.from_host(chunks); S.slot_src.from_host(src);
  S.nslot = (long long)src.size();
  S.bytes = (long long)e0 * (8 * VPE + 4) + 4ll * o0 + 8ll * n + (long long)chunks.size() * sizeof(NmSellChunk);
  S.nchunk = (int)chunks.size();
  S.permuted = permuted;
  if (permuted) S.order.from_host(final_order);
  nm_sell_fill_from(M, S);
}

__global__ void k_sell_fill(long long nslot, const int* __restrict__ src, const double* __restrict__ vals,
                            double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nslot) out[i] = vals[src[i]];
}

void nm_sell_fill_from(NmParcsr& M, NmSell& S) {
  if (S.nchunk == 0 || S.nslot == 0) return;
  NmCtx& c = nm_ctx();
  const double* vals = M.format == NM_FMT_KRON3 ? M.mval.p : M.a.p;
  k_sell_fill<<<nm_div_up(S.nslot, 256), 256, 0, c.stream>>>(S.nslot, S.slot_src.p, vals, S.val.p);
  c.launches++;
}

void nm_sell_clone(const NmParcsr& src, NmParcsr& dst) {
  const NmSell& S = src.sell;
  NmSell& D = dst.sell;
  D.nchunk = 0;
  if (S.nchunk == 0) return;
  NmCtx& c = nm_ctx();
  auto cp = [&](auto& d, const auto& s) {
    d.alloc(s.n);
    NM_CUDA(cudaMemcpyAsync(d.p, s.p, s.n * sizeof(*s.p), cudaMemcpyDeviceToDevice, c.stream));
  };
  D.val.alloc(S.val.n);
  cp(D.col, S.col); cp(D.off, S.off); cp(D.rows, S.rows); cp(D.rowlen, S.rowlen); cp(D.chunks, S.chunks);
  cp(D.slot_src, S.slot_src);
  D.nslot = S.nslot; D.bytes = S.bytes; D.nchunk = S.nchunk;
  nm_sell_fill_from(dst, D);
}
