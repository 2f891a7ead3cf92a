// Host-side construction of the warp-sliced ELL slabs streamed by k_slab (nm_slab.cuh).
//
// Stores exactly the matrix handed to pevsl_parcsrcreate_f90 (src/mod_matvec.f90:69-71,146-148) after the Jacobi
// scaling of Bdiagscaling / Apdiagscaling (:252-441): nothing dropped or approximated.  Only the ORDER changes:
//   1. index rows (nodes for B = M (x) I3, scalar rows for Ap~) are ordered by a Cuthill-McKee sweep so that
//      consecutive rows share columns; the iteration's vectors live in that order (NmChebIter permutes b in and
//      x out once per solve);
//   2. the order is cut into chunks of at most T rows (T = threads per CTA), bounded by padded entries (stage
//      size) and DISTINCT columns (the x values a chunk needs are staged once in shared memory);
//   3. inside a chunk rows are sorted by length (stable) and grouped into slices of 32; a slice is padded to its
//      longest row and stored step-major: entry k of lane j at eoff + 32 k + j (conflict-free LDS);
//   4. column ids become 16-bit indices into the chunk's ascending distinct-column list;
//   5. the order of the entries INSIDE a row (free: it only reassociates the row sum, deterministically) is picked
//      so that the 16 lanes of a half-warp gather x from 16 different shared-memory bank pairs.
// Per entry the stream is 8 (value) + 2 (index) bytes plus ~0.8 byte of column list, against 12 for CSR.
#include "nm_slab.cuh"
#include <algorithm>
#include <numeric>

static inline size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// Everything the device needs, built on the host (no CUDA calls: also driven by the CPU tests through
// nm_slab_host_selftest).
struct NmSlabHost {
  std::vector<unsigned char> blob;
  std::vector<NmPackDesc> desc;
  std::vector<int> cta_first, slot_src, order;
  std::vector<unsigned> slot_off8;
  int nchunk = 0, grid = 0, threads = 0, max_chunks_per_cta = 0, stage_bytes = 0, xs_doubles = 0, nstage = 0, smem_bytes = 0;
  int ws = 0, nxs = 2, nprod = 0;
  long long padded_entries = 0;
  std::vector<int> desc_cid;               // chunk id (pack-order index) of every descriptor (the CTA ranges are reordered)
  bool deps_ok = false;                    // every blob carries the ids of the chunks that own its columns, symmetric
};

// rp/idx: row pointers and column ids (< ncolb) of the n index rows; R scalar rows/columns per index entry;
// sm_count: SMs of the device the launch geometry is made for.  Returns false when the matrix is not packable.
static bool slab_build_host(NmSlabHost& H, const std::vector<int>& rp, const std::vector<int>& idx, int n, int R,
                            int ncolb, int sm_count) {
  H.nchunk = 0;
  // defaults from the B200 sweep (tools/sweep_slab.py, profiles/r1d_sweep_slab.json): 256 lanes per chunk (8 consumer
  // warps), <= 12 entries per lane, 2 stages + 2 x buffers (2 CTAs per SM), 4 producer warps
  int T = nm_env_int("NM_SLAB_THREADS", 256);
  if (T != 64 && T != 128 && T != 256 && T != 512) T = 256;
  const int NW = T / 32;
  const int lcap = std::max(2, nm_env_int("NM_SLAB_SPLIT", 12));               // entries one lane walks at most
  const int ecap = std::max(32, nm_env_int("NM_SLAB_ENTRIES", 3584));          // (estimated) padded entries per chunk
  const int dcap = std::max(16, std::min(16384, nm_env_int("NM_SLAB_DISTINCT", R == 3 ? 640 : 1536)));
  auto len_of = [&](int row) { return rp[row + 1] - rp[row]; };
  // lanes a row is shared by (1..32) and the entries each of them walks
  auto lanes_of = [&](int len) { return std::max(1, (len + lcap - 1) / lcap); };
  auto vlen_of = [&](int row) { const int len = len_of(row), g = lanes_of(len); return (len + g - 1) / g; };
  // ---- 1+2. row order and chunks in one sweep: a chunk is grown breadth-first from a seed over the rows not
  // yet placed (a compact blob of mesh neighbours: its rows share most of their columns, so the DISTINCT columns
  // staged per chunk stay few), until its warps are full or the entry / distinct-column cap is reached; the next
  // seed comes from the frontier left behind, so consecutive chunks are neighbours as well.  A row's lanes must sit
  // in ONE warp (they are combined by shuffles): rows are placed first-fit into the chunk's T/32 warps, preferring
  // a warp whose lanes walk the same number of entries (a warp is padded to its longest lane).
  struct Chunk { int first, nr, nd, nslice, lane_off; };
  std::vector<Chunk> chunks;
  std::vector<int> order;                   // final row order: chunk after chunk, warp after warp
  std::vector<short> lane_row_all;          // per chunk 32*nslice lanes: local row (position in the chunk) or -1
  order.reserve(n);
  {
    const bool grow = nm_env_int("NM_PACK_ORDER", 1) != 0;
    std::vector<int> stamp(ncolb, -1);
    std::vector<char> placed(n, 0), queued(n, 0);
    std::vector<int> frontier, q;
    std::vector<std::vector<int>> bin_rows(NW);
    std::vector<int> bin_free(NW), bin_vlen(NW);
    size_t fhead = 0;
    int scan = 0, id = 0;
    int nplaced = 0;
    while (nplaced < n) {
      Chunk c{nplaced, 0, 0, 0, (int)lane_row_all.size()};
      int ne = 0;
      for (int w = 0; w < NW; ++w) { bin_rows[w].clear(); bin_free[w] = 32; bin_vlen[w] = 0; }
      bool full = false;
      q.clear();
      size_t qhead = 0;
      while (!full) {
        if (qhead == q.size()) {                    // (re)seed: frontier first, then the lowest unplaced row
          int seed = -1;
          while (grow && fhead < frontier.size()) {
            const int v = frontier[fhead++];
            if (!placed[v] && !queued[v]) { seed = v; break; }
          }
          if (seed < 0) {
            while (scan < n && (placed[scan] || queued[scan])) ++scan;
            if (scan == n) break;
            seed = scan;
          }
          queued[seed] = 1;
          q.push_back(seed);
        }
        const int row = q[qhead];
        const int len = len_of(row), g = lanes_of(len), v = vlen_of(row);
        if (g > 32) return false;                     // a row longer than a warp can share: not packable
        int fresh = 0;
        for (int p = rp[row]; p < rp[row + 1]; ++p) if (stamp[idx[p]] != id) ++fresh;
        if (fresh > dcap) return false;
        // warp for the row: same lane length > empty warp > any warp with room
        int best = -1, score = -1;
        for (int w = 0; w < NW; ++w) {
          if (bin_free[w] < g) continue;
          const int sc = bin_vlen[w] == v ? 3 : (bin_free[w] == 32 ? 2 : 1);
          if (sc > score) { score = sc; best = w; }
        }
        const int est = g * v;
        if (best < 0 || (c.nr > 0 && (ne + est > ecap || c.nd + fresh > dcap))) { full = true; break; }
        ++qhead;
        for (int p = rp[row]; p < rp[row + 1]; ++p) stamp[idx[p]] = id;
        c.nd += fresh; ne += est; c.nr++;
        bin_rows[best].push_back(row); bin_free[best] -= g; bin_vlen[best] = std::max(bin_vlen[best], v);
        placed[row] = 1;
        ++nplaced;
        if (grow)
          for (int p = rp[row]; p < rp[row + 1]; ++p) {
            const int u = idx[p];
            if (u < n && !placed[u] && !queued[u]) { queued[u] = 1; q.push_back(u); }
          }
      }
      for (size_t i = qhead; i < q.size(); ++i) { queued[q[i]] = 0; frontier.push_back(q[i]); }
      if (c.nr == 0) return false;
      if (c.nd > 65535) return false;
      ++id;
      // Optional (NM_SLAB_REPACK=1): re-pack the chunk's rows into its warps by decreasing lane length (first-fit
      // decreasing) so that a warp holds lanes of equal length.  Off by default: on the bench workload it trades
      // padding 1.096 -> 1.073 for modelled x-read wavefronts 1.27 -> 1.31 (warps no longer hold mesh neighbours that
      // read the same columns), a wash on the load/store pipe.
      if (nm_env_int("NM_SLAB_REPACK", 0)) {
        std::vector<int> all;
        for (int w = 0; w < NW; ++w) all.insert(all.end(), bin_rows[w].begin(), bin_rows[w].end());
        std::stable_sort(all.begin(), all.end(), [&](int a, int b) {
          const int va = vlen_of(a), vb = vlen_of(b);
          if (va != vb) return va > vb;
          return lanes_of(len_of(a)) > lanes_of(len_of(b));
        });
        std::vector<std::vector<int>> nb(NW);
        std::vector<int> nfree(NW, 32), nvl(NW, 0);
        bool ok = true;
        for (int row : all) {
          const int g = lanes_of(len_of(row));
          int best = -1;
          for (int w = 0; w < NW; ++w) if (nfree[w] >= g) { best = w; break; }
          if (best < 0) { ok = false; break; }
          nb[best].push_back(row); nfree[best] -= g; nvl[best] = std::max(nvl[best], vlen_of(row));
        }
        auto cost = [&](const std::vector<int>& vl, const std::vector<std::vector<int>>& rowsv) {
          int cst = 0;
          for (int w = 0; w < NW; ++w) if (!rowsv[w].empty()) cst += 32 * vl[w];
          return cst;
        };
        if (ok && cost(nvl, nb) < cost(bin_vlen, bin_rows)) { bin_rows = nb; bin_vlen = nvl; }
      }
      // lanes: non-empty warps, longest lanes first (stable); inside a warp rows in placement order
      std::vector<int> ws;
      for (int w = 0; w < NW; ++w) if (!bin_rows[w].empty()) ws.push_back(w);
      std::stable_sort(ws.begin(), ws.end(), [&](int a, int b) { return bin_vlen[a] > bin_vlen[b]; });
      c.nslice = (int)ws.size();
      int local = 0;
      for (int w : ws) {
        int used = 0;
        for (int row : bin_rows[w]) {
          const int g = lanes_of(len_of(row));
          for (int l = 0; l < g; ++l) lane_row_all.push_back((short)local);
          used += g;
          order.push_back(row);
          ++local;
        }
        for (; used < 32; ++used) lane_row_all.push_back((short)-1);
      }
      chunks.push_back(c);
    }
  }
  std::vector<int>& final_order = order;
  std::vector<int> newid(n);
  for (int i = 0; i < n; ++i) newid[final_order[i]] = i;
  auto colid = [&](int c) { return c < n ? newid[c] : c; };                  // ghosts (>= n) keep their id
  // ---- 2b. chunk dependencies (dataflow execution of the persistent kernel): the chunks that own the columns a chunk
  // gathers.  A Chebyshev step of chunk c may start as soon as THOSE chunks have finished the previous step; for a
  // symmetric pattern the relation is symmetric, which is what makes the two alternating direction buffers safe
  // without a grid barrier (a chunk cannot overwrite values a neighbour still has to gather).
  const int nchunk = (int)chunks.size();
  std::vector<std::vector<int>> deps(nchunk);
  {
    std::vector<int> chunk_of(n);
    for (int i = 0; i < nchunk; ++i)
      for (int j = 0; j < chunks[i].nr; ++j) chunk_of[chunks[i].first + j] = i;
    std::vector<int> mark(nchunk, -1);
    for (int i = 0; i < nchunk; ++i) {
      for (int j = 0; j < chunks[i].nr; ++j) {
        const int row = final_order[chunks[i].first + j];
        for (int p = rp[row]; p < rp[row + 1]; ++p) {
          if (idx[p] >= n) continue;                                        // ghost column: flag-in-data slot
          const int c2 = chunk_of[newid[idx[p]]];
          if (mark[c2] != i) { mark[c2] = i; deps[i].push_back(c2); }
        }
      }
      std::sort(deps[i].begin(), deps[i].end());
    }
    bool sym = true;
    for (int i = 0; i < nchunk && sym; ++i)
      for (int c2 : deps[i])
        if (!std::binary_search(deps[c2].begin(), deps[c2].end(), i)) { sym = false; break; }
    H.deps_ok = sym && nm_env_int("NM_SLAB_DEPS", 1) != 0;
    if (!H.deps_ok) for (auto& d : deps) d.clear();
  }
  // ---- 3. blobs.  Lane t of a chunk walks one share of a row; slice s = lanes 32s..32s+31 (one warp), padded
  // to its longest lane.
  std::vector<NmPackDesc> desc(nchunk);
  std::vector<size_t> start(nchunk);
  std::vector<int> nep_of(nchunk);
  size_t total = 0, max_blob = 0;
  int max_nd = 0;
  long long pentries = 0;
  for (int i = 0; i < nchunk; ++i) {
    const Chunk& c = chunks[i];
    const int* rows = final_order.data() + c.first;
    const short* lr = lane_row_all.data() + c.lane_off;
    int nep = 0;
    for (int s2 = 0; s2 < c.nslice; ++s2) {
      int w = 0;
      for (int l = 0; l < 32; ++l) if (lr[32 * s2 + l] >= 0) w = std::max(w, vlen_of(rows[lr[32 * s2 + l]]));
      nep += 32 * w;
    }
    nep_of[i] = nep;
    size_t b = 32 + up16(8 * (size_t)c.nslice) + 8 * (size_t)nep + 4 * (size_t)c.nd + 2 * (size_t)nep + 2 * (size_t)(32 * c.nslice) +
               4 * deps[i].size();
    b = up16(b);
    start[i] = total;
    desc[i].off16 = (unsigned)(total / 16);
    desc[i].bytes = (unsigned)b;
    total += b;
    max_blob = std::max(max_blob, b);
    max_nd = std::max(max_nd, c.nd);
    pentries += nep;
  }
  NM_REQUIRE(total / 16 < 0xffffffffull && (total + 7) / 8 < 0xffffffffull, "slab: matrix too large for 32-bit offsets");
  std::vector<unsigned char>& blob = H.blob;
  blob.assign(total, 0);
  std::vector<unsigned>& slot_off8 = H.slot_off8;
  std::vector<int>& slot_src = H.slot_src;
  slot_off8.clear(); slot_src.clear();
  slot_off8.reserve((size_t)rp[n]);
  slot_src.reserve((size_t)rp[n]);
  std::vector<int> cols, lidx_of(ncolb, -1);
  std::vector<char> ghost_chunk(nchunk, 0);
  std::vector<std::vector<int>> rem;
  const int bank_mode = std::max(0, std::min(3, nm_env_int("NM_PACK_BANK_AWARE", 3)));
  for (int i = 0; i < nchunk; ++i) {
    const Chunk& c = chunks[i];
    unsigned char* base = blob.data() + start[i];
    const int nslice = c.nslice, nep = nep_of[i];
    const size_t nslot_before = slot_src.size();
    size_t chunk_entries = 0;
    for (int j = 0; j < c.nr; ++j) chunk_entries += (size_t)len_of(final_order[c.first + j]);
    const int* rows = final_order.data() + c.first;
    const short* lane_row = lane_row_all.data() + c.lane_off;
    int gmax = 1;
    for (int j = 0; j < c.nr; ++j) gmax = std::max(gmax, lanes_of(len_of(rows[j])));
    int has_ghost = 0;                                            // any column owned by another rank (id >= n)
    for (int j = 0; j < c.nr && !has_ghost; ++j)
      for (int p = rp[rows[j]]; p < rp[rows[j] + 1]; ++p)
        if (idx[p] >= n) { has_ghost = 1; break; }
    ghost_chunk[i] = (char)has_ghost;
    NmSlabHeader h{c.nr, c.nd, nslice, c.first, nep, gmax, has_ghost, (int)deps[i].size()};
    memcpy(base, &h, sizeof(h));
    const size_t o_tbl = 32;
    const size_t o_val = 32 + up16(8 * (size_t)nslice);
    const size_t o_cols = o_val + 8 * (size_t)nep;
    const size_t o_idx = o_cols + 4 * (size_t)c.nd;
    const size_t o_lane = o_idx + 2 * (size_t)nep;
    unsigned* tbl = (unsigned*)(base + o_tbl);
    int* bcols = (int*)(base + o_cols);
    unsigned short* bidx = (unsigned short*)(base + o_idx);
    unsigned short* blane = (unsigned short*)(base + o_lane);
    if (!deps[i].empty()) memcpy(base + o_lane + 2 * (size_t)(32 * nslice), deps[i].data(), 4 * deps[i].size());
    // per-lane word of the kernel: bit 15 = first lane of its row (does the epilogue), bits 10..14 = steps s of
    // the shuffle tree at which lane + 2^s belongs to the same row (add its partial), bits 0..9 = local row
    for (int t = 0; t < 32 * nslice; ++t) {
      const int j = lane_row[t];
      if (j < 0) { blane[t] = 0; continue; }
      const int lane = t & 31;
      unsigned wd = (unsigned)j & 0x3ffu;
      if (lane == 0 || lane_row[t - 1] != j) wd |= 0x8000u;
      for (int s2 = 0; s2 < 5; ++s2)
        if (lane + (1 << s2) < 32 && lane_row[t + (1 << s2)] == j) wd |= 1u << (10 + s2);
      blane[t] = (unsigned short)wd;
    }
    // distinct columns, ascending in the numbering the vectors use (neighbouring ids share cache lines)
    cols.clear();
    for (int j = 0; j < c.nr; ++j)
      for (int p = rp[rows[j]]; p < rp[rows[j] + 1]; ++p)
        if (lidx_of[idx[p]] < 0) { lidx_of[idx[p]] = 0; cols.push_back(idx[p]); }
    std::sort(cols.begin(), cols.end(), [&](int a, int b) { return colid(a) < colid(b); });
    NM_REQUIRE((int)cols.size() == c.nd, "slab: distinct-column count mismatch");
    for (int j = 0; j < c.nd; ++j) { lidx_of[cols[j]] = j; bcols[j] = colid(cols[j]); }
    const size_t val8 = (start[i] + o_val) / 8;                   // blob position of the value region in doubles
    // The ORDER in which the lanes of a row walk its entries is free (it only reassociates the row sum,
    // deterministically); it decides how many shared-memory wavefronts the x reads of a half-warp step cost: 16 lanes
    // x 8 bytes are one wavefront when they hit 16 different bank pairs (bank pair = local column mod 16) OR the same
    // address.  NM_PACK_BANK_AWARE:
    //   3 (default) share-aware greedy: per step the most constrained lane chooses first; it prefers an entry whose
    //     column another lane of the half-warp already reads at this step (same address: free), then an empty bank
    //     pair, then the least loaded one; the lanes of a split row draw from one pool.  Modelled x wavefronts on the
    //     bench workload (nm_slab_host_selftest): 1.27 x the conflict-free count;
    //   2 edge colouring of the bipartite multigraph lanes x bank pairs with the steps as colours (ignores sharing): 1.81 x;
    //   1 lane-by-lane first free bank pair (round 1c; ncu: 31% of the shared wavefronts were replays): 1.9-2.0 x;
    //   0 CSR order: 2.4 x.
    int eoff = 0;
    for (int s = 0; s < nslice; ++s) {
      const int t0 = 32 * s;
      int w = 0;
      for (int l = 0; l < 32; ++l) if (lane_row[t0 + l] >= 0) w = std::max(w, vlen_of(rows[lane_row[t0 + l]]));
      tbl[2 * s] = (unsigned)eoff;
      tbl[2 * s + 1] = (unsigned)w;
      // lane -> its entries (CSR positions): round-robin over the row's lanes, or (share-aware greedy) one pool per
      // row kept by its first lane and drawn from by all of them
      std::vector<int> lane_ent[32];
      int pool_of[32];
      for (int l = 0; l < 32;) {
        const int j = lane_row[t0 + l];
        pool_of[l] = l;
        if (j < 0) { ++l; continue; }
        int g = 1;
        while (l + g < 32 && lane_row[t0 + l + g] == j) ++g;
        int e = 0;
        for (int p = rp[rows[j]]; p < rp[rows[j] + 1]; ++p, ++e) lane_ent[bank_mode == 3 ? l : l + e % g].push_back(p);
        for (int q = 0; q < g; ++q) pool_of[l + q] = bank_mode == 3 ? l : l + q;
        l += g;
      }
      // step of every entry: sched[l][k] = CSR position walked by lane l at step k, or -1 (padding)
      std::vector<int> sched(32 * (size_t)std::max(w, 1), -1);
      for (int h0 = 0; h0 < 32; h0 += 16) {
        if (bank_mode == 2) {
          // edges of the half-warp
          struct Edge { int lane, bank, src, col; };
          std::vector<Edge> E;
          int bankdeg[16] = {0};
          for (int l = h0; l < h0 + 16; ++l)
            for (int p : lane_ent[l]) { const int bk = lidx_of[idx[p]] & 15; E.push_back({l - h0, bk, p, -1}); bankdeg[bk]++; }
          int C = w;
          for (int q = 0; q < 16; ++q) C = std::max(C, bankdeg[q]);
          std::vector<int> lane_col(16 * (size_t)C, -1), bank_col(16 * (size_t)C, -1);     // [vertex][colour] -> edge
          for (int ei = 0; ei < (int)E.size(); ++ei) {
            const int u = E[ei].lane, v = E[ei].bank;
            int ca = 0, cb = 0;
            while (lane_col[(size_t)u * C + ca] >= 0) ++ca;                              // free at the lane
            while (bank_col[(size_t)v * C + cb] >= 0) ++cb;                              // free at the bank pair
            if (bank_col[(size_t)v * C + ca] >= 0) {
              // free colour ca at v: swap ca <-> cb along the alternating path that starts at v with its ca edge
              std::vector<int> path;
              int bankv = v, col = ca;
              for (;;) {
                const int e1 = bank_col[(size_t)bankv * C + col];                        // bank --col--> lane
                if (e1 < 0) break;
                path.push_back(e1);
                const int other = col == ca ? cb : ca;
                const int e2 = lane_col[(size_t)E[e1].lane * C + other];                  // lane --other--> bank
                if (e2 < 0) break;
                path.push_back(e2);
                bankv = E[e2].bank;
              }
              for (int pe : path) {                                                       // clear, then set swapped
                lane_col[(size_t)E[pe].lane * C + E[pe].col] = -1;
                bank_col[(size_t)E[pe].bank * C + E[pe].col] = -1;
              }
              for (int pe : path) {
                E[pe].col = E[pe].col == ca ? cb : ca;
                lane_col[(size_t)E[pe].lane * C + E[pe].col] = pe;
                bank_col[(size_t)E[pe].bank * C + E[pe].col] = pe;
              }
            }
            E[ei].col = ca;
            lane_col[(size_t)u * C + ca] = ei;
            bank_col[(size_t)v * C + ca] = ei;
          }
          // colours < w are steps; the few entries coloured >= w (an overloaded bank pair) take free steps of their lane
          std::vector<int> over;
          for (int ei = 0; ei < (int)E.size(); ++ei) {
            if (E[ei].col < w) sched[(size_t)(h0 + E[ei].lane) * w + E[ei].col] = E[ei].src;
            else over.push_back(ei);
          }
          for (int ei : over) {
            const int l = h0 + E[ei].lane;
            int k = 0;
            while (k < w && sched[(size_t)l * w + k] >= 0) ++k;
            NM_REQUIRE(k < w, "slab: lane with more entries than its slice is wide");
            sched[(size_t)l * w + k] = E[ei].src;
          }
        } else if (bank_mode == 3) {
          // share-aware greedy: lanes that read the SAME x value in the same step share a wavefront, so a lane first
          // looks for an entry whose column another lane of the half-warp already reads at this step, then for an
          // empty bank pair, then for the least loaded one
          for (int k = 0; k < w; ++k) {
            int naddr[16] = {0};
            int addr[16][16];
            int lanes[16];
            for (int q = 0; q < 16; ++q) lanes[q] = h0 + q;
            // most constrained lanes (fewest candidates left) choose first
            std::stable_sort(lanes, lanes + 16, [&](int a, int b) { return lane_ent[pool_of[a]].size() < lane_ent[pool_of[b]].size(); });
            for (int q0 = 0; q0 < 16; ++q0) {
              const int l = lanes[q0];
              std::vector<int>& rj = lane_ent[pool_of[l]];
              if (lane_row[t0 + l] < 0 || rj.empty()) continue;
              int pick = 0, best = 1 << 30;
              for (int q = 0; q < (int)rj.size(); ++q) {
                const int li = lidx_of[idx[rj[q]]], bk = li & 15;
                int score = 2 + 2 * naddr[bk];
                for (int z = 0; z < naddr[bk]; ++z) if (addr[bk][z] == li) { score = 0; break; }
                if (score == 2) score = 1;
                if (score < best) { best = score; pick = q; if (score == 0) break; }
              }
              const int src = rj[pick];
              rj.erase(rj.begin() + pick);
              const int li = lidx_of[idx[src]], bk = li & 15;
              bool dup = false;
              for (int z = 0; z < naddr[bk]; ++z) dup = dup || addr[bk][z] == li;
              if (!dup) addr[bk][naddr[bk]++] = li;
              sched[(size_t)l * w + k] = src;
            }
          }
        } else {
          for (int k = 0; k < w; ++k) {
            unsigned taken = 0;
            for (int l = h0; l < h0 + 16; ++l) {
              std::vector<int>& rj = lane_ent[l];
              if (rj.empty()) continue;
              int pick = 0;
              if (bank_mode == 1)
                for (int q = 0; q < (int)rj.size(); ++q)
                  if (!(taken & (1u << (lidx_of[idx[rj[q]]] & 15)))) { pick = q; break; }
              const int src = rj[pick];
              rj.erase(rj.begin() + pick);
              taken |= 1u << (lidx_of[idx[src]] & 15);
              sched[(size_t)l * w + k] = src;
            }
          }
        }
        // write the half-warp's steps; padding lanes (value 0.0) point at a bank pair that is still free
        for (int k = 0; k < w; ++k) {
          unsigned taken = 0;
          for (int l = h0; l < h0 + 16; ++l) {
            const int src = sched[(size_t)l * w + k];
            if (src < 0) continue;
            const int p = eoff + 32 * k + l;
            const int li = lidx_of[idx[src]];
            taken |= 1u << (li & 15);
            bidx[p] = (unsigned short)li;
            slot_off8.push_back((unsigned)(val8 + p));
            slot_src.push_back(src);
          }
          for (int l = h0; l < h0 + 16; ++l) {
            if (sched[(size_t)l * w + k] >= 0) continue;
            int li = 0;
            for (int q = 0; q < 16 && q < c.nd; ++q)
              if (!(taken & (1u << q))) { li = q; break; }
            taken |= 1u << (li & 15);
            bidx[eoff + 32 * k + l] = (unsigned short)li;
          }
        }
      }
      eoff += 32 * w;
    }
    NM_REQUIRE((long long)slot_src.size() - nslot_before == (long long)chunk_entries, "slab: entries lost while scheduling");
    NM_REQUIRE(eoff == nep, "slab: padded entry count mismatch");
    for (int j = 0; j < c.nd; ++j) lidx_of[cols[j]] = -1;
  }
  // ---- 4. launch geometry: CTAs get contiguous chunk ranges of ~equal bytes
  H.threads = T;
  H.xs_doubles = R * max_nd;
  H.stage_bytes = (int)up16(max_blob);
  // NM_SLAB_WS (default 1): warp-specialised kernel -- producer warps + full/empty mbarriers, deeper rings
  H.ws = nm_env_int("NM_SLAB_WS", 1) != 0;
  H.nprod = std::max(1, std::min(8, nm_env_int("NM_SLAB_PRODUCERS", 4)));
  H.nstage = std::max(2, std::min(8, nm_env_int("NM_SLAB_STAGES", 2)));   // >= 2: blob it+1 is awaited while chunk it is walked
  H.nxs = H.ws ? std::max(2, std::min(8, nm_env_int("NM_SLAB_XS", 2))) : 2;
  const int budget = 226 * 1024;
  auto smem_of = [&]() {
    const int fixed = NM_SLAB_MAXDESC * (int)sizeof(NmPackDesc) + (H.ws ? 256 : 64) + 8 * H.nxs * H.xs_doubles;
    return (int)up16(fixed) + H.nstage * H.stage_bytes;
  };
  while (smem_of() > budget && (H.nstage > 2 || H.nxs > 2)) {
    if (H.nxs > 2 && H.nxs >= H.nstage) H.nxs--; else if (H.nstage > 2) H.nstage--; else H.nxs--;
  }
  H.smem_bytes = smem_of();
  if (H.smem_bytes > budget) return false;
  const int fit = std::max(1, (228 * 1024) / (H.smem_bytes + 1024));          // 1 KB reserved per CTA
  const int per_sm = std::max(1, std::min(fit, nm_env_int("NM_SLAB_CTAS_PER_SM", fit)));
  int grid = std::min(nchunk, sm_count * per_sm);
  grid = std::max(1, std::min(grid, nm_env_int("NM_SLAB_MAXGRID", grid)));     // tests: force several chunks per CTA
  const int maxper = NM_SLAB_MAXDESC;
  grid = std::max(grid, nm_div_up(nchunk, maxper));
  H.cta_first.assign(grid + 1, 0);
  H.max_chunks_per_cta = 0;
  {
    // balanced contiguous split by bytes: every CTA gets 1..maxper chunks
    int ci = 0;
    size_t done = 0;
    for (int g = 0; g < grid; ++g) {
      H.cta_first[g] = ci;
      const size_t target = (size_t)((double)total * (g + 1) / grid);
      const int after = grid - g - 1;                                  // CTAs still to be served
      int cnt = 1;
      done += desc[ci].bytes;
      while (ci + cnt < nchunk - after && cnt < maxper) {
        const bool forced = (long long)(nchunk - (ci + cnt)) > (long long)after * maxper;
        if (!forced && done + desc[ci + cnt].bytes / 2 > target) break;
        done += desc[ci + cnt].bytes;
        ++cnt;
      }
      ci += cnt;
      H.max_chunks_per_cta = std::max(H.max_chunks_per_cta, cnt);
    }
    H.cta_first[grid] = ci;
    NM_REQUIRE(ci == nchunk, "slab: chunk split lost chunks (%d of %d)", ci, nchunk);
  }
  // Several GPUs: a chunk with ghost columns costs more than its bytes (its ghost values are polled out of the
  // flag-in-data slots through registers, its boundary rows are stored to the peers), and the chunks next to a
  // partition boundary are neighbours in the breadth-first order -- a contiguous split hands ALL of them to a few CTAs,
  // whose tail then sets the step time (200k-tet workload on 2 GPUs: B~ step 32 us for half the rows of the 28 us
  // single-GPU step, with or without waiting for the peer: profiles/r2e_*).  So the two kinds are dealt separately:
  // every CTA gets a contiguous share of the interior chunks (by bytes) followed by a contiguous share of the boundary
  // chunks (by bytes).  The processing order inside a CTA stays "interior first": the values the boundary chunks need
  // were sent a whole step earlier.  (The row positions are in the blob headers: the order of the descriptors is free.)
  H.desc_cid.resize(nchunk);
  {
    std::vector<int> in_id, bd_id;
    for (int ci = 0; ci < nchunk; ++ci) (ghost_chunk[ci] ? bd_id : in_id).push_back(ci);
    std::vector<std::vector<int>> mine(grid);
    bool dealt = !bd_id.empty() && nm_env_int("NM_SLAB_DEAL_GHOST", 1) != 0;
    if (dealt) {
      auto deal = [&](const std::vector<int>& ids) {
        size_t tot = 0, cum = 0;
        for (int ci : ids) tot += desc[ci].bytes;
        for (int ci : ids) {
          const int g = (int)std::min<size_t>((size_t)grid - 1, (size_t)((double)(cum + desc[ci].bytes / 2) / (double)std::max<size_t>(tot, 1) * grid));
          mine[g].push_back(ci);
          cum += desc[ci].bytes;
        }
      };
      deal(in_id); deal(bd_id);
      for (int g = 0; g < grid && dealt; ++g) dealt = !mine[g].empty() && (int)mine[g].size() <= maxper;
    }
    if (!dealt) {                                                // one GPU, or a matrix too small to deal: contiguous ranges
      for (int g = 0; g < grid; ++g) {
        mine[g].clear();
        for (int ci = H.cta_first[g]; ci < H.cta_first[g + 1]; ++ci) if (!ghost_chunk[ci]) mine[g].push_back(ci);
        for (int ci = H.cta_first[g]; ci < H.cta_first[g + 1]; ++ci) if (ghost_chunk[ci]) mine[g].push_back(ci);
      }
    }
    std::vector<NmPackDesc> nd(nchunk);
    int pos = 0;
    H.max_chunks_per_cta = 0;
    for (int g = 0; g < grid; ++g) {
      H.cta_first[g] = pos;
      for (int ci : mine[g]) { nd[pos] = desc[ci]; H.desc_cid[pos] = ci; ++pos; }
      H.max_chunks_per_cta = std::max(H.max_chunks_per_cta, (int)mine[g].size());
    }
    H.cta_first[grid] = pos;
    NM_REQUIRE(pos == nchunk, "slab: chunk deal lost chunks (%d of %d)", pos, nchunk);
    desc = nd;
  }
  H.grid = grid;
  H.padded_entries = pentries;
  H.desc = desc;
  H.order = final_order;
  H.nchunk = nchunk;
  return true;
}

void nm_slab_build_into(NmParcsr& M, NmSlab& S, const std::vector<int>& rp, const std::vector<int>& idx, int n) {
  S.nchunk = 0;
  if (n == 0 || M.nrow != M.ncol || M.format == NM_FMT_ROW3) return;
  const int R = M.format == NM_FMT_CSR ? 1 : 3;
  const int ncolb = (M.ncol + M.halo.nghost + R - 1) / R;                     // column ids are < ncolb
  NmSlabHost H;
  if (!slab_build_host(H, rp, idx, n, R, ncolb, nm_ctx().sm_count)) return;
  S.threads = H.threads; S.xs_doubles = H.xs_doubles; S.stage_bytes = H.stage_bytes; S.nstage = H.nstage;
  S.smem_bytes = H.smem_bytes; S.grid = H.grid; S.max_chunks_per_cta = H.max_chunks_per_cta;
  S.ws = H.ws != 0; S.nxs = H.nxs; S.nprod = H.nprod;
  S.pdl = nm_env_int("NM_SLAB_PDL", 1) != 0;
  S.bytes = (long long)H.blob.size();
  S.entries = (long long)H.slot_src.size(); S.padded_entries = H.padded_entries;
  S.nslot = (long long)H.slot_src.size();
  S.blob.alloc(H.blob.size()); S.blob.upload(H.blob.data(), H.blob.size());
  S.desc.from_host(H.desc);
  S.cta_first.from_host(H.cta_first);
  S.slot_off8.from_host(H.slot_off8);
  S.slot_src.from_host(H.slot_src);
  S.order.from_host(H.order);
  S.desc_cid.from_host(H.desc_cid);
  S.deps_ok = H.deps_ok;
  S.cflag.alloc(H.nchunk); S.cflag.zero();
  S.nchunk = H.nchunk;
  if (nm_env_int("NM_SLAB_TRACE", 0)) { S.trace.alloc((size_t)S.grid * NM_SLAB_MAXDESC * 8); S.trace.zero(); }
  nm_slab_fill_from(M, S);
}

// CPU self-test hook (no GPU needed): packs the n x ncolb index pattern with R scalar components per entry, fills
// the values, then WALKS THE BLOBS exactly as k_slab does (same decode, same slice / lane / step order, same ring
// and CTA assignment) to form y = A x in pack order.  Returns the pack order and geometry for the caller to check.
extern "C" int nm_slab_host_selftest(int n, int ncolb, int R, const int* rp_, const int* idx_, const double* vals,
                                     const double* x /* R*ncolb */, double* y /* R*n, pack order */, int* order_out,
                                     int* info /* 10: nchunk, grid, threads, smem_bytes, nstage, max_chunks_per_cta, padded, sum nd, x wavefronts (model), ideal */) {
  NM_API_BEGIN
  std::vector<int> rp(rp_, rp_ + n + 1), idx(idx_, idx_ + rp_[n]);
  NmSlabHost H;
  NM_REQUIRE(slab_build_host(H, rp, idx, n, R, ncolb, 148), "slab: not packable");
  double* blob8 = (double*)H.blob.data();
  for (size_t i = 0; i < H.slot_src.size(); ++i) blob8[H.slot_off8[i]] = vals[H.slot_src[i]];
  std::vector<int> newid(n);
  for (int i = 0; i < n; ++i) newid[H.order[i]] = i;
  std::vector<char> seen(n, 0);
  std::vector<double> xs;
  long long model_wavefronts = 0, model_ideal = 0;
  for (int g = 0; g < H.grid; ++g) {
    NM_REQUIRE(H.cta_first[g + 1] - H.cta_first[g] >= 1 && H.cta_first[g + 1] - H.cta_first[g] <= NM_SLAB_MAXDESC,
               "slab: CTA %d has %d chunks", g, H.cta_first[g + 1] - H.cta_first[g]);
    for (int ci = H.cta_first[g]; ci < H.cta_first[g + 1]; ++ci) {
      const unsigned char* st = H.blob.data() + 16ull * H.desc[ci].off16;
      NM_REQUIRE((int)H.desc[ci].bytes <= H.stage_bytes && H.desc[ci].bytes % 16 == 0, "slab: blob size");
      const NmSlabHeader h = *(const NmSlabHeader*)st;
      const unsigned* tbl = (const unsigned*)(st + 32);
      const double* sv = (const double*)(st + 32 + ((8 * h.nslice + 15) & ~15));
      const int* scols = (const int*)(sv + h.nep);
      const unsigned short* sidx = (const unsigned short*)(scols + h.nd);
      NM_REQUIRE(h.nr >= 1 && h.nr <= 32 * h.nslice && 32 * h.nslice <= H.threads && R * h.nd <= H.xs_doubles, "slab: header");
      NM_REQUIRE((const unsigned char*)(sidx + h.nep) <= st + H.desc[ci].bytes, "slab: blob overrun");
      xs.assign((size_t)R * h.nd, 0.0);
      for (int j = 0; j < R * h.nd; ++j) {
        const int node = j / R;
        // columns are in the vectors' numbering (pack order for owned columns): x is given in the CALLER's order
        const int c = scols[node];
        NM_REQUIRE(c >= 0 && c < ncolb, "slab: column id");
        const int orig = c < n ? H.order[c] : c;
        xs[j] = x[(size_t)R * orig + (j - R * node)];
      }
      const unsigned short* slane = sidx + h.nep;
      NM_REQUIRE((const unsigned char*)(slane + 32 * h.nslice) + 4 * (size_t)h.pad2 <= st + H.desc[ci].bytes, "slab: blob overrun (lanes / deps)");
      {
        const int* dp = (const int*)(slane + 32 * h.nslice);
        for (int q = 0; q < h.pad2; ++q) NM_REQUIRE(dp[q] >= 0 && dp[q] < H.nchunk, "slab: dependency id");
        NM_REQUIRE(H.desc_cid[ci] >= 0 && H.desc_cid[ci] < H.nchunk, "slab: chunk id");
      }
      int owners = 0;
      for (int warp = 0; warp < H.threads / 32; ++warp) {
        double acc[32][3];
        unsigned lw[32];
        for (int lane = 0; lane < 32; ++lane) {
          acc[lane][0] = acc[lane][1] = acc[lane][2] = 0.0;
          lw[lane] = warp < h.nslice ? slane[32 * warp + lane] : 0u;
          if (warp < h.nslice) {
            const unsigned eoff = tbl[2 * warp], w = tbl[2 * warp + 1];
            for (unsigned k = 0; k < w; ++k) {
              if (lane == 0 || lane == 16) {
                // modelled shared-memory wavefronts of one x component read by this half-warp at this step: 8-byte
                // accesses, 16 bank pairs; lanes reading the same address share a wavefront
                int cnt[16] = {0}, seen_li[16][16], nseen[16] = {0}, worst = 1;
                for (int l2 = lane; l2 < lane + 16; ++l2) {
                  const int li2 = sidx[eoff + 32 * k + l2], bp = (R * li2) & 15;
                  bool dup = false;
                  for (int q = 0; q < nseen[bp]; ++q) dup = dup || seen_li[bp][q] == li2;
                  if (!dup) { seen_li[bp][nseen[bp]++] = li2; cnt[bp]++; worst = std::max(worst, cnt[bp]); }
                }
                model_wavefronts += worst;
                model_ideal += 1;
              }
              const double m = sv[eoff + 32 * k + lane];
              const int li = sidx[eoff + 32 * k + lane];
              NM_REQUIRE(li < h.nd, "slab: local index");
              for (int c = 0; c < R; ++c) acc[lane][c] += m * xs[(size_t)R * li + c];
            }
          }
        }
        for (int s2 = 0; (1 << s2) < h.gmax; ++s2) {            // __shfl_down_sync semantics (out of range: own value)
          const int o = 1 << s2;
          double nxt[32][3];
          for (int lane = 0; lane < 32; ++lane)
            for (int c = 0; c < R; ++c) {
              const double other = lane + o < 32 ? acc[lane + o][c] : acc[lane][c];
              nxt[lane][c] = (lw[lane] & (1u << (10 + s2))) ? acc[lane][c] + other : acc[lane][c];
            }
          memcpy(acc, nxt, sizeof(acc));
        }
        for (int lane = 0; lane < 32; ++lane) {
          if (!(lw[lane] & 0x8000u)) continue;
          const unsigned rank = lw[lane] & 0x3ffu;
          NM_REQUIRE((int)rank < h.nr, "slab: local row %u of %d", rank, h.nr);
          const int row = h.first + (int)rank;
          NM_REQUIRE(row >= 0 && row < n && !seen[row], "slab: row %d visited twice or out of range", row);
          seen[row] = 1;
          ++owners;
          for (int c = 0; c < R; ++c) y[(size_t)R * row + c] = acc[lane][c];
        }
      }
      NM_REQUIRE(owners == h.nr, "slab: %d leader lanes for %d rows", owners, h.nr);
    }
  }
  for (int i = 0; i < n; ++i) NM_REQUIRE(seen[i], "slab: row %d never visited", i);
  if (order_out) std::copy(H.order.begin(), H.order.end(), order_out);
  if (info) {
    info[0] = H.nchunk; info[1] = H.grid; info[2] = H.threads; info[3] = H.smem_bytes; info[4] = H.nstage;
    info[5] = H.max_chunks_per_cta; info[6] = (int)std::min<long long>(H.padded_entries, 0x7fffffff);
    long long snd = 0;
    for (int ci = 0; ci < H.nchunk; ++ci) snd += ((const NmSlabHeader*)(H.blob.data() + 16ull * H.desc[ci].off16))->nd;
    info[7] = (int)std::min<long long>(snd, 0x7fffffff);                     // sum of distinct columns over chunks
    info[8] = (int)std::min<long long>(model_wavefronts, 0x7fffffff);        // modelled wavefronts of one x component
    info[9] = (int)std::min<long long>(model_ideal, 0x7fffffff);             // ... and the conflict-free count
  }
  NM_API_END
}

__global__ void k_slab_fill(long long nslot, const unsigned* __restrict__ off8, const int* __restrict__ src,
                            const double* __restrict__ vals, double* __restrict__ blob) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < nslot) blob[off8[i]] = vals[src[i]];
}

void nm_slab_fill_from(NmParcsr& M, NmSlab& S) {
  if (S.nchunk == 0 || S.nslot == 0) return;
  NmCtx& c = nm_ctx();
  const double* vals = M.format == NM_FMT_KRON3 ? M.mval.p : M.a.p;
  k_slab_fill<<<nm_div_up(S.nslot, 256), 256, 0, c.stream>>>(S.nslot, S.slot_off8.p, S.slot_src.p, vals, (double*)S.blob.p);
  c.launches++;
}
