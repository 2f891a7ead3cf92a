// Mesh topology, DOF numbering and CSR sparsity patterns on the host (integer work, done once):
// the part of the reference that fixes the bit-exact numbering contract.
//   vertex status / v2v        src/mod_geometry.f90:284-421
//   P2 edge nodes              src/mod_geometry.f90:428-689, 1566-1717
//   solid numbering + pattern  src/mod_cg_create_matrix.f90:1269-1455   (matrixstruct)
//   fluid / fluid-solid        src/mod_cg_create_matrix.f90:1458-2033   (matrixstruct_general)
// A rank owns the nodes with part[] == rank in ascending id (src/mod_geometry.f90:953-965,1129) and the
// rows of those nodes; DOFs are numbered rank after rank (src/mod_cg_create_matrix.f90:1291-1303).
#include "nm_fem.h"
#include <algorithm>
#include <numeric>

static const int P2_PAIRS[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}};   // e12,e13,e23,e14,e24,e34
static const int P2_ORD[10] = {0, 2, 5, 9, 1, 3, 4, 6, 7, 8};                          // src/mod_geometry.f90:1707

static void block_dist(int n, int nproc, std::vector<int>& d) {      // src/mod_geometry.f90:110-120
  d.assign(nproc + 1, 0);
  const int r = n % nproc, q = n / nproc;
  for (int i = 0; i < nproc; ++i) d[i + 1] = d[i] + q + (i < r ? 1 : 0);
}

NmFem* nm_fem_build(int ntet, int nvert, const int* ele, const int* neigh, const double* node, int porder,
                    const double* vs, int nproc, const int* part, int rank) {
  NM_REQUIRE(porder == 1 || porder == 2, "pOrder must be 1 or 2 (src/mod_para.f90:111)");
  NM_REQUIRE(nproc >= 1 && rank >= 0 && rank < nproc, "bad rank %d of %d", rank, nproc);
  std::unique_ptr<NmFem> F(new NmFem());
  F->ntet = ntet; F->nvert = nvert; F->porder = porder; F->pNp = porder == 1 ? 4 : 10;
  F->nproc = nproc; F->rank = rank;
  const int pNp = F->pNp;
  F->ele.assign(ele, ele + 4 * (size_t)ntet);
  F->neigh.assign(neigh, neigh + 4 * (size_t)ntet);
  F->node.assign(node, node + 3 * (size_t)nvert);
  for (size_t i = 0; i < F->ele.size(); ++i)
    NM_REQUIRE(F->ele[i] >= 0 && F->ele[i] < nvert, "element vertex id out of range (0-based ids expected)");
  // ---- fluid elements and vertex status (src/mod_geometry.f90:284-312)
  F->efl.assign(ntet, 0);
  std::vector<int> nfl(nvert, 0), ntot(nvert, 0);
  for (int e = 0; e < ntet; ++e) {
    double mx = 0.0;
    for (int k = 0; k < pNp; ++k) mx = std::max(mx, vs[(size_t)e * pNp + k]);
    F->efl[e] = mx < 1.0e-6;
    for (int k = 0; k < 4; ++k) { ntot[ele[4 * e + k]]++; nfl[ele[4 * e + k]] += F->efl[e]; }
  }
  std::vector<int> vstat_v(nvert);
  int mx = 0;
  for (int v = 0; v < nvert; ++v) {
    vstat_v[v] = nfl[v] == 0 ? 0 : (nfl[v] == ntot[v] ? 1 : 2);
    mx = std::max(mx, vstat_v[v]);
  }
  F->fsexist = (mx == 2); F->purefluid = (mx == 1);                  // :329-341
  F->fluidcase = F->fsexist || F->purefluid;
  // ---- element -> node table (P2: edge nodes numbered as src/mod_geometry.f90:452-580)
  if (porder == 1) {
    F->nn = nvert;
    F->t2n.assign(F->ele.begin(), F->ele.end());
    F->vstat = vstat_v;
  } else {
    std::vector<long long> keys((size_t)6 * ntet);
    for (int e = 0; e < ntet; ++e)
      for (int k = 0; k < 6; ++k) {
        const int a = ele[4 * e + P2_PAIRS[k][0]], b = ele[4 * e + P2_PAIRS[k][1]];
        keys[(size_t)6 * e + k] = (long long)std::min(a, b) * nvert + std::max(a, b);
      }
    std::vector<long long> uk(keys);
    std::sort(uk.begin(), uk.end());
    uk.erase(std::unique(uk.begin(), uk.end()), uk.end());
    const int ne = (int)uk.size();
    // order: (owner rank under the initial block distribution, scanning vertex, neighbour)
    std::vector<int> vtxdist;
    block_dist(nvert, nproc, vtxdist);
    auto pid = [&](int v) { return (int)(std::upper_bound(vtxdist.begin(), vtxdist.end(), v) - vtxdist.begin()) - 1; };
    struct EK { int owner, i, nb, idx; };
    std::vector<EK> ek(ne);
    for (int k = 0; k < ne; ++k) {
      const int a = (int)(uk[k] / nvert), b = (int)(uk[k] % nvert);
      const int pa = pid(a), pb = pid(b);
      const int i = (pa == pb) ? a : (pa > pb ? a : b);
      ek[k] = {std::max(pa, pb), i, i == a ? b : a, k};
    }
    std::sort(ek.begin(), ek.end(), [](const EK& x, const EK& y) {
      if (x.owner != y.owner) return x.owner < y.owner;
      if (x.i != y.i) return x.i < y.i;
      return x.nb < y.nb;
    });
    std::vector<int> eid(ne);                      // sorted-key index -> edge number
    for (int k = 0; k < ne; ++k) eid[ek[k].idx] = k;
    F->nn = nvert + ne;
    F->t2n.resize((size_t)10 * ntet);
    for (int e = 0; e < ntet; ++e) {
      for (int k = 0; k < 4; ++k) F->t2n[(size_t)10 * e + P2_ORD[k]] = ele[4 * e + k];
      for (int k = 0; k < 6; ++k) {
        const int pos = (int)(std::lower_bound(uk.begin(), uk.end(), keys[(size_t)6 * e + k]) - uk.begin());
        F->t2n[(size_t)10 * e + P2_ORD[4 + k]] = nvert + eid[pos];
      }
    }
    F->vstat.assign(F->nn, 0);
    std::copy(vstat_v.begin(), vstat_v.end(), F->vstat.begin());
    F->edges.resize((size_t)2 * ne);
    for (int k = 0; k < ne; ++k) {
      const int a = (int)(uk[k] / nvert), b = (int)(uk[k] % nvert);
      F->vstat[nvert + eid[k]] = std::min(vstat_v[a], vstat_v[b]) + 3;      // :672-680
      F->edges[2 * (size_t)eid[k]] = a; F->edges[2 * (size_t)eid[k] + 1] = b;
    }
  }
  const int nn = F->nn;
  // ---- node -> element lists and node adjacency (sorted, unique, incl. self; :343-421, 1045-1060)
  F->n2e_ptr.assign(nn + 1, 0);
  for (size_t i = 0; i < F->t2n.size(); ++i) F->n2e_ptr[F->t2n[i] + 1]++;
  for (int i = 0; i < nn; ++i) F->n2e_ptr[i + 1] += F->n2e_ptr[i];
  F->n2e.resize(F->t2n.size());
  {
    std::vector<int> fill(F->n2e_ptr.begin(), F->n2e_ptr.end() - 1);
    for (int e = 0; e < ntet; ++e)
      for (int k = 0; k < pNp; ++k) F->n2e[fill[F->t2n[(size_t)e * pNp + k]]++] = e;   // ascending element id
  }
  F->v2v_ptr.assign(nn + 1, 0);
  {
    std::vector<int> tmp;
    std::vector<std::vector<int>> adj;       // built in chunks to bound memory
    F->v2v.clear();
    F->v2v.reserve((size_t)nn * (porder == 1 ? 15 : 30));
    for (int i = 0; i < nn; ++i) {
      tmp.clear();
      for (int p = F->n2e_ptr[i]; p < F->n2e_ptr[i + 1]; ++p) {
        const int e = F->n2e[p];
        for (int k = 0; k < pNp; ++k) tmp.push_back(F->t2n[(size_t)e * pNp + k]);
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      F->v2v.insert(F->v2v.end(), tmp.begin(), tmp.end());
      F->v2v_ptr[i + 1] = (int)F->v2v.size();
    }
  }
  // ---- numbering (src/mod_cg_create_matrix.f90:1278-1303, 1467-1499, 1581-1613)
  F->part.assign(nn, 0);
  if (part) {
    F->part.assign(part, part + nn);
    for (int i = 0; i < nn; ++i) NM_REQUIRE(F->part[i] >= 0 && F->part[i] < nproc, "part[%d] = %d out of range", i, F->part[i]);
  } else {
    NM_REQUIRE(nproc == 1, "a partition vector is required for nproc > 1 (ParMETIS is not part of this library)");
  }
  F->order.resize(nn);
  std::iota(F->order.begin(), F->order.end(), 0);
  std::stable_sort(F->order.begin(), F->order.end(), [&](int a, int b) { return F->part[a] < F->part[b]; });
  F->vnum.resize(nn); F->pnum.resize(nn); F->vstt.resize(nn); F->pstt.assign(nn, -1);
  for (int i = 0; i < nn; ++i) {
    const int s = F->vstat[i];
    F->vnum[i] = (F->fluidcase && (s == 2 || s == 5)) ? 6 : 3;
    F->pnum[i] = (F->fluidcase && s != 0 && s != 3) ? 1 : 0;
  }
  F->vtxdist.assign(nproc + 1, 0);
  for (int i = 0; i < nn; ++i) F->vtxdist[F->part[i] + 1]++;
  for (int r = 0; r < nproc; ++r) F->vtxdist[r + 1] += F->vtxdist[r];
  F->sizdist.assign(nproc + 1, 0); F->psizdist.assign(nproc + 1, 0);
  {
    long long cs = 0, cp = 0;
    int r = 0;
    for (int k = 0; k < nn; ++k) {
      while (k == F->vtxdist[r + 1]) { ++r; F->sizdist[r] = (int)cs; F->psizdist[r] = (int)cp; }
      const int i = F->order[k];
      F->vstt[i] = (int)cs; cs += F->vnum[i];
      if (F->pnum[i]) { F->pstt[i] = (int)cp; cp += 1; }
      NM_REQUIRE(cs < 2147483647ll, "global DOF count overflows int32 (SURVEY.md App. F)");
    }
    for (int q = r + 1; q <= nproc; ++q) { F->sizdist[q] = (int)cs; F->psizdist[q] = (int)cp; }
    F->N = (int)cs; F->Np = (int)cp;
  }
  // ---- patterns of the rows this rank owns
  const int k0 = F->vtxdist[rank], k1 = F->vtxdist[rank + 1];
  const int row0 = F->sizdist[rank], prow0 = F->psizdist[rank];
  const int nrow = F->sizdist[rank + 1] - row0, nprow = F->psizdist[rank + 1] - prow0;
  auto init = [&](NmPattern& P, int nr, const std::vector<int>& rd, const std::vector<int>& cd) {
    P.present = true; P.nrow = nr; P.rowdist = rd; P.coldist = cd; P.ia.assign(nr + 1, 0); P.ja.clear();
  };
  NmPattern& A = F->pat[NM_MAT_A]; NmPattern& B = F->pat[NM_MAT_B];
  NmPattern& E = F->pat[NM_MAT_E]; NmPattern& ET = F->pat[NM_MAT_ET]; NmPattern& Ap = F->pat[NM_MAT_AP];
  init(A, nrow, F->sizdist, F->sizdist); init(B, nrow, F->sizdist, F->sizdist);
  if (F->fluidcase) {
    init(E, nrow, F->sizdist, F->psizdist); init(ET, nprow, F->psizdist, F->sizdist); init(Ap, nprow, F->psizdist, F->psizdist);
  }
  std::vector<int> cols;
  auto close_row = [&](NmPattern& P, int lrow) {       // sorted ascending (simplessort, src/mod_utility.f90:829)
    std::sort(cols.begin(), cols.end());
    NM_REQUIRE(std::adjacent_find(cols.begin(), cols.end()) == cols.end(), "duplicate column in a pattern row");
    P.ja.insert(P.ja.end(), cols.begin(), cols.end());
    P.ia[lrow + 1] = (int)P.ja.size();
    NM_REQUIRE(P.ja.size() < 2147483647ull, "local nnz overflows int32 (SURVEY.md App. F)");
  };
  const std::vector<int>&vstt = F->vstt, &vnum = F->vnum, &pstt = F->pstt, &pnum = F->pnum;
  for (int k = k0; k < k1; ++k) {
    const int i = F->order[k];
    const int* nb = &F->v2v[F->v2v_ptr[i]];
    const int nnb = F->v2v_ptr[i + 1] - F->v2v_ptr[i];
    const int lr = vstt[i] - row0;
    if (!F->fluidcase) {
      for (int p = 0; p < 3; ++p) {                    // A: 3 identical rows (:1373-1388)
        cols.clear();
        for (int t = 0; t < nnb; ++t) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + q);
        close_row(A, lr + p);
      }
      for (int p = 0; p < 3; ++p) {                    // B: component p only (:1417-1434)
        cols.clear();
        for (int t = 0; t < nnb; ++t) cols.push_back(vstt[nb[t]] + p);
        close_row(B, lr + p);
      }
      continue;
    }
    const int lp = pnum[i] ? pstt[i] - prow0 : -1;
    if (pnum[i] == 0) {                                // pure solid node (:1987-2022)
      for (int p = 0; p < 3; ++p) {
        cols.clear();
        for (int t = 0; t < nnb; ++t) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + q);
        close_row(A, lr + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t) cols.push_back(vstt[nb[t]] + p);
        close_row(B, lr + p);
        cols.clear();
        close_row(E, lr + p);
      }
    } else if (vnum[i] == 3) {                         // pure fluid node (:1820-1859)
      for (int t = 0; t < nnb; ++t)
        NM_REQUIRE(pnum[nb[t]] == 1, "Error: pure fluid node with a solid neighbour (src/mod_cg_create_matrix.f90:1694)");
      for (int p = 0; p < 3; ++p) {
        cols.clear();
        for (int t = 0; t < nnb; ++t) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + vnum[nb[t]] - 3 + q);
        close_row(A, lr + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t) cols.push_back(vstt[nb[t]] + vnum[nb[t]] - 3 + p);
        close_row(B, lr + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t) cols.push_back(pstt[nb[t]]);
        close_row(E, lr + p);
      }
      cols.clear();
      for (int t = 0; t < nnb; ++t) cols.push_back(pstt[nb[t]]);
      close_row(Ap, lp);
      cols.clear();
      for (int t = 0; t < nnb; ++t) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + vnum[nb[t]] - 3 + q);
      close_row(ET, lp);
    } else {                                           // fluid-solid interface node, 6 rows (:1861-1985)
      for (int p = 0; p < 3; ++p) {                    // solid-side rows
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (pnum[nb[t]] == 0 || vnum[nb[t]] == 6) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + q);
        close_row(A, lr + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (pnum[nb[t]] == 0 || vnum[nb[t]] == 6) cols.push_back(vstt[nb[t]] + p);
        close_row(B, lr + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (vnum[nb[t]] == 6) cols.push_back(pstt[nb[t]]);
        close_row(E, lr + p);
      }
      for (int p = 0; p < 3; ++p) {                    // fluid-side rows
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (pnum[nb[t]] == 1) for (int q = 0; q < 3; ++q) cols.push_back(vstt[nb[t]] + vnum[nb[t]] - 3 + q);
        close_row(A, lr + 3 + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (pnum[nb[t]] == 1) cols.push_back(vstt[nb[t]] + vnum[nb[t]] - 3 + p);
        close_row(B, lr + 3 + p);
        cols.clear();
        for (int t = 0; t < nnb; ++t)
          if (pnum[nb[t]] == 1) cols.push_back(pstt[nb[t]]);
        close_row(E, lr + 3 + p);
      }
      cols.clear();
      for (int t = 0; t < nnb; ++t)
        if (pnum[nb[t]] == 1) cols.push_back(pstt[nb[t]]);
      close_row(Ap, lp);
      cols.clear();
      for (int t = 0; t < nnb; ++t)
        if (pnum[nb[t]] == 1) for (int q = 0; q < vnum[nb[t]]; ++q) cols.push_back(vstt[nb[t]] + q);
      close_row(ET, lp);
    }
  }
  // ---- elements this rank integrates: those touching an owned node, ascending (Clelist, :1016-1026)
  {
    std::vector<char> mark(ntet, 0);
    for (int k = k0; k < k1; ++k) {
      const int i = F->order[k];
      for (int p = F->n2e_ptr[i]; p < F->n2e_ptr[i + 1]; ++p) mark[F->n2e[p]] = 1;
    }
    for (int e = 0; e < ntet; ++e)
      if (mark[e]) F->lelist.push_back(e);
  }
  return F.release();
}
