// qp_sparse_cta_host.hpp -- host-side symbolic analysis for the ON-CHIP sparse QP kernel (qp_sparse_cta.cuh): one CTA per
// instance, the factor, Abar and every vector resident in shared memory.
//
// Replaces, like qp_sparse_host.hpp, what the reference gets from Eigen for QuadraticProgramSparse problems
// (pettni/smooth_feedback @ 9a08971):
//   SimplicialLDLT::analyzePattern   call site include/smooth/feedback/qp_solver.hpp:424
//   sparse KKT fill                  include/smooth/feedback/qp_solver.hpp:380-397
// but organises the factor for a CTA instead of a warp.  The tiled kernel's per-iteration cost is one warp's dependent
// chain through ~850 sweep steps (minimum degree orders an MPC problem along the time axis: elimination tree height 346
// at n = 422).  Here
//   * the ordering is chosen for a SHORT elimination tree: nested dissection (BFS level-structure separators, constrained
//     minimum degree inside the parts) competes with plain minimum degree on a cost model; MPC K = 50: 13 supernodes in
//     4 levels instead of a chain of 13;
//   * columns with nested structure are merged into SUPERNODES (fundamental + relaxed amalgamation, a few explicit zeros)
//     and stored as dense blocks: a packed strict-lower s x s diagonal block and a row-major t x s block of the rows below;
//     index data shrinks from one (col, slot) pair per factor entry to a few integers per block row;
//   * after the numeric factorisation the unit-lower diagonal blocks are INVERTED in place, so a triangular sweep is
//     2 barrier-separated stages of independent dot products per supernode level (4 levels -> 16 stages per solve of
//     L D L^T) instead of one dependent step per row.
// Everything here is computed once per pattern and uploaded as flat integer tables.

#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <functional>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace sfb {

constexpr int kCtaNT = 256;  // threads of the on-chip kernel's CTA

enum CtaStageKind { kStageFwdDiag = 0, kStageFwdPush = 1, kStageBwdPull = 2, kStageBwdDiag = 3 };

// offset of row i inside a packed lower-triangular diagonal block whose rows are padded to multiples of pad = 1 << lpad scalars:
// row i holds the entries (i, 0 .. i) -- the diagonal included -- followed by zeros up to the next multiple of pad
inline int cta_diag_off(int i, int lpad)
{
  if (i <= 0) return 0;
  const int G = i >> lpad, pad = 1 << lpad;
  return ((G * (G + 1) / 2) << (2 * lpad)) + (i & (pad - 1)) * ((G + 1) << lpad);
}

struct CtaSymbolic
{
  int n = 0, m = 0, nnzP = 0, nnzA = 0;
  int pad = 4, lpad = 2;  // block rows and supernode starts are multiples of pad scalars (16-byte vector loads: 4 fp32 / 2 fp64)
  int np = 0;             // padded column ids (holes behind supernodes whose size is not a multiple of pad); all n-vectors have np entries
  int nW = 0;             // value slots of the factor (D_k lives on the diagonal of its block during the factorisation); 1 / D_k at nW + k (k: padded id)
  int ns = 0, nlev = 0, smax = 0;
  int ordering = 0;       // 0 = minimum degree, 1 = nested dissection
  int nnzL_true = 0;      // structural entries of L before padding
  long long flops = 0;
  std::vector<int> perm, iperm;  // perm[padded id] = original column or -1 (hole); iperm[original] = padded id
  // supernodes (padded ids c0 .. c0 + s - 1, sp = s rounded up to pad, t rows below): diagonal block at dbase (entry (i, j), j <= i,
  // at dbase + cta_diag_off(i) + j), below block at bbase (row r, column j at bbase + r sp + j), row ids rlist[rbase .. rbase + t)
  std::vector<int> sn_c0, sn_s, sn_sp, sn_t, sn_dbase, sn_bbase, sn_rbase, sn_level;
  std::vector<int> rlist;
  std::vector<int> col_sn;  // per padded id, -1 for holes
  // P and A in padded ids and their gather mirrors (same meaning as in SparseSymbolic; pointer arrays have np + 1 entries)
  std::vector<int> P_rowp, P_colp, P_tgt;
  std::vector<int> A_rowptr, A_col;
  std::vector<int> AT_ptr, AT_row, AT_slot, PR_ptr, PR_col, PR_slot, PS_ptr, PS_col, PS_slot, PC_ptr, PC_slot;
  // assembly of Abar^T diag(w) Abar: rows of A are coloured so that rows of one colour touch disjoint columns; the pairs
  // (a <= b) of the entries of every row of one colour are contiguous:  W[tgt] += w_row A[ea] A[eb].
  // Dealt to the kCtaNT threads of the kernel: rounds of kCtaNT (ea | eb << 16, tgt | row << 16) descriptors, every colour
  // padded to whole rounds (padding: second word == -1); asm_round_sync[r] != 0: colour boundary, a barrier follows the round
  std::vector<int> asm_round_desc, asm_round_sync;
  // numeric factorisation: per supernode the pairs (a <= b) of its below rows and the slot of entry (R[b], R[a]): a | b << 8 | target << 16
  std::vector<int> ext_ptr, ext_word;
  // rounds: supernodes of one level (at most kCtaNT / 32 per round) advance column by column together.  4 ints per round:
  // longest supernode, supernodes | warps << 16, first entry in fac_ext, first entry in fac_warp; fac_warp: one word per warp of
  // the round, supernode | rank << 16 | warps of the supernode << 24
  std::vector<int> fac_rounds, fac_ext, fac_warp;
  std::vector<int> sn_tab;  // 8 ints per supernode: c0, s, t, dbase, bbase, rbase, sp, prbase
  std::vector<int> prow;    // slot of every row of a supernode's panel (s rows of the diagonal block, then the t rows below), from prbase
  // triangular sweeps: stages separated by CTA barriers; 4 ints per stage: kind, first output, outputs, log2(lanes per output).
  //   forward diagonal stage:  outputs fdiag_out[first ..] = row k | supernode << 16
  //   forward push stage:      outputs push_out[first ..] = dst | first task << 16 (the next entry bounds the task list),
  //                            push_task = supernode | below row << 16
  //   backward stages:         outputs bwd_out[first ..] = first column k of a group of pad columns | supernode << 16
  std::vector<int> stages, fdiag_out, bwd_out, push_out, push_task;
  std::vector<int> diag_off;  // cta_diag_off(i) for i <= smax
  // 16-bit copies (two per int) of index arrays, and the column mirror of A as one word per entry: slot | row << 16
  std::vector<int> A_col16, A_rowptr16, AT_ptr16, rlist16, diag_off16, prow16, AT_word;
  // everything the kernel copies into shared memory, concatenated (each table 16-byte aligned); smem_off[CtaIntTable]
  std::vector<int> smem_ints, smem_off;
  bool big_tables_global = false;
  std::string error;
};

namespace cta_detail {

using Adj = std::vector<std::set<int>>;

// structure of L for the elimination order perm (perm[new] = old): st[k] = rows below the diagonal of column k, ascending
inline std::vector<std::vector<int>> symbolic(const Adj& adj, const std::vector<int>& perm)
{
  const int n = (int)adj.size();
  std::vector<int> ip(n);
  for (int k = 0; k < n; ++k) ip[perm[k]] = k;
  std::vector<std::set<int>> g(n);
  for (int k = 0; k < n; ++k)
    for (int b : adj[perm[k]]) g[k].insert(ip[b]);
  std::vector<std::vector<int>> st(n);
  for (int k = 0; k < n; ++k) {
    for (int x : g[k])
      if (x > k) st[k].push_back(x);
    // only the parent needs the clique: struct(k) \ {parent} is a subset of struct(parent)
    if (!st[k].empty()) {
      const int p = st[k][0];
      for (size_t a = 1; a < st[k].size(); ++a) g[p].insert(st[k][a]);
    }
  }
  return st;
}

// minimum degree on the subgraph induced by verts; vertices in `later` stay as never-eliminated neighbours
inline std::vector<int> local_mindeg(const Adj& adj, const std::vector<int>& verts, const std::vector<char>& in_later)
{
  std::unordered_map<int, std::set<int>> g;
  std::vector<char> mine(adj.size(), 0);
  for (int v : verts) mine[v] = 1;
  for (int v : verts) {
    auto& s = g[v];
    for (int w : adj[v])
      if (mine[w] || in_later[w]) s.insert(w);
  }
  std::set<std::pair<int, int>> queue;
  for (int v : verts) queue.insert({(int)g[v].size(), v});
  std::vector<int> order;
  while (!queue.empty()) {
    const int v = queue.begin()->second;
    queue.erase(queue.begin());
    mine[v] = 0;
    order.push_back(v);
    const std::vector<int> nb(g[v].begin(), g[v].end());
    for (int a : nb) {
      if (!mine[a]) continue;
      auto& ga = g[a];
      queue.erase({(int)ga.size(), a});
      ga.erase(v);
      for (int b : nb)
        if (b != a) ga.insert(b);
      queue.insert({(int)ga.size(), a});
    }
  }
  return order;
}

struct NestedDissection
{
  const Adj& adj;
  int leaf;
  std::vector<char> in_later, in_set;
  std::vector<int> level;
  std::vector<int> order;

  NestedDissection(const Adj& a, int leaf_) : adj(a), leaf(leaf_), in_later(a.size(), 0), in_set(a.size(), 0), level(a.size(), -1) {}

  // BFS level structure of the subgraph `verts` (in_set marks it) from start
  std::vector<std::vector<int>> bfs(const std::vector<int>& verts, int start)
  {
    for (int v : verts) level[v] = -1;
    std::vector<std::vector<int>> lv(1, std::vector<int>{start});
    level[start] = 0;
    for (;;) {
      std::vector<int> nxt;
      for (int v : lv.back())
        for (int w : adj[v])
          if (in_set[w] && level[w] < 0) { level[w] = (int)lv.size(); nxt.push_back(w); }
      if (nxt.empty()) break;
      lv.push_back(std::move(nxt));
    }
    return lv;
  }

  void emit_mindeg(const std::vector<int>& verts)
  {
    const std::vector<int> o = local_mindeg(adj, verts, in_later);
    order.insert(order.end(), o.begin(), o.end());
  }

  void run(std::vector<int> verts)
  {
    if ((int)verts.size() <= leaf) { emit_mindeg(verts); return; }
    for (int v : verts) in_set[v] = 1;
    // connected components are independent subtrees
    {
      std::vector<std::vector<int>> comps;
      std::vector<char> seen(adj.size(), 0);
      for (int s : verts) {
        if (seen[s]) continue;
        std::vector<int> comp{s};
        seen[s] = 1;
        for (size_t i = 0; i < comp.size(); ++i)
          for (int w : adj[comp[i]])
            if (in_set[w] && !seen[w]) { seen[w] = 1; comp.push_back(w); }
        comps.push_back(std::move(comp));
      }
      if (comps.size() > 1) {
        for (int v : verts) in_set[v] = 0;
        for (auto& c : comps) run(std::move(c));
        return;
      }
    }
    // pseudo-peripheral start vertex, then the level that balances the halves with a small separator
    int s = verts[0];
    std::vector<std::vector<int>> lv;
    for (int rep = 0; rep < 4; ++rep) {
      lv = bfs(verts, s);
      int s2 = lv.back()[0];
      for (int v : lv.back())
        if (adj[v].size() < adj[s2].size()) s2 = v;
      if (s2 == s) break;
      s = s2;
    }
    lv = bfs(verts, s);
    if (lv.size() < 3) {
      for (int v : verts) in_set[v] = 0;
      emit_mindeg(verts);
      return;
    }
    const int tot = (int)verts.size();
    long long best = -1;
    int bestL = 1, before = (int)lv[0].size();
    for (int L = 1; L + 1 < (int)lv.size(); ++L) {
      const int sep = (int)lv[L].size();
      const int a = before, b = tot - before - sep;
      const long long score = 4LL * sep + std::abs(a - b);
      if (best < 0 || score < best) { best = score; bestL = L; }
      before += sep;
    }
    std::vector<int> A, B, sep = lv[bestL];
    for (int L = 0; L < bestL; ++L) A.insert(A.end(), lv[L].begin(), lv[L].end());
    for (int L = bestL + 1; L < (int)lv.size(); ++L) B.insert(B.end(), lv[L].begin(), lv[L].end());
    for (int v : verts) in_set[v] = 0;
    for (int v : sep) in_later[v] = 1;
    run(std::move(A));
    run(std::move(B));
    for (int v : sep) in_later[v] = 0;
    emit_mindeg(sep);
  }
};

// postorder of the elimination tree (children with the larger structure last, so that chains of nested columns become
// contiguous); returns the new perm
inline std::vector<int> postorder(const Adj& adj, const std::vector<int>& perm)
{
  const int n = (int)adj.size();
  const auto st = symbolic(adj, perm);
  std::vector<std::vector<int>> ch(n);
  std::vector<int> roots;
  for (int k = 0; k < n; ++k) {
    if (st[k].empty()) roots.push_back(k);
    else ch[st[k][0]].push_back(k);
  }
  for (auto& c : ch) std::stable_sort(c.begin(), c.end(), [&](int a, int b) { return st[a].size() < st[b].size(); });
  std::vector<int> out;
  out.reserve(n);
  std::vector<std::pair<int, size_t>> stack;
  for (int r : roots) {
    stack.push_back({r, 0});
    while (!stack.empty()) {
      auto& top = stack.back();
      if (top.second < ch[top.first].size()) {
        const int c = ch[top.first][top.second++];
        stack.push_back({c, 0});
      } else {
        out.push_back(perm[top.first]);
        stack.pop_back();
      }
    }
  }
  return out;
}

struct SnPlan  // supernode partition of an ordering
{
  std::vector<int> perm;                 // final elimination order (supernode columns contiguous)
  std::vector<std::vector<int>> below;   // per supernode: rows below (final permuted indices, ascending)
  std::vector<int> c0, s, level, parent;
  int nlev = 0, nW = 0, nnzL_true = 0;
  long long flops = 0;
  double cost = 0;
};

// fundamental supernodes of `perm` (postordered), relaxed amalgamation, renumbering, level structure
inline SnPlan plan_supernodes(const Adj& adj, const std::vector<int>& perm_in, int zabs, double zrel, int scap)
{
  const int n = (int)adj.size();
  SnPlan pl;
  const auto st = symbolic(adj, perm_in);
  for (const auto& c : st) pl.nnzL_true += (int)c.size();
  // fundamental supernodes: chains k -> k + 1 with struct(k) = {k + 1} + struct(k + 1)
  std::vector<std::vector<int>> cols, bel;
  std::vector<int> sn_of(n, 0);
  for (int k = 0; k < n;) {
    const int a = k;
    while (k + 1 < n && !st[k].empty() && st[k][0] == k + 1 && st[k].size() == st[k + 1].size() + 1) ++k;
    std::vector<int> c;
    for (int j = a; j <= k; ++j) { c.push_back(j); sn_of[j] = (int)cols.size(); }
    cols.push_back(std::move(c));
    bel.push_back(st[k]);
    ++k;
  }
  const int ns0 = (int)cols.size();
  std::vector<int> par(ns0, -1);
  for (int S = 0; S < ns0; ++S)
    if (!bel[S].empty()) par[S] = sn_of[bel[S][0]];
  // relaxed amalgamation, children before parents: merging S into its parent gives every column of S the parent's
  // structure: s_S (s_P + t_P - t_S) explicit zeros
  std::vector<char> alive(ns0, 1);
  auto find = [&](int p) { while (p >= 0 && !alive[p]) p = par[p]; return p; };
  for (int S = 0; S < ns0; ++S) {
    const int P = find(par[S]);
    if (P < 0) continue;
    const long long sS = (long long)cols[S].size(), tS = (long long)bel[S].size(), sP = (long long)cols[P].size(), tP = (long long)bel[P].size();
    const long long added = sS * (sP + tP - tS);
    const long long panel = sS * (sS - 1) / 2 + sS * tS;
    if (sS + sP <= scap && (added <= zabs || (double)added <= zrel * (double)panel)) {
      std::vector<int> merged = cols[S];
      merged.insert(merged.end(), cols[P].begin(), cols[P].end());
      cols[P] = std::move(merged);
      alive[S] = 0;
      par[S] = P;
    }
  }
  // renumber: DFS over the surviving supernode tree, a supernode's columns consecutive (children's columns merged in
  // front of the parent's own, which keeps every column after its descendants)
  std::vector<int> live;
  for (int S = 0; S < ns0; ++S)
    if (alive[S]) live.push_back(S);
  std::vector<std::vector<int>> kids(ns0);
  std::vector<int> roots;
  for (int S : live) {
    const int P = find(par[S]);
    par[S] = P;
    if (P < 0) roots.push_back(S);
    else kids[P].push_back(S);
  }
  std::vector<int> sn_order;
  {
    std::vector<std::pair<int, size_t>> stack;
    for (int r : roots) {
      stack.push_back({r, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second < kids[top.first].size()) {
          const int c = kids[top.first][top.second++];
          stack.push_back({c, 0});
        } else {
          sn_order.push_back(top.first);
          stack.pop_back();
        }
      }
    }
  }
  std::vector<int> newidx(n, -1);  // old permuted index -> final permuted index
  pl.perm.resize(n);
  std::vector<int> new_of_sn(ns0, -1);
  int next = 0;
  for (size_t q = 0; q < sn_order.size(); ++q) {
    const int S = sn_order[q];
    new_of_sn[S] = (int)q;
    pl.c0.push_back(next);
    pl.s.push_back((int)cols[S].size());
    for (int c : cols[S]) {
      newidx[c] = next;
      pl.perm[next] = perm_in[c];
      ++next;
    }
  }
  const int ns = (int)sn_order.size();
  pl.below.resize(ns);
  pl.parent.assign(ns, -1);
  pl.level.assign(ns, 0);
  for (int q = 0; q < ns; ++q) {
    const int S = sn_order[q];
    for (int r : bel[S]) pl.below[q].push_back(newidx[r]);
    std::sort(pl.below[q].begin(), pl.below[q].end());
    if (par[S] >= 0) pl.parent[q] = new_of_sn[par[S]];
  }
  for (int q = 0; q < ns; ++q)
    if (pl.parent[q] >= 0) pl.level[pl.parent[q]] = std::max(pl.level[pl.parent[q]], pl.level[q] + 1);
  // a supernode must come after everything that pushes into it: levels are recomputed from the below rows, not only the tree parent
  for (int q = 0; q < ns; ++q) pl.nlev = std::max(pl.nlev, pl.level[q] + 1);
  for (int q = 0; q < ns; ++q) {
    const long long s = pl.s[q], t = (long long)pl.below[q].size();
    pl.nW += (int)(s * (s - 1) / 2 + s * t);
    for (long long k = 0; k < s; ++k) {
      const long long L = s - 1 - k + t;
      pl.flops += L * (L + 1) / 2;
    }
  }
  // cost model (cycles per solve of the on-chip kernel, coarse): 30 iterations of 4 nlev sweep stages + work, 2 factorisations
  long long maxs = 0;
  for (int q = 0; q < ns; ++q) maxs = std::max<long long>(maxs, pl.s[q]);
  const double sweep = 4.0 * pl.nlev * 220.0 + 2.0 * pl.nW / 256.0 * 8.0;
  const double factor = n * 150.0 + ns * 300.0 + maxs * 200.0 + pl.flops / 256.0 * 10.0;
  pl.cost = 34.0 * sweep + 2.0 * factor;
  return pl;
}

}  // namespace cta_detail

// shared-memory footprint of the on-chip kernel per CTA: scalars of type T (factor slots + 1/D, Abar, 6 n-vectors,
// 8 m-vectors, reduction scratch) followed by the integer tables (CtaSymbolic::smem_ints)
constexpr int kCtaNV = 6;
constexpr int kCtaMV = 8;
constexpr int kCtaRed = 128;
constexpr int kCtaPivot = 16;  // scalars of a parked 4 x 4 pivot-block factorisation (10 + 4, padded)
enum CtaIntTable { kI_Acol = 0, kI_ATword, kI_Arowptr, kI_ATptr, kI_rlist, kI_diagoff, kI_prow, kI_sntab, kI_stages, kI_fdiag, kI_bwd, kI_pushout,
                   kI_pushtask, kI_extptr, kI_extword, kI_facrounds, kI_facext, kI_facwarp, kI_count };
inline size_t cta_round4(size_t v) { return (v + 3) / 4 * 4; }
inline size_t cta_smem_scalars(const CtaSymbolic& S)
{
  return cta_round4((size_t)S.nW + (size_t)S.np) + cta_round4(S.nnzA) + (size_t)kCtaNV * cta_round4(S.np) +
         (size_t)kCtaMV * cta_round4(S.m) + kCtaRed + 2 * (kCtaNT / 32) * kCtaPivot;
}
inline size_t cta_smem_bytes(const CtaSymbolic& S, size_t scalar) { return cta_smem_scalars(S) * scalar + S.smem_ints.size() * sizeof(int); }

// md_perm: the minimum-degree order already computed by sparse_analyze (perm[new] = old); lpad: log2 of the row padding
inline bool cta_analyze(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                        const int32_t* A_colidx, const std::vector<int>& md_perm, CtaSymbolic& S, int lpad = 2, int force_ordering = -1,
                        bool big_tables_global = false)
{
  using namespace cta_detail;
  S = CtaSymbolic();
  S.n = n;
  S.m = m;
  S.lpad = lpad;
  S.pad = 1 << lpad;
  const int pad = S.pad;
  S.nnzP = P_colptr[n];
  S.nnzA = m > 0 ? A_rowptr[m] : 0;
  if (n >= 0x7fff || m >= 0xffff || S.nnzA >= 0xffff) { S.error = "too large for 16-bit schedule fields"; return false; }
  // ---- pattern of M (original indices), as in sparse_analyze
  Adj adj(n);
  for (int j = 0; j < n; ++j)
    for (int e = P_colptr[j]; e < P_colptr[j + 1]; ++e) {
      const int r = P_rowidx[e];
      if (j > r) { adj[r].insert(j); adj[j].insert(r); }
    }
  for (int i = 0; i < m; ++i)
    for (int e1 = A_rowptr[i]; e1 < A_rowptr[i + 1]; ++e1)
      for (int e2 = e1 + 1; e2 < A_rowptr[i + 1]; ++e2) {
        adj[A_colidx[e1]].insert(A_colidx[e2]);
        adj[A_colidx[e2]].insert(A_colidx[e1]);
      }
  // ---- candidate orderings
  constexpr int kZabs = 32;
  constexpr double kZrel = 0.25;
  constexpr int kScap = 64;
  SnPlan best;
  bool have = false;
  for (int ord = 0; ord < 2; ++ord) {
    if (force_ordering >= 0 && ord != force_ordering) continue;
    std::vector<int> p;
    if (ord == 0) p = md_perm;
    else {
      NestedDissection nd(adj, 32);
      std::vector<int> all(n);
      for (int v = 0; v < n; ++v) all[v] = v;
      nd.run(all);
      p = nd.order;
    }
    if ((int)p.size() != n) { S.error = "internal: ordering is not a permutation"; return false; }
    SnPlan pl = plan_supernodes(adj, postorder(adj, p), kZabs, kZrel, kScap);
    if (!have || pl.cost < best.cost) { best = std::move(pl); S.ordering = ord; have = true; }
  }
  const SnPlan& pl = best;  // compact permuted indices 0 .. n - 1 (pl.perm[compact] = original)
  S.ns = (int)pl.s.size();
  S.flops = pl.flops;
  S.nnzL_true = pl.nnzL_true;
  // ---- padded ids and slot layout
  std::vector<int> pid(n, -1), csn(n, -1);  // compact -> padded id / supernode
  {
    int next = 0, slot = 0;
    for (int q = 0; q < S.ns; ++q) {
      const int s = pl.s[q], sp = (s + pad - 1) / pad * pad, t = (int)pl.below[q].size();
      S.sn_c0.push_back(next);
      S.sn_s.push_back(s);
      S.sn_sp.push_back(sp);
      S.sn_t.push_back(t);
      for (int c = 0; c < s; ++c) { pid[pl.c0[q] + c] = next + c; csn[pl.c0[q] + c] = q; }
      next += sp;
      S.sn_dbase.push_back(slot);
      slot += cta_diag_off(s, lpad);
      S.sn_bbase.push_back(slot);
      slot += sp * t;
      S.smax = std::max(S.smax, s);
    }
    S.np = next;
    S.nW = slot;
  }
  if ((long long)S.nW + (long long)S.np >= 0xffff) { S.error = "factor too large for 16-bit schedule fields"; return false; }
  S.perm.assign(S.np, -1);
  S.iperm.assign(n, -1);
  S.col_sn.assign(S.np, -1);
  for (int k = 0; k < n; ++k) {
    S.perm[pid[k]] = pl.perm[k];
    S.iperm[pl.perm[k]] = pid[k];
    S.col_sn[pid[k]] = csn[k];
  }
  for (int q = 0; q < S.ns; ++q) {
    S.sn_rbase.push_back((int)S.rlist.size());
    for (int r : pl.below[q]) S.rlist.push_back(pid[r]);
  }
  std::vector<std::unordered_map<int, int>> rowpos(S.ns);  // padded row id -> position in the below block
  for (int q = 0; q < S.ns; ++q)
    for (int r = 0; r < S.sn_t[q]; ++r) rowpos[q][S.rlist[S.sn_rbase[q] + r]] = r;
  auto target = [&](int pr, int pc) -> int {  // PADDED ids, any order
    const int lo = std::min(pr, pc), hi = std::max(pr, pc);
    const int q = S.col_sn[lo];
    const int j = lo - S.sn_c0[q];
    if (hi < S.sn_c0[q] + S.sn_s[q]) return S.sn_dbase[q] + cta_diag_off(hi - S.sn_c0[q], lpad) + j;
    const auto f = rowpos[q].find(hi);
    return f == rowpos[q].end() ? -2 : S.sn_bbase[q] + f->second * S.sn_sp[q] + j;
  };
  {  // the padded structure must contain the true one
    const auto st = symbolic(adj, pl.perm);
    for (int k = 0; k < n; ++k)
      for (int r : st[k])
        if (target(pid[r], pid[k]) == -2) { S.error = "internal: supernodal structure does not cover L"; return false; }
  }
  // ---- P, A in padded ids, assembly targets, mirrors
  S.P_tgt.assign(S.nnzP, -1);
  S.P_rowp.resize(S.nnzP);
  S.P_colp.resize(S.nnzP);
  for (int j = 0; j < n; ++j)
    for (int e = P_colptr[j]; e < P_colptr[j + 1]; ++e) {
      const int r = P_rowidx[e];
      S.P_rowp[e] = S.iperm[r];
      S.P_colp[e] = S.iperm[j];
      if (j >= r) {
        S.P_tgt[e] = target(S.iperm[r], S.iperm[j]);
        if (S.P_tgt[e] == -2) { S.error = "internal: P entry outside the symbolic factor"; return false; }
      }
    }
  S.A_rowptr.assign(m + 1, 0);
  if (m > 0) S.A_rowptr.assign(A_rowptr, A_rowptr + m + 1);
  S.A_col.resize(S.nnzA);
  for (int e = 0; e < S.nnzA; ++e) S.A_col[e] = S.iperm[A_colidx[e]];
  auto build_rows = [&](int nrows, const std::vector<std::array<int, 3>>& trip, std::vector<int>& ptr, std::vector<int>& col,
                        std::vector<int>& slt) {
    ptr.assign(nrows + 1, 0);
    for (const auto& t : trip) ptr[t[0] + 1]++;
    for (int r = 0; r < nrows; ++r) ptr[r + 1] += ptr[r];
    col.resize(trip.size());
    slt.resize(trip.size());
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (const auto& t : trip) {
      const int p = fill[t[0]]++;
      col[p] = t[1];
      slt[p] = t[2];
    }
  };
  {
    std::vector<std::array<int, 3>> trip;
    for (int i = 0; i < m; ++i)
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) trip.push_back({S.A_col[e], i, e});
    build_rows(S.np, trip, S.AT_ptr, S.AT_row, S.AT_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) trip.push_back({S.P_rowp[e], S.P_colp[e], e});
    build_rows(S.np, trip, S.PR_ptr, S.PR_col, S.PR_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) {
      if (S.P_tgt[e] < 0) continue;
      trip.push_back({S.P_rowp[e], S.P_colp[e], e});
      if (S.P_rowp[e] != S.P_colp[e]) trip.push_back({S.P_colp[e], S.P_rowp[e], e});
    }
    build_rows(S.np, trip, S.PS_ptr, S.PS_col, S.PS_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) trip.push_back({S.P_colp[e], S.P_rowp[e], e});
    std::vector<int> dummy;
    build_rows(S.np, trip, S.PC_ptr, dummy, S.PC_slot);
  }
  // ---- assembly pairs, rows coloured greedily (a colour = rows with pairwise disjoint column sets)
  {
    std::vector<int> color(m, -1);
    std::vector<std::vector<char>> used;  // per colour: columns touched
    for (int i = 0; i < m; ++i) {
      int c = 0;
      for (;; ++c) {
        if (c == (int)used.size()) used.emplace_back(S.np, 0);
        bool clash = false;
        for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1] && !clash; ++e) clash = used[c][S.A_col[e]] != 0;
        if (!clash) break;
      }
      color[i] = c;
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) used[c][S.A_col[e]] = 1;
    }
    const int ncol = (int)used.size();
    for (int c = 0; c < ncol; ++c) {
      std::vector<int> ab, tr;
      for (int i = 0; i < m; ++i) {
        if (color[i] != c) continue;
        for (int e1 = S.A_rowptr[i]; e1 < S.A_rowptr[i + 1]; ++e1)
          for (int e2 = e1; e2 < S.A_rowptr[i + 1]; ++e2) {
            const int t = target(S.A_col[e1], S.A_col[e2]);
            if (t == -2) { S.error = "internal: A^T A entry outside the symbolic factor"; return false; }
            if (e1 != e2 && S.A_col[e1] == S.A_col[e2]) { S.error = "duplicate column index in a row of A"; return false; }
            ab.push_back((int)((uint32_t)e1 | ((uint32_t)e2 << 16)));
            tr.push_back((int)((uint32_t)t | ((uint32_t)i << 16)));
          }
      }
      const int p1 = (int)ab.size();
      for (int p = 0; p < p1; p += kCtaNT) {
        for (int k = 0; k < kCtaNT; ++k) {
          S.asm_round_desc.push_back(p + k < p1 ? ab[p + k] : 0);
          S.asm_round_desc.push_back(p + k < p1 ? tr[p + k] : -1);
        }
        S.asm_round_sync.push_back(p + kCtaNT >= p1 ? 1 : 0);
      }
    }
  }
  // ---- external update pairs of every supernode: a | b << 8 | target << 16
  S.ext_ptr.assign(1, 0);
  for (int q = 0; q < S.ns; ++q) {
    const int t = S.sn_t[q];
    if (t > 255) { S.error = "supernode with more than 255 rows below"; return false; }
    const int* R = S.rlist.data() + S.sn_rbase[q];
    for (int a = 0; a < t; ++a)
      for (int b = a; b < t; ++b) {
        const int tg = target(R[b], R[a]);
        if (tg == -2) { S.error = "internal: fill entry outside the supernodal structure"; return false; }
        S.ext_word.push_back((int)((uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)tg << 16)));
      }
    S.ext_ptr.push_back((int)S.ext_word.size());
  }
  // ---- levels.  A supernode's level is above every supernode that has one of its columns among its below rows.
  {
    std::vector<int> lev(S.ns, 0);
    for (int q = 0; q < S.ns; ++q)
      for (int r = 0; r < S.sn_t[q]; ++r) {
        const int a = S.col_sn[S.rlist[S.sn_rbase[q] + r]];
        lev[a] = std::max(lev[a], lev[q] + 1);  // q < a always (below rows are later columns), so one ascending pass suffices
      }
    S.sn_level = lev;
    S.nlev = 0;
    for (int q = 0; q < S.ns; ++q) S.nlev = std::max(S.nlev, lev[q] + 1);
  }
  for (int q = 0; q < S.ns; ++q) {
    const int row[8] = {S.sn_c0[q], S.sn_s[q], S.sn_t[q], S.sn_dbase[q], S.sn_bbase[q], S.sn_rbase[q], S.sn_sp[q], (int)S.prow.size()};
    S.sn_tab.insert(S.sn_tab.end(), row, row + 8);
    for (int i = 0; i < S.sn_s[q]; ++i) S.prow.push_back(S.sn_dbase[q] + cta_diag_off(i, lpad));
    for (int r = 0; r < S.sn_t[q]; ++r) S.prow.push_back(S.sn_bbase[q] + r * S.sn_sp[q]);
  }
  for (int i = 0; i <= S.smax; ++i) S.diag_off.push_back(cta_diag_off(i, lpad));
  // ---- sweep stages
  auto add_stage = [&](int kind, int first, int count, int longest) {  // longest: the most vector loads one output needs
    if (count == 0) return;
    int lg = 0;
    while (lg < 5 && (2 << lg) * count <= kCtaNT && (1 << lg) < longest) ++lg;
    S.stages.push_back(kind);
    S.stages.push_back(first);
    S.stages.push_back(count);
    S.stages.push_back(lg);
  };
  std::vector<int> bfirst(S.nlev + 1, 0);
  for (int L = 0; L < S.nlev; ++L) {
    bfirst[L] = (int)S.bwd_out.size();
    for (int q = 0; q < S.ns; ++q)
      if (S.sn_level[q] == L)
        for (int kc = 0; kc < S.sn_s[q]; kc += pad) S.bwd_out.push_back((int)((uint32_t)(S.sn_c0[q] + kc) | ((uint32_t)q << 16)));
  }
  bfirst[S.nlev] = (int)S.bwd_out.size();
  for (int L = 0; L < S.nlev; ++L) {  // forward: diagonal blocks of level L, then their pushes into the rows above
    int first = (int)S.fdiag_out.size(), longest = 0;
    for (int q = 0; q < S.ns; ++q)
      if (S.sn_level[q] == L)
        for (int i = 0; i < S.sn_s[q]; ++i) {
          S.fdiag_out.push_back((int)((uint32_t)(S.sn_c0[q] + i) | ((uint32_t)q << 16)));
          longest = std::max(longest, (i + pad - 1) / pad);
        }
    add_stage(kStageFwdDiag, first, (int)S.fdiag_out.size() - first, longest);
    std::vector<std::array<int, 3>> tasks;  // dst, supernode, row of its below block
    for (int q = 0; q < S.ns; ++q) {
      if (S.sn_level[q] != L) continue;
      for (int r = 0; r < S.sn_t[q]; ++r) tasks.push_back({S.rlist[S.sn_rbase[q] + r], q, r});
    }
    std::stable_sort(tasks.begin(), tasks.end(), [](const std::array<int, 3>& a, const std::array<int, 3>& b) { return a[0] < b[0]; });
    first = (int)S.push_out.size();
    int count = 0;
    longest = 0;
    for (size_t k = 0; k < tasks.size();) {
      size_t e = k;
      int len = 0;
      while (e < tasks.size() && tasks[e][0] == tasks[k][0]) { len += S.sn_sp[tasks[e][1]] / pad; ++e; }
      S.push_out.push_back((int)((uint32_t)tasks[k][0] | ((uint32_t)S.push_task.size() << 16)));
      for (size_t j = k; j < e; ++j) S.push_task.push_back((int)((uint32_t)tasks[j][1] | ((uint32_t)tasks[j][2] << 16)));
      longest = std::max(longest, len);
      ++count;
      k = e;
    }
    if (count > 0) S.push_out.push_back((int)(0xffffu | ((uint32_t)S.push_task.size() << 16)));  // sentinel: end of the last output's tasks
    if (S.push_task.size() >= 0xffff) { S.error = "too many push tasks for 16-bit schedule fields"; return false; }
    add_stage(kStageFwdPush, first, count, longest);
  }
  for (int L = S.nlev - 1; L >= 0; --L) {  // backward: pull from the rows above, then the transposed diagonal blocks
    int tmax = 0, smaxL = 0;
    for (int q = 0; q < S.ns; ++q)
      if (S.sn_level[q] == L) { tmax = std::max(tmax, S.sn_t[q]); smaxL = std::max(smaxL, S.sn_s[q]); }
    add_stage(kStageBwdPull, bfirst[L], bfirst[L + 1] - bfirst[L], std::max(tmax, 1));
    add_stage(kStageBwdDiag, bfirst[L], bfirst[L + 1] - bfirst[L], std::max(smaxL - 1, 1));
  }
  // ---- factorisation rounds: supernodes of one level advance column by column together, each with its own warps
  {
    constexpr int NWARP = kCtaNT / 32;
    for (int L = 0; L < S.nlev; ++L) {
      std::vector<std::pair<long long, int>> sn;  // (work, supernode)
      for (int q = 0; q < S.ns; ++q) {
        if (S.sn_level[q] != L) continue;
        long long wk = 0;
        for (long long kc = 0; kc < S.sn_s[q]; ++kc) wk += (S.sn_s[q] - kc - 1) * (S.sn_s[q] - kc + S.sn_t[q]) + 64;
        sn.push_back({wk, q});
      }
      std::sort(sn.begin(), sn.end(), [](const std::pair<long long, int>& a, const std::pair<long long, int>& b) { return a.first > b.first; });
      for (size_t k0 = 0; k0 < sn.size(); k0 += NWARP) {
        const int cnt = (int)std::min<size_t>(NWARP, sn.size() - k0);
        std::vector<int> nw(cnt, 1), cap(cnt, 1);
        for (int k = 0; k < cnt; ++k) {  // no more warps than vector elements of the first trailing panel, one per lane
          const int q = sn[k0 + k].second;
          cap[k] = std::max(1, std::min(NWARP, ((S.sn_s[q] + S.sn_t[q]) * (S.sn_sp[q] / pad) + 31) / 32));
        }
        for (int extra = NWARP - cnt; extra > 0; --extra) {
          int bi = -1;
          for (int k = 0; k < cnt; ++k)
            if (nw[k] < cap[k] && (bi < 0 || sn[k0 + k].first * nw[bi] > sn[k0 + bi].first * nw[k])) bi = k;
          if (bi < 0) break;
          nw[bi]++;
        }
        int totw = 0;
        for (int k = 0; k < cnt; ++k) totw += nw[k];
        int maxs = 0;
        for (int k = 0; k < cnt; ++k) maxs = std::max(maxs, S.sn_s[sn[k0 + k].second]);
        S.fac_rounds.push_back(maxs);
        S.fac_rounds.push_back(cnt | (totw << 16));  // supernodes | warps with an entry << 16
        S.fac_rounds.push_back((int)S.fac_ext.size());
        S.fac_rounds.push_back((int)S.fac_warp.size());
        for (int k = 0; k < cnt; ++k) {
          S.fac_ext.push_back(sn[k0 + k].second);
          for (int r = 0; r < nw[k]; ++r) S.fac_warp.push_back((int)((uint32_t)sn[k0 + k].second | ((uint32_t)r << 16) | ((uint32_t)nw[k] << 24)));
        }
      }
    }
  }
  // ---- tables for shared memory
  auto pack16 = [](const std::vector<int>& v, std::vector<int>& out) {
    out.assign((v.size() + 1) / 2, 0);
    for (size_t k = 0; k < v.size(); ++k) out[k / 2] |= (int)((uint32_t)(v[k] & 0xffff) << (16 * (k & 1)));
  };
  pack16(S.A_col, S.A_col16);
  pack16(S.A_rowptr, S.A_rowptr16);
  pack16(S.AT_ptr, S.AT_ptr16);
  pack16(S.rlist, S.rlist16);
  pack16(S.diag_off, S.diag_off16);
  pack16(S.prow, S.prow16);
  S.AT_word.resize(S.nnzA);
  for (int e = 0; e < S.nnzA; ++e) S.AT_word[e] = (int)((uint32_t)S.AT_slot[e] | ((uint32_t)S.AT_row[e] << 16));
  {
    const std::vector<int>* tabs[kI_count] = {&S.A_col16, &S.AT_word, &S.A_rowptr16, &S.AT_ptr16, &S.rlist16, &S.diag_off16, &S.prow16, &S.sn_tab, &S.stages,
                                              &S.fdiag_out, &S.bwd_out, &S.push_out, &S.push_task, &S.ext_ptr, &S.ext_word, &S.fac_rounds,
                                              &S.fac_ext, &S.fac_warp};
    // big_tables_global: the column mirror of A and the external-update pairs stay in global memory (read through L1), which
    // brings the fp32 working set under half an SM's shared memory: two CTAs per SM
    S.big_tables_global = big_tables_global;
    for (int k = 0; k < kI_count; ++k) {
      S.smem_off.push_back((int)S.smem_ints.size());
      if (big_tables_global && (k == kI_ATword || k == kI_extword)) continue;
      S.smem_ints.insert(S.smem_ints.end(), tabs[k]->begin(), tabs[k]->end());
      while (S.smem_ints.size() % 4) S.smem_ints.push_back(0);
    }
  }
  return true;
}

inline int cta_lo16(int x) { return (int)((uint32_t)x & 0xffffu); }
inline int cta_hi16(int x) { return (int)((uint32_t)x >> 16); }

// Host execution of the schedules on one instance (double): assembles M = shift I + triu-mirrored Pbar + A^T diag(w) A from
// the tables, factorises, inverts the diagonal blocks and runs the sweeps as the device kernel does (stage by stage,
// sequentially, vector padding included).  Test infrastructure for the CPU suite: validates every table without a GPU.
struct CtaHostExec
{
  const CtaSymbolic& S;
  std::vector<double> W;  // nW + np
  explicit CtaHostExec(const CtaSymbolic& s) : S(s), W((size_t)s.nW + (size_t)s.np, 0.0) {}
  static int u16at(const std::vector<int>& packed, int k) { return (int)(((uint32_t)packed[k / 2] >> (16 * (k & 1))) & 0xffffu); }
  int prow(int q, int p) const { return u16at(S.prow16, S.sn_tab[8 * q + 7] + p); }  // slot of panel row p of supernode q

  // Pv: values of P in pattern order (already scaled), Av: values of A in CSR order, w: row weights
  void assemble(double shift, const double* Pv, const double* Av, const double* w)
  {
    std::fill(W.begin(), W.end(), 0.0);
    for (int e = 0; e < S.n; ++e) {  // the diagonal of every real column, through the table the kernel uses
      const int k = cta_lo16(S.fdiag_out[e]), q = cta_hi16(S.fdiag_out[e]), i = k - S.sn_tab[8 * q];
      W[prow(q, i) + i] = shift;
    }
    for (int e = 0; e < S.nnzP; ++e)
      if (S.P_tgt[e] >= 0) W[S.P_tgt[e]] += Pv[e];
    for (size_t r = 0; r < S.asm_round_sync.size(); ++r)
      for (int k = 0; k < kCtaNT; ++k) {
        const int ab = S.asm_round_desc[2 * (r * kCtaNT + k)], tr = S.asm_round_desc[2 * (r * kCtaNT + k) + 1];
        if (tr == -1) continue;
        W[cta_lo16(tr)] += (w[cta_hi16(tr)] * Av[cta_lo16(ab)]) * Av[cta_hi16(ab)];
      }
  }
  // the factorisation rounds, executed sequentially; false on a non-positive pivot or an inconsistent round table
  bool factor()
  {
    const int nW = S.nW, pad = S.pad;
    bool ok = true;
    std::vector<int> done(S.ns, 0);
    for (size_t r = 0; r < S.fac_rounds.size() / 4; ++r) {
      const int cnt = cta_lo16(S.fac_rounds[4 * r + 1]), e0 = S.fac_rounds[4 * r + 2], w0 = S.fac_rounds[4 * r + 3];
      int nwarps = 0;
      for (int k = 0; k < cnt; ++k) {
        const int q = S.fac_ext[e0 + k];
        const int c0 = S.sn_tab[8 * q], s = S.sn_tab[8 * q + 1], t = S.sn_tab[8 * q + 2], db = S.sn_tab[8 * q + 3], bb = S.sn_tab[8 * q + 4],
                  sp = S.sn_tab[8 * q + 6];
        if (s > S.fac_rounds[4 * r]) ok = false;
        done[q]++;
        const int nw = (int)((uint32_t)S.fac_warp[w0 + nwarps] >> 24);
        for (int rr = 0; rr < nw; ++rr) {
          const uint32_t wd = (uint32_t)S.fac_warp[w0 + nwarps + rr];
          if ((int)(wd & 0xffff) != q || (int)((wd >> 16) & 0xff) != rr || (int)(wd >> 24) != nw) ok = false;
        }
        nwarps += nw;
        for (int p = 0; p < s + t; ++p)
          if (prow(q, p) != (p < s ? db + cta_diag_off(p, S.lpad) : bb + (p - s) * sp)) ok = false;
        for (int kc = 0; kc < s; ++kc) {
          const double d = W[prow(q, kc) + kc];
          if (!(d > 0)) ok = false;
          const double dinv = 1.0 / d;
          W[nW + c0 + kc] = dinv;
          // the kernel's uniform update: every row p > kc, every whole vector of columns from the one holding kc + 1 on; columns
          // <= kc and >= s see a zero multiplier, the padding of a row (columns > p) collects garbage that the cleanup removes
          for (int p = kc + 1; p < s + t; ++p) {
            const double vi = W[prow(q, p) + kc];
            const int jend = (p < s) ? (p / pad + 1) * pad : sp;
            for (int j = (kc + 1) / pad * pad; j < jend; ++j) {
              const double vjs = (j > kc && j < s) ? W[prow(q, j) + kc] * dinv : 0.0;
              if (j > kc) W[prow(q, p) + j] -= vi * vjs;
            }
          }
        }
      }
      if (nwarps != cta_hi16(S.fac_rounds[4 * r + 1]) || nwarps > kCtaNT / 32) ok = false;
      for (int k = 0; k < cnt; ++k) {  // external updates with the unscaled columns: U D^-1 U^T
        const int q = S.fac_ext[e0 + k];
        const int c0 = S.sn_tab[8 * q], bb = S.sn_tab[8 * q + 4], sp = S.sn_tab[8 * q + 6];
        for (int p = S.ext_ptr[q]; p < S.ext_ptr[q + 1]; ++p) {
          const uint32_t wd = (uint32_t)S.ext_word[p];
          const int a = wd & 0xff, b = (wd >> 8) & 0xff, tg = wd >> 16;
          double acc = 0;
          for (int c = 0; c < sp; ++c) acc += (W[bb + a * sp + c] * W[nW + c0 + c]) * W[bb + b * sp + c];  // padding: zeros, 1 / D of a hole: zero
          W[tg] -= acc;
        }
      }
    }
    for (int q = 0; q < S.ns; ++q)
      if (done[q] != 1) ok = false;
    // scale the panels, L = (unscaled columns) D^-1, and clear the diagonal and the padding of every row of a diagonal block
    for (int q = 0; q < S.ns; ++q) {
      const int c0 = S.sn_c0[q], s = S.sn_s[q], t = S.sn_t[q], sp = S.sn_sp[q];
      for (int p = 0; p < s + t; ++p) {
        const int jend = (p < s) ? (p / pad + 1) * pad : sp;
        for (int j = 0; j < jend; ++j) {
          double& e = W[prow(q, p) + j];
          e = (p < s && j >= p) ? 0.0 : e * W[nW + c0 + j];
        }
      }
    }
    // invert the unit-lower diagonal blocks in place, right-looking; stored: X' = -(strict lower part of L_SS^-1).
    // Step k: X'(i, c) -= L(i, k) X'(k, c) for i > k over the whole vectors of row k
    for (int k = 1; k + 1 < S.smax; ++k)
      for (int q = 0; q < S.ns; ++q) {
        const int s = S.sn_s[q];
        for (int i = k + 1; i < s; ++i) {
          const double lik = W[prow(q, i) + k];
          for (int c = 0; c < (k + pad - 1) / pad * pad; ++c) W[prow(q, i) + c] -= lik * W[prow(q, k) + c];
        }
      }
    return ok;
  }
  // v (padded ids, length np, holes zero) <- (L D L^T)^-1 v.  Every dot product runs over whole padded rows, as the vector
  // loads of the kernel do: the padding must hold zeros.
  void solve(std::vector<double>& v) const
  {
    const int np = S.np, nW = S.nW, pad = S.pad;
    std::vector<double> y(np, 0.0);
    for (size_t st = 0; st < S.stages.size() / 4; ++st) {
      const int kind = S.stages[4 * st], first = S.stages[4 * st + 1], count = S.stages[4 * st + 2];
      std::vector<std::pair<int, double>> writes;
      for (int o = first; o < first + count; ++o) {
        if (kind == kStageFwdPush) {
          const int dst = cta_lo16(S.push_out[o]), t0 = cta_hi16(S.push_out[o]), t1 = cta_hi16(S.push_out[o + 1]);
          double acc = 0;
          for (int k = t0; k < t1; ++k) {
            const int q = cta_lo16(S.push_task[k]), r = cta_hi16(S.push_task[k]);
            const int c0 = S.sn_tab[8 * q], bb = S.sn_tab[8 * q + 4], sp = S.sn_tab[8 * q + 6];
            for (int c = 0; c < sp; ++c) acc += W[bb + r * sp + c] * y[c0 + c];
          }
          writes.push_back({dst, v[dst] - acc});
        } else if (kind == kStageFwdDiag) {
          const int k = cta_lo16(S.fdiag_out[o]), q = cta_hi16(S.fdiag_out[o]);
          const int c0 = S.sn_tab[8 * q], i = k - c0;
          double acc = 0;
          for (int c = 0; c < (i + pad - 1) / pad * pad; ++c) acc += W[prow(q, i) + c] * v[c0 + c];
          writes.push_back({k, v[k] - acc});
        } else {
          const int k0 = cta_lo16(S.bwd_out[o]), q = cta_hi16(S.bwd_out[o]);
          const int c0 = S.sn_tab[8 * q], s = S.sn_tab[8 * q + 1], t = S.sn_tab[8 * q + 2], bb = S.sn_tab[8 * q + 4], rb = S.sn_tab[8 * q + 5],
                    sp = S.sn_tab[8 * q + 6];
          const int kc = k0 - c0;
          for (int c = 0; c < pad; ++c) {  // whole vectors: the holes come out as zero
            double acc = 0;
            if (kind == kStageBwdPull) {
              for (int r = 0; r < t; ++r) acc += W[bb + r * sp + kc + c] * y[u16at(S.rlist16, rb + r)];
              writes.push_back({k0 + c, W[nW + k0 + c] * y[k0 + c] - acc});
            } else {
              for (int i = kc + 1; i < s; ++i) acc += W[prow(q, i) + kc + c] * v[c0 + i];  // columns >= i of row i: zeros
              writes.push_back({k0 + c, v[k0 + c] - acc});
            }
          }
        }
      }
      std::vector<double>& out = (kind == kStageFwdDiag || kind == kStageBwdDiag) ? y : v;
      for (auto& wv : writes) out[wv.first] = wv.second;
    }
    v = y;
  }
};

}  // namespace sfb
