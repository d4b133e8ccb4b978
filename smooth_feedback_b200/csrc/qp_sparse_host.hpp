// qp_sparse_host.hpp -- host-side symbolic analysis for the batched sparse QP path (shared sparsity pattern).
//
// Replaces what the reference gets from Eigen for QuadraticProgramSparse problems
// (pettni/smooth_feedback @ 9a08971):
//   SimplicialLDLT::analyzePattern   call site include/smooth/feedback/qp_solver.hpp:424   (ordering + symbolic factor)
//   sparse KKT fill                  include/smooth/feedback/qp_solver.hpp:380-397
// The engine does not factorise the (n+m) x (n+m) quasi-definite KKT matrix: the diagonal (2,2) block -1/rho is
// eliminated analytically (as in the dense kernel) and the n x n matrix
//      M = c Sx triu(P) Sx (mirrored) + sigma I + Abar^T R Abar
// is factorised as L D L^T.  Its pattern depends only on the patterns of P and A, which every instance of a batch
// shares (MPC: one OCP structure, many agents / time steps), so ordering, fill, assembly targets and the
// right-looking update schedule are computed ONCE here and uploaded as flat index arrays.
//
// Ordering: greedy minimum degree on the elimination graph (exact degrees; n is a few hundred to a few thousand).

#pragma once

#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

namespace sfb {

struct SparseSymbolic
{
  int n = 0, m = 0, nnzP = 0, nnzA = 0, nnzL = 0;
  std::vector<int> perm, iperm;       // perm[new] = old, iperm[old] = new
  std::vector<int> P_colptr, P_row;   // as given (original indices)
  std::vector<int> P_rowp, P_colp;    // per stored entry: permuted row / column
  std::vector<int> P_tgt;             // per stored entry: slot in W (L slots, then nnzL + i for D_i), -1 if below the diagonal
  std::vector<int> A_rowptr, A_col;   // CSR, column index permuted
  std::vector<int> A_pair_ptr, A_pair_tgt;  // per row: targets of the pairs (a <= b) of its entries, a-major
  std::vector<int> A_pair_ab, F_ab;         // the pair itself, (a << 16) | b, offsets relative to the row / column start
  std::vector<int> L_colptr, L_row;   // strictly lower triangle of L, CSC, permuted indices, rows ascending
  std::vector<int> F_ptr, F_tgt;      // per column k: targets of the pairs (a <= b) of struct(k), a-major
  // gather-form mirrors (no read-modify-write in the hot loops): each lists, per output element, the (input index, value
  // slot) pairs it sums over
  std::vector<int> LR_ptr, LR_col, LR_slot;   // rows of L:           v[k] -= sum W[slot] v[col]        (forward solve)
  std::vector<int> AT_ptr, AT_row, AT_slot;   // columns of A:        out[j] = sum A[slot] in[row]      (Abar^T w)
  std::vector<int> PR_ptr, PR_col, PR_slot;   // rows of P as stored: out[r] = sum P[slot] in[col]      (P x)
  std::vector<int> PS_ptr, PS_col, PS_slot;   // rows of sym(triu P): out[r] = sum Pbar[slot] in[col]   (polish)
  std::vector<int> LB_ptr, LB_row, LB_slot;   // columns of L in REVERSE column order (backward solve as one forward stream)
  // padded sweep schedules (TW == 8 kernel): a sweep is a list of STEPS of exactly kStepWidth entries; a row of L takes
  // ceil(len / kStepWidth) consecutive steps.  meta = (row << 1) | last_step_of_row; col / slot are padded with 0 / -1.
  static constexpr int kStepWidth = 32;
  static constexpr int kStepPad = 4;  // == kSpDepth on the device
  std::vector<int> FS_meta, FS_col, FS_slot;  // forward:  rows ascending, rows without entries skipped
  std::vector<int> BS_meta, BS_col, BS_slot;  // backward: rows descending, every row present (1 / D scaling)
  // padded row / column streams of A for the pipelined SpMV passes (TW < 32): every row of A padded to WR entries, every
  // column to WA entries (col / row index 0 and slot -1 on padding), rows and columns padded to a multiple of 32
  int WR = 0, WA = 0, m_pad = 0, n_pad = 0;   // WR == 0: rows / columns too long, the kernel keeps the generic passes
  std::vector<int> RP_col, RP_slot, ATP_row, ATP_slot;
  std::vector<int> PC_ptr;                    // P_colptr re-indexed by PERMUTED column (entries regrouped in PC_slot)
  std::vector<int> PC_slot;
  long long flops = 0;                // multiply-adds of the numeric factorisation
  std::string error;
};

inline bool sparse_analyze(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                           const int32_t* A_colidx, SparseSymbolic& S)
{
  S = SparseSymbolic();
  S.n = n;
  S.m = m;
  if (n <= 0 || m < 0 || !P_colptr || (m > 0 && !A_rowptr)) { S.error = "bad sizes or null pattern"; return false; }
  // validate BEFORE anything is copied or dereferenced through the caller's arrays
  if (P_colptr[0] != 0 || (m > 0 && A_rowptr[0] != 0)) { S.error = "pattern pointers must start at 0"; return false; }
  for (int j = 0; j < n; ++j)
    if (P_colptr[j + 1] < P_colptr[j]) { S.error = "P_colptr not monotone"; return false; }
  for (int i = 0; i < m; ++i)
    if (A_rowptr[i + 1] < A_rowptr[i]) { S.error = "A_rowptr not monotone"; return false; }
  S.nnzP = P_colptr[n];
  S.nnzA = m > 0 ? A_rowptr[m] : 0;
  if ((S.nnzP > 0 && !P_rowidx) || (S.nnzA > 0 && !A_colidx)) { S.error = "index array is NULL"; return false; }
  for (int e = 0; e < S.nnzP; ++e)
    if (P_rowidx[e] < 0 || P_rowidx[e] >= n) { S.error = "P row index out of range"; return false; }
  for (int e = 0; e < S.nnzA; ++e)
    if (A_colidx[e] < 0 || A_colidx[e] >= n) { S.error = "A column index out of range"; return false; }
  S.P_colptr.assign(P_colptr, P_colptr + n + 1);
  S.P_row.assign(P_rowidx, P_rowidx + S.nnzP);
  S.A_rowptr.assign(m + 1, 0);
  if (m > 0) S.A_rowptr.assign(A_rowptr, A_rowptr + m + 1);
  std::vector<int> A_col_orig(A_colidx, A_colidx + S.nnzA);

  // ---- pattern of M (original indices) ----
  std::vector<std::set<int>> adj(n);
  for (int j = 0; j < n; ++j)
    for (int e = P_colptr[j]; e < P_colptr[j + 1]; ++e) {
      const int r = P_rowidx[e];
      if (j >= r && j != r) { adj[r].insert(j); adj[j].insert(r); }  // only col >= row enters the KKT (qp_solver.hpp:384)
    }
  for (int i = 0; i < m; ++i)
    for (int e1 = A_rowptr[i]; e1 < A_rowptr[i + 1]; ++e1)
      for (int e2 = e1 + 1; e2 < A_rowptr[i + 1]; ++e2) {
        const int a = A_colidx[e1], b = A_colidx[e2];
        if (a == b) { S.error = "duplicate column index in a row of A (pattern must be compressed)"; return false; }
        adj[a].insert(b);
        adj[b].insert(a);
      }

  // ---- greedy minimum-degree ordering + symbolic elimination ----
  S.perm.assign(n, -1);
  S.iperm.assign(n, -1);
  std::vector<std::vector<int>> struct_old(n);  // per elimination step: higher neighbours (original indices)
  {
    std::vector<std::set<int>> g = adj;
    std::vector<char> alive(n, 1);
    std::set<std::pair<int, int>> queue;  // (degree, vertex)
    for (int v = 0; v < n; ++v) queue.insert({(int)g[v].size(), v});
    for (int step = 0; step < n; ++step) {
      const auto it = queue.begin();
      const int v = it->second;
      queue.erase(it);
      alive[v] = 0;
      S.perm[step] = v;
      S.iperm[v] = step;
      struct_old[step].assign(g[v].begin(), g[v].end());
      const std::vector<int>& nb = struct_old[step];
      for (int a : nb) {
        queue.erase({(int)g[a].size(), a});
        g[a].erase(v);
        for (int b : nb)
          if (b != a) g[a].insert(b);
        queue.insert({(int)g[a].size(), a});
      }
      g[v].clear();
    }
  }
  // L in CSC over permuted indices
  S.L_colptr.assign(n + 1, 0);
  std::vector<std::vector<int>> st(n);
  for (int k = 0; k < n; ++k) {
    st[k].reserve(struct_old[k].size());
    for (int v : struct_old[k]) st[k].push_back(S.iperm[v]);
    std::sort(st[k].begin(), st[k].end());
    S.L_colptr[k + 1] = S.L_colptr[k] + (int)st[k].size();
  }
  S.nnzL = S.L_colptr[n];
  S.L_row.resize(S.nnzL);
  std::vector<std::unordered_map<int, int>> slot(n);  // slot[col][row] -> index into L values
  for (int k = 0; k < n; ++k) {
    slot[k].reserve(st[k].size() * 2 + 1);
    for (size_t t = 0; t < st[k].size(); ++t) {
      S.L_row[S.L_colptr[k] + t] = st[k][t];
      slot[k][st[k][t]] = S.L_colptr[k] + (int)t;
    }
  }
  auto target = [&](int pr, int pc) -> int {  // permuted indices, any order
    if (pr == pc) return S.nnzL + pr;
    const int lo = std::min(pr, pc), hi = std::max(pr, pc);
    const auto f = slot[lo].find(hi);
    return f == slot[lo].end() ? -2 : f->second;
  };

  // ---- assembly targets ----
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
  S.A_col.resize(S.nnzA);
  for (int e = 0; e < S.nnzA; ++e) S.A_col[e] = S.iperm[A_col_orig[e]];
  S.A_pair_ptr.assign(m + 1, 0);
  for (int i = 0; i < m; ++i) {
    const int k = S.A_rowptr[i + 1] - S.A_rowptr[i];
    S.A_pair_ptr[i + 1] = S.A_pair_ptr[i] + k * (k + 1) / 2;
  }
  S.A_pair_tgt.resize(S.A_pair_ptr[m]);
  for (int i = 0; i < m; ++i) {
    int p = S.A_pair_ptr[i];
    for (int e1 = S.A_rowptr[i]; e1 < S.A_rowptr[i + 1]; ++e1)
      for (int e2 = e1; e2 < S.A_rowptr[i + 1]; ++e2) {
        const int t = target(S.A_col[e1], S.A_col[e2]);
        if (t == -2) { S.error = "internal: A^T A entry outside the symbolic factor"; return false; }
        S.A_pair_ab.push_back(((e1 - S.A_rowptr[i]) << 16) | (e2 - S.A_rowptr[i]));
        S.A_pair_tgt[p++] = t;
      }
    if (S.A_rowptr[i + 1] - S.A_rowptr[i] > 0xffff) { S.error = "more than 65535 entries in a row of A"; return false; }
  }

  // ---- gather-form mirrors ----
  auto build_rows = [&](int nrows, const std::vector<std::array<int, 3>>& trip, std::vector<int>& ptr, std::vector<int>& col,
                        std::vector<int>& slt) {  // trip = (row, col, slot), stable order within a row
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
    for (int k = 0; k < n; ++k)
      for (int e = S.L_colptr[k]; e < S.L_colptr[k + 1]; ++e) trip.push_back({S.L_row[e], k, e});
    build_rows(n, trip, S.LR_ptr, S.LR_col, S.LR_slot);
    trip.clear();
    for (int k = n - 1; k >= 0; --k)
      for (int e = S.L_colptr[k]; e < S.L_colptr[k + 1]; ++e) trip.push_back({n - 1 - k, S.L_row[e], e});
    build_rows(n, trip, S.LB_ptr, S.LB_row, S.LB_slot);
    trip.clear();
    for (int i = 0; i < m; ++i)
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) trip.push_back({S.A_col[e], i, e});
    build_rows(n, trip, S.AT_ptr, S.AT_row, S.AT_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) trip.push_back({S.P_rowp[e], S.P_colp[e], e});
    build_rows(n, trip, S.PR_ptr, S.PR_col, S.PR_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) {
      if (S.P_tgt[e] < 0) continue;
      trip.push_back({S.P_rowp[e], S.P_colp[e], e});
      if (S.P_rowp[e] != S.P_colp[e]) trip.push_back({S.P_colp[e], S.P_rowp[e], e});
    }
    build_rows(n, trip, S.PS_ptr, S.PS_col, S.PS_slot);
    trip.clear();
    for (int e = 0; e < S.nnzP; ++e) trip.push_back({S.P_colp[e], S.P_rowp[e], e});
    std::vector<int> dummy;
    build_rows(n, trip, S.PC_ptr, dummy, S.PC_slot);
  }

  // ---- padded SpMV streams ----
  {
    int wr = 0, wa = 0;
    for (int i = 0; i < m; ++i) wr = std::max(wr, S.A_rowptr[i + 1] - S.A_rowptr[i]);
    for (int j = 0; j < n; ++j) wa = std::max(wa, S.AT_ptr[j + 1] - S.AT_ptr[j]);
    wr = (wr + 7) & ~7;  // whole chunks of 8 entries (kSpU on the device)
    wa = (wa + 7) & ~7;
    if (m > 0 && wr > 0 && wr <= 32 && wa > 0 && wa <= 64) {
      S.WR = wr; S.WA = wa;
      S.m_pad = (m + 31) / 32 * 32;
      S.n_pad = (n + 31) / 32 * 32;
      S.RP_col.assign((size_t)S.m_pad * wr, n);  // padding gathers the dummy slot v[n] == 0
      S.RP_slot.assign((size_t)S.m_pad * wr, -1);
      for (int i = 0; i < m; ++i)
        for (int e = S.A_rowptr[i], t = 0; e < S.A_rowptr[i + 1]; ++e, ++t) {
          S.RP_col[(size_t)i * wr + t] = S.A_col[e];
          S.RP_slot[(size_t)i * wr + t] = e;
        }
      S.ATP_row.assign((size_t)S.n_pad * wa, 0);
      S.ATP_slot.assign((size_t)S.n_pad * wa, -1);
      for (int j = 0; j < n; ++j)
        for (int e = S.AT_ptr[j], t = 0; e < S.AT_ptr[j + 1]; ++e, ++t) {
          S.ATP_row[(size_t)j * wa + t] = S.AT_row[e];
          S.ATP_slot[(size_t)j * wa + t] = S.AT_slot[e];
        }
    }
  }

  // ---- padded sweep schedules ----
  {
    auto emit = [&, n](int k, const std::vector<std::pair<int, int>>& ent, bool keep_empty, std::vector<int>& meta,
                    std::vector<int>& col, std::vector<int>& slt) {
      const int W = SparseSymbolic::kStepWidth;
      const int steps = ent.empty() ? (keep_empty ? 1 : 0) : (int)((ent.size() + W - 1) / W);
      for (int st = 0; st < steps; ++st) {
        meta.push_back((k << 1) | (st == steps - 1 ? 1 : 0));
        for (int i = 0; i < W; ++i) {
          const size_t idx = (size_t)st * W + i;
          col.push_back(idx < ent.size() ? ent[idx].first : n);  // padding gathers the dummy slot v[n] == 0
          slt.push_back(idx < ent.size() ? ent[idx].second : -1);
        }
      }
    };
    std::vector<std::pair<int, int>> ent;
    for (int k = 0; k < n; ++k) {
      ent.clear();
      for (int e = S.LR_ptr[k]; e < S.LR_ptr[k + 1]; ++e) ent.push_back({S.LR_col[e], S.LR_slot[e]});
      emit(k, ent, false, S.FS_meta, S.FS_col, S.FS_slot);
    }
    for (int k = n - 1; k >= 0; --k) {
      ent.clear();
      for (int e = S.L_colptr[k]; e < S.L_colptr[k + 1]; ++e) ent.push_back({S.L_row[e], e});
      emit(k, ent, true, S.BS_meta, S.BS_col, S.BS_slot);
    }
    // pad both lists to a multiple of kStepPad with no-op steps (no entries, not a last step): the device loop is
    // unrolled kStepPad times without bounds checks
    auto pad = [&, n](std::vector<int>& meta, std::vector<int>& col, std::vector<int>& slt) {
      while (meta.size() % SparseSymbolic::kStepPad != 0) {
        meta.push_back(0);
        for (int i = 0; i < SparseSymbolic::kStepWidth; ++i) { col.push_back(n); slt.push_back(-1); }
      }
    };
    pad(S.FS_meta, S.FS_col, S.FS_slot);
    pad(S.BS_meta, S.BS_col, S.BS_slot);
  }

  // ---- right-looking update schedule ----
  S.F_ptr.assign(n + 1, 0);
  for (int k = 0; k < n; ++k) {
    const long long s = (long long)st[k].size();
    S.F_ptr[k + 1] = S.F_ptr[k] + (int)(s * (s + 1) / 2);
    S.flops += s * (s + 1) / 2;
  }
  S.F_tgt.resize(S.F_ptr[n]);
  for (int k = 0; k < n; ++k) {
    int p = S.F_ptr[k];
    for (size_t a = 0; a < st[k].size(); ++a)
      for (size_t b = a; b < st[k].size(); ++b) {
        const int t = target(st[k][b], st[k][a]);
        if (t == -2) { S.error = "internal: fill entry outside the symbolic factor"; return false; }
        S.F_ab.push_back((int)((a << 16) | b));
        S.F_tgt[p++] = t;
      }
    if (st[k].size() > 0xffff) { S.error = "more than 65535 entries in a column of L"; return false; }
  }
  return true;
}

// Self-check of every schedule the device kernel relies on (used by sfb_qp_sparse_symbolic, i.e. by the CPU tests):
// each factor slot is streamed exactly once by the forward and once by the backward step list, each entry of A exactly
// once by the padded row and column streams, every target lies inside the factor, pairs address their own row / column.
inline bool sparse_validate(const SparseSymbolic& S, std::string& why)
{
  const int n = S.n, m = S.m, nL = S.nnzL, nW = S.nnzL + S.n;
  auto fail = [&](const char* msg) { why = msg; return false; };
  std::vector<char> seen(n, 0);
  for (int k = 0; k < n; ++k) {
    if (S.perm[k] < 0 || S.perm[k] >= n || seen[S.perm[k]]) return fail("perm is not a permutation");
    seen[S.perm[k]] = 1;
    if (S.iperm[S.perm[k]] != k) return fail("iperm is not the inverse of perm");
  }
  for (int k = 0; k < n; ++k)
    for (int e = S.L_colptr[k]; e < S.L_colptr[k + 1]; ++e)
      if (S.L_row[e] <= k || S.L_row[e] >= n || (e > S.L_colptr[k] && S.L_row[e] <= S.L_row[e - 1])) return fail("L column not strictly below the diagonal / not ascending");
  for (int t : S.P_tgt) if (t < -1 || t >= nW) return fail("P target out of range");
  for (int t : S.A_pair_tgt) if (t < 0 || t >= nW) return fail("A pair target out of range");
  for (int t : S.F_tgt) if (t < 0 || t >= nW) return fail("factor update target out of range");
  if (S.A_pair_ab.size() != S.A_pair_tgt.size() || S.F_ab.size() != S.F_tgt.size()) return fail("pair lists of different length");
  for (int i = 0; i < m; ++i) {
    const int len = S.A_rowptr[i + 1] - S.A_rowptr[i];
    for (int p = S.A_pair_ptr[i]; p < S.A_pair_ptr[i + 1]; ++p) {
      const int a = S.A_pair_ab[p] >> 16, b = S.A_pair_ab[p] & 0xffff;
      if (a > b || b >= len) return fail("A pair outside its row");
    }
  }
  for (int k = 0; k < n; ++k) {
    const int len = S.L_colptr[k + 1] - S.L_colptr[k];
    std::set<int> tg;
    for (int p = S.F_ptr[k]; p < S.F_ptr[k + 1]; ++p) {
      const int a = S.F_ab[p] >> 16, b = S.F_ab[p] & 0xffff;
      if (a > b || b >= len) return fail("factor pair outside its column");
      if (!tg.insert(S.F_tgt[p]).second) return fail("two updates of one column share a target (the kernel batches them)");
    }
  }
  auto once = [&](const std::vector<int>& slots, int count, const char* msg) {
    std::vector<int> hits(count, 0);
    for (int sl : slots) {
      if (sl < -1 || sl >= count) return fail(msg);
      if (sl >= 0) hits[sl]++;
    }
    for (int c : hits) if (c != 1) return fail(msg);
    return true;
  };
  if (!once(S.LR_slot, nL, "LR mirror does not cover L exactly once")) return false;
  if (!once(S.LB_slot, nL, "LB mirror does not cover L exactly once")) return false;
  if (!once(S.FS_slot, nL, "forward step list does not cover L exactly once")) return false;
  if (!once(S.BS_slot, nL, "backward step list does not cover L exactly once")) return false;
  if (!once(S.AT_slot, S.nnzA, "AT mirror does not cover A exactly once")) return false;
  if (S.WR > 0) {
    if (!once(S.RP_slot, S.nnzA, "padded row stream does not cover A exactly once")) return false;
    if (!once(S.ATP_slot, S.nnzA, "padded column stream does not cover A exactly once")) return false;
    if (S.WR % 8 || S.WA % 8 || S.m_pad % 32 || S.n_pad % 32 || S.m_pad < m || S.n_pad < n) return fail("padded stream geometry");
  }
  if (S.FS_meta.size() % SparseSymbolic::kStepPad || S.BS_meta.size() % SparseSymbolic::kStepPad) return fail("step lists not padded");
  if (S.FS_col.size() != S.FS_meta.size() * SparseSymbolic::kStepWidth || S.BS_col.size() != S.BS_meta.size() * SparseSymbolic::kStepWidth) return fail("step list geometry");
  // step lists: entries of a step belong to the row in its meta; forward rows ascend, backward rows descend; every backward
  // row closes exactly once
  {
    std::vector<int> closed(n, 0);
    int prev = n;
    for (size_t st = 0; st < S.BS_meta.size(); ++st) {
      const int k = S.BS_meta[st] >> 1;
      bool empty = true;
      for (int i = 0; i < SparseSymbolic::kStepWidth; ++i) {
        const int sl = S.BS_slot[st * SparseSymbolic::kStepWidth + i];
        if (sl < 0) { if (S.BS_col[st * SparseSymbolic::kStepWidth + i] != n) return fail("padding does not point at the dummy slot"); continue; }
        empty = false;
        if (sl < S.L_colptr[k] || sl >= S.L_colptr[k + 1]) return fail("backward step entry outside its column");
        if (S.BS_col[st * SparseSymbolic::kStepWidth + i] != S.L_row[sl]) return fail("backward step column index mismatch");
      }
      if (S.BS_meta[st] & 1) { if (k > prev) return fail("backward rows not descending"); prev = k; closed[k]++; }
      else if (empty && S.BS_meta[st] != 0) return fail("open step without entries");
    }
    for (int c : closed) if (c != 1) return fail("a backward row is not closed exactly once");
  }
  for (size_t st = 0; st < S.FS_meta.size(); ++st) {
    const int k = S.FS_meta[st] >> 1;
    for (int i = 0; i < SparseSymbolic::kStepWidth; ++i) {
      const int sl = S.FS_slot[st * SparseSymbolic::kStepWidth + i];
      if (sl < 0) continue;
      if (S.L_row[sl] != k) return fail("forward step entry outside its row");
      const int col = S.FS_col[st * SparseSymbolic::kStepWidth + i];
      if (sl < S.L_colptr[col] || sl >= S.L_colptr[col + 1]) return fail("forward step column index mismatch");
    }
  }
  return true;
}

}  // namespace sfb
