// ic0_symbolic.cpp -- host symbolic phase of the IC(0) preconditioner (north_star item 4).
//
// An incomplete Cholesky factor with the pattern of the lower triangle of H is inherently
// sequential in the natural order of a stencil-like H (the triangular-solve DAG is thousands of
// levels deep), so the unknowns are re-ordered by a greedy multicolouring of H's graph: rows of
// one colour are mutually non-adjacent, the factor and both triangular solves become
// `ncolors` fully parallel phases (14-18 colours on the LF workload).  The reference has no
// iterative solver; EigenSolverWrapper could wrap Eigen::ConjugateGradient with
// IncompleteCholesky, nobody does.
#include <algorithm>
#include <numeric>

#include "ic0.h"

namespace spb {

int greedy_coloring(const HPattern& H, std::vector<int>& color) {
  const int n = H.n;
  color.assign(n, -1);
  std::vector<int> mark;
  int ncolors = 0;
  for (int v = 0; v < n; ++v) {
    mark.assign((size_t)ncolors + 1, 0);
    for (int q = H.colptr[v]; q < H.colptr[(size_t)v + 1]; ++q) {
      const int w = H.rowidx[q];
      if (w != v && color[w] >= 0) mark[color[w]] = 1;
    }
    int c = 0;
    while (c < ncolors && mark[c]) ++c;
    color[v] = c;
    if (c == ncolors) ++ncolors;
  }
  return ncolors;
}

void ic0_symbolic(const HPattern& H, const SellLayout& S, Ic0Symbolic& I) {
  const int n = H.n;
  I.n = n;
  // greedy colouring in natural order
  std::vector<int> color;
  const int ncolors = greedy_coloring(H, color);
  I.ncolors = ncolors;
  // order by (colour, index)
  I.color_ptr.assign((size_t)ncolors + 1, 0);
  for (int v = 0; v < n; ++v) I.color_ptr[(size_t)color[v] + 1]++;
  for (int c = 0; c < ncolors; ++c) I.color_ptr[(size_t)c + 1] += I.color_ptr[c];
  I.perm.resize(n);
  std::vector<int> pinv(n);
  {
    std::vector<int> cur(I.color_ptr.begin(), I.color_ptr.end() - 1);
    for (int v = 0; v < n; ++v) {
      const int k = cur[color[v]]++;
      I.perm[k] = v;
      pinv[v] = k;
    }
  }
  // strict lower triangle of the permuted matrix, rows sorted by column; H storage position of every entry
  I.Lptr.assign((size_t)n + 1, 0);
  I.Lcol.clear();
  I.a_pos.clear();
  I.a_diag.resize(n);
  std::vector<std::pair<int, int64_t>> row;
  for (int i = 0; i < n; ++i) {
    const int r = I.perm[i];
    row.clear();
    int k = 0;
    for (int q = H.colptr[r]; q < H.colptr[(size_t)r + 1]; ++q, ++k) {
      const int j = pinv[H.rowidx[q]];
      if (j < i) row.emplace_back(j, S.pos(r, k));
      else if (j == i) I.a_diag[i] = S.pos(r, k);
    }
    std::sort(row.begin(), row.end());
    for (auto& e : row) {
      I.Lcol.push_back(e.first);
      I.a_pos.push_back(e.second);
    }
    I.Lptr[(size_t)i + 1] = (int)I.Lcol.size();
  }
  // U = L^T (strict): row i of U lists the rows j > i of L that have column i, ascending
  const int nz = (int)I.Lcol.size();
  I.Uptr.assign((size_t)n + 1, 0);
  for (int t = 0; t < nz; ++t) I.Uptr[(size_t)I.Lcol[t] + 1]++;
  for (int i = 0; i < n; ++i) I.Uptr[(size_t)i + 1] += I.Uptr[i];
  I.Ucol.resize(nz);
  I.u_from.resize(nz);
  {
    std::vector<int> cur(I.Uptr.begin(), I.Uptr.end() - 1);
    for (int i = 0; i < n; ++i)
      for (int t = I.Lptr[i]; t < I.Lptr[(size_t)i + 1]; ++t) {
        const int d = cur[I.Lcol[t]]++;
        I.Ucol[d] = i;
        I.u_from[d] = t;
      }
  }
  // sliced-ELL copies for the triangular solves (warp == 32 consecutive permuted rows)
  auto build = [n](const std::vector<int>& ptr, const std::vector<int>& col, Ic0Symbolic::Sell& S) {
    const int ns = (n + 31) / 32;
    S.slice_off.assign((size_t)ns + 1, 0);
    S.rowlen.resize(n);
    for (int i = 0; i < n; ++i) S.rowlen[i] = ptr[(size_t)i + 1] - ptr[i];
    for (int s = 0; s < ns; ++s) {
      int w = 0;
      for (int i = s * 32; i < std::min(n, s * 32 + 32); ++i) w = std::max(w, S.rowlen[i]);
      S.slice_off[(size_t)s + 1] = S.slice_off[s] + (int64_t)w * 32;
    }
    S.col.assign((size_t)S.slice_off[ns], 0);
    S.pos_of_slot.resize(col.size());
    for (int i = 0; i < n; ++i) {
      const int64_t base = S.slice_off[i >> 5] + (i & 31);
      int k = 0;
      for (int t = ptr[i]; t < ptr[(size_t)i + 1]; ++t, ++k) {
        S.col[(size_t)(base + (int64_t)k * 32)] = col[t];
        S.pos_of_slot[t] = (int)(base + (int64_t)k * 32);
      }
    }
  };
  build(I.Lptr, I.Lcol, I.Ls);
  build(I.Uptr, I.Ucol, I.Us);
}

}  // namespace spb
