// ic0.h -- multicolour IC(0) preconditioner: host symbolic data and the device state.
#pragma once
#include "spb_internal.h"

namespace spb {

struct Ic0Symbolic {
  int n = 0, ncolors = 0;
  std::vector<int> perm;        // perm[new] = old, rows grouped by colour
  std::vector<int> color_ptr;   // ncolors+1: row ranges (new numbering) of every colour
  std::vector<int> Lptr, Lcol;  // strict lower triangle of the permuted H, CSR, ascending columns
  std::vector<int64_t> a_pos;   // H storage position of every strict-lower entry
  std::vector<int64_t> a_diag;  // H storage position of every diagonal entry
  std::vector<int> Uptr, Ucol;  // U = L^T (strict), CSR
  std::vector<int> u_from;      // U slot -> L slot
  // triangular-solve storage: sliced ELL (slice height 32) over the permuted rows, like H's SpMV layout
  struct Sell {
    std::vector<int64_t> slice_off;  // nslices+1
    std::vector<int> col, rowlen;    // padded columns (padding: 0, never read thanks to rowlen), n
    std::vector<int> pos_of_slot;    // CSR slot -> storage position
  } Ls, Us;
};

void ic0_symbolic(const HPattern& H, const SellLayout& S, Ic0Symbolic& I);
// Greedy distance-1 colouring of H's graph in natural order; returns the number of colours.
// Columns of J with one colour share no residual row (H = J^T J structurally).
int greedy_coloring(const HPattern& H, std::vector<int>& color);

struct Ic0Device {
  int n = 0, ncolors = 0;
  int64_t nz = 0;
  DevBuf<int> perm, color_ptr, Lptr, Lcol, Uptr, Ucol, u_from;
  DevBuf<int64_t> Ls_off, Us_off;
  DevBuf<int> Ls_col, Ls_len, Ls_pos, Us_col, Us_len, Us_pos;
  DevBuf<double> Ls_val, Us_val;
  DevBuf<int64_t> a_pos, a_diag;
  DevBuf<double> Lval, Uval, invd, y, zc, z;
  DevBuf<int> flags;  // [0]: number of pivots that had to fall back to the unmodified diagonal
  int breakdowns = 0;
};

}  // namespace spb
