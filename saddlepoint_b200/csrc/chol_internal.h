// chol_internal.h -- data shared by the symbolic (host) and numeric (device) phases of the
// supernodal Cholesky.
#pragma once
#include "spb_internal.h"

namespace spb {

constexpr int kSmallChildU = 32;  // update matrices up to this size are added by the per-parent kernel

struct CholSymbolic {
  int n = 0, nsuper = 0, nlevels = 0;
  std::vector<int> perm, pinv;         // perm[new] = old ; pinv[old] = new
  std::vector<int> sn_col0;            // nsuper+1: pivot column ranges (new numbering)
  std::vector<int64_t> struct_ptr;     // nsuper+1
  std::vector<int> struct_idx;         // rows below the pivot block (new numbering, ascending)
  std::vector<int> sn_parent, sn_level;
  std::vector<int> level_ptr, level_sn;  // supernodes grouped by level, leaves first
  std::vector<int64_t> L_off, U_off, W_off;  // panel (r x k), update matrix (u x u, lifetime-packed: see U_total), rhs work (r)
  // Update matrices live from their front's level to their parent's level only, so their storage is
  // packed by lifetime into one arena (first fit): U_total doubles instead of the sum of all u^2.
  int64_t U_total = 0, U_sum = 0;
  std::vector<int> u_zero_ptr;        // nlevels+1: ranges to clear before level l's extend-add
  std::vector<int64_t> u_zero_range;  // (offset, length) pairs
  std::vector<int64_t> a_src, a_dst;   // H storage position -> panel position (lower entries)
  std::vector<int64_t> rel_ptr;        // nsuper+1
  std::vector<int> rel_idx;            // child struct -> parent front local index
  std::vector<int> child_ptr, child_idx; // children of every supernode (ascending)
  std::vector<int64_t> dinv_off;         // nsuper+1: inverted 32x32 diagonal blocks (1024 doubles each)
  std::vector<int> level_rank_ptr;     // nlevels+1 into rank_ptr
  std::vector<int> rank_ptr;           // per (level, child rank): range in rank_children
  std::vector<int> rank_children;
  // Extend-add order: the children of a parent sorted by descending update-matrix size.  The first level_big_ranks[l]
  // ranks of a level (the ones that can hold a child with more than kSmallChildU rows) are added rank by rank, one launch
  // each over all parents; every further child is small and is added by ONE launch per level in which a warp walks the
  // remaining children of its parent in order (ea_ptr/ea_idx; small_parents: the parents of a level that have such children).
  std::vector<int> ea_ptr, ea_idx;            // nsuper+1 / children in extend-add order
  std::vector<int> level_big_ranks;           // nlevels
  std::vector<int> small_parents_ptr, small_parents;  // nlevels+1 / parent supernodes
  int max_k = 0, max_r = 0;
  int64_t nnzL = 0;
  double flops = 0.0;
};

void chol_symbolic(const HPattern& H, const SellLayout& S, CholSymbolic& C);

}  // namespace spb
