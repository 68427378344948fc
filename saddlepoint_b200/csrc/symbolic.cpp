// symbolic.cpp -- one-time host symbolic analysis of the Jacobian pattern.
//
// The reference re-derives the structure of dampJ^T*dampJ inside Eigen's sparse product on
// every LM iteration (LMSolver.h:136) and never calls its own analyze
// (LMSolver.h:69-73, EigenSolverWrapper.h:30-52).  The fixed-pattern idea exists only in
// the legacy GNSolver (MatrixPattern/S2D, GNSolver.h:44-83): a list of contributing J-entry
// pairs per normal-matrix entry.  This file builds that map once, in a form the GPU can
// stream: every strictly-lower entry H(i,j) gets the list of (position in column i,
// position in column j) pairs of J entries that share a residual row, in ascending row
// order -- the accumulation order of the reference's product.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>

#include "spb_internal.h"

namespace spb {

uint64_t pattern_fingerprint(int m, int n, const int* colptr, const int* rowidx) {
  // FNV-1a over sizes, column pointers and row indices.  The arrays are hashed in fixed blocks of 2^20 entries (a
  // function of the data only, not of the thread count) on the host threads, and the block hashes are chained: the
  // solve-entry re-validation of a 30 M-entry pattern costs a few milliseconds instead of ~40.
  constexpr uint64_t kOff = 1469598103934665603ULL, kPrime = 1099511628211ULL;
  constexpr int64_t kBlock = 1 << 20;
  const int64_t nz = colptr[n];
  const int64_t nb_c = ((int64_t)n + 1 + kBlock - 1) / kBlock, nb_r = (nz + kBlock - 1) / kBlock;
  std::vector<uint64_t> bh((size_t)(nb_c + nb_r), 0);
  const int parts = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(), (nb_c + nb_r + 3) / 4));
  run_parts(parts, [&](int p) {
    for (int64_t b = p; b < nb_c + nb_r; b += parts) {
      uint64_t h = kOff ^ (uint64_t)b;
      if (b < nb_c) {
        const int64_t lo = b * kBlock, hi = std::min<int64_t>((int64_t)n + 1, lo + kBlock);
        for (int64_t j = lo; j < hi; ++j) {
          h ^= (uint64_t)(uint32_t)colptr[j];
          h *= kPrime;
        }
      } else {
        const int64_t lo = (b - nb_c) * kBlock, hi = std::min<int64_t>(nz, lo + kBlock);
        for (int64_t q = lo; q < hi; ++q) {
          h ^= (uint64_t)(uint32_t)rowidx[q] + 0x9E3779B97F4A7C15ULL * (uint64_t)q;
          h *= kPrime;
        }
      }
      bh[(size_t)b] = h;
    }
  });
  uint64_t h = kOff;
  auto mix = [&h](uint64_t v) {
    h ^= v;
    h *= kPrime;
  };
  mix((uint64_t)m);
  mix((uint64_t)n);
  for (uint64_t v : bh) mix(v);
  return h;
}

namespace {
// SPB_ANALYZE_TIMING=2: stderr stopwatch of the symbolic sub-phases
struct Lap {
  const char* who;
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit Lap(const char* w) : who(w), on(getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 1), t0(std::chrono::steady_clock::now()) {}
  void operator()(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb analyze timing]   %s: %s %.3f s (threads %d)\n", who, what, std::chrono::duration<double>(t1 - t0).count(), host_threads());
    t0 = t1;
  }
};
// Splits columns [0,n) into `parts` contiguous ranges of roughly equal weight (prefix sums in w, n+1 entries).
template <class W>
std::vector<int> split_columns(int n, const W* w, int parts) {
  std::vector<int> cut((size_t)parts + 1, n);
  cut[0] = 0;
  const double total = (double)(w[n] - w[0]);
  for (int p = 1; p < parts; ++p) {
    const double target = (double)w[0] + total * p / parts;
    cut[p] = (int)(std::lower_bound(w, w + n + 1, target, [](W a, double b) { return (double)a < b; }) - w);
    cut[p] = std::max(cut[p], cut[p - 1]);
    cut[p] = std::min(cut[p], n);
  }
  return cut;
}
// CSR view of the J pattern with, per entry, its position inside its CSC column.
struct RowView {
  std::vector<int64_t> ptr;  // m+1
  std::vector<int> col;
  std::vector<int> pos_in_col;
};
void build_row_view(int m, int n, const int* colptr, const int* rowidx, RowView& R) {
  int64_t nz = colptr[n];
  R.ptr.assign((size_t)m + 1, 0);
  for (int64_t q = 0; q < nz; ++q) {
    int r = rowidx[q];
    if (r < 0 || r >= m) throw ArgFail{"row index out of range in J pattern"};
    R.ptr[(size_t)r + 1]++;
  }
  for (int r = 0; r < m; ++r) R.ptr[(size_t)r + 1] += R.ptr[r];
  R.col.resize(nz);
  R.pos_in_col.resize(nz);
  std::vector<int64_t> cur(R.ptr.begin(), R.ptr.end() - 1);
  for (int j = 0; j < n; ++j) {
    int prev = -1;
    for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
      int r = rowidx[q];
      if (r <= prev) throw ArgFail{"J pattern must have strictly ascending row indices per column"};
      prev = r;
      int64_t d = cur[r]++;
      R.col[d] = j;  // columns visited ascending => ascending inside each row
      R.pos_in_col[d] = q - colptr[j];
    }
  }
}
}  // namespace

namespace {
void h_pattern_impl(int m, int n, const int* colptr, const int* rowidx, const RowView& R, HPattern& H);
void assembly_plan_impl(int m, int n, const int* colptr, const int* rowidx, const RowView& R, const HPattern& H, AssemblyPlan& P);
}  // namespace

void build_h_pattern(int m, int n, const int* colptr, const int* rowidx, HPattern& H) {
  RowView R;
  build_row_view(m, n, colptr, rowidx, R);
  h_pattern_impl(m, n, colptr, rowidx, R, H);
}

void build_assembly_plan(int m, int n, const int* colptr, const int* rowidx, const HPattern& H, AssemblyPlan& P) {
  RowView R;
  build_row_view(m, n, colptr, rowidx, R);
  assembly_plan_impl(m, n, colptr, rowidx, R, H, P);
}

// pattern + SpMV layout + assembly plan with one shared row view (what spb_analyze runs)
void build_analysis(int m, int n, const int* colptr, const int* rowidx, HPattern& H, SellLayout& S, AssemblyPlan& P) {
  RowView R;
  Lap lap("build_analysis");
  build_row_view(m, n, colptr, rowidx, R);
  lap("row view");
  h_pattern_impl(m, n, colptr, rowidx, R, H);
  build_sell(H, S);
  lap("pattern + sell");
  assembly_plan_impl(m, n, colptr, rowidx, R, H, P);
  lap("plan");
}

namespace {
void h_pattern_impl(int m, int n, const int* colptr, const int* rowidx, const RowView& R, HPattern& H) {
  (void)m;
  Lap lap("build_h_pattern");
  H.n = n;
  H.colptr.assign((size_t)n + 1, 0);
  H.nup1.assign(n, 0);
  // column ranges are independent: one thread each, local row lists concatenated afterwards
  const int parts = std::max(1, std::min(host_threads(), n / 32768));  // small problems stay on the calling thread
  const std::vector<int> cut = split_columns(n, colptr, parts);
  std::vector<std::vector<int>> local((size_t)parts);
  std::vector<int> collen(n, 0);
  run_parts(parts, [&](int p) {
    std::vector<int> mark(n, -1), tmp;
    std::vector<int>& out = local[p];
    out.reserve((size_t)(colptr[cut[p + 1]] - colptr[cut[p]]) * 2 + 16);
    for (int j = cut[p]; j < cut[p + 1]; ++j) {
      tmp.clear();
      mark[j] = j;  // the damping row makes the diagonal structural even for empty columns
      tmp.push_back(j);
      for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
        int r = rowidx[q];
        for (int64_t t = R.ptr[r]; t < R.ptr[(size_t)r + 1]; ++t) {
          int i = R.col[t];
          if (mark[i] != j) {
            mark[i] = j;
            tmp.push_back(i);
          }
        }
      }
      std::sort(tmp.begin(), tmp.end());
      int up = 0;
      for (int i : tmp) {
        if (i <= j) ++up;
        out.push_back(i);
      }
      H.nup1[j] = up;
      collen[j] = (int)tmp.size();
    }
  });
  lap("columns");
  int64_t total = 0;
  for (int j = 0; j < n; ++j) {
    total += collen[j];
    if (total > (int64_t)0x7fffffff) throw ArgFail{"H pattern exceeds int32 indexing"};
    H.colptr[(size_t)j + 1] = (int)total;
  }
  H.rowidx.resize((size_t)total);
  run_parts(parts, [&](int p) {
    if (!local[p].empty()) memcpy(H.rowidx.data() + H.colptr[cut[p]], local[p].data(), local[p].size() * sizeof(int));
    std::vector<int>().swap(local[p]);
  });
  lap("concatenate");
}
}  // namespace

void build_sell(const HPattern& H, SellLayout& S) {
  int n = H.n;
  S.n = n;
  S.nslices = (n + 31) / 32;
  S.slice_off.assign((size_t)S.nslices + 1, 0);
  S.rowlen.resize(n);
  for (int r = 0; r < n; ++r) S.rowlen[r] = H.colptr[(size_t)r + 1] - H.colptr[r];
  for (int s = 0; s < S.nslices; ++s) {
    int w = 0;
    for (int r = s * 32; r < std::min(n, s * 32 + 32); ++r) w = std::max(w, S.rowlen[r]);
    S.slice_off[(size_t)s + 1] = S.slice_off[s] + (int64_t)w * 32;
  }
  S.padded = S.slice_off[S.nslices];
  S.col.assign((size_t)S.padded, 0);
  S.col16.assign((size_t)S.padded, 32768);
  S.slice16.assign((size_t)S.nslices, 1);
  // slices are independent: whole slices per host thread
  const int parts = std::max(1, std::min(host_threads(), S.nslices / 4096));
  run_parts(parts, [&](int p) {
    const int r_begin = (int)((int64_t)S.nslices * p / parts) * 32, r_end = std::min(n, (int)((int64_t)S.nslices * (p + 1) / parts) * 32);
    for (int r = r_begin; r < r_end; ++r) {
      int64_t base = S.slice_off[r >> 5] + (r & 31);
      int k = 0;
      for (int q = H.colptr[r]; q < H.colptr[(size_t)r + 1]; ++q, ++k) {
        const int c = H.rowidx[q];
        S.col[(size_t)(base + (int64_t)k * 32)] = c;
        const int d = c - r + 32768;
        if (d < 0 || d > 65535) S.slice16[(size_t)(r >> 5)] = 0;
        else S.col16[(size_t)(base + (int64_t)k * 32)] = (uint16_t)d;
      }
      int w = (int)((S.slice_off[(size_t)(r >> 5) + 1] - S.slice_off[r >> 5]) / 32);
      for (; k < w; ++k) S.col[(size_t)(base + (int64_t)k * 32)] = r;  // padding: own row, delta 0
    }
  });
  // rows past n in the last slice never load; their col16 entries stay at delta 0
}

namespace {
void assembly_plan_impl(int m, int n, const int* colptr, const int* rowidx, const RowView& R, const HPattern& H, AssemblyPlan& P) {
  Lap lap("build_assembly_plan");
  P.m = m;
  P.n = n;
  P.nnzJ = colptr[n];
  int maxcol = 0;
  for (int j = 0; j < n; ++j) maxcol = std::max(maxcol, colptr[j + 1] - colptr[j]);
  if (maxcol > 32768) throw ArgFail{"a Jacobian column with more than 32768 entries is not supported"};
  P.rec_bits = maxcol <= 128 ? 16 : 32;
  // strictly-lower slots
  P.low_colptr.assign((size_t)n + 1, 0);
  for (int j = 0; j < n; ++j) P.low_colptr[(size_t)j + 1] = P.low_colptr[j] + (H.colptr[(size_t)j + 1] - H.colptr[j] - H.nup1[j]);
  P.nslots = P.low_colptr[n];
  P.low_row.resize(P.nslots);
  P.low_rank.resize(P.nslots);
  for (int j = 0; j < n; ++j) {
    int s = P.low_colptr[j];
    for (int q = H.colptr[j] + H.nup1[j]; q < H.colptr[(size_t)j + 1]; ++q, ++s) P.low_row[s] = H.rowidx[q];
  }
  // rank of column j inside row i (== column i of the symmetric pattern), for i > j: count as we sweep j
  {
    std::vector<int> seen(n, 0);  // entries of row i with column < current j already emitted
    for (int j = 0; j < n; ++j) {
      for (int s = P.low_colptr[j]; s < P.low_colptr[(size_t)j + 1]; ++s) {
        int i = P.low_row[s];
        int rk = seen[i]++;
        if (rk > 65535) throw ArgFail{"H row longer than 65535 entries is not supported"};
        P.low_rank[s] = (uint16_t)rk;
      }
    }
  }
  lap("slots and ranks");
  // pair stream: column ranges are independent -- one thread each builds its slice of the
  // slot-major record stream, the slices are concatenated in column order
  struct Tmp {
    int slot, po, qo;
  };
  std::vector<int> slot_npairs(P.nslots, 0);
  // First pass materialises records per column into one big stream (slot-major); chunking
  // and 32-padding are applied in a second pass over slots.
  std::vector<uint32_t> stream;
  const int pobits = P.rec_bits == 16 ? 7 : 15;
  {
    const int parts = std::max(1, std::min(host_threads(), n / 32768));
    const std::vector<int> cut = split_columns(n, P.low_colptr.data(), parts);
    std::vector<std::vector<uint32_t>> lstream((size_t)parts);
    run_parts(parts, [&](int p) {
      std::vector<int> where(n, -1);
      std::vector<Tmp> tmp;
      std::vector<int> cnt, start;
      std::vector<uint32_t> colrecs;  // records of the current column, slot-major
      std::vector<uint32_t>& out = lstream[p];
      out.reserve((size_t)(colptr[cut[p + 1]] - colptr[cut[p]]) * 4 + 16);
      for (int j = cut[p]; j < cut[p + 1]; ++j) {
        int s0 = P.low_colptr[j], s1 = P.low_colptr[(size_t)j + 1];
        int ns = s1 - s0;
        if (ns == 0) continue;
        for (int s = s0; s < s1; ++s) where[P.low_row[s]] = s - s0;
        tmp.clear();
        cnt.assign(ns, 0);
        for (int q = colptr[j]; q < colptr[j + 1]; ++q) {
          int r = rowidx[q];
          int qo = q - colptr[j];
          for (int64_t t = R.ptr[r]; t < R.ptr[(size_t)r + 1]; ++t) {
            int i = R.col[t];
            if (i <= j) continue;
            int sl = where[i];
            tmp.push_back({sl, R.pos_in_col[t], qo});
            cnt[sl]++;
          }
        }
        start.assign((size_t)ns + 1, 0);
        for (int k = 0; k < ns; ++k) start[(size_t)k + 1] = start[k] + cnt[k];
        colrecs.resize(tmp.size());
        std::vector<int>& cur = cnt;  // reuse as cursor
        for (int k = 0; k < ns; ++k) cur[k] = start[k];
        for (const Tmp& e : tmp) {
          int d = cur[e.slot]++;
          uint32_t first = (d == start[e.slot]) ? 1u : 0u;
          colrecs[d] = (first << (2 * pobits)) | ((uint32_t)e.po << pobits) | (uint32_t)e.qo;
        }
        for (int k = 0; k < ns; ++k) slot_npairs[s0 + k] = start[(size_t)k + 1] - start[k];
        out.insert(out.end(), colrecs.begin(), colrecs.end());
        for (int s = s0; s < s1; ++s) where[P.low_row[s]] = -1;
      }
    });
    std::vector<size_t> off((size_t)parts + 1, 0);
    for (int p = 0; p < parts; ++p) off[(size_t)p + 1] = off[p] + lstream[p].size();
    stream.resize(off[parts]);
    run_parts(parts, [&](int p) {
      if (!lstream[p].empty()) memcpy(stream.data() + off[p], lstream[p].data(), lstream[p].size() * sizeof(uint32_t));
      std::vector<uint32_t>().swap(lstream[p]);
    });
  }
  lap("pair stream");
  P.npairs = (int64_t)stream.size();
  // Position of the mirror entry H(j,i) in row j, and compaction to the slots that actually
  // receive products.  With a full Jacobian every structural slot is active; under row
  // partitioning (dist.cu) a rank only owns the pairs of its residual rows and the other slots
  // stay zero in its partial H.
  {
    P.low_rank_up.resize(P.nslots);
    int w = 0;
    std::vector<int> new_colptr((size_t)n + 1, 0);
    for (int j = 0; j < n; ++j) {
      for (int s = P.low_colptr[j]; s < P.low_colptr[(size_t)j + 1]; ++s) {
        if (slot_npairs[s] == 0) continue;
        int up = H.nup1[j] + (s - P.low_colptr[j]);
        if (up > 65535) throw ArgFail{"H row longer than 65535 entries is not supported"};
        P.low_row[w] = P.low_row[s];
        P.low_rank[w] = P.low_rank[s];
        P.low_rank_up[w] = (uint16_t)up;
        slot_npairs[w] = slot_npairs[s];
        ++w;
      }
      new_colptr[(size_t)j + 1] = w;
    }
    P.low_colptr.swap(new_colptr);
    P.nslots = w;
    P.low_row.resize(w);
    P.low_rank.resize(w);
    P.low_rank_up.resize(w);
    slot_npairs.resize(w);
  }
  // Chunking.  A chunk is a run of whole H columns (split inside a column only when one column
  // alone exceeds the caps) together with its *stage list*: the J columns whose values its pairs
  // touch (the chunk's own columns j and the row-columns i of its slots).  The kernel copies
  // those columns into shared memory once (coalesced) and the per-pair gathers become shared
  // memory reads; a chunk whose stage list does not fit (very long J columns) is flagged
  // unstaged and gathers from global memory instead.
  P.chunk_slot.clear();
  P.chunk_rec.clear();
  P.chunk_npairs.clear();
  P.chunk_stage_ptr.clear();
  P.chunk_staged.clear();
  P.stage_cols.clear();
  P.stage_off.clear();
  P.stage_run.clear();
  P.slot_il.assign(P.nslots, 0);
  P.slot_jl.assign(P.nslots, 0);
  P.slot_start.assign(P.nslots, 0);
  P.slot_cnt.assign(P.nslots, 0);
  P.proc_perm.resize(P.nslots);
  P.grp_slot.clear();
  P.rec16.clear();
  P.rec32.clear();
  {
    // Phase A (sequential, light): chunk boundaries and stage lists.  Phase B (parallel over
    // chunks): copy the records with their 32-padding, group -> slot map, per-slot start/count
    // and the processing order.
    struct Open {
      int cs, s_end, np;
      int64_t src;
    };
    std::vector<Open> chunks;
    std::vector<int> local_of(n, -1);   // column -> index in the current chunk's stage list
    std::vector<int> cur_list;          // stage list being built
    int64_t src = 0;                    // cursor in `stream`
    int cs = 0;                         // first slot of the open chunk
    int c_np = 0, c_ns = 0, c_vals = 0;
    auto add_col = [&](int col) {
      if (local_of[col] < 0) {
        local_of[col] = (int)cur_list.size();
        cur_list.push_back(col);
        c_vals += (colptr[col + 1] - colptr[col] + (colptr[col] & 1) + 1) & ~1;  // 16-byte aligned window
      }
      return local_of[col];
    };
    auto close_chunk = [&](int s_end) {
      if (s_end == cs) return;
      const int np = c_np;
      chunks.push_back({cs, s_end, np, src});
      P.chunk_slot.push_back(cs);
      P.chunk_npairs.push_back(np);
      P.chunk_stage_ptr.push_back((int)P.stage_cols.size());
      const bool staged = c_vals <= kStageVals && (int)cur_list.size() <= kStageCols;
      P.chunk_staged.push_back(staged ? 1 : 0);
      // The list is stored in ascending column order so that runs of consecutive J columns -- which
      // are contiguous in the CSC value array -- are staged by ONE bulk copy each (the TMA unit is
      // issue-rate bound on small copies: one copy per column was the assembly kernel's limit).
      // A run's window starts 16-byte aligned (one element early when the head column is
      // odd-aligned) and is padded to an even length; inside a run the columns are back to back.
      {
        std::vector<int> sorted(cur_list);
        std::sort(sorted.begin(), sorted.end());
        for (size_t t = 0; t < sorted.size(); ++t) local_of[sorted[t]] = (int)t;  // new local index
        for (int s = cs; s < s_end; ++s) {  // slots recorded insertion-order indices: remap
          P.slot_il[s] = (uint16_t)local_of[cur_list[P.slot_il[s]]];
          P.slot_jl[s] = (uint16_t)local_of[cur_list[P.slot_jl[s]]];
        }
        int cur = 0, run_head = -1, run_len = 0;
        auto close_run = [&]() {
          if (run_head < 0) return;
          const int padded = (run_len + 1) & ~1;
          P.stage_run[(size_t)run_head] = (uint16_t)std::min(padded, 65535);
          cur += padded;
          run_head = -1;
        };
        for (size_t t = 0; t < sorted.size(); ++t) {
          const int col = sorted[t];
          const int len = colptr[col + 1] - colptr[col];
          if (t == 0 || col != sorted[t - 1] + 1) {
            close_run();
            const int a = colptr[col] & 1;
            run_head = (int)P.stage_cols.size();
            run_len = a;
          }
          P.stage_cols.push_back(col);
          P.stage_off.push_back(cur + run_len);
          P.stage_run.push_back(0);
          run_len += len;
          local_of[col] = -1;
        }
        close_run();
      }
      src += np;
      cur_list.clear();
      cs = s_end;
      c_np = c_ns = c_vals = 0;
    };
    for (int j = 0; j < n; ++j) {
      const int s0 = P.low_colptr[j], s1 = P.low_colptr[(size_t)j + 1];
      if (s1 == s0) continue;
      int pj = 0;
      for (int s = s0; s < s1; ++s) pj += slot_npairs[s];
      // would this whole column fit into the open chunk?
      int new_cols = 0, new_vals = 0;
      if (local_of[j] < 0) { ++new_cols; new_vals += colptr[j + 1] - colptr[j] + 2; }
      for (int s = s0; s < s1; ++s) {
        int i = P.low_row[s];
        if (local_of[i] < 0) { ++new_cols; new_vals += colptr[i + 1] - colptr[i] + 2; }
      }
      const bool fits = c_np + pj <= kPairTile && c_ns + (s1 - s0) <= kSlotCap && (int)cur_list.size() + new_cols <= kStageCols &&
                        c_vals + new_vals <= kStageVals;
      if (!fits && c_ns > 0) close_chunk(s0);
      const bool alone_ok = pj <= kPairTile && (s1 - s0) <= kSlotCap;
      if (alone_ok) {
        const int jl = add_col(j);
        for (int s = s0; s < s1; ++s) {
          P.slot_il[s] = (uint16_t)add_col(P.low_row[s]);
          P.slot_jl[s] = (uint16_t)jl;
        }
        c_np += pj;
        c_ns += s1 - s0;
      } else {
        // a column too big for one chunk: split by slots (a single slot may exceed the pair tile:
        // the kernel's multi-tile path handles a one-slot chunk of any length)
        for (int s = s0; s < s1; ++s) {
          if (c_ns > 0 && (c_np + slot_npairs[s] > kPairTile || c_ns + 1 > kSlotCap || (int)cur_list.size() + 2 > kStageCols)) close_chunk(s);
          const int jl = add_col(j);
          P.slot_il[s] = (uint16_t)add_col(P.low_row[s]);
          P.slot_jl[s] = (uint16_t)jl;
          c_np += slot_npairs[s];
          c_ns += 1;
        }
        close_chunk(s1);
      }
    }
    close_chunk(P.nslots);
    lap("chunk boundaries");
    // record offsets (32-padded per chunk)
    const int nch = (int)chunks.size();
    P.chunk_rec.resize(nch);
    int64_t total = 0;
    for (int c = 0; c < nch; ++c) {
      P.chunk_rec[c] = total;
      total += ((int64_t)chunks[c].np + 31) & ~(int64_t)31;
    }
    if (P.rec_bits == 16) P.rec16.assign((size_t)total, 0); else P.rec32.assign((size_t)total, 0);
    P.grp_slot.assign((size_t)(total / 32), 0);
    const int parts = std::max(1, std::min(host_threads(), nch / 256));
    run_parts(parts, [&](int p) {
      const int c_begin = (int)((int64_t)nch * p / parts), c_end = (int)((int64_t)nch * (p + 1) / parts);
      for (int c = c_begin; c < c_end; ++c) {
        const Open& ch = chunks[c];
        const int64_t rec0 = P.chunk_rec[c];
        int sl = -1;  // chunk-local slot of the record being copied
        for (int t = 0; t < ch.np; ++t) {
          const uint32_t rec = stream[(size_t)(ch.src + t)];
          if (rec >> (2 * pobits)) {
            ++sl;
            P.slot_start[(size_t)ch.cs + sl] = (uint16_t)std::min(t, 65535);
          }
          if ((t & 31) == 0) P.grp_slot[(size_t)(rec0 / 32 + t / 32)] = (uint16_t)sl;
          if (P.rec_bits == 16) P.rec16[(size_t)(rec0 + t)] = (uint16_t)rec; else P.rec32[(size_t)(rec0 + t)] = rec;
        }
        // processing order inside the chunk: slots sorted by descending pair count, so the 32
        // slots a warp walks in lock step have (nearly) equal trip counts -- no divergence waste.
        // Records stay slot-major; every slot carries its own (start, count).
        for (int sidx = ch.cs; sidx < ch.s_end; ++sidx) {
          P.slot_cnt[sidx] = (uint16_t)std::min(slot_npairs[sidx], 65535);
          P.proc_perm[sidx] = sidx;
        }
        std::stable_sort(P.proc_perm.begin() + ch.cs, P.proc_perm.begin() + ch.s_end,
                         [&](int a2, int b2) { return slot_npairs[a2] > slot_npairs[b2]; });
        // The kernel walks position p of the chunk in thread p & 255, pass p >> 8.  The second pass runs in REVERSE
        // order, so the warp that got the 32 longest slots in the first pass gets the 32 shortest in the second: the
        // warps of a CTA reach the barrier together (the barrier wait was 30 % of the kernel's stall samples).
        if (ch.s_end - ch.cs > 256) std::reverse(P.proc_perm.begin() + ch.cs + 256, P.proc_perm.begin() + ch.s_end);
      }
    });
  }
  lap("chunking");
  P.chunk_slot.push_back(P.nslots);
  P.chunk_stage_ptr.push_back((int)P.stage_cols.size());
  // column kernel chunks: by entry count and column count
  P.colchunk.clear();
  int j = 0;
  while (j < n) {
    P.colchunk.push_back(j);
    int cj = j;
    int64_t ne = 0;
    while (j < n && (j - cj) < kColCap && (ne == 0 || ne + (colptr[j + 1] - colptr[j]) <= kEntTile)) {
      ne += colptr[j + 1] - colptr[j];
      ++j;
    }
  }
  P.colchunk.push_back(n);
}
}  // namespace

}  // namespace spb
