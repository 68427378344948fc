// symbolic_api.cpp -- host-only entry points of the symbolic phase (no GPU needed): used by
// the CPU test tier to check the pattern and the assembly plan's invariants.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chol_internal.h"

using namespace spb;

extern "C" {

spb_status spb_symbolic_h_pattern(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* nnz,
                                  int32_t* h_colptr, int32_t* h_rowidx) {
  if (!colptr || !nnz || m < 0 || n < 0) return SPB_ERR_ARG;
  try {
    HPattern H;
    build_h_pattern(m, n, colptr, rowidx, H);
    *nnz = H.nnz();
    if (h_colptr) memcpy(h_colptr, H.colptr.data(), sizeof(int32_t) * ((size_t)n + 1));
    if (h_rowidx) memcpy(h_rowidx, H.rowidx.data(), sizeof(int32_t) * (size_t)H.nnz());
    return SPB_OK;
  } catch (const ArgFail&) {
    return SPB_ERR_ARG;
  }
}

// Cholesky symbolic phase on the host: stats[0..7] = nsuper, nlevels, nnz(L), max pivot block,
// max front, flops (rounded), update-matrix arena size (entries), structural violations (must be 0).
// perm (optional, n ints) receives the fill-reducing order.
spb_status spb_symbolic_chol_stats(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* stats, int32_t* perm) {
  if (!colptr || !stats || m < 0 || n < 0) return SPB_ERR_ARG;
  try {
    HPattern H;
    build_h_pattern(m, n, colptr, rowidx, H);
    SellLayout S;
    build_sell(H, S);
    CholSymbolic C;
    chol_symbolic(H, S, C);
    int64_t bad = 0;
    std::vector<char> seen(n, 0);
    for (int v : C.perm) {
      if (v < 0 || v >= n || seen[v]) ++bad; else seen[v] = 1;
    }
    for (int s = 0; s < C.nsuper; ++s) {
      int p = C.sn_parent[s];
      if (p >= 0 && (p <= s || C.sn_level[p] <= C.sn_level[s])) ++bad;
      int prev = C.sn_col0[s + 1] - 1;
      for (int64_t t = C.struct_ptr[s]; t < C.struct_ptr[(size_t)s + 1]; ++t) {
        if (C.struct_idx[(size_t)t] <= prev) ++bad;
        prev = C.struct_idx[(size_t)t];
      }
    }
    if ((int64_t)C.a_src.size() != (H.nnz() + n) / 2) ++bad;
    {
      // update-matrix arena: blocks that are live at the same time must not overlap
      struct Iv { int64_t off, len; int birth, death; };
      std::vector<Iv> iv;
      for (int s = 0; s < C.nsuper; ++s) {
        const int64_t u = C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s];
        if (u == 0) continue;
        if (C.U_off[s] < 0 || C.U_off[s] + u * u > C.U_total) ++bad;
        iv.push_back({C.U_off[s], u * u, C.sn_level[s], C.sn_parent[s] >= 0 ? C.sn_level[C.sn_parent[s]] : C.sn_level[s]});
      }
      std::sort(iv.begin(), iv.end(), [](const Iv& a, const Iv& b) { return a.off < b.off; });
      // sweep: compare each block with the following ones that start before it ends
      for (size_t a = 0; a < iv.size(); ++a)
        for (size_t b = a + 1; b < iv.size() && iv[b].off < iv[a].off + iv[a].len; ++b)
          if (iv[a].birth <= iv[b].death && iv[b].birth <= iv[a].death) ++bad;
    }
    stats[0] = C.nsuper;
    stats[1] = C.nlevels;
    stats[2] = C.nnzL;
    stats[3] = C.max_k;
    stats[4] = C.max_r;
    stats[5] = (int64_t)C.flops;
    stats[6] = C.U_total;  // arena size (lifetime-packed); C.U_sum is the sum of all update matrices
    stats[7] = bad;
    if (perm) memcpy(perm, C.perm.data(), sizeof(int32_t) * (size_t)n);
    return SPB_OK;
  } catch (const ArgFail&) {
    return SPB_ERR_ARG;
  }
}

// stats: 0 nslots, 1 npairs, 2 nchunks, 3 rec_bits, 4 padded SELL entries, 5 nnz(H),
//        6 ncolchunks, 7 number of invariant violations (must be 0)
spb_status spb_symbolic_plan_check(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int64_t* stats) {
  if (!colptr || !stats || m < 0 || n < 0) return SPB_ERR_ARG;
  try {
    const bool timing = getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 0;
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
      if (!timing) return;
      auto t1 = std::chrono::steady_clock::now();
      fprintf(stderr, "[spb analyze timing] %s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
      t0 = t1;
    };
    HPattern H;
    SellLayout S;
    AssemblyPlan P;
    build_analysis(m, n, colptr, rowidx, H, S, P);
    lap("build_analysis (pattern + SpMV layout + assembly plan)");
    int64_t bad = 0;
    const int pobits = P.rec_bits == 16 ? 7 : 15;
    const uint32_t mask = (1u << pobits) - 1u;
    int nch = (int)P.chunk_npairs.size();
    int64_t pairs_seen = 0;
    for (int c = 0; c < nch; ++c) {
      int s0 = P.chunk_slot[c], s1 = P.chunk_slot[c + 1];
      int64_t r0 = P.chunk_rec[c];
      if (r0 % 32) ++bad;
      int sl = -1, prev_row = -1;
      for (int t = 0; t < P.chunk_npairs[c]; ++t) {
        uint32_t rec = P.rec_bits == 16 ? (uint32_t)P.rec16[(size_t)(r0 + t)] : P.rec32[(size_t)(r0 + t)];
        if (rec >> (2 * pobits)) { ++sl; prev_row = -1; }
        if ((t & 31) == 0 && P.grp_slot[(size_t)(r0 / 32 + t / 32)] != (uint16_t)sl) ++bad;
        if (sl < 0 || s0 + sl >= s1) { ++bad; continue; }
        int s = s0 + sl;
        int i = P.low_row[s];
        // column of the slot
        int j = (int)(std::upper_bound(P.low_colptr.begin(), P.low_colptr.end(), s) - P.low_colptr.begin()) - 1;
        int po = (int)((rec >> pobits) & mask), qo = (int)(rec & mask);
        if (colptr[i] + po >= colptr[i + 1] || colptr[j] + qo >= colptr[j + 1]) { ++bad; continue; }
        int rp = rowidx[colptr[i] + po], rq = rowidx[colptr[j] + qo];
        if (rp != rq) ++bad;            // both J entries must share the residual row
        if (rp <= prev_row) ++bad;      // ascending residual row inside a slot
        prev_row = rp;
        if (!(i > j)) ++bad;
        ++pairs_seen;
      }
      if (sl + 1 != s1 - s0) ++bad;
      // stage list: every slot's (i, j) columns are in the chunk's list at the recorded local index,
      // offsets are the running sum of the listed columns' lengths, and the staged flag respects the caps
      const int st0 = P.chunk_stage_ptr[c], st1 = P.chunk_stage_ptr[c + 1];
      // ascending columns; runs of consecutive columns share one 16-byte-aligned, even-length window
      int off = 0, run_len = 0, run_head = -1;
      auto close_run = [&]() {
        if (run_head < 0) return;
        const int padded = (run_len + 1) & ~1;
        if (P.stage_run[run_head] != padded) ++bad;
        off += padded;
        run_head = -1;
      };
      for (int t = st0; t < st1; ++t) {
        const int col = P.stage_cols[t];
        if (t > st0 && col <= P.stage_cols[t - 1]) ++bad;
        if (t == st0 || col != P.stage_cols[t - 1] + 1) {
          close_run();
          run_head = t;
          run_len = colptr[col] & 1;
        } else if (P.stage_run[t] != 0) {
          ++bad;
        }
        if (P.stage_off[t] != off + run_len) ++bad;
        run_len += colptr[col + 1] - colptr[col];
      }
      close_run();
      if (P.chunk_staged[c] && (off > kStageVals || st1 - st0 > kStageCols)) ++bad;
      for (int s = s0; s < s1; ++s) {
        int j = (int)(std::upper_bound(P.low_colptr.begin(), P.low_colptr.end(), s) - P.low_colptr.begin()) - 1;
        if (st0 + P.slot_il[s] >= st1 || P.stage_cols[st0 + P.slot_il[s]] != P.low_row[s]) ++bad;
        if (st0 + P.slot_jl[s] >= st1 || P.stage_cols[st0 + P.slot_jl[s]] != j) ++bad;
      }
      if (s1 - s0 > kSlotCap || (P.chunk_npairs[c] > kPairTile && s1 - s0 != 1)) ++bad;
      // processing order: a permutation of the chunk's slots; pair counts non-increasing over the first 256 positions
      // (first pass of the kernel), then non-decreasing and never above the first pass's minimum (second pass, reversed)
      {
        std::vector<int> seen(s1 - s0, 0);
        int prevc = 1 << 30, first_min = 1 << 30;
        for (int pos = s0; pos < s1; ++pos) {
          int sidx = P.proc_perm[pos];
          if (sidx < s0 || sidx >= s1 || seen[sidx - s0]++) { ++bad; continue; }
          if (pos - s0 < 256) {
            if (P.slot_cnt[sidx] > prevc) ++bad;
            first_min = P.slot_cnt[sidx];
          } else {
            if (pos - s0 > 256 && P.slot_cnt[sidx] < prevc) ++bad;
            if (P.slot_cnt[sidx] > first_min) ++bad;
          }
          prevc = P.slot_cnt[sidx];
        }
      }
    }
    if (pairs_seen != P.npairs) ++bad;
    // rank check: H(i,j) must sit at rank low_rank in row i of the symmetric pattern
    for (int j = 0; j < n; ++j)
      for (int s = P.low_colptr[j]; s < P.low_colptr[j + 1]; ++s) {
        int i = P.low_row[s];
        if (H.rowidx[(size_t)H.colptr[i] + P.low_rank[s]] != j) ++bad;
      }
    stats[0] = P.nslots;
    stats[1] = P.npairs;
    stats[2] = nch;
    stats[3] = P.rec_bits;
    stats[4] = S.padded;
    stats[5] = H.nnz();
    stats[6] = (int64_t)P.colchunk.size() - 1;
    stats[7] = bad;
    return SPB_OK;
  } catch (const ArgFail&) {
    return SPB_ERR_ARG;
  }
}

// The "stream" form of the pair assembly (pairs_v2.cpp) executed on the host in the kernel's order: strictly-lower
// and mirrored entries of H in CSC order (diagonal entries are left 0: the column kernel owns them).
// stats: 0 form applicable, 1 chunks, 2 section bytes, 3 runs, 4 largest value window (bytes), 5 largest section
spb_status spb_symbolic_pairs_v2(int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, const double* Jvals, double* hvals,
                                 int64_t* stats) {
  if (!colptr || !stats || !Jvals || !hvals || m < 0 || n < 0) return SPB_ERR_ARG;
  try {
    HPattern H;
    SellLayout S;
    AssemblyPlan P;
    build_analysis(m, n, colptr, rowidx, H, S, P);
    PairsV2Host V;
    const auto tp0 = std::chrono::steady_clock::now();
    pack_pairs_v2(colptr, P, S, V);
    if (getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 0)
      fprintf(stderr, "[spb analyze timing] pack_pairs_v2 %.3f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
    stats[0] = V.ok ? 1 : 0;
    stats[1] = (int64_t)V.hdr.size();
    stats[2] = (int64_t)V.blob.size();
    stats[3] = (int64_t)V.runs.size();
    stats[4] = V.max_val_bytes;
    stats[5] = V.max_sec_bytes;
    if (!V.ok) return SPB_OK;
    {
      // shared-memory cost model of the consumer loop: an 8-byte load of a half-warp takes as many wavefronts as the
      // largest number of DISTINCT addresses its active lanes have in one of the 16 bank pairs
      int64_t wf = 0, rounds = 0;
      for (const ChunkHdr2& h : V.hdr) {
        const V2Section sv(V.blob.data() + h.blob_off);
        for (int g = 0; g < sv.ng; ++g)
          for (int d = 0; d < (int)(sv.ginfo[g] >> 16); ++d)
            for (int half = 0; half < 2; ++half) {
              bool any = false;
              for (int op = 0; op < 2; ++op) {
                int addr[16], na = 0;
                for (int t = g * 32 + half * 16; t < g * 32 + half * 16 + 16; ++t) {
                  if (d >= (int)sv.cnt[t]) continue;
                  const uint32_t m = sv.meta[t], r = sv.rec[((size_t)(sv.ginfo[g] & 0xffffu) + d) * 32 + (t & 31)];
                  addr[na++] = op ? (int)(sv.stab[2 * ((m >> 8) & 255) + 1] / 8) + (int)(r & 127) : (int)(sv.stab[2 * (m & 255) + 1] / 8) + (int)((r >> 7) & 127);
                }
                any |= na > 0;
                int worst = 0;
                for (int b = 0; b < 16; ++b) {
                  int distinct = 0;
                  for (int x = 0; x < na; ++x) {
                    if ((addr[x] & 15) != b) continue;
                    bool dup = false;
                    for (int y = 0; y < x; ++y) dup |= addr[y] == addr[x];
                    distinct += !dup;
                  }
                  worst = std::max(worst, distinct);
                }
                wf += worst;
              }
              rounds += any;
            }
      }
      stats[6] = wf;          // wavefronts of the two operand loads
      stats[7] = 2 * rounds;  // conflict-free minimum
    }
    std::vector<double> st((size_t)S.padded, 0.0);
    run_pairs_v2_host(V, Jvals, st.data());
    for (int r = 0; r < H.n; ++r) {
      int k = 0;
      for (int q = H.colptr[r]; q < H.colptr[(size_t)r + 1]; ++q, ++k) hvals[q] = st[(size_t)S.pos(r, k)];
    }
    return SPB_OK;
  } catch (const ArgFail&) {
    return SPB_ERR_ARG;
  } catch (const StateFail&) {
    return SPB_ERR_STATE;
  }
}
}
