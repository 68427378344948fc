// pairs_v2.cpp -- host packing of the assembly plan into the "stream" form that assemble_pairs_stream_kernel
// (assembly_v2.cu) consumes, and a host execution of that form in the kernel's own order.
//
// The plan (symbolic.cpp) describes, per chunk of whole H columns, the strictly-lower slots H(i,j), their pair
// records (which entry of J column i meets which entry of J column j, ascending residual row: the accumulation
// order of the reference's sparse product, LMSolver.h:136) and the stage list of J columns the chunk reads.
// Here every chunk becomes ONE contiguous section (tables, 4-byte slot metadata, records in lane-interleaved rows per group of 32 slots)
// plus a short list of bulk-copy runs, so that the kernel's consumer warps touch nothing but shared memory and
// the two H stores per slot.  Index bytes per slot drop from 12 to 5 and no per-slot record offsets are stored:
// slots of similar pair count share a group of 32 lanes whose records are stored lane-interleaved (row d = the d-th
// product of every lane: one 64-byte read per warp), and the lanes of a half-warp are chosen so that their operand
// loads fall into different shared-memory banks.
#include <cstring>

#include "spb_internal.h"

namespace spb {

namespace {
inline uint32_t align16(uint32_t b) { return (b + 15u) & ~15u; }
}  // namespace

void pack_pairs_v2(const int* colptr, const AssemblyPlan& P, const SellLayout& S, PairsV2Host& out) {
  out = PairsV2Host();
  const int nch = (int)P.chunk_npairs.size();
  const int64_t nnzJ = colptr[P.n];
  auto fail = [&](const char* why) {
    out = PairsV2Host();
    out.why = why;
  };
  if (nch == 0) return fail("no chunks");
  if (P.rec_bits != 16) return fail("32-bit pair records");
  if (S.padded >= ((int64_t)1 << 32)) return fail("H storage beyond 32-bit positions");
  static const int penalty = getenv("SPB_ASM_V2_PENALTY") ? atoi(getenv("SPB_ASM_V2_PENALTY")) : 2;
  static const int window = getenv("SPB_ASM_V2_WINDOW") ? std::max(1, atoi(getenv("SPB_ASM_V2_WINDOW"))) : 8;  // measured on LF-1M: window 1 / 8 / 48 -> 0.250 / 0.241 / 0.240 ms per assembly, 0 / 0.17 / 0.58 s of packing
  // pass 1: eligibility
  for (int c = 0; c < nch; ++c) {
    const int s0 = P.chunk_slot[c], s1 = P.chunk_slot[(size_t)c + 1];
    const int st0 = P.chunk_stage_ptr[c], st1 = P.chunk_stage_ptr[(size_t)c + 1];
    const int ns = s1 - s0, nst = st1 - st0, np = P.chunk_npairs[c];
    if (!P.chunk_staged[c]) return fail("a chunk's stage list does not fit shared memory");
    if (np > kPairTile || ns > kSlotCap || ns < 1) return fail("a chunk exceeds one record tile");
    if (nst > 255) return fail("stage list longer than 255 columns");
    for (int s = s0; s < s1; ++s) {
      if (P.low_rank[s] > 255 || P.low_rank_up[s] > 255) return fail("H row longer than 256 entries");
      if (P.slot_cnt[s] > 255 || P.slot_cnt[s] < 1) return fail("a slot with more than 255 products");
    }
  }
  {
    // The stream form feeds a stage through ONE producer warp: it pays off when the J columns a chunk reads come in
    // long consecutive runs (stencil-like problems: ~600 doubles per bulk copy on LF).  Scattered stage lists (mesh
    // problems: ~100 copies of ~6 doubles per chunk) are bound by the issue rate of the copies, and the general
    // kernels issue them from all threads of several co-resident CTAs -- measured on the seamless-integration
    // structure (1.6 M unknowns): 0.26 ms general against 0.49 ms stream form.
    int64_t nrun = 0, nval = 0;
    for (size_t t = 0; t < P.stage_run.size(); ++t)
      if (P.stage_run[t] > 0) {
        ++nrun;
        nval += P.stage_run[t];
      }
    const int min_run = getenv("SPB_ASM_V2_MIN_RUN") ? atoi(getenv("SPB_ASM_V2_MIN_RUN")) : 64;  // read per call: the tests force the form on scattered patterns
    if (nrun == 0 || nval < (int64_t)min_run * nrun) return fail("short bulk-copy runs");
  }
  // pass 2 (parallel over chunks): lane assignment, sections into per-part buffers
  const int parts = std::max(1, std::min(host_threads(), nch / 64));
  std::vector<std::vector<unsigned char>> pblob((size_t)parts);
  std::vector<std::vector<Run2>> pruns((size_t)parts);
  std::vector<uint32_t> part_maxval((size_t)parts, 0), part_maxsec((size_t)parts, 0);
  std::vector<const char*> part_fail((size_t)parts, nullptr);
  out.hdr.resize((size_t)nch);
  run_parts(parts, [&](int p) {
    std::vector<int> order, cand;
    std::vector<uint16_t> paddr;
    std::vector<int> hw_addr;                     // half-warp being filled: [d][op][lane]
    std::vector<uint8_t> hw_n, hw_max, hw_bank;   // [d][op], [d][op], [d][op][16]
    std::vector<uint8_t> taken;
    std::vector<unsigned char>& blob = pblob[p];
    std::vector<Run2>& runsv = pruns[p];
    for (int c = (int)((int64_t)nch * p / parts); c < (int)((int64_t)nch * (p + 1) / parts); ++c) {
      ChunkHdr2& h = out.hdr[c];
      const int s0 = P.chunk_slot[c], s1 = P.chunk_slot[(size_t)c + 1];
      const int st0 = P.chunk_stage_ptr[c], st1 = P.chunk_stage_ptr[(size_t)c + 1];
      const int ns = s1 - s0, nst = st1 - st0;
      const int64_t rec0 = P.chunk_rec[c];
      // shared-memory address (doubles) of operand op of product d of slot s, tabulated once per chunk
      paddr.resize((size_t)2 * P.chunk_npairs[c]);
      for (int s = s0; s < s1; ++s) {
        const int oi = P.stage_off[st0 + P.slot_il[s]], oj = P.stage_off[st0 + P.slot_jl[s]];
        for (int d = 0; d < (int)P.slot_cnt[s]; ++d) {
          const uint32_t r = P.rec16[(size_t)(rec0 + P.slot_start[s] + d)];
          paddr[2 * ((size_t)P.slot_start[s] + d)] = (uint16_t)(oi + (int)((r >> 7) & 127u));
          paddr[2 * ((size_t)P.slot_start[s] + d) + 1] = (uint16_t)(oj + (int)(r & 127u));
        }
      }
      auto addr_of = [&](int s, int d, int op) { return (int)paddr[2 * ((size_t)P.slot_start[s] + d) + op]; };
      // Lane assignment.  Candidates in descending pair count; every half-warp of 16 lanes starts with the longest
      // slot left and is filled, lane by lane, with the slot among the next `window` candidates that adds the fewest
      // shared-memory wavefronts: an 8-byte load of a half-warp costs as many wavefronts as the largest number of
      // distinct addresses in one of the 16 bank pairs, and consecutive J columns sit ~16 doubles apart, so the
      // plain count order (neighbouring columns in neighbouring lanes) is the worst case.
      cand.resize((size_t)ns);
      for (int k = 0; k < ns; ++k) cand[k] = s0 + k;
      std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return P.slot_cnt[a] > P.slot_cnt[b]; });
      taken.assign((size_t)ns, 0);
      order.clear();
      int head = 0;
      while ((int)order.size() < ns) {
        // state of the half-warp being filled: per (product index, operand) the addresses so far
        while (taken[head]) ++head;
        const int seed = cand[head];
        const int dmax = P.slot_cnt[seed];
        hw_addr.assign((size_t)dmax * 2 * 16, 0);
        hw_n.assign((size_t)dmax * 2, 0);
        hw_max.assign((size_t)dmax * 2, 0);
        hw_bank.assign((size_t)dmax * 2 * 16, 0);
        auto add_slot = [&](int s) {
          for (int d = 0; d < (int)P.slot_cnt[s]; ++d)
            for (int op = 0; op < 2; ++op) {
              const int x = addr_of(s, d, op), k = d * 2 + op;
              bool dup = false;
              for (int y = 0; y < hw_n[k]; ++y) dup |= hw_addr[(size_t)k * 16 + y] == x;
              hw_addr[(size_t)k * 16 + hw_n[k]++] = x;
              if (!dup) {
                const uint8_t nb = ++hw_bank[(size_t)k * 16 + (x & 15)];
                hw_max[k] = std::max(hw_max[k], nb);
              }
            }
        };
        auto cost_of = [&](int s, int bound) {
          int cost = 0;
          for (int d = 0; d < (int)P.slot_cnt[s] && cost < bound; ++d)
            for (int op = 0; op < 2; ++op) {
              // (an address another lane already reads would be a broadcast, not a conflict: ignored here, the estimate
              // only has to rank candidates; add_slot keeps the exact bank counts)
              const int x = addr_of(s, d, op), k = d * 2 + op;
              if (hw_bank[(size_t)k * 16 + (x & 15)] + 1 > hw_max[k]) ++cost;
            }
          return cost;
        };
        taken[head] = 1;
        order.push_back(seed);
        add_slot(seed);
        for (int lane = 1; lane < 16 && (int)order.size() < ns; ++lane) {
          int best = -1, best_cost = 1 << 30, seen = 0;
          for (int q = head + 1; q < ns && seen < window; ++q) {
            if (taken[q]) continue;
            ++seen;
            const int waste = (dmax - (int)P.slot_cnt[cand[q]]) * penalty;  // record rows this lane would idle through
            if (waste >= best_cost) break;  // candidates are in descending count: the rest wastes at least as much
            const int cst = waste + cost_of(cand[q], best_cost - waste);
            if (cst < best_cost) {
              best_cost = cst;
              best = q;
              if (cst == 0) break;
            }
          }
          if (best < 0) break;
          taken[best] = 1;
          order.push_back(cand[best]);
          add_slot(cand[best]);
        }
      }
      // groups of 32 lanes: records of a group in lane-interleaved rows, as many rows as its longest slot
      const int ngroups = (ns + 31) / 32, nsp = ngroups * 32;
      std::vector<int> grow((size_t)ngroups + 1, 0);
      for (int g = 0; g < ngroups; ++g) {
        int mx = 0;
        for (int k = g * 32; k < std::min(ns, g * 32 + 32); ++k) mx = std::max<int>(mx, P.slot_cnt[order[k]]);
        grow[(size_t)g + 1] = grow[g] + mx;
      }
      if (grow[ngroups] > 65535) {
        part_fail[p] = "record rows beyond 16 bits";
        return;
      }
      const uint32_t off_stab = 16, off_meta = off_stab + 8u * nst, off_ginfo = off_meta + 4u * nsp, off_cnt = off_ginfo + 4u * ngroups,
                     off_rec = (off_cnt + (uint32_t)nsp + 1u) & ~1u;
      const uint32_t sec_bytes = align16(off_rec + 2u * 32u * (uint32_t)grow[ngroups]);
      if (sec_bytes > 65520u) {
        part_fail[p] = "section too long";
        return;
      }
      const size_t b0 = blob.size();
      blob.resize(b0 + sec_bytes, 0);
      unsigned char* sec = blob.data() + b0;
      uint16_t* hd16 = reinterpret_cast<uint16_t*>(sec);
      hd16[0] = (uint16_t)ns;
      hd16[1] = (uint16_t)nst;
      hd16[2] = (uint16_t)ngroups;
      hd16[3] = (uint16_t)off_meta;
      hd16[4] = (uint16_t)off_ginfo;
      hd16[5] = (uint16_t)off_cnt;
      hd16[6] = (uint16_t)off_rec;
      uint32_t* stab = reinterpret_cast<uint32_t*>(sec + off_stab);
      uint32_t* meta = reinterpret_cast<uint32_t*>(sec + off_meta);
      uint32_t* ginfo = reinterpret_cast<uint32_t*>(sec + off_ginfo);
      uint8_t* cnt = sec + off_cnt;
      uint16_t* rec = reinterpret_cast<uint16_t*>(sec + off_rec);
      h.blob_off = b0;  // relative to the part's buffer for now
      h.run0 = (uint32_t)runsv.size();
      h.sec16 = (uint16_t)(sec_bytes / 16);
      h.nslots = (uint16_t)ns;
      h.nstage = (uint16_t)nst;
      h.ndiag = (uint16_t)ngroups;
      h.tail_src = 0;
      h.tail_dst = 0;
      part_maxsec[p] = std::max(part_maxsec[p], sec_bytes);
      // stage tables and runs
      uint32_t vals = 0, top = 0;
      int nr = 0;
      for (int t = st0; t < st1; ++t) {
        const int col = P.stage_cols[t];
        stab[2 * (t - st0)] = (uint32_t)(S.slice_off[col >> 5] + (col & 31));
        stab[2 * (t - st0) + 1] = (uint32_t)P.stage_off[t] * 8u;
        const int run = P.stage_run[t];
        if (run > 0) {
          const int a = colptr[col] & 1;
          Run2 r;
          r.src = (uint32_t)(colptr[col] - a);
          r.dst = (uint16_t)(P.stage_off[t] - a);
          r.len = (uint16_t)run;
          top = std::max<uint32_t>(top, (uint32_t)r.dst + r.len);
          if ((int64_t)r.src + r.len > nnzJ) {
            // the padding element of the window lies past the value array: drop the last two elements from the
            // bulk copy and fetch the last real one by hand
            r.len = (uint16_t)(r.len - 2);
            h.tail_src = r.src + r.len + 1u;
            h.tail_dst = (uint16_t)(r.dst + r.len);
          }
          vals += (uint32_t)r.len * 8u;
          runsv.push_back(r);  // zero-length runs stay in the list (the kernel skips them)
          ++nr;
        }
      }
      if (nr > 65535) {
        part_fail[p] = "too many runs";
        return;
      }
      h.nruns = (uint16_t)nr;
      h.val_bytes = vals;
      part_maxval[p] = std::max(part_maxval[p], top * 8u);
      for (int g = 0; g < ngroups; ++g) ginfo[g] = (uint32_t)grow[g] | ((uint32_t)(grow[(size_t)g + 1] - grow[g]) << 16);
      for (int k = 0; k < ns; ++k) {
        const int s = order[k];
        meta[k] = (uint32_t)P.slot_il[s] | ((uint32_t)P.slot_jl[s] << 8) | ((uint32_t)P.low_rank[s] << 16) | ((uint32_t)P.low_rank_up[s] << 24);
        cnt[k] = (uint8_t)P.slot_cnt[s];
        uint16_t* rg = rec + (size_t)grow[k >> 5] * 32 + (k & 31);
        for (int d = 0; d < (int)P.slot_cnt[s]; ++d) rg[(size_t)d * 32] = (uint16_t)(P.rec16[(size_t)(rec0 + P.slot_start[s] + d)] & 0x3fffu);
      }
    }
  });
  for (int p = 0; p < parts; ++p)
    if (part_fail[p]) return fail(part_fail[p]);
  // concatenate the parts
  uint64_t boff = 0, roff = 0;
  std::vector<uint64_t> pb((size_t)parts + 1, 0), pr((size_t)parts + 1, 0);
  for (int p = 0; p < parts; ++p) {
    pb[(size_t)p + 1] = pb[p] + pblob[p].size();
    pr[(size_t)p + 1] = pr[p] + pruns[p].size();
  }
  boff = pb[parts];
  roff = pr[parts];
  if (roff >= ((uint64_t)1 << 32)) return fail("run list beyond 32-bit offsets");
  out.blob.resize((size_t)boff);
  out.runs.resize((size_t)roff);
  run_parts(parts, [&](int p) {
    if (!pblob[p].empty()) memcpy(out.blob.data() + pb[p], pblob[p].data(), pblob[p].size());
    if (!pruns[p].empty()) memcpy(out.runs.data() + pr[p], pruns[p].data(), pruns[p].size() * sizeof(Run2));
    for (int c = (int)((int64_t)nch * p / parts); c < (int)((int64_t)nch * (p + 1) / parts); ++c) {
      out.hdr[c].blob_off += pb[p];
      out.hdr[c].run0 += (uint32_t)pr[p];
    }
    std::vector<unsigned char>().swap(pblob[p]);
  });
  for (int p = 0; p < parts; ++p) {
    out.max_val_bytes = std::max(out.max_val_bytes, part_maxval[p]);
    out.max_sec_bytes = std::max(out.max_sec_bytes, part_maxsec[p]);
  }
  out.ok = true;
}

void run_pairs_v2_host(const PairsV2Host& V, const double* Jv, double* hval) {
  std::vector<double> val(kStageVals + 8);
  for (const ChunkHdr2& h : V.hdr) {
    // what the producer warp does
    for (int r = 0; r < (int)h.nruns; ++r) {
      const Run2& ru = V.runs[(size_t)h.run0 + r];
      for (int e = 0; e < (int)ru.len; ++e) val[(size_t)ru.dst + e] = Jv[(size_t)ru.src + e];
    }
    if (h.tail_src) val[h.tail_dst] = Jv[(size_t)h.tail_src - 1];
    // what the consumer warps do
    const V2Section sv(V.blob.data() + h.blob_off);
    for (int t = 0; t < sv.ns; ++t) {
      const uint32_t m = sv.meta[t];
      const int il = m & 255, jl = (m >> 8) & 255, rk = (m >> 16) & 255, rku = m >> 24;
      const double* vi = val.data() + sv.stab[2 * il + 1] / 8;
      const double* vj = val.data() + sv.stab[2 * jl + 1] / 8;
      const uint16_t* rg = sv.rec + (size_t)(sv.ginfo[t >> 5] & 0xffffu) * 32 + (t & 31);
      double acc = 0.0;
      for (int d = 0; d < (int)sv.cnt[t]; ++d) {
        const uint32_t r = rg[(size_t)d * 32];
        const double prod = vi[(r >> 7) & 127] * vj[r & 127];
        acc = d == 0 ? prod : acc + prod;
      }
      hval[(size_t)sv.stab[2 * il] + (size_t)rk * 32] = acc;
      hval[(size_t)sv.stab[2 * jl] + (size_t)rku * 32] = acc;
    }
  }
}

}  // namespace spb
