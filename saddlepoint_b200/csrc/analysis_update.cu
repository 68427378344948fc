// analysis_update.cu -- incremental re-analysis when the Jacobian gains one residual row.
//
// The reference's iterative-rounding driver (benchmark/seamless_integration/main.cpp:69-80) pins one more variable
// per outer iteration: IterativeRoundingTraits.h:76-79 appends one row to the constraint matrix, so between two
// LMSolver::solve calls the Jacobian gains ONE residual row with one nonzero (all rows below it -- the barrier block --
// move down by one).  A from-scratch analysis of the 500k-face problem costs 1.8 s (+ ~6 s of ordering / supernodes for
// the direct solver) against 0.8 s per LM iteration.  Here the analysed state is PATCHED:
//   * the normal-matrix pattern, the SpMV layout, the slot metadata and the Cholesky / IC(0) symbolic state do not
//     change (a single-entry row only adds to a diagonal entry of J^T J);
//   * the J pattern on the device is rewritten by one kernel (insert one entry, relabel the rows below);
//   * the pair records of the ~25 slots that touch the changed column get the positions behind the insertion point
//     shifted by one;
//   * the stage layouts of all chunks (shared-memory offsets, bulk-copy runs, TMA byte counts) are recomputed from the
//     new column pointers -- they depend on the parity of every column's offset in the value array.
// Anything else (several nonzeros whose pairs are not all in H yet, a column outgrowing the record format, a chunk
// outgrowing its shared-memory stage) falls back to the full analysis; the Cholesky symbolic state survives that too
// whenever the pattern of H is unchanged (context.cu, spb_analyze).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "spb_internal.h"

namespace spb {

namespace {
// new J row indices from the old ones: one entry (row rho) inserted at position pos, rows >= rho move down by one
__global__ void jrow_insert_kernel(const int* __restrict__ old_row, int* __restrict__ new_row, int64_t nnz_new, int64_t pos, int rho) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz_new; e += (int64_t)gridDim.x * blockDim.x) {
    if (e == pos) {
      new_row[e] = rho;
    } else {
      const int r = old_row[e < pos ? e : e - 1];
      new_row[e] = r + (r >= rho ? 1 : 0);
    }
  }
}
__global__ void jcolptr_shift_kernel(int* __restrict__ colptr, int n, int c) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += gridDim.x * blockDim.x)
    if (j > c) colptr[j] += 1;
}

// Stage layouts of all chunks from the (already updated) device column pointers -- the rules of symbolic.cpp's
// close_chunk / assembly.cu's upload_analysis: a run of consecutive listed columns is staged by one bulk copy whose
// window starts 16-byte aligned (one element early when the head column starts at an odd position of the value array)
// and has an even length; inside a run the columns are back to back.  One thread per chunk (a list has ~60 columns).
// flags_out[0]: a staged chunk outgrew the shared-memory stage; flags_out[1]: some chunk is off the pipelined kernel's
// common path.
__global__ void restage_kernel(int nch, const int* __restrict__ chunk_stage_ptr, const int* __restrict__ stage_cols,
                               const int* __restrict__ colptr, int n, StageEntry* __restrict__ se, ChunkDesc* __restrict__ cd,
                               ChunkDesc* __restrict__ cd_sorted, const int* __restrict__ sorted_pos, int* __restrict__ flags_out) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= nch) return;
  const int t0 = chunk_stage_ptr[ch], t1 = chunk_stage_ptr[ch + 1];
  ChunkDesc d = cd[ch];
  const int64_t nnz = colptr[n];
  unsigned flags = d.flags & 1u;
  int cur = 0, run_head = -1, run_len = 0, bytes = 0;
  bool inside = true;
  for (int t = t0; t <= t1; ++t) {
    const int col = t < t1 ? stage_cols[t] : -1;
    if (t == t1 || t == t0 || col != stage_cols[t - 1] + 1) {
      if (run_head >= 0) {  // close the open run
        const int padded = (run_len + 1) & ~1;
        const int start = se[run_head].src, a = start & 1;
        se[run_head].slice_base |= (int64_t)min(padded, 65535) << 48;
        if ((int64_t)start - a + padded > nnz) inside = false;
        bytes += padded * 8;
        cur += padded;
      }
      if (t == t1) break;
      run_head = t;
      run_len = colptr[col] & 1;
    }
    const int start = colptr[col], len = colptr[col + 1] - start;
    StageEntry e;
    e.src = start;
    e.off = (uint16_t)min(cur + run_len, 65535);
    e.len = (uint16_t)min(len, 65535);
    e.slice_base = se[t].slice_base & (((int64_t)1 << 48) - 1);
    se[t] = e;
    if (len > 32) flags |= 2u;
    run_len += len;
  }
  d.stage_bytes = 0;
  if (flags & 1u) {
    if (cur > kStageVals) atomicOr(&flags_out[0], 1);
    if (inside) flags |= 4u;
    d.stage_bytes = bytes;
  }
  d.flags = (uint16_t)flags;
  if (!((flags & 1u) && (flags & 4u) && d.npairs <= kPairTile && d.nstage <= kStageCols)) atomicOr(&flags_out[1], 1);
  cd[ch] = d;
  cd_sorted[sorted_pos[ch]] = d;
}
}  // namespace

// Tries the incremental path; returns false (state untouched) if the change is not of the supported form.
static bool try_append_row(spb_context* h, int m_new, int n, const int* colptr, const int* rowidx) {
  PlanKeep& K = h->keep;
  if (!h->analyzed || !K.valid || h->partitioned || h->halo_mode || h->trip_nnz >= 0) return false;
  if (n != h->n || m_new != h->m + 1 || (int64_t)colptr[n] != h->nnzJ + 1) return false;
  const std::vector<int>& ocp = h->Jcolptr_host;
  const std::vector<int>& ori = K.Jrow;
  const bool timing = getenv("SPB_ANALYZE_TIMING") && atoi(getenv("SPB_ANALYZE_TIMING")) > 0;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[spb analyze-update timing] %s %.4f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  };
  // the changed column: the first whose end pointer moved
  int c = -1;
  for (int j = 0; j < n; ++j)
    if (colptr[j + 1] != ocp[(size_t)j + 1]) {
      c = j;
      break;
    }
  if (c < 0) return false;
  const int q0 = ocp[c], q1 = ocp[(size_t)c + 1];
  // Every other column: the column pointers move by one behind c; every row index is either unchanged (old row < rho)
  // or one larger (old row >= rho).  One parallel pass checks that and brackets rho: lo = largest unchanged old row,
  // hi = smallest shifted old row, so lo < rho <= hi.
  int lo = -1, hi = h->m;
  {
    std::atomic<int> bad{0};
    const int parts = std::max(1, std::min(host_threads(), n / 65536 + 1));
    std::vector<int> plo((size_t)parts, -1), phi((size_t)parts, h->m);
    run_parts(parts, [&](int p) {
      const int j0 = (int)((int64_t)n * p / parts), j1 = (int)((int64_t)n * (p + 1) / parts);
      int l = -1, u = h->m;
      for (int j = j0; j < j1 && !bad.load(std::memory_order_relaxed); ++j) {
        if (colptr[j] != ocp[j] + (j > c ? 1 : 0) || colptr[j + 1] != ocp[(size_t)j + 1] + (j >= c ? 1 : 0)) {
          bad = 1;
          return;
        }
        if (j == c) continue;
        const int* nr = rowidx + colptr[j];
        const int* orow = ori.data() + ocp[j];
        const int len = ocp[(size_t)j + 1] - ocp[j];
        for (int t = 0; t < len; ++t) {
          const int o = orow[t], v = nr[t];
          if (v == o) l = std::max(l, o);
          else if (v == o + 1) u = std::min(u, o);
          else {
            bad = 1;
            return;
          }
        }
      }
      plo[(size_t)p] = l;
      phi[(size_t)p] = u;
    });
    if (bad) return false;
    for (int p = 0; p < parts; ++p) {
      lo = std::max(lo, plo[(size_t)p]);
      hi = std::min(hi, phi[(size_t)p]);
    }
    if (lo >= hi) return false;
  }
  // Column c: the new entry sits at position q iff the entries before q are unchanged (rows < rho), the old entries from
  // q on reappear one position later and one row lower (rows >= rho), and rho = rowidx[q] lies in (lo, hi].  (Inside a run
  // of consecutive rows several positions explain the column alone; the bracket from the other columns decides.  If
  // it leaves several, the rows in between are empty elsewhere and every choice describes the same matrix.)
  int pos = -1, rho = -1;
  {
    const int len = q1 - q0;
    std::vector<uint8_t> suf((size_t)len + 1, 1);  // suf[t]: old entries t.. of the column are new entries t+1.. shifted
    for (int t = len - 1; t >= 0; --t) suf[(size_t)t] = suf[(size_t)t + 1] && rowidx[q0 + t + 1] == ori[(size_t)q0 + t] + 1;
    bool pre = true;  // old entries ..t-1 are new entries ..t-1 unchanged
    for (int t = 0; t <= len && pre; ++t) {
      const int cand = rowidx[q0 + t];
      if (suf[(size_t)t] && cand > lo && cand <= hi && (t == 0 || ori[(size_t)q0 + t - 1] < cand) && (t == len || ori[(size_t)q0 + t] >= cand)) {
        pos = q0 + t;
        rho = cand;
        break;
      }
      if (t < len) pre = rowidx[q0 + t] == ori[(size_t)q0 + t];
    }
  }
  if (pos < 0 || rho < 0 || rho > h->m) return false;
  lap("verify the change");
  // the new row has ONE entry: it adds only to the diagonal of J^T J (the column kernel's job); no new H entry, no
  // new pair.  (A row with several entries would also need new pair records: full analysis.)
  const int newlen = q1 - q0 + 1;
  if (newlen > (h->rec_bits == 16 ? 128 : 32768)) return false;
  const int qrel = pos - q0;  // position of the new entry inside column c: old positions >= qrel move up by one

  // ---- from here on the handle is modified ---------------------------------------------------------------------------
  cudaStream_t st = h->stream;
  h->h_valid = h->factor_valid = false;
  // (1) pair records of the slots that touch column c: positions >= qrel inside column c move up by one
  {
    const int pobits = h->rec_bits == 16 ? 7 : 15;
    const uint32_t mask = (1u << pobits) - 1u;
    auto patch_slot = [&](int s, bool c_is_row /* column c is the slot's row-column i (po) */) {
      const int ch = (int)(std::upper_bound(K.chunk_slot.begin(), K.chunk_slot.end(), s) - K.chunk_slot.begin()) - 1;
      const int cnt = K.slot_cnt[s];
      if (cnt >= 65535) throw StateFail{"slot too long for the incremental update"};
      const int64_t at = K.chunk_rec[ch] + K.slot_start[s];
      if (h->rec_bits == 16) {
        std::vector<uint16_t> r((size_t)cnt);
        SPB_CUDA(cudaMemcpyAsync(r.data(), h->d_rec16.p + at, sizeof(uint16_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        for (auto& v : r) {
          uint32_t po = (v >> pobits) & mask, qo = v & mask, first = v >> (2 * pobits);
          if (c_is_row) po += (int)po >= qrel ? 1u : 0u; else qo += (int)qo >= qrel ? 1u : 0u;
          v = (uint16_t)((first << (2 * pobits)) | (po << pobits) | qo);
        }
        SPB_CUDA(cudaMemcpyAsync(h->d_rec16.p + at, r.data(), sizeof(uint16_t) * (size_t)cnt, cudaMemcpyHostToDevice, st));
        SPB_CUDA(cudaStreamSynchronize(st));
      } else {
        std::vector<uint32_t> r((size_t)cnt);
        SPB_CUDA(cudaMemcpyAsync(r.data(), h->d_rec32.p + at, sizeof(uint32_t) * (size_t)cnt, cudaMemcpyDeviceToHost, st));
        SPB_CUDA(cudaStreamSynchronize(st));
        for (auto& v : r) {
          uint32_t po = (v >> pobits) & mask, qo = v & mask, first = v >> (2 * pobits);
          if (c_is_row) po += (int)po >= qrel ? 1u : 0u; else qo += (int)qo >= qrel ? 1u : 0u;
          v = (first << (2 * pobits)) | (po << pobits) | qo;
        }
        SPB_CUDA(cudaMemcpyAsync(h->d_rec32.p + at, r.data(), sizeof(uint32_t) * (size_t)cnt, cudaMemcpyHostToDevice, st));
        SPB_CUDA(cudaStreamSynchronize(st));
      }
    };
    // slots of H column c (c is the slot's column j: qo)
    for (int s = K.low_colptr[c]; s < K.low_colptr[(size_t)c + 1]; ++s) patch_slot(s, false);
    // slots (i = c, j < c): the columns j are the rows < c of H's column c (symmetric pattern)
    for (int q = h->H.colptr[c]; q < h->H.colptr[c] + h->H.nup1[c] - 1; ++q) {
      const int j = h->H.rowidx[q];
      const int* lo = K.low_row.data() + K.low_colptr[j];
      const int* hi = K.low_row.data() + K.low_colptr[(size_t)j + 1];
      const int* it = std::lower_bound(lo, hi, c);
      if (it != hi && *it == c) patch_slot((int)(it - K.low_row.data()), true);  // absent: a structural slot without pairs
    }
  }
  lap("patch pair records");
  // (2) J pattern on the device
  {
    DevBuf<int> nr;
    SPB_CUDA(nr.alloc((size_t)h->nnzJ + 1));
    const int g = std::max(1, std::min((int)((h->nnzJ + 256) / 256), h->num_sms * 8));
    jrow_insert_kernel<<<g, 256, 0, st>>>(h->d_Jrow.p, nr.p, h->nnzJ + 1, (int64_t)pos, rho);
    jcolptr_shift_kernel<<<std::max(1, std::min((n + 256) / 256, h->num_sms * 8)), 256, 0, st>>>(h->d_Jcolptr.p, n, c);
    h->launches += 2;
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaStreamSynchronize(st));
    std::swap(h->d_Jrow.p, nr.p);
    std::swap(h->d_Jrow.n, nr.n);
  }
  lap("J pattern on the device");
  // (3) stage entries and chunk descriptors (both copies), recomputed on the device from the new column pointers
  const int nch = (int)K.chunk_npairs.size();
  h->v2_ok = false;  // the packed stream form (assembly_v2.cu) describes the old J layout: the patched general kernels take over
  {
    if (!h->d_keep_stage_cols.p) {  // first update on this analysis: the stage lists go to the device once
      SPB_CUDA(h->d_keep_stage_cols.upload(K.stage_cols, st));
      SPB_CUDA(h->d_keep_stage_ptr.upload(K.chunk_stage_ptr, st));
      SPB_CUDA(h->d_keep_sorted_pos.upload(K.sorted_pos, st));
      SPB_CUDA(h->d_keep_flags.alloc(2));
    }
    SPB_CUDA(cudaMemsetAsync(h->d_keep_flags.p, 0, 2 * sizeof(int), st));
    restage_kernel<<<(nch + 127) / 128, 128, 0, st>>>(nch, h->d_keep_stage_ptr.p, h->d_keep_stage_cols.p, h->d_Jcolptr.p, n, h->d_stage_ent.p,
                                                    h->d_chunk_desc.p, h->d_chunk_desc_sorted.p, h->d_keep_sorted_pos.p, h->d_keep_flags.p);
    h->launches += 1;
    SPB_CUDA(cudaGetLastError());
    int fl[2] = {0, 0};
    SPB_CUDA(cudaMemcpyAsync(fl, h->d_keep_flags.p, sizeof(fl), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    if (fl[0]) throw StateFail{"a chunk outgrew its shared-memory stage"};  // the full analysis re-plans
    h->pairs_fast_path = !fl[1];
  }
  lap("stage layouts (device)");
  // streamed assembly from pinned host memory: how much of J a chunk needs (its last listed column, plus the padding
  // element of a bulk-copy window).  The launch order is kept; the needs are kept non-decreasing along it.
  {
    int64_t prev = 0;
    const int64_t nnz_new = colptr[n];
    std::vector<int64_t> need((size_t)nch);
    for (int ch = 0; ch < nch; ++ch) {
      const int t1 = K.chunk_stage_ptr[(size_t)ch + 1];
      need[(size_t)ch] = t1 > K.chunk_stage_ptr[ch] ? std::min<int64_t>((int64_t)colptr[K.stage_cols[(size_t)t1 - 1] + 1] + 1, nnz_new) : 0;
    }
    for (int ch = 0; ch < nch; ++ch) h->chunk_need[(size_t)K.sorted_pos[ch]] = need[(size_t)ch];
    for (int i = 0; i < nch; ++i) {
      prev = std::max(prev, h->chunk_need[(size_t)i]);
      h->chunk_need[(size_t)i] = prev;
    }
  }
  lap("chunk needs for the streamed assembly");
  // (4) host copies (the row array on the host threads)
  for (int j = c + 1; j <= n; ++j) h->Jcolptr_host[(size_t)j] += 1;
  {
    const int64_t nnz_new = colptr[n];
    K.Jrow.resize((size_t)nnz_new);
    const int parts = std::max(1, std::min(host_threads(), (int)(nnz_new / 4000000) + 1));
    run_parts(parts, [&](int p) {
      const int64_t a = nnz_new * p / parts, b = nnz_new * (p + 1) / parts;
      std::memcpy(K.Jrow.data() + a, rowidx + a, sizeof(int) * (size_t)(b - a));
    });
  }
  h->m = m_new;
  h->m_global = m_new;
  h->nnzJ = colptr[n];
  h->fingerprint = pattern_fingerprint(m_new, n, colptr, rowidx);
  lap("host copies, fingerprint");
  return true;
}

}  // namespace spb

using namespace spb;

// how: 0 the pattern is the analysed one, 1 patched incrementally, 2 analysed from scratch
extern "C" spb_status spb_analyze_update(spb_handle h, int32_t m, int32_t n, const int32_t* colptr, const int32_t* rowidx, int* how) {
  if (!h || !colptr || (colptr[n] > 0 && !rowidx)) return SPB_ERR_ARG;
  if (how) *how = 2;
  try {
    cudaGetLastError();
    SPB_CUDA(cudaSetDevice(h->device));
    int same = 0;
    spb_pattern_matches(h, m, n, colptr, rowidx, &same);
    if (same) {
      if (how) *how = 0;
      return SPB_OK;
    }
    static const bool off = getenv("SPB_ANALYZE_UPDATE") && atoi(getenv("SPB_ANALYZE_UPDATE")) == 0;
    if (!off && try_append_row(h, m, n, colptr, rowidx)) {
      if (how) *how = 1;
      return SPB_OK;
    }
  } catch (const CudaFail& e) {
    h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    h->analyzed = false;
    return SPB_ERR_CUDA;
  } catch (const StateFail&) {
    // an unsupported corner met after the state was touched: the full analysis below rebuilds everything
  } catch (const std::exception& e) {
    h->err = std::string("internal error: ") + e.what();
    h->analyzed = false;
    return SPB_ERR_STATE;
  }
  return spb_analyze(h, m, n, colptr, rowidx);
}
