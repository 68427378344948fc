// assembly.cu -- fixed-pattern numeric assembly of H = dampJ^T*dampJ, rhs = -(J^T f), the
// Marquardt damping vector and ||rhs||_inf, for sm_100a.
//
// Replaces, per LM iteration (all on the host in the reference):
//   rhs = -(J.transpose()*EVec)               LMSolver.h:117   (sequential sum per column)
//   fooOptimality = rhs.lpNorm<Infinity>()    LMSolver.h:119
//   dampVector(j) += lambda*v*v ; sqrt ; augmented triplet rebuild   DiagonalDamping.h:31-43,73-85
//   dampJ.transpose()*dampJ                   LMSolver.h:136   (Gustavson product, pattern redone)
//
// Arithmetic contract: every H entry, every rhs entry and every damping entry is accumulated
// in the reference's own order (ascending residual row, damping term last on the diagonal)
// with separate IEEE multiply and add (__dmul_rn/__dadd_rn: no FMA contraction, matching the
// reference's x86-64 -O3 build), so the results are bit-identical to the CPU path.
//
// Two kernels, both "coalesced stage -> ordered segment sum":
//  * assemble_pairs: one CTA per chunk of strictly-lower H slots (whole H columns).  The chunk's
//    stage list -- the J columns its pairs touch: its own columns j and the row-columns i of
//    its slots -- is copied into shared memory once (one warp per column, coalesced), as are
//    the chunk's pair records ([first:1][po][qo], 2 bytes for J columns <=128 entries).  One
//    thread per slot then walks its records in order, multiplies the two staged J values and
//    adds the products left to right (first product assigned), and writes H(i,j) and its mirror
//    H(j,i) straight into the SpMV storage (sliced ELL), so no second pass over H is needed.
//    All loads that do not depend on another load (slot metadata, records, stage entries) are
//    issued before the first barrier; the only dependent chain is stage entry -> J values.
//    Chunks whose stage list does not fit in shared memory (very long J columns) gather from
//    global memory instead; a single slot with more pairs than a tile (dense column pairs) is
//    walked by one thread from global memory.
//  * assemble_cols: one CTA per chunk of J columns.  Values, row indices and the gathered
//    residuals are staged coalesced in (bank-skewed) shared memory, then one thread per
//    column walks its column in order producing sum v*f, sum (lambda*v)*v and sum v*v, and
//    emits rhs, d, sqrt(d)^2 onto the diagonal, the Jacobi diagonal, and a block max of
//    |rhs| folded with atomicMax on the (non-negative) bit pattern.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "spb_internal.h"
#include "tma.cuh"

namespace spb {

constexpr int64_t kSliceBaseMask = ((int64_t)1 << 48) - 1;  // StageEntry::slice_base: low 48 bits position, high 16 bits run length

// ordered sum of one slot's products: first product assigned, the rest added left to right
template <typename RecT, int POBITS, bool SMEM>
__device__ __forceinline__ double slot_sum(const RecT* rec, const double* vi, const double* vj, int beg, int end) {
  constexpr uint32_t MASK = (1u << POBITS) - 1u;
  uint32_t r = (uint32_t)rec[beg];
  double a = SMEM ? vi[(r >> POBITS) & MASK] : __ldg(vi + ((r >> POBITS) & MASK));
  double b = SMEM ? vj[r & MASK] : __ldg(vj + (r & MASK));
  double acc = __dmul_rn(a, b);
  int q = beg + 1;
  // Four pairs at a time: records, then both operands of every pair, are loaded before the first
  // dependent use, so the longest slot of a chunk (the CTA's critical path between two barriers)
  // costs one shared-memory round trip per four pairs instead of two per pair.  The products are
  // still added strictly left to right.
  for (; q + 4 <= end; q += 4) {
    const uint32_t r0 = (uint32_t)rec[q], r1 = (uint32_t)rec[q + 1], r2 = (uint32_t)rec[q + 2], r3 = (uint32_t)rec[q + 3];
    const double a0 = SMEM ? vi[(r0 >> POBITS) & MASK] : __ldg(vi + ((r0 >> POBITS) & MASK));
    const double a1 = SMEM ? vi[(r1 >> POBITS) & MASK] : __ldg(vi + ((r1 >> POBITS) & MASK));
    const double a2 = SMEM ? vi[(r2 >> POBITS) & MASK] : __ldg(vi + ((r2 >> POBITS) & MASK));
    const double a3 = SMEM ? vi[(r3 >> POBITS) & MASK] : __ldg(vi + ((r3 >> POBITS) & MASK));
    const double b0 = SMEM ? vj[r0 & MASK] : __ldg(vj + (r0 & MASK));
    const double b1 = SMEM ? vj[r1 & MASK] : __ldg(vj + (r1 & MASK));
    const double b2 = SMEM ? vj[r2 & MASK] : __ldg(vj + (r2 & MASK));
    const double b3 = SMEM ? vj[r3 & MASK] : __ldg(vj + (r3 & MASK));
    acc = __dadd_rn(acc, __dmul_rn(a0, b0));
    acc = __dadd_rn(acc, __dmul_rn(a1, b1));
    acc = __dadd_rn(acc, __dmul_rn(a2, b2));
    acc = __dadd_rn(acc, __dmul_rn(a3, b3));
  }
  for (; q < end; ++q) {
    r = (uint32_t)rec[q];
    a = SMEM ? vi[(r >> POBITS) & MASK] : __ldg(vi + ((r >> POBITS) & MASK));
    b = SMEM ? vj[r & MASK] : __ldg(vj + (r & MASK));
    acc = __dadd_rn(acc, __dmul_rn(a, b));
  }
  return acc;
}

template <typename RecT, int POBITS>
__global__ void __launch_bounds__(256, 8) assemble_pairs_kernel(const RecT* __restrict__ recs, const ChunkDesc* __restrict__ chunks,
                                                             const StageEntry* __restrict__ stage, const SlotMeta* __restrict__ slots,
                                                             const uint32_t* __restrict__ slot_start, const double* __restrict__ Jv,
                                                             double* __restrict__ hval, int bulk_ok) {
  __shared__ __align__(16) double sval[kStageVals];
  __shared__ uint64_t bar;
  __shared__ RecT srec[kPairTile];
  __shared__ int s_src[kStageCols];
  __shared__ unsigned s_offlen[kStageCols];
  __shared__ int64_t s_sbase[kStageCols];
  const int tid = threadIdx.x;
  const ChunkDesc cd = chunks[blockIdx.x];
  const int s0 = cd.slot0, ns = cd.nslots, np = cd.npairs, st0 = cd.stage0, nst = cd.nstage;
  const int64_t rec0 = cd.rec0;
  const bool staged = (cd.flags & 1) != 0;
  const bool bulk = bulk_ok && (cd.flags & 4);  // every column's 16-byte-aligned window is inside the J value array
  const bool small_list = nst <= kStageCols;
  const bool one_tile = np <= kPairTile;
  constexpr int SPT = kSlotCap / 256;  // slots per thread
  if (bulk) {
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();  // the barrier object must be initialised before any copy can signal it
    if (tid == 0) mbar_expect_tx(&bar, (uint32_t)cd.stage_bytes);
  }

  // Independent loads first: per-slot metadata, the chunk's pair records (coalesced, into shared
  // memory) and the stage list.  The only dependent chain left is stage entry -> J values.
  SlotMeta mreg[SPT];
  int beg[SPT], end[SPT];
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    const int t = tid + k * 256;
    beg[k] = end[k] = 0;
    if (t < ns) {
      mreg[k] = slots[s0 + t];
      const uint32_t sc = slot_start[s0 + t];  // (start << 16) | count; a multi-tile slot is alone in its chunk
      beg[k] = (int)(sc >> 16);
      end[k] = ns == 1 ? np : beg[k] + (int)(sc & 0xffffu);
    }
  }
  if (one_tile)
    for (int t = tid; t < np; t += 256) srec[t] = recs[rec0 + t];
  if (small_list) {
    for (int t = tid; t < nst; t += 256) {
      const StageEntry e = stage[st0 + t];
      s_src[t] = e.src;
      s_offlen[t] = ((unsigned)e.off << 16) | e.len;
      s_sbase[t] = e.slice_base & kSliceBaseMask;
      if (staged && bulk) {
        // TMA: one bulk async copy per RUN of consecutive listed J columns (16-byte-aligned window),
        // issued by the thread that just read the run's head entry, all completing on one mbarrier
        // -- no per-element instructions, and the copy is in flight while the other loads land.
        const int run = (int)((uint64_t)e.slice_base >> 48);
        const int a = e.src & 1;
        if (run > 0) tma_bulk_g2s(sval + ((int)e.off - a), Jv + (e.src - a), (uint32_t)run * 8u, &bar);
      }
    }
  }
  __syncthreads();
  if (staged && bulk) {
    mbar_wait(&bar, 0);
  } else if (staged) {
    // Copy the listed J columns into shared memory: one warp per column and pass (columns of up
    // to 32 entries in one go), four columns in flight per thread before the first store.
    const int sub = tid & 31, cg = tid >> 5;
    for (int t0 = 0; t0 < nst; t0 += 32) {
      double v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u * 8 + cg;
        dst[u] = -1;
        if (t < nst) {
          const unsigned ol = s_offlen[t];
          if (sub < (int)(ol & 0xffffu)) {
            dst[u] = (int)(ol >> 16) + sub;
            v[u] = Jv[s_src[t] + sub];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) sval[dst[u]] = v[u];
    }
    if (cd.flags & 2) {  // columns longer than 32 entries: the rest of each column
      for (int t = cg; t < nst; t += 8) {
        const unsigned ol = s_offlen[t];
        const int off = (int)(ol >> 16), len = (int)(ol & 0xffffu), src = s_src[t];
        for (int e = 32 + sub; e < len; e += 32) sval[off + e] = Jv[src + e];
      }
    }
    __syncthreads();
  }

  // One thread per slot: products added in ascending residual row, first product assigned
  // (exactly the accumulation order of the reference's sparse product).
#pragma unroll
  for (int k = 0; k < SPT; ++k) {
    const int t = tid + k * 256;
    if (t < ns) {
      const SlotMeta mt = mreg[k];
      int bx, by;
      int64_t bi, bj;
      if (small_list) {
        bx = staged ? (int)(s_offlen[mt.il] >> 16) : s_src[mt.il];
        by = staged ? (int)(s_offlen[mt.jl] >> 16) : s_src[mt.jl];
        bi = s_sbase[mt.il];
        bj = s_sbase[mt.jl];
      } else {
        const StageEntry ei = stage[st0 + mt.il], ej = stage[st0 + mt.jl];
        bx = ei.src;
        by = ej.src;
        bi = ei.slice_base & kSliceBaseMask;
        bj = ej.slice_base & kSliceBaseMask;
      }
      double acc;
      if (staged && one_tile) acc = slot_sum<RecT, POBITS, true>(srec, sval + bx, sval + by, beg[k], end[k]);
      else if (one_tile) acc = slot_sum<RecT, POBITS, false>(srec, Jv + bx, Jv + by, beg[k], end[k]);
      else if (staged) acc = slot_sum<RecT, POBITS, true>(recs + rec0, sval + bx, sval + by, beg[k], end[k]);
      else acc = slot_sum<RecT, POBITS, false>(recs + rec0, Jv + bx, Jv + by, beg[k], end[k]);
      hval[bi + (int64_t)mt.rank * 32] = acc;
      hval[bj + (int64_t)mt.rank_up * 32] = acc;
    }
  }
}

// ---- pipelined variant: a persistent CTA walks a contiguous range of chunks --------------------
// assemble_pairs_kernel above is one dependent chain per CTA (descriptor -> stage entries / records /
// slot metadata -> TMA copies -> sums -> stores, ~3.5 us) and relies on 8 co-resident CTAs per SM to
// hide it.  Here the chain is software-pipelined three chunks deep inside one CTA: while chunk k is
// summed from shared-memory buffer k&1, the TMA copies of chunk k+1 are in flight into the other
// buffer and the stage entries and slot metadata of chunk k+2 are on their way into registers; the
// pair records follow one chunk behind through cp.async.  Used when every chunk is on the common
// path (staged in shared memory, TMA-safe windows, one record tile, short stage list); the sums,
// their order and the stores are the same code as above, so the results are the same bits.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

constexpr int kDescRing = 64;
template <typename RecT>
constexpr size_t pipelined_smem_bytes() {
  return sizeof(double) * 2 * kStageVals + sizeof(int64_t) * 2 * kStageCols + sizeof(ChunkDesc) * kDescRing + 16 + sizeof(unsigned) * 2 * kStageCols +
         sizeof(RecT) * 2 * kPairTile;
}

template <typename RecT, int POBITS>
__global__ void __launch_bounds__(256, 4) assemble_pairs_pipelined_kernel(const RecT* __restrict__ recs, const ChunkDesc* __restrict__ chunks,
                                                                          int nchunks, const StageEntry* __restrict__ stage,
                                                                          const SlotMeta* __restrict__ slots,
                                                                          const uint32_t* __restrict__ slot_start,
                                                                          const double* __restrict__ Jv, double* __restrict__ hval,
                                                                          long long* dbg, int strided) {
  // dynamic shared memory (above the 48 KB static limit): two value buffers, two record buffers, two stage
  // arrays, the descriptor ring and the two mbarriers -- see pipelined_smem_bytes()
  extern __shared__ __align__(16) unsigned char pipe_smem[];
  double(*sval)[kStageVals] = reinterpret_cast<double(*)[kStageVals]>(pipe_smem);
  int64_t(*s_sbase)[kStageCols] = reinterpret_cast<int64_t(*)[kStageCols]>(pipe_smem + sizeof(double) * 2 * kStageVals);
  ChunkDesc* ring = reinterpret_cast<ChunkDesc*>(reinterpret_cast<unsigned char*>(s_sbase) + sizeof(int64_t) * 2 * kStageCols);
  uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(ring) + sizeof(ChunkDesc) * kDescRing);
  unsigned(*s_offlen)[kStageCols] = reinterpret_cast<unsigned(*)[kStageCols]>(reinterpret_cast<unsigned char*>(bar) + 16);
  RecT(*srec)[kPairTile] = reinterpret_cast<RecT(*)[kPairTile]>(reinterpret_cast<unsigned char*>(s_offlen) + sizeof(unsigned) * 2 * kStageCols);
  long long tk = 0, acc_t[6] = {0, 0, 0, 0, 0, 0};
  const bool dbg_on = dbg && blockIdx.x == 7 && (threadIdx.x == 0 || threadIdx.x == 64 || threadIdx.x == 224);
#define ASM_TICK(i)                     \
  if (dbg_on) {                         \
    const long long now_ = clock64();   \
    acc_t[i] += now_ - tk;              \
    tk = now_;                          \
  }
  const int tid = threadIdx.x;
  // Which chunks this CTA walks.  strided: chunk blockIdx.x + k*gridDim.x -- at any moment the whole grid works on one
  // window of ~gridDim.x consecutive chunks, so the J columns a chunk shares with its neighbours (and with the chunks
  // one torus row further, for stencil-like problems) are re-read from L2, not from DRAM.  contiguous: a private range
  // per CTA (the grid's working set is then spread over the whole J array and does not fit the 126 MB L2 at C2 size).
  const int c0 = strided ? (int)blockIdx.x : (int)(((int64_t)nchunks * blockIdx.x) / gridDim.x);
  const int cstep = strided ? (int)gridDim.x : 1;
  const int nk = strided ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x
                         : (int)(((int64_t)nchunks * (blockIdx.x + 1)) / gridDim.x) - c0;
  if (nk <= 0) return;
  constexpr int SPT = kSlotCap / 256;  // slots per thread
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  // descriptors of the first kDescRing chunks of the range (refilled half a ring at a time below)
  for (int t = tid; t < min(nk, kDescRing) * 2; t += 256)
    reinterpret_cast<uint4*>(ring)[t] = reinterpret_cast<const uint4*>(chunks + c0 + (int64_t)(t >> 1) * cstep)[t & 1];
  __syncthreads();

  StageEntry eNext;          // this thread's stage entry of the chunk whose TMA copies are issued next
  SlotMeta mCur[SPT], mNext[SPT], mNew[SPT];
  uint32_t scCur[SPT], scNext[SPT], scNew[SPT];
  eNext.src = 0;
  eNext.off = eNext.len = 0;
  eNext.slice_base = 0;
#pragma unroll
  for (int q = 0; q < SPT; ++q) {
    mCur[q] = mNext[q] = mNew[q] = SlotMeta{0, 0, 0, 0};
    scCur[q] = scNext[q] = scNew[q] = 0;
  }
  // loads of chunk k that depend only on its descriptor: stage entry and slot metadata into registers
  auto load_meta = [&](int k, StageEntry& e, SlotMeta (&m)[SPT], uint32_t (&sc)[SPT]) {
    const ChunkDesc& d = ring[k % kDescRing];
    if (tid < d.nstage) e = stage[d.stage0 + tid];
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      const int t = tid + q * 256;
      if (t < d.nslots) {
        m[q] = slots[d.slot0 + t];
        sc[q] = slot_start[d.slot0 + t];
      }
    }
  };
  auto load_recs = [&](int k) {  // pair records of chunk k -> srec[k&1], asynchronously (chunks are 32-record padded)
    if (k < nk) {
      const ChunkDesc& d = ring[k % kDescRing];
      const int n16 = (d.npairs * (int)sizeof(RecT) + 15) / 16;
      const char* src = reinterpret_cast<const char*>(recs + d.rec0);
      char* dst = reinterpret_cast<char*>(srec[k & 1]);
      for (int t = tid; t < n16; t += 256) cp_async16(dst + t * 16, src + t * 16);
    }
    cp_async_commit();
  };
  auto issue_tma = [&](int k, const StageEntry& e) {  // stage list of chunk k -> shared memory buffer k&1
    const ChunkDesc& d = ring[k % kDescRing];
    const int b = k & 1;
    if (tid == 0) mbar_expect_tx(&bar[b], (uint32_t)d.stage_bytes);
    if (tid < d.nstage) {
      s_offlen[b][tid] = ((unsigned)e.off << 16) | e.len;
      s_sbase[b][tid] = e.slice_base & kSliceBaseMask;
      const int run = (int)((uint64_t)e.slice_base >> 48);  // this column heads a run of consecutive columns: one copy for all
      const int a = e.src & 1;
      if (run > 0) tma_bulk_g2s(sval[b] + ((int)e.off - a), Jv + (e.src - a), (uint32_t)run * 8u, &bar[b]);
    }
  };

  // prologue: chunk 0 fully issued, chunk 1's metadata on its way
  load_meta(0, eNext, mCur, scCur);
  load_recs(0);
  issue_tma(0, eNext);
  if (nk > 1) load_meta(1, eNext, mNext, scNext);
  load_recs(1);

  if (dbg_on) tk = clock64();
  for (int k = 0; k < nk; ++k) {
    const int b = k & 1;
    if (k + 1 < nk) issue_tma(k + 1, eNext);                          // buffer b^1 was released by the barrier closing iteration k-1
    if (k + 2 < nk) load_meta(k + 2, eNext, mNew, scNew);             // consumed in iteration k+1 / k+2
    ASM_TICK(0)
    mbar_wait(&bar[b], (uint32_t)((k >> 1) & 1));
    ASM_TICK(1)
    asm volatile("cp.async.wait_group 1;" ::: "memory");               // records of chunk k (those of k+1 may still be in flight)
    __syncthreads();
    ASM_TICK(2)
    {
      const ChunkDesc& d = ring[k % kDescRing];
      const int ns = d.nslots, np = d.npairs;
#pragma unroll
      for (int q = 0; q < SPT; ++q) {
        const int t = tid + q * 256;
        if (t < ns) {
          const SlotMeta mt = mCur[q];
          const int beg = (int)(scCur[q] >> 16);
          const int end = ns == 1 ? np : beg + (int)(scCur[q] & 0xffffu);
          const int bx = (int)(s_offlen[b][mt.il] >> 16), by = (int)(s_offlen[b][mt.jl] >> 16);
          const int64_t bi = s_sbase[b][mt.il], bj = s_sbase[b][mt.jl];
          const double acc = slot_sum<RecT, POBITS, true>(srec[b], sval[b] + bx, sval[b] + by, beg, end);
          hval[bi + (int64_t)mt.rank * 32] = acc;
          hval[bj + (int64_t)mt.rank_up * 32] = acc;
        }
      }
    }
    ASM_TICK(3)
    __syncthreads();  // buffer b (values, records, stage arrays) is free again
    ASM_TICK(4)
    // descriptor ring: once half a ring has been consumed, refill it with the descriptors half a ring ahead
    if ((k + 1) % (kDescRing / 2) == 0) {
      const int base = k + 1 + kDescRing / 2;  // first chunk of the half being refilled
      for (int t = tid; t < kDescRing; t += 256) {
        const int kk = base + t / 2;
        if (kk < nk) reinterpret_cast<uint4*>(ring)[(kk % kDescRing) * 2 + (t & 1)] = reinterpret_cast<const uint4*>(chunks + c0 + (int64_t)kk * cstep)[t & 1];
      }
      __syncthreads();
    }
    load_recs(k + 2);
#pragma unroll
    for (int q = 0; q < SPT; ++q) {
      mCur[q] = mNext[q];
      scCur[q] = scNext[q];
      mNext[q] = mNew[q];
      scNext[q] = scNew[q];
    }
    ASM_TICK(5)
  }
  if (dbg_on) {
    long long* o = dbg + (threadIdx.x == 0 ? 0 : threadIdx.x == 64 ? 8 : 16);
    for (int i = 0; i < 6; ++i) o[i] = acc_t[i];
    o[6] = nk;
  }
#undef ASM_TICK
}

__device__ __forceinline__ int skew(int i) { return i + (i >> 4); }
constexpr int kColsTileDoubles = kEntTile + kEntTile / 16 + 1;
constexpr size_t kColsSmemBytes = sizeof(double) * 2 * kColsTileDoubles;

__global__ void __launch_bounds__(256, 4) assemble_cols_kernel(
    const int* __restrict__ colchunk, const int* __restrict__ Jcolptr, const int* __restrict__ Jrow,
    const double* __restrict__ Jv, const double* __restrict__ f, double lambda, const int* __restrict__ nup1,
    const int64_t* __restrict__ slice_off, double* __restrict__ hval, double* __restrict__ hdiag,
    double* __restrict__ rhs, double* __restrict__ dvec, unsigned long long* __restrict__ foo_bits, int defer_damping,
    const double* __restrict__ lam_vec, int lam_n_per) {
  extern __shared__ __align__(16) double cols_smem[];  // two staged tiles (values, gathered residuals), bank-skewed
  double* sv = cols_smem;
  double* sf = cols_smem + kColsTileDoubles;
  __shared__ double s_max[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int j0 = colchunk[c], j1 = colchunk[c + 1];
  const int e0 = Jcolptr[j0], e1 = Jcolptr[j1];
  constexpr int CPT = kColCap / 256;  // columns per thread
  double a_vf[CPT], a_d[CPT], a_vv[CPT], lam[CPT];
  bool started[CPT];
  int cs[CPT], ce[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int j = j0 + tid + k * 256;
    // batched independent problems (block-diagonal Jacobian): every problem has its own lambda
    lam[k] = (lam_vec && j < j1) ? lam_vec[j / lam_n_per] : lambda;
    a_vf[k] = 0.0;
    a_d[k] = 0.0;
    a_vv[k] = 0.0;
    started[k] = false;
    cs[k] = j < j1 ? Jcolptr[j] : 0;
    ce[k] = j < j1 ? Jcolptr[j + 1] : 0;
  }
  for (int a = e0; a < e1 || a == e0; a += kEntTile) {
    const int b = min(e1, a + kEntTile);
    {
      // all of a thread's row indices first, then all residual gathers, then the stores: written as one loop the
      // dependent chain Jrow -> f (a random access into an L2-resident vector) is paid once per element, eight
      // times in a row
      constexpr int EPT = kEntTile / 256;
      double vv[EPT], ff[EPT];
      int rr[EPT];
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int e = a + tid + u * 256;
        const bool ok = e < b;
        vv[u] = ok ? Jv[e] : 0.0;
        rr[u] = (ok && f) ? Jrow[e] : -1;
      }
#pragma unroll
      for (int u = 0; u < EPT; ++u) ff[u] = rr[u] >= 0 ? __ldg(f + rr[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < EPT; ++u) {
        const int e = a + tid + u * 256;
        if (e < b) {
          sv[skew(e - a)] = vv[u];
          sf[skew(e - a)] = ff[u];
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int lo = max(cs[k], a), hi = min(ce[k], b);
      for (int e = lo; e < hi; ++e) {
        const double v = sv[skew(e - a)], fr = sf[skew(e - a)];
        a_vf[k] = __dadd_rn(a_vf[k], __dmul_rn(v, fr));              // tmp += v*f   (LMSolver.h:117)
        a_d[k] = __dadd_rn(a_d[k], __dmul_rn(__dmul_rn(lam[k], v), v));  // (lambda*v)*v (DiagonalDamping.h:35)
        const double vv = __dmul_rn(v, v);
        a_vv[k] = started[k] ? __dadd_rn(a_vv[k], vv) : vv;          // first product assigned, then added
        started[k] = true;
      }
    }
    __syncthreads();
    if (e1 == e0) break;
  }
  double mymax = 0.0;
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int j = j0 + tid + k * 256;
    if (j < j1) {
      const double sd = sqrt(a_d[k]);                    // sqrt(dampVector(i))  (DiagonalDamping.h:40)
      const double dd = __dmul_rn(sd, sd);               // enters H last, through the augmented row
      // defer_damping (row-partitioned mode): this rank only has a partial d; the damping term
      // is added after the partial sums have been combined (apply_damping_kernel)
      const double hjj = defer_damping ? (started[k] ? a_vv[k] : 0.0) : (started[k] ? __dadd_rn(a_vv[k], dd) : dd);
      const int64_t pd = slice_off[j >> 5] + (int64_t)(nup1[j] - 1) * 32 + (j & 31);
      hval[pd] = hjj;
      hdiag[j] = hjj;
      dvec[j] = a_d[k];
      if (f) {
        const double g = -a_vf[k];
        rhs[j] = g;
        const double ag = fabs(g);
        if (ag > mymax) mymax = ag;  // NaN never wins, like a max-reduction
      }
    }
  }
  if (f && !defer_damping) {
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = mymax;
    __syncthreads();
    if (tid == 0) {
      double mx = s_max[0];
      for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
      atomicMax(foo_bits, (unsigned long long)__double_as_longlong(mx));
    }
  }
}

// row-partitioned mode, after the allreduce: H_jj += fl(sqrt(d_j))^2 (damping last), Jacobi
// diagonal, ||rhs||_inf
__global__ void __launch_bounds__(256) apply_damping_kernel(const double* __restrict__ dvec, const double* __restrict__ rhs,
                                                            const int* __restrict__ nup1, const int64_t* __restrict__ slice_off, int n,
                                                            double* __restrict__ hval, double* __restrict__ hdiag,
                                                            unsigned long long* __restrict__ foo_bits) {
  __shared__ double s_max[8];
  const int j = blockIdx.x * 256 + threadIdx.x;
  double mymax = 0.0;
  if (j < n) {
    const double sd = sqrt(dvec[j]);
    const int64_t pd = slice_off[j >> 5] + (int64_t)(nup1[j] - 1) * 32 + (j & 31);
    const double hjj = __dadd_rn(hval[pd], __dmul_rn(sd, sd));
    hval[pd] = hjj;
    hdiag[j] = hjj;
    if (rhs) {
      const double ag = fabs(rhs[j]);
      if (ag > mymax) mymax = ag;
    }
  }
  if (rhs) {
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
      double mx = s_max[0];
      for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
      atomicMax(foo_bits, (unsigned long long)__double_as_longlong(mx));
    }
  }
}

// subdomain (halo) mode, after the halo sums of rhs, d and diag(H): on the OWNED unknowns the damping
// fl(sqrt(d_j))^2 enters this rank's local H_jj (the neighbours' local copies of the entry stay
// undamped: H = sum of the local matrices), the Jacobi diagonal becomes the full H_jj, and ||rhs||_inf
// is taken over the owned entries (combined with an allreduce-max afterwards)
__global__ void __launch_bounds__(256) halo_damping_kernel(const double* __restrict__ dvec, const double* __restrict__ rhs,
                                                           const int* __restrict__ nup1, const int64_t* __restrict__ slice_off, int own_lo,
                                                           int own_hi, double* __restrict__ hval, double* __restrict__ hdiag,
                                                           unsigned long long* __restrict__ foo_bits) {
  __shared__ double s_max[8];
  const int j = own_lo + blockIdx.x * 256 + threadIdx.x;
  double mymax = 0.0;
  if (j < own_hi) {
    const double sd = sqrt(dvec[j]);
    const double dd = __dmul_rn(sd, sd);
    const int64_t pd = slice_off[j >> 5] + (int64_t)(nup1[j] - 1) * 32 + (j & 31);
    hval[pd] = __dadd_rn(hval[pd], dd);
    hdiag[j] = __dadd_rn(hdiag[j], dd);
    if (rhs) {
      const double ag = fabs(rhs[j]);
      if (ag > mymax) mymax = ag;
    }
  }
  if (rhs) {
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
      double mx = s_max[0];
      for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
      atomicMax(foo_bits, (unsigned long long)__double_as_longlong(mx));
    }
  }
}

__global__ void gather_vals_kernel(const double* __restrict__ hval, const int64_t* __restrict__ pos, int64_t nnz,
                                   double* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) out[t] = hval[pos[t]];
}
__global__ void scatter_vals_kernel(const double* __restrict__ in, const int64_t* __restrict__ pos, int64_t nnz,
                                    double* __restrict__ hval) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nnz) hval[pos[t]] = in[t];
}
__global__ void extract_diag_kernel(const double* __restrict__ hval, const int* __restrict__ nup1,
                                    const int64_t* __restrict__ slice_off, int n, double* __restrict__ hdiag) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) hdiag[j] = hval[slice_off[j >> 5] + (int64_t)(nup1[j] - 1) * 32 + (j & 31)];
}

// ---------------------------------------------------------------------------------------
void upload_h_layout(spb_context* h) {
  cudaStream_t st = h->stream;
  SPB_CUDA(h->d_sell_col.upload(h->S.col, st));
  SPB_CUDA(h->d_sell_col16.upload(h->S.col16, st));
  SPB_CUDA(h->d_slice16.upload(h->S.slice16, st));
  SPB_CUDA(h->d_rowlen.upload(h->S.rowlen, st));
  SPB_CUDA(h->d_slice_off.upload(h->S.slice_off, st));
  SPB_CUDA(h->d_nup1.upload(h->H.nup1, st));
  SPB_CUDA(h->d_hval.alloc((size_t)h->S.padded));
  SPB_CUDA(cudaMemsetAsync(h->d_hval.p, 0, (size_t)h->S.padded * sizeof(double), st));
  SPB_CUDA(h->d_hdiag.alloc(h->n));
  SPB_CUDA(h->d_rhs.alloc(h->n));
  SPB_CUDA(h->d_dvec.alloc(h->n));
  // an assembly without residuals (no EVec, or zero residual rows) leaves the gradient untouched: it must read as zero
  if (h->n > 0) SPB_CUDA(cudaMemsetAsync(h->d_rhs.p, 0, (size_t)h->n * sizeof(double), st));
  SPB_CUDA(cudaStreamSynchronize(st));
  std::vector<int>().swap(h->S.col);  // the big host copies are no longer needed
  std::vector<uint16_t>().swap(h->S.col16);
  h->d_pos_of_slot.release();
}

void upload_analysis(spb_context* h, const int* colptr, const int* rowidx, AssemblyPlan& P) {
  cudaStream_t st = h->stream;
  // the "stream" form of the pair assembly (default when it applies; see assembly_v2.cu) is packed on its own host thread
  // while this one packs and uploads the general form: the plan is only read until both are done
  PairsV2Host V2;
  std::exception_ptr v2_err;
  std::thread v2_thread([&] {
    try {
      pack_pairs_v2(colptr, P, h->S, V2);
    } catch (...) {
      v2_err = std::current_exception();
    }
  });
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if (t.joinable()) t.join();
    }
  } v2_join{v2_thread};
  std::vector<int> cp(colptr, colptr + h->n + 1), ri(rowidx, rowidx + colptr[h->n]);
  SPB_CUDA(h->d_Jcolptr.upload(cp, st));
  SPB_CUDA(h->d_Jrow.upload(ri, st));
  h->rec_bits = P.rec_bits;
  if (P.rec_bits == 16) SPB_CUDA(h->d_rec16.upload(P.rec16, st)); else SPB_CUDA(h->d_rec32.upload(P.rec32, st));
  {
    // packed device forms: one descriptor per chunk, one entry per staged column, one record per slot
    const int nch = (int)P.chunk_npairs.size();
    std::vector<ChunkDesc> cd((size_t)nch);
    bool fast = true;  // every chunk on the common path: the pipelined kernel applies
    for (int c = 0; c < nch; ++c) {
      cd[c].slot0 = P.chunk_slot[c];
      cd[c].nslots = P.chunk_slot[(size_t)c + 1] - P.chunk_slot[c];
      cd[c].npairs = P.chunk_npairs[c];
      cd[c].stage0 = P.chunk_stage_ptr[c];
      cd[c].nstage = (uint16_t)std::min(P.chunk_stage_ptr[(size_t)c + 1] - P.chunk_stage_ptr[c], 65535);
      cd[c].flags = P.chunk_staged[c];
      cd[c].stage_bytes = 0;
      if (cd[c].flags) {
        bool inside = true;
        int bytes = 0;
        for (int t = P.chunk_stage_ptr[c]; t < P.chunk_stage_ptr[(size_t)c + 1]; ++t) {
          const int col = P.stage_cols[t];
          const int start = colptr[col], len = colptr[col + 1] - start, a = start & 1;
          if (len > 32) cd[c].flags |= 2;
          const int run = P.stage_run[t];  // doubles of the bulk copy headed by this column (0: inside a run)
          if (run > 0) {
            if ((int64_t)start - a + run > (int64_t)colptr[h->n]) inside = false;  // window would run past the value array
            bytes += run * 8;
          }
        }
        if (inside) cd[c].flags |= 4;
        cd[c].stage_bytes = bytes;
      }
      cd[c].rec0 = P.chunk_rec[c];
      if (!((cd[c].flags & 1) && (cd[c].flags & 4) && cd[c].npairs <= kPairTile && cd[c].nstage <= kStageCols)) fast = false;
    }
    h->pairs_fast_path = fast;
    // the per-entry packing loops run on the host threads (hundreds of millions of slots at C5 size)
    std::vector<StageEntry> se(P.stage_cols.size());
    std::vector<SlotMeta> sm((size_t)P.nslots);
    std::vector<uint32_t> sc((size_t)P.nslots);
    {
      const int parts = std::max(1, std::min(host_threads(), P.nslots / 262144));
      const size_t nse = se.size();
      run_parts(parts, [&](int p) {
        for (size_t t = nse * p / parts; t < nse * (p + 1) / parts; ++t) {
          const int col = P.stage_cols[t];
          se[t].src = colptr[col];
          se[t].off = (uint16_t)std::min(P.stage_off[t], 65535);
          se[t].len = (uint16_t)std::min(colptr[col + 1] - colptr[col], 65535);
          se[t].slice_base = (h->S.slice_off[col >> 5] + (col & 31)) | ((int64_t)P.stage_run[t] << 48);
        }
        const int64_t ns64 = P.nslots;
        for (int64_t pos = ns64 * p / parts; pos < ns64 * (p + 1) / parts; ++pos) {  // processing order (see symbolic.cpp)
          const int s = P.proc_perm[(size_t)pos];
          sm[(size_t)pos] = SlotMeta{P.slot_il[s], P.slot_jl[s], P.low_rank[s], P.low_rank_up[s]};
          sc[(size_t)pos] = ((uint32_t)P.slot_start[s] << 16) | P.slot_cnt[s];
        }
      });
    }
    SPB_CUDA(h->d_chunk_desc.upload(cd, st));
    SPB_CUDA(h->d_stage_ent.upload(se, st));
    SPB_CUDA(h->d_slot_meta.upload(sm, st));
    SPB_CUDA(h->d_slot_start.upload(sc, st));
    SPB_CUDA(h->d_colchunk.upload(P.colchunk, st));
    SPB_CUDA(cudaStreamSynchronize(st));
  }
  h->nchunks = (int)P.chunk_npairs.size();
  h->ncolchunks = (int)P.colchunk.size() - 1;
  h->occ_pairs = 0;  // the record width may have changed: the pipelined kernel's instantiation (and its shared-memory opt-in) follows it
  {
    // for the streamed host path: how much of the J value array a chunk needs (its stage list's last column, plus
    // the padding element of a bulk-copy window).  Chunks are launched in order of that need (a periodic problem's
    // first chunks reference the last columns), so a second copy of the descriptors sorted by it is kept.
    const int nch = h->nchunks;
    std::vector<int64_t> need((size_t)nch, 0);
    for (int c = 0; c < nch; ++c) {
      int64_t nd = 0;
      for (int t = P.chunk_stage_ptr[c]; t < P.chunk_stage_ptr[(size_t)c + 1]; ++t) nd = std::max<int64_t>(nd, (int64_t)colptr[P.stage_cols[t] + 1] + 1);
      need[c] = std::min<int64_t>(nd, colptr[h->n]);
    }
    std::vector<int> order((size_t)nch);
    for (int c = 0; c < nch; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return need[a] < need[b]; });
    std::vector<ChunkDesc> sorted((size_t)nch);
    std::vector<ChunkDesc> host_cd((size_t)nch);
    SPB_CUDA(cudaMemcpy(host_cd.data(), h->d_chunk_desc.p, sizeof(ChunkDesc) * (size_t)nch, cudaMemcpyDeviceToHost));
    h->chunk_need.assign((size_t)nch, 0);
    h->keep.sorted_pos.assign((size_t)nch, 0);
    for (int i = 0; i < nch; ++i) {
      sorted[i] = host_cd[order[i]];
      h->chunk_need[i] = need[order[i]];
      h->keep.sorted_pos[(size_t)order[i]] = i;
    }
    SPB_CUDA(h->d_chunk_desc_sorted.upload(sorted, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    h->colchunk_host = P.colchunk;
    h->Jcolptr_host.assign(colptr, colptr + h->n + 1);
  }
  v2_thread.join();
  if (v2_err) std::rethrow_exception(v2_err);
  upload_pairs_v2(h, V2);
  h->nslots = P.nslots;
  h->npairs = P.npairs;
  // host-side pieces kept for incremental re-analysis (spb_analyze_update)
  h->keep.valid = false;
  h->d_keep_stage_cols.release();
  h->d_keep_stage_ptr.release();
  h->d_keep_sorted_pos.release();
  if ((int64_t)colptr[h->n] <= kPlanKeepMaxNnz && !h->partitioned) {
    PlanKeep& K = h->keep;
    K.Jrow.assign(rowidx, rowidx + colptr[h->n]);
    K.stage_cols.swap(P.stage_cols);
    K.chunk_stage_ptr.swap(P.chunk_stage_ptr);
    K.chunk_slot.swap(P.chunk_slot);
    K.chunk_npairs.swap(P.chunk_npairs);
    K.chunk_rec.swap(P.chunk_rec);
    K.chunk_staged.swap(P.chunk_staged);
    K.low_colptr.swap(P.low_colptr);
    K.low_row.swap(P.low_row);
    K.slot_start.swap(P.slot_start);
    K.slot_cnt.swap(P.slot_cnt);
    K.valid = true;
  }
  SPB_CUDA(h->d_foo_bits.alloc(1));
  SPB_CUDA(cudaStreamSynchronize(st));
}

// column kernel on stream `cst`
static void cols_smem_optin() {
  static const bool done = [] {
    return cudaFuncSetAttribute(assemble_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColsSmemBytes) == cudaSuccess;
  }();
  if (!done) throw StateFail{"shared-memory opt-in of the column kernel failed"};
}
static void launch_cols(spb_context* h, cudaStream_t cst, const double* d_Jv, const double* d_f, double lambda, double* hv, bool part) {
  cols_smem_optin();
  double* rhs = part ? h->d_rhs_part.p : h->d_rhs.p;
  double* dv = part ? h->d_dvec_part.p : h->d_dvec.p;
  const int defer = (part || h->halo_mode) ? 1 : 0;
  {
    assemble_cols_kernel<<<h->ncolchunks, 256, kColsSmemBytes, cst>>>(h->d_colchunk.p, h->d_Jcolptr.p, h->d_Jrow.p, d_Jv, d_f, lambda, h->d_nup1.p,
                                                         h->d_slice_off.p, hv, h->d_hdiag.p, rhs, dv, h->d_foo_bits.p, defer, h->d_lambda_vec,
                                                         h->lambda_n_per);
  }
}

void run_assembly(spb_context* h, const double* d_Jv, const double* d_f, double lambda, double* foo_host) {
  cudaStream_t st = h->stream;
  stage_begin(h, ST_ASSEMBLY);
  int nl = 0;
  if (d_f) SPB_CUDA(cudaMemsetAsync(h->d_foo_bits.p, 0, sizeof(unsigned long long), st));
  const bool part = h->partitioned;
  double* hv = part ? h->d_hval_part.p : h->d_hval.p;
  // With the stream-form pair kernel (one CTA per SM, shared-memory-pipe bound, HBM half idle) the column kernel
  // (latency bound, writes disjoint outputs: diagonal of H, rhs, d) runs NEXT to it on a side stream: its CTAs fit
  // into the shared memory and registers the pair kernel leaves free.
  static const bool overlap_on = !(getenv("SPB_ASM_OVERLAP") && atoi(getenv("SPB_ASM_OVERLAP")) == 0);
  const bool side = overlap_on && h->ncolchunks > 0 && h->nchunks > 0 && h->v2_ok && ((((uintptr_t)d_Jv) & 15) == 0);
  cudaStream_t cst = st;
  if (side) {
    if (!h->asm_side_stream) {
      SPB_CUDA(cudaStreamCreateWithFlags(&h->asm_side_stream, cudaStreamNonBlocking));
      SPB_CUDA(cudaEventCreateWithFlags(&h->asm_fork_ev, cudaEventDisableTiming));
      SPB_CUDA(cudaEventCreateWithFlags(&h->asm_join_ev, cudaEventDisableTiming));
    }
    cst = h->asm_side_stream;
    SPB_CUDA(cudaEventRecord(h->asm_fork_ev, st));
    SPB_CUDA(cudaStreamWaitEvent(cst, h->asm_fork_ev, 0));
  }
  if (h->ncolchunks > 0 && !side) {
    launch_cols(h, st, d_Jv, d_f, lambda, hv, part);
    ++nl;
  }
  if (h->nchunks > 0) {
    const int bulk_ok = (((uintptr_t)d_Jv) & 15) == 0 ? 1 : 0;  // TMA bulk copies need a 16-byte aligned source
    static const bool pipe_off = getenv("SPB_ASM_PIPELINE") && atoi(getenv("SPB_ASM_PIPELINE")) == 0;
    h->last_asm_form = 1;
    if (run_pairs_v2(h, d_Jv, hv)) {
      h->last_asm_form = 3;
      // the stream form ran; the column kernel goes second so that the pair kernel's one-CTA-per-SM grid is resident first
      if (side) {
        launch_cols(h, cst, d_Jv, d_f, lambda, hv, part);
        ++nl;
        SPB_CUDA(cudaEventRecord(h->asm_join_ev, cst));
        SPB_CUDA(cudaStreamWaitEvent(st, h->asm_join_ev, 0));
      }
    } else if (bulk_ok && h->pairs_fast_path && !pipe_off) {
      h->last_asm_form = 2;
      if (h->occ_pairs == 0) {
        int occ = 0;
        if (h->rec_bits == 16) {
          SPB_CUDA(cudaFuncSetAttribute(assemble_pairs_pipelined_kernel<uint16_t, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipelined_smem_bytes<uint16_t>()));
          SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assemble_pairs_pipelined_kernel<uint16_t, 7>, 256, pipelined_smem_bytes<uint16_t>()));
        } else {
          SPB_CUDA(cudaFuncSetAttribute(assemble_pairs_pipelined_kernel<uint32_t, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipelined_smem_bytes<uint32_t>()));
          SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assemble_pairs_pipelined_kernel<uint32_t, 15>, 256, pipelined_smem_bytes<uint32_t>()));
        }
        h->occ_pairs = std::max(1, occ);
      }
      const int G = std::max(1, std::min(h->nchunks, h->num_sms * h->occ_pairs));
      static const bool asm_dbg = getenv("SPB_ASM_TIMING") && atoi(getenv("SPB_ASM_TIMING")) > 0;
      long long* dbgp = nullptr;
      if (asm_dbg) {
        SPB_CUDA(h->d_dbg.alloc(24));
        dbgp = h->d_dbg.p;
      }
      static const int strided = getenv("SPB_ASM_STRIDED") ? atoi(getenv("SPB_ASM_STRIDED")) : 1;
      if (h->rec_bits == 16)
        assemble_pairs_pipelined_kernel<uint16_t, 7><<<G, 256, pipelined_smem_bytes<uint16_t>(), st>>>(h->d_rec16.p, h->d_chunk_desc.p, h->nchunks, h->d_stage_ent.p,
                                                                        h->d_slot_meta.p, h->d_slot_start.p, d_Jv, hv, dbgp, strided);
      else
        assemble_pairs_pipelined_kernel<uint32_t, 15><<<G, 256, pipelined_smem_bytes<uint32_t>(), st>>>(h->d_rec32.p, h->d_chunk_desc.p, h->nchunks, h->d_stage_ent.p,
                                                                         h->d_slot_meta.p, h->d_slot_start.p, d_Jv, hv, dbgp, strided);
      if (dbgp) {
        long long tt[24];
        cudaStreamSynchronize(st);
        cudaMemcpy(tt, dbgp, sizeof(tt), cudaMemcpyDeviceToHost);
        for (int w = 0; w < 3; ++w) {
          const long long* t = tt + 8 * w;
          fprintf(stderr, "[spb asm timing] thread %d cycles/chunk over %lld chunks: issue+meta %lld mbar_wait %lld recs+sync %lld sums %lld sync %lld tail %lld\n",
                  w == 0 ? 0 : w == 1 ? 64 : 224, t[6], t[0] / t[6], t[1] / t[6], t[2] / t[6], t[3] / t[6], t[4] / t[6], t[5] / t[6]);
        }
      }
    } else if (h->rec_bits == 16)
      assemble_pairs_kernel<uint16_t, 7><<<h->nchunks, 256, 0, st>>>(h->d_rec16.p, h->d_chunk_desc.p, h->d_stage_ent.p, h->d_slot_meta.p,
                                                                     h->d_slot_start.p, d_Jv, hv, bulk_ok);
    else
      assemble_pairs_kernel<uint32_t, 15><<<h->nchunks, 256, 0, st>>>(h->d_rec32.p, h->d_chunk_desc.p, h->d_stage_ent.p, h->d_slot_meta.p,
                                                                      h->d_slot_start.p, d_Jv, hv, bulk_ok);
    ++nl;
  }
  if (part) {
    // combine the ranks' partial normal equations: one grouped allreduce over NVLink
    dist_group_begin(h);
    dist_allreduce_sum(h, h->d_hval_part.p, h->d_hval.p, (size_t)h->S.padded);
    dist_allreduce_sum(h, h->d_dvec_part.p, h->d_dvec.p, (size_t)h->n);
    if (d_f) dist_allreduce_sum(h, h->d_rhs_part.p, h->d_rhs.p, (size_t)h->n);
    dist_group_end(h);
    apply_damping_kernel<<<(h->n + 255) / 256, 256, 0, st>>>(h->d_dvec.p, d_f ? h->d_rhs.p : nullptr, h->d_nup1.p, h->d_slice_off.p, h->n,
                                                             h->d_hval.p, h->d_hdiag.p, h->d_foo_bits.p);
    ++nl;
  }
  if (h->halo_mode && !part) {
    if (h->n != h->own_hi + h->halo) throw StateFail{"halo description does not match the analysed local problem"};
    // sums over the ranks that share an unknown: gradient, damping vector and diagonal travel as halo slabs
    if (d_f) halo_reverse_add(h, h->d_rhs.p);
    halo_reverse_add(h, h->d_dvec.p);
    halo_reverse_add(h, h->d_hdiag.p);
    const int own = h->own_hi - h->own_lo;
    halo_damping_kernel<<<(own + 255) / 256, 256, 0, st>>>(h->d_dvec.p, d_f ? h->d_rhs.p : nullptr, h->d_nup1.p, h->d_slice_off.p, h->own_lo,
                                                         h->own_hi, h->d_hval.p, h->d_hdiag.p, h->d_foo_bits.p);
    ++nl;
    if (d_f) {
      dist_allreduce_max(h, (const double*)h->d_foo_bits.p, (double*)h->d_foo_bits.p, 1);  // non-negative doubles: bits order like values
      halo_forward(h, h->d_rhs.p);
    }
    halo_forward(h, h->d_dvec.p);
    halo_forward(h, h->d_hdiag.p);  // the PCG updates the ghost copies of its vectors itself and needs their Jacobi diagonal
  }
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_ASSEMBLY, nl);
  if (foo_host) {
    if (d_f) {
      SPB_CUDA(cudaMemcpyAsync(h->h_scalar, h->d_foo_bits.p, sizeof(double), cudaMemcpyDeviceToHost, st));
      SPB_CUDA(cudaStreamSynchronize(st));
      *foo_host = h->h_scalar[0];
    } else {
      *foo_host = 0.0;
    }
  }
  h->h_valid = true;
  h->factor_valid = false;
}

// spb_assemble from PINNED host memory: the J values are uploaded in parts on a copy stream and every part of
// the assembly (column kernel chunks, pair kernel chunks) starts as soon as the values it reads have arrived,
// so the 0.4 ms of assembly disappear behind the 2.4 ms PCIe transfer (C2) instead of following it.
bool run_assembly_streamed(spb_context* h, const double* host_Jv, const double* host_f, double lambda, double* foo_host) {
  constexpr int kParts = 8;
  static const bool off = getenv("SPB_ASM_STREAMED") && atoi(getenv("SPB_ASM_STREAMED")) == 0;
  if (off || !h->analyzed || h->partitioned || h->halo_mode || h->nchunks == 0 || !h->pairs_fast_path || h->nnzJ < (1 << 20)) return false;
  {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host_Jv) != cudaSuccess || at.type != cudaMemoryTypeHost) {
      cudaGetLastError();
      return false;  // pageable memory: the copy would not be asynchronous anyway
    }
  }
  const int64_t nnzJ = h->nnzJ;
  SPB_CUDA(h->d_Jv_stage.alloc((size_t)nnzJ));
  if (host_f) SPB_CUDA(h->d_f_stage.alloc((size_t)h->m));
  if ((((uintptr_t)h->d_Jv_stage.p) & 15) != 0) return false;
  if (!h->copy_stream) {
    SPB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (auto& e : h->copy_ev) SPB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    SPB_CUDA(cudaEventCreateWithFlags(&h->asm_done_ev, cudaEventDisableTiming));
  }
  cudaStream_t st = h->stream, cs = h->copy_stream;
  double* dJ = h->d_Jv_stage.p;
  // the staging buffers are still being read by the previous call's kernels
  if (h->asm_done_recorded) SPB_CUDA(cudaStreamWaitEvent(cs, h->asm_done_ev, 0));
  if (host_f) SPB_CUDA(cudaMemcpyAsync(h->d_f_stage.p, host_f, (size_t)h->m * sizeof(double), cudaMemcpyHostToDevice, cs));
  SPB_CUDA(cudaEventRecord(h->copy_ev[kParts], cs));
  // part boundaries: whole J columns, about nnzJ/kParts values each
  int colb[kParts + 1];
  colb[0] = 0;
  for (int k = 1; k < kParts; ++k) {
    const int64_t target = nnzJ * k / kParts;
    colb[k] = (int)(std::lower_bound(h->Jcolptr_host.begin(), h->Jcolptr_host.end(), (int)target) - h->Jcolptr_host.begin());
    colb[k] = std::max(colb[k], colb[k - 1]);
    colb[k] = std::min(colb[k], h->n);
  }
  colb[kParts] = h->n;
  for (int k = 0; k < kParts; ++k) {
    const int64_t v0 = h->Jcolptr_host[colb[k]], v1 = h->Jcolptr_host[colb[k + 1]];
    if (v1 > v0) SPB_CUDA(cudaMemcpyAsync(dJ + v0, host_Jv + v0, (size_t)(v1 - v0) * sizeof(double), cudaMemcpyHostToDevice, cs));
    SPB_CUDA(cudaEventRecord(h->copy_ev[k], cs));
  }
  stage_begin(h, ST_ASSEMBLY);
  int nl = 0;
  const double* d_f = host_f ? h->d_f_stage.p : nullptr;
  if (d_f) SPB_CUDA(cudaMemsetAsync(h->d_foo_bits.p, 0, sizeof(unsigned long long), st));
  SPB_CUDA(cudaStreamWaitEvent(st, h->copy_ev[kParts], 0));
  if (h->occ_pairs == 0) {
    int occ = 0;
    if (h->rec_bits == 16) {
      SPB_CUDA(cudaFuncSetAttribute(assemble_pairs_pipelined_kernel<uint16_t, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipelined_smem_bytes<uint16_t>()));
      SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assemble_pairs_pipelined_kernel<uint16_t, 7>, 256, pipelined_smem_bytes<uint16_t>()));
    } else {
      SPB_CUDA(cudaFuncSetAttribute(assemble_pairs_pipelined_kernel<uint32_t, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pipelined_smem_bytes<uint32_t>()));
      SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, assemble_pairs_pipelined_kernel<uint32_t, 15>, 256, pipelined_smem_bytes<uint32_t>()));
    }
    h->occ_pairs = std::max(1, occ);
  }
  h->last_asm_form = 2;
  cols_smem_optin();
  int cc_done = 0, ch_done = 0;  // column-kernel chunks / pair-kernel chunks already launched
  for (int k = 0; k < kParts; ++k) {
    SPB_CUDA(cudaStreamWaitEvent(st, h->copy_ev[k], 0));
    const int64_t have = h->Jcolptr_host[colb[k + 1]];  // J values [0, have) are on the device after this part
    // column kernel: chunks whose columns are complete
    int cc_end = cc_done;
    while (cc_end < h->ncolchunks && h->colchunk_host[(size_t)cc_end + 1] <= colb[k + 1]) ++cc_end;
    if (k == kParts - 1) cc_end = h->ncolchunks;
    if (cc_end > cc_done) {
      assemble_cols_kernel<<<cc_end - cc_done, 256, kColsSmemBytes, st>>>(h->d_colchunk.p + cc_done, h->d_Jcolptr.p, h->d_Jrow.p, dJ, d_f, lambda, h->d_nup1.p,
                                                            h->d_slice_off.p, h->d_hval.p, h->d_hdiag.p, h->d_rhs.p, h->d_dvec.p, h->d_foo_bits.p, 0, nullptr, 1);
      ++nl;
      cc_done = cc_end;
    }
    // pair kernel: chunks whose stage lists are complete
    int ch_end = (int)(std::upper_bound(h->chunk_need.begin() + ch_done, h->chunk_need.end(), have) - h->chunk_need.begin());
    if (k == kParts - 1) ch_end = h->nchunks;
    if (ch_end > ch_done) {
      const int cnt = ch_end - ch_done;
      const int G = std::max(1, std::min(cnt, h->num_sms * h->occ_pairs));
      if (h->rec_bits == 16)
        assemble_pairs_pipelined_kernel<uint16_t, 7><<<G, 256, pipelined_smem_bytes<uint16_t>(), st>>>(h->d_rec16.p, h->d_chunk_desc_sorted.p + ch_done, cnt, h->d_stage_ent.p,
                                                                        h->d_slot_meta.p, h->d_slot_start.p, dJ, h->d_hval.p, nullptr, 1);
      else
        assemble_pairs_pipelined_kernel<uint32_t, 15><<<G, 256, pipelined_smem_bytes<uint32_t>(), st>>>(h->d_rec32.p, h->d_chunk_desc_sorted.p + ch_done, cnt, h->d_stage_ent.p,
                                                                         h->d_slot_meta.p, h->d_slot_start.p, dJ, h->d_hval.p, nullptr, 1);
      ++nl;
      ch_done = ch_end;
    }
  }
  SPB_CUDA(cudaEventRecord(h->asm_done_ev, st));
  h->asm_done_recorded = true;
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_ASSEMBLY, nl);
  if (foo_host) {
    if (d_f) {
      SPB_CUDA(cudaMemcpyAsync(h->h_scalar, h->d_foo_bits.p, sizeof(double), cudaMemcpyDeviceToHost, st));
      SPB_CUDA(cudaStreamSynchronize(st));
      *foo_host = h->h_scalar[0];
    } else {
      *foo_host = 0.0;
    }
  }
  h->h_valid = true;
  h->factor_valid = false;
  return true;
}

void ensure_pos_of_slot(spb_context* h) {
  if (h->d_pos_of_slot.p) return;
  const HPattern& H = h->H;
  std::vector<int64_t> pos((size_t)H.nnz());
  for (int r = 0; r < H.n; ++r) {
    int64_t base = h->S.slice_off[r >> 5] + (r & 31);
    int k = 0;
    for (int q = H.colptr[r]; q < H.colptr[(size_t)r + 1]; ++q, ++k) pos[q] = base + (int64_t)k * 32;
  }
  SPB_CUDA(h->d_pos_of_slot.upload(pos, h->stream));
  SPB_CUDA(cudaStreamSynchronize(h->stream));
}

void gather_h_values(spb_context* h, double* d_out) {
  ensure_pos_of_slot(h);
  int64_t nnz = h->H.nnz();
  if (nnz == 0) return;
  gather_vals_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(h->d_hval.p, h->d_pos_of_slot.p, nnz, d_out);
  h->launches++;
  SPB_CUDA(cudaGetLastError());
}

// Legacy triplet interface: every H entry (CSC order) is the sum of its triplets in insertion order,
// the way Eigen's setFromTriplets accumulates duplicates [Eigen-upstream]; one thread per entry.
__global__ void __launch_bounds__(256) sum_triplets_kernel(const int64_t* __restrict__ ptr, const int* __restrict__ src,
                                                           const double* __restrict__ vals, int64_t nnz, double* __restrict__ out) {
  const int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (q >= nnz) return;
  const int64_t a = ptr[q], b = ptr[q + 1];
  double s = a < b ? vals[src[a]] : 0.0;
  for (int64_t t = a + 1; t < b; ++t) s = __dadd_rn(s, vals[src[t]]);
  out[q] = s;
}

void sum_triplets(spb_context* h, const double* d_trip_vals, double* d_csc_vals) {
  const int64_t nnz = h->H.nnz();
  if (nnz == 0) return;
  sum_triplets_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(h->d_trip_ptr.p, h->d_trip_src.p, d_trip_vals, nnz, d_csc_vals);
  h->launches++;
  SPB_CUDA(cudaGetLastError());
}

void scatter_h_values(spb_context* h, const double* d_in) {
  ensure_pos_of_slot(h);
  int64_t nnz = h->H.nnz();
  if (nnz > 0) {
    scatter_vals_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, h->stream>>>(d_in, h->d_pos_of_slot.p, nnz, h->d_hval.p);
    extract_diag_kernel<<<(h->n + 255) / 256, 256, 0, h->stream>>>(h->d_hval.p, h->d_nup1.p, h->d_slice_off.p, h->n,
                                                                 h->d_hdiag.p);
    h->launches += 2;
  }
  SPB_CUDA(cudaGetLastError());
  h->h_valid = true;
  h->factor_valid = false;
}

}  // namespace spb
