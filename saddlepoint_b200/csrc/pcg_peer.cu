// pcg_peer.cu -- Jacobi-preconditioned conjugate gradients as ONE persistent cooperative kernel with a SINGLE
// reduction per iteration, on one GPU and -- fused with its collectives over NVLink peer memory -- on the ranks
// of the subdomain (halo) mode of config C5 (sm_100a, NVSwitch).
//
// Recurrences: Chronopoulos & Gear's rearrangement of CG (u = M^-1 r, w = A u, s = A p kept by recurrence):
//     gamma = (r,u), delta = (w,u), rr = (r,r)                      <- the only reduction
//     beta = gamma/gamma_old, alpha = gamma / (delta - beta*gamma/alpha_old)
//     p = u + beta p ; s = w + beta s ; x += alpha p ; u -= alpha dinv s (r = u/dinv is not stored) ; w = A u
// Same iterates as textbook CG in exact arithmetic; on the LM systems of this repo the iteration counts are
// identical and the solutions agree to 4e-16 (tests/test_gpu_solve.py).  An iteration is two phases and two grid
// barriers (pcg_persistent_kernel in pcg.cu: three phases, three barriers, two reductions):
//   A  per-row vector updates (every row is owned by the lane that multiplies it: lane == row of a 32-row slice)
//   B  w = A u with the three dot products accumulated on the fly.
//
// Multi-GPU (PEER = true).  pcg_solve_halo (pcg.cu) runs an iteration as four launches, a grouped
// ncclSend/ncclRecv halo swap and two scalar ncclAllReduce calls: at 8 ranks ~120 us of a 287 us iteration is
// launch + NCCL latency although nothing of size n moves.  Here nothing leaves the kernel:
//   * the rows of the two exchanged bands are multiplied FIRST, spread over all warps of the grid, and written
//     straight into the neighbours' windows (peer-mapped device memory, coalesced stores over NVLink) as 8-byte words
//     that carry 32 payload bits and the exchange tag: no fence, no "slab complete" flag, no ticket.  The slab is in
//     flight while the interior rows are multiplied; afterwards every band row adds the neighbour's view as soon as
//     its two words carry the tag (q = mine + theirs: the owner and the ghost holder end up with the same bits, a + b
//     being commutative);
//   * the reduction is an all-to-all in the same flag-in-word form: block 0 writes this rank's three partial sums into
//     slot [rank] of every rank's window, every block polls its own window and adds the slots in rank order: all
//     blocks of all ranks hold bit-identical alpha / beta / stop decisions;
//   * vector updates run redundantly on the ghost entries from identical inputs, so no vector is ever exchanged.
// Waiting is bounded: a spin that exceeds the timeout raises an abort word in every rank's window; the kernels
// leave at their next grid-uniform check and the solve returns SPB_ERR_STATE.
//
// Reference semantics kept: the solve stands in for EigenSolverWrapper::solve (EigenSolverWrapper.h:72-78) on the
// damped normal matrix of LMSolver.h:136; a non-positive curvature is the iterative analogue of LLT's failed
// pivot (EigenSolverWrapper.h:59), NaN systems return a NaN step.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>

#include "spb_internal.h"
#include "pcg_device.cuh"

namespace cg = cooperative_groups;

namespace spb {

// ---- system-scope loads / stores on peer-visible words ------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

struct PeerView {
  char* win;           // own window
  char* const* peers;  // windows of all ranks (device array)
  PeerLayout L;
  int rank, nranks;
  unsigned long long timeout_ns;
};

__device__ __noinline__ void peer_raise_abort_raw(char* const* peers, int nranks, size_t abort_off) {
  for (int r = 0; r < nranks; ++r) st_relaxed_sys_u32((unsigned int*)(peers[r] + abort_off), 1u);
  __threadfence_system();
}
__device__ __forceinline__ void peer_raise_abort(const PeerView& V) { peer_raise_abort_raw(V.peers, V.nranks, V.L.abort_word); }

// Spin until pred(load(addr)); false on abort / timeout (the caller carries on and reports at the next grid-uniform
// check).  MODE 0: acquire load, value >= want.  MODE 1: relaxed load, high 32 bits == want (flag-in-word).
template <int MODE>
__device__ __noinline__ bool peer_wait_raw(const unsigned long long* addr, unsigned long long want, unsigned long long* got,
                                           const unsigned int* abort_w, unsigned long long timeout_ns, char* const* peers, int nranks,
                                           size_t abort_off) {
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    const unsigned long long v = MODE == 0 ? ld_acquire_sys_u64(addr) : ld_relaxed_sys_u64(addr);
    if (MODE == 0 ? (v >= want) : ((v >> 32) == want)) {
      *got = v;
      return true;
    }
    if ((++spins & 63u) == 0) {
      if (ld_relaxed_sys_u32(abort_w)) return false;
      const unsigned long long now = peer_global_ns();
      if (t0 == 0) {
        t0 = now;
      } else if (now - t0 > timeout_ns) {
        peer_raise_abort_raw(peers, nranks, abort_off);
        return false;
      }
    }
  }
}
__device__ __forceinline__ bool peer_wait_ge(const PeerView& V, const unsigned long long* addr, unsigned long long want) {
  unsigned long long got;
  return peer_wait_raw<0>(addr, want, &got, (const unsigned int*)(V.win + V.L.abort_word), V.timeout_ns, V.peers, V.nranks, V.L.abort_word);
}

// Sum of K <= 3 doubles over all ranks.  Precondition: every block of this rank holds the same local values v[]
// (finalised redundantly from the per-block partials).  `round` counts these calls since the windows were set up and
// is the same number on every rank; slots are double-buffered by its parity (a rank can be at most one round ahead
// of the slowest reader: posting round r+2 needs everybody's round r+1, which a rank posts only after all its blocks
// have read round r).  Result: the sum in rank order -- the same bits in every block of every rank.
template <int K>
__device__ __forceinline__ bool peer_allsum(const PeerView& V, unsigned long long round, double (&v)[K], unsigned int* sh /* shared, nranks*2K */,
                                            double* bc /* shared, K */) {
  const int tid = threadIdx.x;
  const int par = (int)(round & 1ull);
  const unsigned long long tag = (round + 1) & 0xffffffffull;
  const int nw = V.nranks * 2 * K;  // words to deliver / to collect
  if (blockIdx.x == 0 && tid < nw) {  // thread t delivers word (t % 2K) to rank t / 2K
    const int dst = tid / (2 * K), word = tid % (2 * K);
    PeerSlot* s = reinterpret_cast<PeerSlot*>(V.peers[dst] + V.L.slots) + (size_t)par * V.nranks + V.rank;
    double val = v[0];
#pragma unroll
    for (int k = 1; k < K; ++k)
      if ((word >> 1) == k) val = v[k];
    const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
    const unsigned long long half = (word & 1) ? (bits >> 32) : (bits & 0xffffffffull);
    st_relaxed_sys_u64(&s->w[word], half | (tag << 32));
  }
  const PeerSlot* mine = reinterpret_cast<const PeerSlot*>(V.win + V.L.slots) + (size_t)par * V.nranks;
  int ok = 1;
  if (tid < nw) {
    const int src = tid / (2 * K), word = tid % (2 * K);
    unsigned long long got = 0;
    ok = peer_wait_raw<1>(&mine[src].w[word], tag, &got, (const unsigned int*)(V.win + V.L.abort_word), V.timeout_ns, V.peers, V.nranks,
                          V.L.abort_word)
             ? 1
             : 0;
    sh[tid] = (unsigned int)(got & 0xffffffffull);
  }
  ok = __syncthreads_and(ok);
  if (tid < K) {
    double a = 0.0;
    for (int r = 0; r < V.nranks; ++r) {
      const unsigned long long bits = (unsigned long long)sh[r * 2 * K + 2 * tid] | ((unsigned long long)sh[r * 2 * K + 2 * tid + 1] << 32);
      a += __longlong_as_double((long long)bits);
    }
    bc[tid] = a;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = bc[k];
  __syncthreads();
  return ok != 0;
}

// K per-block partial arrays (each G long, consecutive) summed redundantly by every block in one fixed order
template <int K>
__device__ __forceinline__ void block_totals(const double* part, int G, double* smk /* K*8 */, double* bck /* K */, double (&out)[K]) {
  double a[K];
#pragma unroll
  for (int k = 0; k < K; ++k) a[k] = 0.0;
  for (int i = threadIdx.x; i < G; i += 256) {
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] += __ldcg(part + (size_t)k * G + i);
  }
#pragma unroll
  for (int k = 0; k < K; ++k) a[k] = warp_sum(a[k]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smk[k * 8 + (threadIdx.x >> 5)] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < K) {
    const double* src = smk + 8 * threadIdx.x;
    double r = src[0];
    for (int w = 1; w < 8; ++w) r += src[w];
    bck[threadIdx.x] = r;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) out[k] = bck[k];
  __syncthreads();
}
// K sums over a 256-thread block; results valid in thread 0
template <int K>
__device__ __forceinline__ void block_sums(double (&v)[K], double* smk /* K*8 */) {
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smk[k * 8 + (threadIdx.x >> 5)] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double r = smk[k * 8];
      for (int w = 1; w < 8; ++w) r += smk[k * 8 + w];
      v[k] = r;
    }
  }
  __syncthreads();
}

struct SrPcgParams {
  const int64_t* slice_off;
  const int* col;
  const uint16_t* col16;
  const uint8_t* s16;
  const int* rowlen;
  const double* hval;
  float* hval32;     // FP32 copy of the value stream, or nullptr (always FP64)
  double relax_tol;  // switch to it once ||r||/||b|| <= relax_tol
  const double* b;
  const double* dinv;
  double* x;
  double* u;  // dinv * r (the residual r itself is never stored)
  double* p;
  double* s;  // H p (by recurrence)
  double* w;  // H u
  double* partial;  // 6 * gridDim.x
  PcgScalars* sc;
  int n, nslices, maxit;
  int sweep;  // 1: slices dealt round-robin to the warps of the grid; 0: contiguous block ranges
  double rtol;
  unsigned long long* phase_ns;  // debug (SPB_PCG_TIMING=1): per block, 8 accumulated phase durations; else null
  // PEER only
  PeerCtl* ctl;
  int lo, hi, halo;  // owned local range [lo, hi), ghosts [0, lo) and [hi, n)
  char* left;        // window of the left / right neighbour
  char* right;
  PeerView V;
};

// one row sum of w = H u; vmode 0: FP64 values, 1: FP64 values and write their FP32 copy (the solve's first product),
// 2: the FP32 copy (relaxed-precision products once the residual is small; see PcgParams::hval32 in pcg.cu)
__device__ __forceinline__ double sr_row_acc(const SrPcgParams& P, int vmode, int slice, int lane) {
  const bool c16 = P.col16 && P.s16[slice];
  if (vmode == 2)
    return c16 ? spmv_slice_acc<false, false, true, float, false>(P.slice_off, P.col, P.rowlen, P.hval32, P.u, P.n, slice, lane, 0, P.col16)
               : spmv_slice_acc<false, false, false, float, false>(P.slice_off, P.col, P.rowlen, P.hval32, P.u, P.n, slice, lane, 0, P.col16);
  if (vmode == 1)
    return c16 ? spmv_slice_acc<false, false, true, double, true>(P.slice_off, P.col, P.rowlen, P.hval, P.u, P.n, slice, lane, 0, P.col16, P.hval32)
               : spmv_slice_acc<false, false, false, double, true>(P.slice_off, P.col, P.rowlen, P.hval, P.u, P.n, slice, lane, 0, P.col16, P.hval32);
  return c16 ? spmv_slice_acc<false, false, true>(P.slice_off, P.col, P.rowlen, P.hval, P.u, P.n, slice, lane, 0, P.col16)
             : spmv_slice_acc<false, false, false>(P.slice_off, P.col, P.rowlen, P.hval, P.u, P.n, slice, lane, 0, P.col16);
}

#define SPB_PTICK(k)                                         \
  if (P.phase_ns && tid == 0) {                              \
    const unsigned long long now_ = peer_global_ns();        \
    P.phase_ns[(size_t)blk * 8 + (k)] += now_ - tick_;       \
    tick_ = now_;                                            \
  }
// rows of band slice index k (0 <= k < nbs): the left band's slices first, then the right band's
#define SPB_BAND_SLICE(k) ((k) < sbl ? (k) : sbr + ((k)-sbl))

template <bool PEER, int MINB>
__global__ void __launch_bounds__(256, MINB) pcg_sr_kernel(SrPcgParams P) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double smk[32];
  __shared__ double bck[4];
  __shared__ unsigned int shw[PEER ? 256 : 1];
  const int G = gridDim.x, tid = threadIdx.x, blk = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int n = P.n;
  // per-block partial arrays, G doubles each:  [0] r.u  [1] w.u  [2] r.r  [3] "a wait failed in this block"  |  [4] x.x  [5] fail.
  // The fail flags are summed with the dot products and therefore seen identically by every block: the only place a
  // lost peer may end the loop.
  double* part = P.partial;
  // ---- which slices are mine -------------------------------------------------------------------------------
  // PEER: local rows [0, band) are shared with the left neighbour, [rb, n) with the right one (band = 2*halo).  The
  // slices holding them ("band slices": [0, sbl) and [sbr, nslices)) are dealt round-robin to ALL warps of the grid
  // and multiplied first; the interior slices [sbl, sbr) are split in contiguous block ranges.
  const int lo = PEER ? P.lo : 0, hi = PEER ? P.hi : n;
  const int band = PEER ? 2 * P.halo : 0, rb = PEER ? hi - P.halo : n;
  const int sbl = PEER ? (band + 31) / 32 : 0, sbr = PEER ? rb / 32 : P.nslices;
  const int nbs = sbl + (P.nslices - sbr);  // band slices
  const int gw = blk * 8 + warp, GW = G * 8;
  // Interior slices: dealt round-robin to the warps of the grid, continuing the deal of the band slices (unit t of
  // [band slices | interior slices] goes to warp t mod GW), so a warp that multiplied a band slice gets one interior
  // slice less, and at any moment the whole grid reads one contiguous window of the H stream.  P.sweep == 0: contiguous
  // private block ranges instead (the one-GPU kernel's layout).
  int i_first, i_end, i_step;
  if (P.sweep) {
    const int t0 = gw >= nbs ? gw : gw + ((nbs - gw + GW - 1) / GW) * GW;  // this warp's first unit that is not a band slice
    i_first = sbl + (t0 - nbs);
    i_end = sbr;
    i_step = GW;
  } else {
    i_first = sbl + (int)(((int64_t)(sbr - sbl) * blk) / G) + warp;
    i_end = sbl + (int)(((int64_t)(sbr - sbl) * (blk + 1)) / G);
    i_step = 8;
  }
  const bool block_has_band = PEER && (blk * 8 < nbs);
  unsigned long long seq = 0, round = 0;
  if (PEER) {  // counters written by block 0 at the very end of the previous solve (before any grid barrier of this one)
    seq = P.ctl->band_seq;
    round = P.ctl->round;
  }
  double failed = 0.0;
  bool lost = false;  // grid-uniform: some block of this rank gave up waiting

  // ---- init: x = 0, r = b, u = dinv*b, p = s = 0 on every local row ----------------------------------------------
  // (r,u) and (r,r) of the NEXT reduction are accumulated where r and u are produced (here and in phase A) and
  // carried in registers across the grid barrier; phase B only adds (w,u)
  double acc_ru = 0.0, acc_rr = 0.0;
  {
    auto init_row = [&](int row) {
      if (row < n) {
        const double bi = P.b[row];
        const double ui = bi * P.dinv[row];
        P.x[row] = 0.0;
        P.u[row] = ui;
        P.p[row] = 0.0;
        P.s[row] = 0.0;
        if (row >= lo && row < hi) {
          acc_ru = fma(bi, ui, acc_ru);
          acc_rr = fma(bi, bi, acc_rr);
        }
      }
    };
    if (PEER)
      for (int k = gw; k < nbs; k += GW) init_row(SPB_BAND_SLICE(k) * 32 + lane);
    for (int slice = i_first; slice < i_end; slice += i_step) init_row(slice * 32 + lane);
  }
  grid.sync();
  double gam_old = 0.0, alpha = 0.0, bb = 0.0, thresh = 0.0, rr = 0.0, denom = 0.0;
  int iters = 0, breakdown = 0;
  bool use32 = false;
  unsigned long long tick_ = P.phase_ns ? peer_global_ns() : 0ull;
  for (;;) {
    // ---- phase B: w = H_loc u, with (r,u), (w,u), (r,r) over the owned rows -------------------------------------
    double acc3[3] = {acc_ru, 0.0, acc_rr};
    const int vmode = use32 ? 2 : ((P.hval32 && iters == 0) ? 1 : 0);
    const int par = (int)(seq & 1ull);
    if (PEER) {
      // band rows first: to memory and straight into the neighbours' windows.  Every double travels as two 8-byte words
      // that carry 32 payload bits and the exchange tag (the flag-in-word form of NCCL's LL protocol): a word is complete
      // the moment its tag matches, so the writer needs no fence, no "slab complete" flag and no block barrier, and the
      // reader can take each row as soon as it has landed.  32 lanes write 512 contiguous bytes over NVLink.
      unsigned long long* const tl = reinterpret_cast<unsigned long long*>(P.left + P.V.L.band_from_right) + (size_t)par * band * 2;  // my view of the left band -> left's "from right"
      unsigned long long* const tr = reinterpret_cast<unsigned long long*>(P.right + P.V.L.band_from_left) + (size_t)par * band * 2;
      const unsigned long long tagw = ((seq + 1) & 0xffffffffull) << 32;
      for (int k = gw; k < nbs; k += GW) {
        const int slice = SPB_BAND_SLICE(k);
        const double acc = sr_row_acc(P, vmode, slice, lane);
        const int row = slice * 32 + lane;
        if (row < n) {
          P.w[row] = acc;
          const unsigned long long bits = (unsigned long long)__double_as_longlong(acc);
          unsigned long long* dst = row < band ? tl + 2 * (size_t)row : (row >= rb ? tr + 2 * (size_t)(row - rb) : nullptr);
          if (dst) {
            st_relaxed_sys_u64(dst, (bits & 0xffffffffull) | tagw);
            st_relaxed_sys_u64(dst + 1, (bits >> 32) | tagw);
          }
        }
      }
    }
    SPB_PTICK(0)
    for (int slice = i_first; slice < i_end; slice += i_step) {
      const double acc = sr_row_acc(P, vmode, slice, lane);
      const int row = slice * 32 + lane;
      if (row < n) {
        P.w[row] = acc;
        acc3[1] = fma(acc, P.u[row], acc3[1]);  // interior rows are all owned
      }
    }
    SPB_PTICK(1)
    if (PEER) {
      // the neighbours' views of the band rows have been in flight meanwhile: w = mine + theirs
      if (block_has_band) {
        const unsigned long long* const fl = reinterpret_cast<const unsigned long long*>(P.V.win + P.V.L.band_from_left) + (size_t)par * band * 2;
        const unsigned long long* const fr = reinterpret_cast<const unsigned long long*>(P.V.win + P.V.L.band_from_right) + (size_t)par * band * 2;
        const unsigned long long tag = (seq + 1) & 0xffffffffull;
        const unsigned int* const abort_w = reinterpret_cast<const unsigned int*>(P.V.win + P.V.L.abort_word);
        int bad = 0;
        for (int k = gw; k < nbs; k += GW) {
          const int row = SPB_BAND_SLICE(k) * 32 + lane;
          if (row < n) {
            double wi = P.w[row];
            if (row < band || row >= rb) {
              const unsigned long long* src = row < band ? fl + 2 * (size_t)row : fr + 2 * (size_t)(row - rb);
              unsigned long long w0 = 0, w1 = 0;
              if (!peer_wait_raw<1>(src, tag, &w0, abort_w, P.V.timeout_ns, P.V.peers, P.V.nranks, P.V.L.abort_word)) bad = 1;
              if (!peer_wait_raw<1>(src + 1, tag, &w1, abort_w, P.V.timeout_ns, P.V.peers, P.V.nranks, P.V.L.abort_word)) bad = 1;
              wi += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
              P.w[row] = wi;
            }  // else: an interior row of a slice that straddles a band boundary
            if (row >= lo && row < hi) acc3[1] = fma(wi, P.u[row], acc3[1]);
          }
        }
        if (__syncthreads_or(bad)) failed = 1.0;
      }
      ++seq;
    }
    block_sums<3>(acc3, smk);
    if (tid == 0) {
      part[blk] = acc3[0];
      part[G + blk] = acc3[1];
      part[2 * G + blk] = acc3[2];
      if (PEER) part[3 * G + blk] = failed;
    }
    SPB_PTICK(2)
    grid.sync();
    SPB_PTICK(3)
    double gam, delta;
    if (PEER) {
      double t[4];
      block_totals<4>(part, G, smk, bck, t);
      if (t[3] > 0.0) {
        lost = true;
        break;
      }
      double v3[3] = {t[0], t[1], t[2]};
      if (!peer_allsum<3>(P.V, round++, v3, shw, bck)) failed = 1.0;
      gam = v3[0];
      delta = v3[1];
      rr = v3[2];
    } else {
      double t[3];
      block_totals<3>(part, G, smk, bck, t);
      gam = t[0];
      delta = t[1];
      rr = t[2];
    }
    SPB_PTICK(4)
    double beta = 0.0;
    if (iters == 0) {
      bb = rr;  // r = b
      thresh = P.rtol * P.rtol * bb;
      if (!(bb > thresh) || P.maxit <= 0) break;  // zero (or NaN) right-hand side
      denom = delta;
    } else {
      if (!(rr > thresh) || iters >= P.maxit) break;
      if (P.hval32 && rr <= P.relax_tol * P.relax_tol * bb) use32 = true;  // same bits in every block of every rank
      beta = gam / gam_old;
      denom = delta - beta * gam / alpha;  // = p.Hp of the new direction
    }
    if (!(denom > 0.0)) {  // not SPD (the iterative analogue of a failed pivot), or NaN in H (quirk H: NaN step)
      breakdown = (denom != denom) ? 2 : 1;
      break;
    }
    alpha = gam / denom;
    gam_old = gam;
    // ---- phase A: p = u + beta p, s = w + beta s, x += alpha p, r -= alpha s, u = dinv r on every local row ---------
    acc_ru = 0.0;
    acc_rr = 0.0;
    {
      auto update_row = [&](int row) {
        if (row < n) {
          const double ui = P.u[row], wi = P.w[row], pi = P.p[row], si = P.s[row], xi = P.x[row], di = P.dinv[row];
          const double pn = fma(beta, pi, ui);
          const double sn = fma(beta, si, wi);
          // The residual itself is not stored: only its preconditioned form u = dinv r is kept (it is what the next product
          // gathers), updated as u -= alpha dinv s; r = u / dinv is formed where the dot products need it.  One vector less
          // to stream through L2 per iteration and one less dirty in it (ncu: the sixth read-write vector pushed 15 MB per
          // iteration of write-backs into the H stream's DRAM time).
          const double un = fma(-alpha * di, sn, ui);
          const double rn = un / di;
          P.p[row] = pn;
          P.s[row] = sn;
          P.x[row] = fma(alpha, pn, xi);
          P.u[row] = un;
          if (row >= lo && row < hi) {
            acc_ru = fma(rn, un, acc_ru);
            acc_rr = fma(rn, rn, acc_rr);
          }
        }
      };
      if (PEER)
        for (int k = gw; k < nbs; k += GW) update_row(SPB_BAND_SLICE(k) * 32 + lane);
      for (int slice = i_first; slice < i_end; slice += i_step) update_row(slice * 32 + lane);
    }
    ++iters;
    SPB_PTICK(5)
    grid.sync();
    SPB_PTICK(6)
  }
  // ---- epilogue: ||x||^2 over the owned entries of all ranks (LMSolver.h:146), NaN system => NaN step ----------
  const bool nan_system = (bb != bb) || breakdown == 2;
  double xx = 0.0;
  {
    double a1[1] = {0.0};
    auto xx_row = [&](int row) {
      if (row < n) {
        double v = P.x[row];
        if (nan_system) {
          v = __longlong_as_double(0x7ff8000000000000LL);
          P.x[row] = v;
        }
        if (row >= lo && row < hi) a1[0] = fma(v, v, a1[0]);
      }
    };
    if (PEER)
      for (int k = gw; k < nbs; k += GW) xx_row(SPB_BAND_SLICE(k) * 32 + lane);
    for (int slice = i_first; slice < i_end; slice += i_step) xx_row(slice * 32 + lane);
    block_sums<1>(a1, smk);
    // own arrays: a block that left the loop at a break may be followed by blocks still summing the loop's arrays
    if (tid == 0) {
      part[4 * G + blk] = a1[0];
      part[5 * G + blk] = failed;
    }
    grid.sync();
    double t2[2];
    block_totals<2>(part + (size_t)4 * G, G, smk, bck, t2);
    xx = t2[0];
    if (PEER) {
      if (t2[1] > 0.0) lost = true;
      if (!lost) {
        double v1[1] = {t2[0]};
        if (!peer_allsum<1>(P.V, round++, v1, shw, bck)) lost = true;  // nothing follows: a late failure only marks the result
        xx = v1[0];
      }
    }
  }
  if (PEER && lost) {
    if (tid == 0) peer_raise_abort(P.V);  // the other ranks must not wait for this one
    atomicExch(&P.ctl->failed, 1);
  }
  if (blk == 0 && tid == 0) {
    PcgScalars* sc = P.sc;
    sc->rz = gam_old;
    sc->pq = denom;
    sc->rr = rr;
    sc->bb = bb;
    sc->thresh = thresh;
    sc->iters = iters;
    sc->breakdown = breakdown;
    sc->maxit = P.maxit;
    sc->xx = xx;
    sc->done = 1;
    if (PEER) {
      P.ctl->band_seq = seq;
      P.ctl->round = round;
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
// blocks per SM the kernel is compiled for: 4 (64 registers) or 3 (80 registers: the eight-deep unrolled loads of the
// SpMV inner loop are then all issued before the first use); SPB_SR_MINB overrides
static int sr_minb() {
  static const int v = getenv("SPB_SR_MINB") ? atoi(getenv("SPB_SR_MINB")) : 3;
  return v == 4 ? 4 : 3;
}
template <bool PEER>
static void* sr_kernel_ptr() {
  return sr_minb() == 4 ? (void*)pcg_sr_kernel<PEER, 4> : (void*)pcg_sr_kernel<PEER, 3>;
}
template <bool PEER>
static int sr_grid(spb_context* h, int nslices) {
  static int occ_cached = 0;
  if (occ_cached == 0) {
    int occ = 0;
    if (sr_minb() == 4) SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_sr_kernel<PEER, 4>, 256, 0));
    else SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_sr_kernel<PEER, 3>, 256, 0));
    occ_cached = std::max(1, occ);
  }
  return std::max(1, std::min((nslices + 7) / 8, h->num_sms * std::min(occ_cached, 8)));
}

static void sr_common(spb_context* h, SrPcgParams& P, const double* d_b, double* d_x) {
  const int n = h->n;
  cudaStream_t st = h->stream;
  {
    const size_t need = (size_t)h->num_sms * 8 * 6 + 64;
    if (h->d_partial.n < need + 8) SPB_CUDA(h->d_partial.alloc(need + 8));
    if (!h->d_sc.p) {
      SPB_CUDA(h->d_sc.alloc(1));
      SPB_CUDA(cudaMemsetAsync(h->d_sc.p, 0, sizeof(PcgScalars), st));
    }
  }
  SPB_CUDA(h->d_p.alloc(n));
  SPB_CUDA(h->d_q.alloc(n));
  SPB_CUDA(h->d_u.alloc(n));
  SPB_CUDA(h->d_w.alloc(n));
  P.slice_off = h->d_slice_off.p;
  P.col = h->d_sell_col.p;
  P.col16 = h->pcg_col16 ? h->d_sell_col16.p : nullptr;
  P.s16 = h->d_slice16.p;
  P.rowlen = h->d_rowlen.p;
  P.hval = h->d_hval.p;
  P.hval32 = nullptr;
  P.relax_tol = 0.0;
  if (h->pcg_relax_factor > 0.0) {
    SPB_CUDA(h->d_hval32.alloc((size_t)h->S.padded));
    P.hval32 = h->d_hval32.p;
    P.relax_tol = std::min(1.0, h->pcg_rtol * h->pcg_relax_factor);
  }
  P.b = d_b;
  P.dinv = h->d_dinv.p;
  P.x = d_x;
  P.u = h->d_u.p;
  P.p = h->d_p.p;
  P.s = h->d_q.p;
  P.w = h->d_w.p;
  P.partial = h->d_partial.p;
  P.sc = h->d_sc.p;
  P.n = n;
  P.nslices = (n + 31) / 32;
  P.maxit = h->pcg_maxit;
  P.rtol = h->pcg_rtol;
  P.phase_ns = nullptr;
  static const int sweep = getenv("SPB_SPMV_SWEEP") ? atoi(getenv("SPB_SPMV_SWEEP")) : 1;
  P.sweep = sweep;
  P.ctl = nullptr;
  P.lo = 0;
  P.hi = n;
  P.halo = 0;
  P.left = P.right = nullptr;
  P.V = PeerView{nullptr, nullptr, PeerLayout{}, 0, 1, 0ull};
}

static void sr_print_phases(const char* tag, int rank, spb_context* h, const DevBuf<unsigned long long>& d_phase, int G) {
  std::vector<unsigned long long> t((size_t)G * 8);
  SPB_CUDA(cudaMemcpy(t.data(), d_phase.p, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
  static const char* names[7] = {"band_spmv+publish", "spmv", "band_fix+block_sums", "sync1", "totals+allsum", "update", "sync2"};
  const double it = std::max(1, h->h_sc->iters + 1);
  char line[1024];
  int o = snprintf(line, sizeof(line), "[spb %s timing] rank=%d iters=%d G=%d us/iteration mean(min..max over blocks):", tag, rank, h->h_sc->iters, G);
  for (int k = 0; k < 7; ++k) {
    double sum = 0, lo = 1e30, hi = 0;
    for (int b = 0; b < G; ++b) {
      const double v = (double)t[(size_t)b * 8 + k] / it * 1e-3;
      sum += v;
      lo = std::min(lo, v);
      hi = std::max(hi, v);
    }
    o += snprintf(line + o, sizeof(line) - (size_t)o, " %s=%.2f(%.2f..%.2f)", names[k], sum / G, lo, hi);
  }
  if (const char* dir = getenv("SPB_PCG_TIMING_DIR")) {  // one file per rank (the ranks' stderr lines interleave under torchrun)
    char path[512];
    snprintf(path, sizeof(path), "%s/pcg_timing_rank%d.txt", dir, rank);
    if (FILE* f = fopen(path, "a")) {
      fprintf(f, "%s\n", line);
      fclose(f);
    }
  } else {
    fprintf(stderr, "%s\n", line);
  }
}

static spb_status sr_finish(spb_context* h, int* iters) {
  h->pcg_last_iters = h->h_sc->iters;
  h->solve_xx_valid = true;
  if (iters) *iters = h->h_sc->iters;
  if (h->h_sc->breakdown == 2 || h->h_sc->bb != h->h_sc->bb) return SPB_OK;  // NaN system: NaN step (quirk H)
  if (h->h_sc->breakdown) return SPB_NOT_SPD;
  if (h->h_sc->rr > h->h_sc->thresh) return SPB_NOT_CONVERGED;
  return SPB_OK;
}

spb_status pcg_solve_sr(spb_context* h, const double* d_b, double* d_x, int* iters) {
  const int n = h->n;
  cudaStream_t st = h->stream;
  if (iters) *iters = 0;
  if (n == 0) return SPB_OK;
  SrPcgParams P;
  sr_common(h, P, d_b, d_x);
  const int G = sr_grid<false>(h, P.nslices);
  static const bool phase_timing = getenv("SPB_PCG_TIMING") && atoi(getenv("SPB_PCG_TIMING")) > 0;
  DevBuf<unsigned long long> d_phase;
  if (phase_timing) {
    SPB_CUDA(d_phase.alloc((size_t)G * 8));
    SPB_CUDA(cudaMemsetAsync(d_phase.p, 0, sizeof(unsigned long long) * (size_t)G * 8, st));
    P.phase_ns = d_phase.p;
  }
  void* args[] = {&P};
  stage_begin(h, ST_PCG);
  SPB_CUDA(cudaLaunchCooperativeKernel(sr_kernel_ptr<false>(), dim3(G), dim3(256), args, 0, st));
  stage_end(h, ST_PCG, 1);
  SPB_CUDA(cudaMemcpyAsync(h->h_sc, h->d_sc.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  if (phase_timing) sr_print_phases("pcg-sr", 0, h, d_phase, G);
  return sr_finish(h, iters);
}

spb_status pcg_solve_halo_peer(spb_context* h, const double* d_b, double* d_x, int* iters) {
  const int n = h->n;
  cudaStream_t st = h->stream;
  if (iters) *iters = 0;
  if (n == 0) return SPB_OK;
  PeerState* S = h->peer;
  if (!S || !S->ready) throw StateFail{"peer windows are not set up (spb_dist_ring_halo)"};
  if (S->failed) throw StateFail{"an earlier solve lost a peer (exchange timed out): re-create the halo description"};
  if (n != h->own_hi + h->halo) throw StateFail{"halo description does not match the analysed local problem"};
  if (S->band != 2 * h->halo) throw StateFail{"peer windows were sized for another halo width"};
  if (S->nranks * 6 > 256) throw StateFail{"the fused all-to-all supports at most 42 ranks"};
  SrPcgParams P;
  sr_common(h, P, d_b, d_x);
  P.ctl = S->d_ctl.p;
  P.lo = h->own_lo;
  P.hi = h->own_hi;
  P.halo = h->halo;
  P.left = S->peer[(size_t)(h->nranks > 1 ? h->left_rank : 0)];
  P.right = S->peer[(size_t)(h->nranks > 1 ? h->right_rank : 0)];
  P.V.win = S->win;
  P.V.peers = S->d_peer.p;
  P.V.L = S->L;
  P.V.rank = S->rank;
  P.V.nranks = S->nranks;
  static const long timeout_ms = getenv("SPB_PEER_TIMEOUT_MS") ? std::max(1L, atol(getenv("SPB_PEER_TIMEOUT_MS"))) : 20000L;
  P.V.timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
  const int G = sr_grid<true>(h, P.nslices);
  static const bool phase_timing = getenv("SPB_PCG_TIMING") && atoi(getenv("SPB_PCG_TIMING")) > 0;
  DevBuf<unsigned long long> d_phase;
  if (phase_timing) {
    SPB_CUDA(d_phase.alloc((size_t)G * 8));
    SPB_CUDA(cudaMemsetAsync(d_phase.p, 0, sizeof(unsigned long long) * (size_t)G * 8, st));
    P.phase_ns = d_phase.p;
  }
  void* args[] = {&P};
  stage_begin(h, ST_PCG);
  SPB_CUDA(cudaLaunchCooperativeKernel(sr_kernel_ptr<true>(), dim3(G), dim3(256), args, 0, st));
  stage_end(h, ST_PCG, 1);
  SPB_CUDA(cudaMemcpyAsync(h->h_sc, h->d_sc.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
  PeerCtl ctl;
  SPB_CUDA(cudaMemcpyAsync(&ctl, S->d_ctl.p, sizeof(PeerCtl), cudaMemcpyDeviceToHost, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  if (phase_timing) sr_print_phases("halo-peer", S->rank, h, d_phase, G);
  if (ctl.failed) {
    S->failed = true;
    throw StateFail{"halo exchange timed out waiting for a peer rank (SPB_PEER_TIMEOUT_MS)"};
  }
  return sr_finish(h, iters);
}

}  // namespace spb
