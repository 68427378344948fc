// pcg_peer.cu -- the subdomain (halo) Jacobi-PCG of config C5 as ONE persistent cooperative kernel per
// rank, fused with its collectives over NVLink peer memory (sm_100a, NVSwitch).
//
// What it replaces: pcg_solve_halo (pcg.cu) runs an iteration as four kernel launches, a grouped
// ncclSend/ncclRecv halo swap and two scalar ncclAllReduce calls -- at 8 ranks ~120 us of a 287 us iteration
// is launch + NCCL latency, nothing of size n moves.  Here an iteration is three grid barriers and nothing
// leaves the kernel:
//   * the SpMV phase writes the rows of the two exchanged bands straight into the neighbours' windows
//     (peer-mapped device memory, plain coalesced stores over NVLink) while it computes them; the blocks
//     that own band rows process them first in program order, publish "band delivered" with a
//     release-store on the neighbour once the last of them is done, and only then wait for the
//     neighbour's slab -- so the transfer overlaps the interior rows of every other block;
//   * dot products are combined by a flag+value all-to-all: block 0 writes this rank's partial (up to three
//     doubles) into slot [rank] of every rank's window, every block polls its own window's slots and adds
//     them in rank order, so all ranks (and all blocks) hold bit-identical alpha / beta / stop decisions;
//   * vector updates run redundantly on the ghost entries from identical inputs, so p is never exchanged
//     (same scheme as pcg_solve_halo).
// Waiting is bounded: a spin that exceeds the timeout raises an abort word in every rank's window; the
// kernels then leave at their next grid-uniform check and the solve returns SPB_ERR_STATE.
//
// Reference semantics kept: the solve stands in for EigenSolverWrapper::solve (EigenSolverWrapper.h:72-78) on
// the damped normal matrix of LMSolver.h:136; acceptance (LMSolver.h:161-171) sees the same step as the
// one-GPU solve to ~1e-16 relative (tests/test_gpu_halo.py, scripts/halo_check.py).
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
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_f64(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
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

// spin until *addr >= want; false on abort / timeout (the caller carries on and reports at the next grid-uniform check)
__device__ __noinline__ bool peer_wait_ge_raw(const unsigned long long* addr, unsigned long long want, const unsigned int* abort_w,
                                              unsigned long long timeout_ns, char* const* peers, int nranks, size_t abort_off) {
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    if (ld_acquire_sys_u64(addr) >= want) return true;
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
  return peer_wait_ge_raw(addr, want, (const unsigned int*)(V.win + V.L.abort_word), V.timeout_ns, V.peers, V.nranks, V.L.abort_word);
}

// Sum of K doubles over all ranks.  Precondition: every block of this rank holds the same local values v[]
// (finalised redundantly from the per-block partials).  `round` counts these calls since the windows were set up
// and is the same number on every rank; slots are double-buffered by its parity (a rank can be at most one round
// ahead of the slowest reader: posting round r+2 needs everybody's round r+1, which a rank posts only after all
// its blocks have read round r).  Result: the sum in rank order -- the same bits in every block of every rank.
template <int K>
__device__ __forceinline__ bool peer_allsum(const PeerView& V, unsigned long long round, double (&v)[K], double* bc /* shared, >= K */) {
  const int tid = threadIdx.x;
  const int par = (int)(round & 1ull);
  if (blockIdx.x == 0 && tid < V.nranks) {  // thread t delivers to rank t
    PeerSlot* s = reinterpret_cast<PeerSlot*>(V.peers[tid] + V.L.slots) + (size_t)par * V.nranks + V.rank;
#pragma unroll
    for (int k = 0; k < K; ++k) st_relaxed_sys_f64(&s->v[k], v[k]);
    __threadfence_system();
    st_release_sys_u64(&s->seq, round + 1);
  }
  const PeerSlot* mine = reinterpret_cast<const PeerSlot*>(V.win + V.L.slots) + (size_t)par * V.nranks;
  int ok = 1;
  if (tid < V.nranks) ok = peer_wait_ge(V, &mine[tid].seq, round + 1) ? 1 : 0;
  ok = __syncthreads_and(ok);
  if (tid < K) {
    double a = 0.0;
    for (int r = 0; r < V.nranks; ++r) a += ld_relaxed_sys_f64(&mine[r].v[tid]);
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

struct HaloPcgParams {
  const int64_t* slice_off;
  const int* col;
  const uint16_t* col16;
  const uint8_t* s16;
  const int* rowlen;
  const double* hval;
  const double* b;
  const double* dinv;
  double* x;
  double* r;
  double* p;
  double* q;
  double* partial;  // 7 * gridDim.x
  PcgScalars* sc;
  PeerCtl* ctl;
  int n, nslices, maxit;
  double rtol;
  int lo, hi, halo;  // owned local range [lo, hi), ghosts [0, lo) and [hi, n)
  char* left;        // window of the left / right neighbour
  char* right;
  PeerView V;
  unsigned long long* phase_ns;  // debug (SPB_PCG_TIMING=1): per block, 9 accumulated phase durations; else null
};

#define SPB_PTICK(k)                                         \
  if (P.phase_ns && tid == 0) {                              \
    const unsigned long long now_ = peer_global_ns();        \
    P.phase_ns[(size_t)blk * 9 + (k)] += now_ - tick_;       \
    tick_ = now_;                                            \
  }

template <int MINB>
__global__ void __launch_bounds__(256, MINB) pcg_halo_peer_kernel(HaloPcgParams P) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sm[8];
  __shared__ double smk[32];
  __shared__ double bck[4];
  const int G = gridDim.x, tid = threadIdx.x, blk = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int n = P.n, lo = P.lo, hi = P.hi;
  const int stride = G * 256;
  const PeerView& V = P.V;
  // per-block partial arrays, G doubles each; a phase's sums are followed by its "a wait failed in this block" flag,
  // summed with them and therefore seen identically by every block (the only place a lost peer may end the loop):
  //   [0] p.q  [1] fail   |   [2] r.z  [3] r.r (init: b.b)  [4] fail   |   [5] x.x  [6] fail
  // Phases that may overlap in time (a fast block writing its next partial while a slow one still sums the previous
  // phase's) use different arrays.
  double* part = P.partial;
  const int band = 2 * P.halo, rb = hi - P.halo;  // shared bands: local [0, band) with the left, [rb, n) with the right neighbour
  const int s_begin = (int)(((int64_t)P.nslices * blk) / G);
  const int s_end = (int)(((int64_t)P.nslices * (blk + 1)) / G);
  const int r_begin = s_begin * 32, r_end = min(n, s_end * 32);
  const int nL = max(0, min(r_end, band) - r_begin);  // this block's rows of the left / right band
  const int nR = max(0, r_end - max(r_begin, rb));
  // counters written by block 0 at the very end of the previous solve (before any grid barrier of this one)
  unsigned long long seq = P.ctl->band_seq, round = P.ctl->round;
  double failed = 0.0;

  // init: r = b, p = z = dinv*r, x = 0 on every local entry (ghosts included); r.z and b.b over the owned ones
  {
    double rz = 0.0, bb = 0.0;
    for (int i = blk * 256 + tid; i < n; i += stride) {
      const double bi = P.b[i];
      const double z = bi * P.dinv[i];
      P.x[i] = 0.0;
      P.r[i] = bi;
      P.p[i] = z;
      if (i >= lo && i < hi) {
        rz = fma(bi, z, rz);
        bb = fma(bi, bi, bb);
      }
    }
    const double s1 = block_sum_256(rz, sm);
    const double s2 = block_sum_256(bb, sm);
    if (tid == 0) {
      part[2 * G + blk] = s1;
      part[3 * G + blk] = s2;
    }
  }
  grid.sync();
  double rz, bb;
  {
    double t2[2];
    block_totals<2>(part + (size_t)2 * G, G, smk, bck, t2);
    if (!peer_allsum<2>(V, round++, t2, bck)) failed = 1.0;
    rz = t2[0];
    bb = t2[1];
  }
  const double thresh = P.rtol * P.rtol * bb;
  double rr = bb, pq = 0.0;
  int iters = 0, breakdown = 0;
  bool run = (bb > thresh) && P.maxit > 0;
  bool lost = false;  // grid-uniform: some block of this rank gave up waiting
  grid.sync();        // the partial arrays are about to be reused
  unsigned long long tick_ = P.phase_ns ? peer_global_ns() : 0ull;
  while (run) {
    // ---- phase 1: q = H_loc p on all local rows; band rows go to the neighbours as they are produced --------
    double loc = 0.0;
    {
      const int par = (int)(seq & 1ull);
      // my view of the left band lands in the left neighbour's "from right" buffer, and vice versa
      double* const tl = reinterpret_cast<double*>(P.left + V.L.band_from_right) + (size_t)par * band;
      double* const tr = reinterpret_cast<double*>(P.right + V.L.band_from_left) + (size_t)par * band;
      for (int slice = s_begin + warp; slice < s_end; slice += 8) {
        const double acc = (P.col16 && P.s16[slice])
                               ? spmv_slice_acc<false, false, true>(P.slice_off, P.col, P.rowlen, P.hval, P.p, n, slice, lane, 0, P.col16)
                               : spmv_slice_acc<false, false, false>(P.slice_off, P.col, P.rowlen, P.hval, P.p, n, slice, lane, 0, P.col16);
        const int row = slice * 32 + lane;
        if (row < n) {
          P.q[row] = acc;
          if (row < band) tl[row] = acc;                     // 32 lanes = 256 contiguous bytes over NVLink
          else if (row >= rb) tr[row - rb] = acc;
          else loc = fma(P.p[row], acc, loc);                // interior rows are all owned
        }
      }
      SPB_PTICK(0)
      if (nL + nR > 0) {
        __threadfence_system();  // this thread's slab stores are ordered before the ticket below
        __syncthreads();
        if (tid == 0) {
          // the block that completes a band (all rows written by all blocks that share it) publishes it
          if (nL > 0) {
            const unsigned long long t = atomicAdd(&P.ctl->sent_left, (unsigned long long)nL) + (unsigned long long)nL;
            if (t == (unsigned long long)band * (seq + 1)) {
              __threadfence_system();
              st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.left + V.L.flag_from_right), seq + 1);
            }
          }
          if (nR > 0) {
            const unsigned long long t = atomicAdd(&P.ctl->sent_right, (unsigned long long)nR) + (unsigned long long)nR;
            if (t == (unsigned long long)band * (seq + 1)) {
              __threadfence_system();
              st_release_sys_u64(reinterpret_cast<unsigned long long*>(P.right + V.L.flag_from_left), seq + 1);
            }
          }
        }
        // the neighbour's view of the same rows: q = mine + theirs (a + b is commutative: the owner and the ghost
        // holder end up with the same bits)
        int ok = 1;
        if (tid == 0) {
          if (nL > 0 && !peer_wait_ge(V, reinterpret_cast<const unsigned long long*>(V.win + V.L.flag_from_left), seq + 1)) ok = 0;
          if (nR > 0 && !peer_wait_ge(V, reinterpret_cast<const unsigned long long*>(V.win + V.L.flag_from_right), seq + 1)) ok = 0;
        }
        ok = __syncthreads_and(ok);
        if (!ok) failed = 1.0;
        const double* const fl = reinterpret_cast<const double*>(V.win + V.L.band_from_left) + (size_t)par * band;
        const double* const fr = reinterpret_cast<const double*>(V.win + V.L.band_from_right) + (size_t)par * band;
        if (nL > 0) {
          const int e = min(r_end, band);
          for (int i = r_begin + tid; i < e; i += 256) {
            const double qi = P.q[i] + __ldcg(fl + i);
            P.q[i] = qi;
            if (i >= lo) loc = fma(P.p[i], qi, loc);
          }
        }
        if (nR > 0) {
          for (int i = max(r_begin, rb) + tid; i < r_end; i += 256) {
            const double qi = P.q[i] + __ldcg(fr + (i - rb));
            P.q[i] = qi;
            if (i < hi) loc = fma(P.p[i], qi, loc);
          }
        }
      }
      ++seq;
      const double bs = block_sum_256(loc, sm);
      if (tid == 0) {
        part[blk] = bs;
        part[G + blk] = failed;
      }
    }
    SPB_PTICK(1)
    grid.sync();
    SPB_PTICK(2)
    {
      double t[2];
      block_totals<2>(part, G, smk, bck, t);
      if (t[1] > 0.0) {
        lost = true;
        break;
      }
      double v1[1] = {t[0]};
      if (!peer_allsum<1>(V, round++, v1, bck)) failed = 1.0;
      pq = v1[0];
    }
    SPB_PTICK(3)
    if (!(pq > 0.0)) {  // not SPD (pq <= 0), or NaN in H (quirk H: the LLT would hand back a NaN step)
      breakdown = (pq != pq) ? 2 : 1;
      break;
    }
    const double alpha = rz / pq;
    // ---- phase 2: x += alpha p, r -= alpha q on every local entry; r.z and r.r over the owned ones ---------
    {
      double a1 = 0.0, a2 = 0.0;
      int i = blk * 256 + tid;
      for (; i + 3 * stride < n; i += 4 * stride) {
        double pv[4], qv[4], xv[4], rv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i + u * stride;
          pv[u] = P.p[j];
          qv[u] = P.q[j];
          xv[u] = P.x[j];
          rv[u] = P.r[j];
          dv[u] = P.dinv[j];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i + u * stride;
          P.x[j] = fma(alpha, pv[u], xv[u]);
          const double ri = fma(-alpha, qv[u], rv[u]);
          P.r[j] = ri;
          if (j >= lo && j < hi) {
            const double z = ri * dv[u];
            a1 = fma(ri, z, a1);
            a2 = fma(ri, ri, a2);
          }
        }
      }
      for (; i < n; i += stride) {
        P.x[i] = fma(alpha, P.p[i], P.x[i]);
        const double ri = fma(-alpha, P.q[i], P.r[i]);
        P.r[i] = ri;
        if (i >= lo && i < hi) {
          const double z = ri * P.dinv[i];
          a1 = fma(ri, z, a1);
          a2 = fma(ri, ri, a2);
        }
      }
      const double s1 = block_sum_256(a1, sm);
      const double s2 = block_sum_256(a2, sm);
      if (tid == 0) {
        part[2 * G + blk] = s1;
        part[3 * G + blk] = s2;
        part[4 * G + blk] = failed;
      }
    }
    SPB_PTICK(4)
    grid.sync();
    SPB_PTICK(5)
    double rz_new;
    {
      double t[3];
      block_totals<3>(part + (size_t)2 * G, G, smk, bck, t);
      if (t[2] > 0.0) {
        lost = true;
        break;
      }
      double v2[2] = {t[0], t[1]};
      if (!peer_allsum<2>(V, round++, v2, bck)) failed = 1.0;
      rz_new = v2[0];
      rr = v2[1];
    }
    SPB_PTICK(6)
    const double beta = rz_new / rz;
    rz = rz_new;
    ++iters;
    if (!(rr > thresh) || iters >= P.maxit) break;
    // ---- phase 3: p = dinv*r + beta p on every local entry --------------------------------------------------
    {
      int i = blk * 256 + tid;
      for (; i + 3 * stride < n; i += 4 * stride) {
        double pv[4], rv[4], dv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = i + u * stride;
          pv[u] = P.p[j];
          rv[u] = P.r[j];
          dv[u] = P.dinv[j];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) P.p[i + u * stride] = fma(beta, pv[u], rv[u] * dv[u]);
      }
      for (; i < n; i += stride) P.p[i] = fma(beta, P.p[i], P.r[i] * P.dinv[i]);
    }
    SPB_PTICK(7)
    grid.sync();
    SPB_PTICK(8)
  }
  // ---- epilogue: ||x||^2 over the owned entries of all ranks (LMSolver.h:146), NaN system => NaN step ----------
  const bool nan_system = (bb != bb) || breakdown == 2;
  double xx = 0.0;
  {
    double a = 0.0;
    for (int i = blk * 256 + tid; i < n; i += stride) {
      double v = P.x[i];
      if (nan_system) {
        v = __longlong_as_double(0x7ff8000000000000LL);
        P.x[i] = v;
      }
      if (i >= lo && i < hi) a = fma(v, v, a);
    }
    const double s1 = block_sum_256(a, sm);
    // own arrays: a block that left the loop at a break may be followed by blocks still summing the loop's arrays
    if (tid == 0) {
      part[5 * G + blk] = s1;
      part[6 * G + blk] = failed;
    }
    grid.sync();
    double t2[2];
    block_totals<2>(part + (size_t)5 * G, G, smk, bck, t2);
    if (t2[1] > 0.0) lost = true;
    if (!lost) {
      double v1[1] = {t2[0]};
      if (!peer_allsum<1>(V, round++, v1, bck)) lost = true;  // nothing follows: a late failure only marks the result
      xx = v1[0];
    }
  }
  if (lost && tid == 0) peer_raise_abort(V);  // the other ranks must not wait for this one
  if (lost) atomicExch(&P.ctl->failed, 1);
  if (blk == 0 && tid == 0) {
    PcgScalars* sc = P.sc;
    sc->rz = rz;
    sc->pq = pq;
    sc->rr = rr;
    sc->bb = bb;
    sc->thresh = thresh;
    sc->iters = iters;
    sc->breakdown = breakdown;
    sc->maxit = P.maxit;
    sc->xx = xx;
    sc->done = 1;
    P.ctl->band_seq = seq;
    P.ctl->round = round;
  }
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
  // partial arrays: 7 segments of G doubles
  {
    const size_t need = (size_t)h->num_sms * 8 * 7 + 64;
    if (h->d_partial.n < need + 8) SPB_CUDA(h->d_partial.alloc(need + 8));
    if (!h->d_sc.p) {
      SPB_CUDA(h->d_sc.alloc(1));
      SPB_CUDA(cudaMemsetAsync(h->d_sc.p, 0, sizeof(PcgScalars), st));
    }
  }
  SPB_CUDA(h->d_r.alloc(n));
  SPB_CUDA(h->d_p.alloc(n));
  SPB_CUDA(h->d_q.alloc(n));
  const int nslices = (n + 31) / 32;
  static int occ_cached = 0;
  if (occ_cached == 0) {
    int occ = 0;
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pcg_halo_peer_kernel<4>, 256, 0));
    occ_cached = std::max(1, occ);
  }
  const int G = std::max(1, std::min((nslices + 7) / 8, h->num_sms * std::min(occ_cached, 8)));
  HaloPcgParams P;
  P.slice_off = h->d_slice_off.p;
  P.col = h->d_sell_col.p;
  P.col16 = h->pcg_col16 ? h->d_sell_col16.p : nullptr;
  P.s16 = h->d_slice16.p;
  P.rowlen = h->d_rowlen.p;
  P.hval = h->d_hval.p;
  P.b = d_b;
  P.dinv = h->d_dinv.p;
  P.x = d_x;
  P.r = h->d_r.p;
  P.p = h->d_p.p;
  P.q = h->d_q.p;
  P.partial = h->d_partial.p;
  P.sc = h->d_sc.p;
  P.ctl = S->d_ctl.p;
  P.n = n;
  P.nslices = nslices;
  P.maxit = h->pcg_maxit;
  P.rtol = h->pcg_rtol;
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
  static const bool phase_timing = getenv("SPB_PCG_TIMING") && atoi(getenv("SPB_PCG_TIMING")) > 0;
  DevBuf<unsigned long long> d_phase;
  P.phase_ns = nullptr;
  if (phase_timing) {
    SPB_CUDA(d_phase.alloc((size_t)G * 9));
    SPB_CUDA(cudaMemsetAsync(d_phase.p, 0, sizeof(unsigned long long) * (size_t)G * 9, st));
    P.phase_ns = d_phase.p;
  }
  void* args[] = {&P};
  stage_begin(h, ST_PCG);
  SPB_CUDA(cudaLaunchCooperativeKernel((void*)pcg_halo_peer_kernel<4>, dim3(G), dim3(256), args, 0, st));
  stage_end(h, ST_PCG, 1);
  SPB_CUDA(cudaMemcpyAsync(h->h_sc, h->d_sc.p, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
  PeerCtl ctl;
  SPB_CUDA(cudaMemcpyAsync(&ctl, S->d_ctl.p, sizeof(PeerCtl), cudaMemcpyDeviceToHost, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  if (phase_timing) {
    std::vector<unsigned long long> t((size_t)G * 9);
    SPB_CUDA(cudaMemcpy(t.data(), d_phase.p, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
    static const char* names[9] = {"spmv", "exchange", "sync1", "allsum1", "update", "sync2", "allsum2", "direction", "sync3"};
    const double it = std::max(1, h->h_sc->iters);
    fprintf(stderr, "[spb halo-peer timing] rank=%d iters=%d G=%d us/iteration mean(min..max over blocks):", S->rank, h->h_sc->iters, G);
    for (int k = 0; k < 9; ++k) {
      double sum = 0, lo = 1e30, hi = 0;
      for (int b = 0; b < G; ++b) {
        const double v = (double)t[(size_t)b * 9 + k] / it * 1e-3;
        sum += v;
        lo = std::min(lo, v);
        hi = std::max(hi, v);
      }
      fprintf(stderr, " %s=%.2f(%.2f..%.2f)", names[k], sum / G, lo, hi);
    }
    fprintf(stderr, "\n");
  }
  if (ctl.failed) {
    S->failed = true;
    throw StateFail{"halo exchange timed out waiting for a peer rank (SPB_PEER_TIMEOUT_MS)"};
  }
  h->pcg_last_iters = h->h_sc->iters;
  h->solve_xx_valid = true;
  if (iters) *iters = h->h_sc->iters;
  if (h->h_sc->breakdown == 2 || h->h_sc->bb != h->h_sc->bb) return SPB_OK;  // NaN system: NaN step (quirk H)
  if (h->h_sc->breakdown) return SPB_NOT_SPD;
  if (h->h_sc->rr > h->h_sc->thresh) return SPB_NOT_CONVERGED;
  return SPB_OK;
}

}  // namespace spb
