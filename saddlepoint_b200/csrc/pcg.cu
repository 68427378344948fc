// pcg.cu -- sliced-ELL FP64 SpMV, deterministic reductions, and the Jacobi-preconditioned
// conjugate-gradient solver built on them (sm_100a).
//
// The reference solves (J^T J + D) dx = rhs with Eigen::SimplicialLLT
// (EigenSolverWrapper.h:58,76).  This file is the iterative alternative north_star item (4)
// asks for: same matrix, same right-hand side, converged to a residual small enough that the
// step agrees with the direct solve to ~1e-11 relative.  It is also where the LM loop's
// vector reductions live (EVec.squaredNorm(), direction.norm(), LMSolver.h:146,155,157,175).
//
// Design for B200:
//  * H is stored once, in the layout the SpMV wants (sliced ELL, slice height = warp width):
//    a warp owns 32 consecutive rows, lane == row, and every load instruction of the value
//    and column streams is a fully coalesced 256-byte / 128-byte transaction; the x gathers
//    hit L1/L2 (x is 8 MB at 1M unknowns, the 126 MB L2 keeps it resident).  Per-row sums run
//    in ascending column order.
//  * No host round trip inside the iteration: alpha/beta and the convergence test are
//    computed by the last-arriving block of the reduction kernels ("ticket" pattern, fixed
//    summation order => bitwise deterministic), kernels of already-converged iterations
//    exit on a device flag, and the host only polls a pinned mirror once per batch.
#include <algorithm>
#include <cmath>

#include "spb_internal.h"

namespace spb {

// ---- block reduction helpers ---------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over a 256-thread block; result valid in thread 0
__device__ __forceinline__ double block_sum_256(double v, double* sm /*8*/) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) {
    r = sm[0];
    for (int w = 1; w < 8; ++w) r += sm[w];
  }
  __syncthreads();
  return r;
}
// the last block to take a ticket sums `count` partials in a fixed order (one warp)
__device__ __forceinline__ bool take_ticket(unsigned int* ticket, unsigned int nblocks) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    last = (t == nblocks - 1);
    if (last) *ticket = 0;
  }
  __syncthreads();
  if (last) __threadfence();
  return last;
}
__device__ __forceinline__ double ordered_sum(const double* partial, int count) {
  // called by warp 0 of the last block; deterministic: lane-strided sums then a shuffle tree
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += 32) s += partial[i];
  return warp_sum(s);
}

// ---- generic reductions --------------------------------------------------------------
// op 0: sum of squares; op 1: max |v| (NaN never wins)
template <int OP>
__global__ void __launch_bounds__(256) reduce_kernel(const double* __restrict__ v, int64_t len, double* __restrict__ partial,
                                                     double* __restrict__ result, unsigned int* ticket) {
  __shared__ double sm[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < len; i += (int64_t)gridDim.x * 256) {
    double a = v[i];
    if (OP == 0) acc = fma(a, a, acc); else acc = fmax(acc, fabs(a));
  }
  double b;
  if (OP == 0) {
    b = block_sum_256(acc, sm);
  } else {
    acc = warp_max(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    b = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < 8; ++w) b = fmax(b, sm[w]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = b;
  if (take_ticket(ticket, gridDim.x) && threadIdx.x < 32) {
    double s;
    if (OP == 0) {
      s = ordered_sum(partial, gridDim.x);
    } else {
      s = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s = fmax(s, partial[i]);
      s = warp_max(s);
    }
    if (threadIdx.x == 0) *result = s;
  }
}

__global__ void add_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int64_t len) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) out[i] = a[i] + b[i];
}

static int reduce_grid(int64_t len) {
  int64_t g = (len + 256 * 8 - 1) / (256 * 8);
  return (int)std::max<int64_t>(1, std::min<int64_t>(g, 1184));  // 8 x 148 SMs
}

static void ensure_pcg_buffers(spb_context* h) {
  size_t need = std::max<size_t>(1184 * 2, (size_t)((h->n + 255) / 256) * 2 + 16);
  SPB_CUDA(h->d_partial.alloc(need + 8));
  if (!h->d_sc.p) {
    SPB_CUDA(h->d_sc.alloc(1));
    SPB_CUDA(cudaMemsetAsync(h->d_sc.p, 0, sizeof(PcgScalars), h->stream));
  }
}

static void dev_reduce(spb_context* h, int op, const double* d_v, int64_t len, double* host_out) {
  ensure_pcg_buffers(h);
  cudaStream_t st = h->stream;
  double* result = h->d_partial.p + (h->d_partial.n - 1);
  stage_begin(h, ST_REDUCE);
  if (len <= 0) {
    *host_out = 0.0;
    stage_end(h, ST_REDUCE, 0);
    return;
  }
  int g = reduce_grid(len);
  unsigned int* ticket = &h->d_sc.p->ticket_a;
  if (op == 0) reduce_kernel<0><<<g, 256, 0, st>>>(d_v, len, h->d_partial.p, result, ticket);
  else reduce_kernel<1><<<g, 256, 0, st>>>(d_v, len, h->d_partial.p, result, ticket);
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_REDUCE, 1);
  SPB_CUDA(cudaMemcpyAsync(h->h_scalar, result, sizeof(double), cudaMemcpyDeviceToHost, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  *host_out = h->h_scalar[0];
}
void dev_squared_norm(spb_context* h, const double* d_v, int64_t len, double* host_out) { dev_reduce(h, 0, d_v, len, host_out); }
void dev_inf_norm(spb_context* h, const double* d_v, int64_t len, double* host_out) { dev_reduce(h, 1, d_v, len, host_out); }
void dev_add(spb_context* h, const double* a, const double* b, double* out, int64_t len) {
  if (len <= 0) return;
  add_kernel<<<(unsigned)((len + 255) / 256), 256, 0, h->stream>>>(a, b, out, len);
  h->launches++;
  SPB_CUDA(cudaGetLastError());
}

// ---- SpMV ----------------------------------------------------------------------------
// y = H x; lane == row inside a 32-row slice.  WITH_DOT additionally produces the PCG's
// p.q partials and, in the last block, alpha = rz / pq.
template <bool WITH_DOT>
__global__ void __launch_bounds__(256) spmv_sell_kernel(const int64_t* __restrict__ slice_off, const int* __restrict__ col,
                                                        const int* __restrict__ rowlen, const double* __restrict__ hval,
                                                        const double* __restrict__ x, double* __restrict__ y, int n,
                                                        double* __restrict__ partial, PcgScalars* sc) {
  __shared__ double sm[8];
  if (WITH_DOT) {
    if (sc->done) return;
  }
  const int row = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int slice = row >> 5;
  double acc = 0.0;
  if (slice * 32 < n) {
    const int64_t off = slice_off[slice];
    const int width = (int)((slice_off[slice + 1] - off) >> 5);
    const int len = row < n ? rowlen[row] : 0;
    const int* cp = col + off + lane;
    const double* vp = hval + off + lane;
    int k = 0;
    for (; k + 4 <= width; k += 4) {
      const int c0 = __ldcs(cp + (k + 0) * 32), c1 = __ldcs(cp + (k + 1) * 32), c2 = __ldcs(cp + (k + 2) * 32), c3 = __ldcs(cp + (k + 3) * 32);
      const double v0 = __ldcs(vp + (k + 0) * 32), v1 = __ldcs(vp + (k + 1) * 32), v2 = __ldcs(vp + (k + 2) * 32), v3 = __ldcs(vp + (k + 3) * 32);
      const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
      if (k + 0 < len) acc = fma(v0, x0, acc);
      if (k + 1 < len) acc = fma(v1, x1, acc);
      if (k + 2 < len) acc = fma(v2, x2, acc);
      if (k + 3 < len) acc = fma(v3, x3, acc);
    }
    for (; k < width; ++k) {
      const int c0 = __ldcs(cp + k * 32);
      const double v0 = __ldcs(vp + k * 32);
      const double x0 = __ldg(x + c0);
      if (k < len) acc = fma(v0, x0, acc);
    }
    if (row < n) y[row] = acc;
  }
  if (WITH_DOT) {
    double loc = row < n ? x[row] * acc : 0.0;
    double b = block_sum_256(loc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = b;
    if (take_ticket(&sc->ticket_a, gridDim.x) && threadIdx.x < 32) {
      double pq = ordered_sum(partial, gridDim.x);
      if (threadIdx.x == 0) {
        sc->pq = pq;
        if (!(pq > 0.0)) {  // not SPD (or NaN): stop, report breakdown
          sc->breakdown = 1;
          sc->done = 1;
          sc->alpha = 0.0;
        } else {
          sc->alpha = sc->rz / pq;
        }
      }
    }
  }
}

// r = b, z = dinv*r, p = z, x = 0; rz = r.z, bb = b.b
__global__ void __launch_bounds__(256) pcg_init_kernel(const double* __restrict__ b, const double* __restrict__ hdiag,
                                                       double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
                                                       int n, double* __restrict__ partial, PcgScalars* sc, double rtol, int maxit) {
  __shared__ double sm[8];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double rz = 0.0, bb = 0.0;
  if (i < n) {
    const double bi = b[i];
    const double z = bi * (1.0 / hdiag[i]);
    x[i] = 0.0;
    r[i] = bi;
    p[i] = z;
    rz = bi * z;
    bb = bi * bi;
  }
  double s1 = block_sum_256(rz, sm);
  double s2 = block_sum_256(bb, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s1;
    partial[gridDim.x + blockIdx.x] = s2;
  }
  if (take_ticket(&sc->ticket_b, gridDim.x) && threadIdx.x < 32) {
    double a = ordered_sum(partial, gridDim.x);
    double c = ordered_sum(partial + gridDim.x, gridDim.x);
    if (threadIdx.x == 0) {
      sc->rz = a;
      sc->bb = c;
      sc->rr = c;
      sc->thresh = rtol * rtol * c;
      sc->iters = 0;
      sc->maxit = maxit;
      sc->breakdown = 0;
      sc->alpha = 0.0;
      sc->beta = 0.0;
      sc->done = (c > sc->thresh && maxit > 0) ? 0 : 1;  // bb==0 or NaN: nothing to iterate
    }
  }
}

// x += alpha p; r -= alpha q; partial rz' = sum r*(dinv*r), rr = sum r*r; last block: beta, stop test
__global__ void __launch_bounds__(256) pcg_update_kernel(const double* __restrict__ hdiag, const double* __restrict__ p,
                                                         const double* __restrict__ q, double* __restrict__ x,
                                                         double* __restrict__ r, int n, double* __restrict__ partial, PcgScalars* sc) {
  __shared__ double sm[8];
  if (sc->done) return;
  const double alpha = sc->alpha;
  const int i = blockIdx.x * 256 + threadIdx.x;
  double rz = 0.0, rr = 0.0;
  if (i < n) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    const double z = ri * (1.0 / hdiag[i]);
    rz = ri * z;
    rr = ri * ri;
  }
  double s1 = block_sum_256(rz, sm);
  double s2 = block_sum_256(rr, sm);
  if (threadIdx.x == 0) {
    partial[blockIdx.x] = s1;
    partial[gridDim.x + blockIdx.x] = s2;
  }
  if (take_ticket(&sc->ticket_b, gridDim.x) && threadIdx.x < 32) {
    double a = ordered_sum(partial, gridDim.x);
    double c = ordered_sum(partial + gridDim.x, gridDim.x);
    if (threadIdx.x == 0) {
      sc->beta = a / sc->rz;
      sc->rz = a;
      sc->rr = c;
      int it = sc->iters + 1;
      sc->iters = it;
      if (!(c > sc->thresh) || it >= sc->maxit) sc->done = 1;
    }
  }
}

// p = dinv*r + beta p
__global__ void __launch_bounds__(256) pcg_direction_kernel(const double* __restrict__ hdiag, const double* __restrict__ r,
                                                            double* __restrict__ p, int n, const PcgScalars* sc) {
  if (sc->done) return;
  const double beta = sc->beta;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] = fma(beta, p[i], r[i] * (1.0 / hdiag[i]));
}

void dev_spmv(spb_context* h, const double* d_x, double* d_y) {
  if (h->n == 0) return;
  stage_begin(h, ST_SPMV);
  spmv_sell_kernel<false><<<(h->n + 255) / 256, 256, 0, h->stream>>>(h->d_slice_off.p, h->d_sell_col.p, h->d_rowlen.p,
                                                                      h->d_hval.p, d_x, d_y, h->n, nullptr, nullptr);
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_SPMV, 1);
}

spb_status pcg_solve(spb_context* h, const double* d_b, double* d_x, int* iters) {
  const int n = h->n;
  cudaStream_t st = h->stream;
  if (iters) *iters = 0;
  if (n == 0) return SPB_OK;
  ensure_pcg_buffers(h);
  SPB_CUDA(h->d_r.alloc(n));
  SPB_CUDA(h->d_p.alloc(n));
  SPB_CUDA(h->d_q.alloc(n));
  const int grid = (n + 255) / 256;
  PcgScalars* sc = h->d_sc.p;
  stage_begin(h, ST_PCGVEC);
  pcg_init_kernel<<<grid, 256, 0, st>>>(d_b, h->d_hdiag.p, d_x, h->d_r.p, h->d_p.p, n, h->d_partial.p, sc, h->pcg_rtol, h->pcg_maxit);
  stage_end(h, ST_PCGVEC, 1);
  int launched = 0;
  int batch = 16;
  for (;;) {
    for (int k = 0; k < batch; ++k) {
      stage_begin(h, ST_SPMV);
      spmv_sell_kernel<true><<<grid, 256, 0, st>>>(h->d_slice_off.p, h->d_sell_col.p, h->d_rowlen.p, h->d_hval.p, h->d_p.p,
                                                   h->d_q.p, n, h->d_partial.p, sc);
      stage_end(h, ST_SPMV, 1);
      stage_begin(h, ST_PCGVEC);
      pcg_update_kernel<<<grid, 256, 0, st>>>(h->d_hdiag.p, h->d_p.p, h->d_q.p, d_x, h->d_r.p, n, h->d_partial.p, sc);
      pcg_direction_kernel<<<grid, 256, 0, st>>>(h->d_hdiag.p, h->d_r.p, h->d_p.p, n, sc);
      stage_end(h, ST_PCGVEC, 2);
    }
    launched += batch;
    SPB_CUDA(cudaGetLastError());
    SPB_CUDA(cudaMemcpyAsync(h->h_sc, sc, sizeof(PcgScalars), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    if (h->h_sc->done) break;
    if (launched >= h->pcg_maxit + batch) break;  // cannot happen: done is set at maxit
  }
  if (iters) *iters = h->h_sc->iters;
  if (h->h_sc->breakdown) return SPB_NOT_SPD;
  if (h->h_sc->rr > h->h_sc->thresh) return SPB_NOT_CONVERGED;
  return SPB_OK;
}

}  // namespace spb
