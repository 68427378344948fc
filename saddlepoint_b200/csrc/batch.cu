// batch.cu -- config C4: MANY independent small problems advanced together (sm_100a).
//
// The reference solves one problem per LMSolver::solve call (LMSolver.h:86-191); a batch of B problems is B host
// loops.  Here the batch is ONE block-diagonal problem -- problem p owns unknowns [p*n_per, (p+1)*n_per) and
// residual rows [p*m_per, (p+1)*m_per), the Jacobian is block diagonal -- and every stage of the LM iteration is
// one launch over all problems:
//   * one fused assembly (assembly.cu) with a per-problem lambda;
//   * ONE persistent cooperative PCG kernel for all problems (single-reduction CG as in pcg_peer.cu) whose dot
//     products are segmented by problem: phase B leaves three partial sums per 32-row slice (two sets when a slice
//     straddles a problem boundary), phase R turns them into every problem's alpha / beta / stop decision (one warp
//     per problem, fixed summation order), phase A updates the rows of the problems that are still iterating.  Slices
//     whose problems have converged are skipped, so the stream shrinks as problems finish;
//   * a device-side LM controller (one thread per problem) that holds lambda, the iteration counter and the exit
//     path of every problem and applies LMSolver.h's tests problem by problem: first-order optimality (:123-129, only
//     when verbose: quirk A), failed factorisation = non-positive curvature (:136-140), direction small (:147-152),
//     acceptance prev > new (:161), function tolerance (:164-168), the lambda rule of DiagonalDamping.h:65-68 and the
//     iteration bound (:188, quirk C).  Finished problems are frozen (masked) while the others go on; the host only
//     reads one integer (problems still active) per LM iteration.
// Per-problem results (iteration counts, exit paths, energies, x) equal those of B separate spb_lm_solve calls with
// dedupe_evaluations (tests/test_gpu_batch.py compares them with the CPU oracle).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

#include <cooperative_groups.h>
#include <math_constants.h>

#include "spb_internal.h"
#include "pcg_device.cuh"

namespace cg = cooperative_groups;

namespace spb {

// per-problem PCG scalars (struct of arrays on the device)
struct BatchPcgState {
  double* gam_old;
  double* alpha;
  double* beta;
  double* bb;
  double* rr;
  double* denom;
  int* iters;
  int* status;  // 0 iterating, 1 converged, 2 not SPD, 3 NaN system, 4 maxit, 5 masked out (problem not active)
};

struct BatchPcgParams {
  const int64_t* slice_off;
  const int* col;
  const uint16_t* col16;
  const uint8_t* s16;
  const int* rowlen;
  const double* hval;
  const double* b;
  const double* dinv;
  // exact 2x2 block preconditioner (block2 != 0; H block diagonal over aligned pairs, n_per even): M = the pair's block
  // [hdiag_i, hoff; hoff, hdiag_i^1], M^-1 = [dinv_i, doff; doff, dinv_i^1]
  const double* doff;
  const double* hoff;
  const double* hdiag;
  int block2;
  double* x;
  double* u;
  double* p;
  double* s;
  double* w;
  double* part;  // [nslices][2][3]
  BatchPcgState S;
  const int* active;  // per problem: 1 = solve it, 0 = leave x = 0 and skip
  int* counter;       // [2]: problems still iterating (double-buffered by iteration parity)
  int n, nslices, nprob, n_per, maxit;
  double rtol;
};

__device__ __forceinline__ int prob_of_row(int row, int n_per) { return row / n_per; }

// One persistent cooperative launch solves H_p x_p = b_p for every active problem p of the block-diagonal system.
__global__ void __launch_bounds__(256, 4) batch_pcg_kernel(BatchPcgParams P) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, tid = threadIdx.x, blk = blockIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int gw = blk * 8 + warp, GW = G * 8;
  const int n = P.n, n_per = P.n_per;
  // init: x = 0, u = dinv b, p = s = 0; gamma and rr partials per slice
  for (int slice = gw; slice < P.nslices; slice += GW) {
    const int row = slice * 32 + lane;
    const int p0 = prob_of_row(slice * 32, n_per);
    const int split = (p0 + 1) * n_per - slice * 32;  // lanes >= split belong to problem p0 + 1
    double g0 = 0.0, r0 = 0.0;
    const bool in_range = row < n;
    const bool act = in_range && P.active[min(lane < split ? p0 : p0 + 1, P.nprob - 1)] != 0;
    const double bi = act ? P.b[row] : 0.0;
    // block2: u = M^-1 b with the pair's inverse block; the partner is the neighbouring lane (same problem: n_per is even),
    // so every lane of the warp takes part in the exchange
    const double bo = P.block2 ? __shfl_xor_sync(0xffffffffu, bi, 1) : 0.0;
    if (in_range) {
      const double di = P.dinv[row];
      double ui = act ? bi * di : 0.0;
      if (P.block2) ui = act ? fma(P.doff[row], bo, bi * di) : 0.0;
      P.x[row] = 0.0;
      P.u[row] = ui;
      P.p[row] = 0.0;
      P.s[row] = 0.0;
      g0 = bi * ui;
      r0 = bi * bi;
    }
    double* out = P.part + (size_t)slice * 6;
    const double ga = warp_sum(lane < split ? g0 : 0.0), ra = warp_sum(lane < split ? r0 : 0.0);
    const double gb = split < 32 ? warp_sum(lane >= split ? g0 : 0.0) : 0.0, rb = split < 32 ? warp_sum(lane >= split ? r0 : 0.0) : 0.0;
    if (lane == 0) {
      out[0] = ga;
      out[2] = ra;
      out[3] = gb;
      out[5] = rb;
    }
  }
  if (blk == 0 && tid < 2) P.counter[tid] = 0;
  grid.sync();
  for (int it = 0;; ++it) {
    // ---- phase B: w = H u on the slices of problems that are still iterating; (w,u) per slice --------------------
    for (int slice = gw; slice < P.nslices; slice += GW) {
      const int p0 = prob_of_row(slice * 32, n_per);
      const int split = (p0 + 1) * n_per - slice * 32;
      const bool a0 = P.S.status[p0] == 0 || it == 0;  // before the first reduction every active problem takes part
      const bool a1 = split < 32 && p0 + 1 < P.nprob && (P.S.status[p0 + 1] == 0 || it == 0);
      if (!a0 && !a1) continue;
      const double acc = (P.col16 && P.s16[slice])
                             ? spmv_slice_acc<false, false, true>(P.slice_off, P.col, P.rowlen, P.hval, P.u, n, slice, lane, 0, P.col16)
                             : spmv_slice_acc<false, false, false>(P.slice_off, P.col, P.rowlen, P.hval, P.u, n, slice, lane, 0, P.col16);
      const int row = slice * 32 + lane;
      double d0 = 0.0;
      if (row < n) {
        P.w[row] = acc;
        d0 = acc * P.u[row];
      }
      double* out = P.part + (size_t)slice * 6;
      const double da = warp_sum(lane < split ? d0 : 0.0);
      const double db = split < 32 ? warp_sum(lane >= split ? d0 : 0.0) : 0.0;
      if (lane == 0) {
        out[1] = da;
        out[4] = db;
      }
    }
    grid.sync();
    // ---- phase R: one warp per problem: totals in slice order -> alpha, beta, stop decision ------------------------
    for (int pr = gw; pr < P.nprob; pr += GW) {
      int st = it == 0 ? (P.active[pr] ? 0 : 5) : P.S.status[pr];
      if (st != 0) {
        if (it == 0 && lane == 0) {
          P.S.status[pr] = st;
          P.S.iters[pr] = 0;
          P.S.bb[pr] = 0.0;
          P.S.rr[pr] = 0.0;
        }
        continue;
      }
      const int r_lo = pr * n_per, r_hi = min(n, (pr + 1) * n_per);
      const int s_lo = r_lo >> 5, s_hi = (r_hi - 1) >> 5;
      double g = 0.0, d = 0.0, r = 0.0;
      for (int sl = s_lo + lane; sl <= s_hi; sl += 32) {
        const int k = (prob_of_row(sl * 32, n_per) == pr) ? 0 : 3;
        const double* in = P.part + (size_t)sl * 6 + k;
        g += in[0];
        d += in[1];
        r += in[2];
      }
      g = warp_sum(g);
      d = warp_sum(d);
      r = warp_sum(r);
      if (lane == 0) {
        int iters = it == 0 ? 0 : P.S.iters[pr];
        double bb = it == 0 ? r : P.S.bb[pr];
        const double thresh = P.rtol * P.rtol * bb;
        double beta = 0.0, denom = d, alpha = 0.0;
        if (it == 0) {
          P.S.bb[pr] = bb;
          if (bb != bb) st = 3;
          else if (!(bb > thresh) || P.maxit <= 0) st = 1;
        } else {
          if (!(r > thresh)) st = 1;
          else if (iters >= P.maxit) st = 4;
          else {
            beta = g / P.S.gam_old[pr];
            denom = d - beta * g / P.S.alpha[pr];
          }
        }
        if (st == 0 && !(denom > 0.0)) st = (denom != denom) ? 3 : 2;
        if (st == 0) {
          alpha = g / denom;
          P.S.gam_old[pr] = g;
          P.S.alpha[pr] = alpha;
          P.S.beta[pr] = beta;
          P.S.iters[pr] = iters + 1;
          atomicAdd(&P.counter[it & 1], 1);
        }
        P.S.rr[pr] = r;
        P.S.denom[pr] = denom;
        P.S.status[pr] = st;
      }
    }
    grid.sync();
    const int still = *((volatile int*)&P.counter[it & 1]);
    if (still == 0) break;
    if (blk == 0 && tid == 0) P.counter[(it + 1) & 1] = 0;
    // ---- phase A: vector updates on the rows of the problems that iterate; gamma and rr partials per slice ----------
    for (int slice = gw; slice < P.nslices; slice += GW) {
      const int p0 = prob_of_row(slice * 32, n_per);
      const int split = (p0 + 1) * n_per - slice * 32;
      const int row = slice * 32 + lane;
      const int pr = lane < split ? p0 : p0 + 1;
      const bool live = row < n && pr < P.nprob && P.S.status[pr] == 0;
      if (!__any_sync(0xffffffffu, live)) continue;
      double g0 = 0.0, r0 = 0.0;
      if (P.block2) {
        // u <- u - alpha M^-1 s and r = M u with the pair's blocks; the partner values come from the neighbouring lane
        // (pairs never straddle a slice or, n_per being even, a problem), so every lane of the warp takes part
        double alpha = 0.0, beta = 0.0, ui = 0.0, wi = 0.0, pi = 0.0, si = 0.0, xi = 0.0, di = 0.0, dof = 0.0, hd = 0.0, hof = 0.0;
        if (live) {
          alpha = P.S.alpha[pr];
          beta = P.S.beta[pr];
          ui = P.u[row];
          wi = P.w[row];
          pi = P.p[row];
          si = P.s[row];
          xi = P.x[row];
          di = P.dinv[row];
          dof = P.doff[row];
          hd = P.hdiag[row];
          hof = P.hoff[row];
        }
        const double pn = fma(beta, pi, ui);
        const double sn = fma(beta, si, wi);
        const double so = __shfl_xor_sync(0xffffffffu, sn, 1);
        const double un = ui - alpha * fma(dof, so, di * sn);
        const double uo = __shfl_xor_sync(0xffffffffu, un, 1);
        const double rn = fma(hof, uo, hd * un);
        if (live) {
          P.p[row] = pn;
          P.s[row] = sn;
          P.x[row] = fma(alpha, pn, xi);
          P.u[row] = un;
          g0 = rn * un;
          r0 = rn * rn;
        }
      } else if (live) {
        const double alpha = P.S.alpha[pr], beta = P.S.beta[pr];
        const double ui = P.u[row], wi = P.w[row], pi = P.p[row], si = P.s[row], xi = P.x[row], di = P.dinv[row];
        const double pn = fma(beta, pi, ui);
        const double sn = fma(beta, si, wi);
        const double un = fma(-alpha * di, sn, ui);
        const double rn = un / di;
        P.p[row] = pn;
        P.s[row] = sn;
        P.x[row] = fma(alpha, pn, xi);
        P.u[row] = un;
        g0 = rn * un;
        r0 = rn * rn;
      }
      double* out = P.part + (size_t)slice * 6;
      const double ga = warp_sum(lane < split ? g0 : 0.0), ra = warp_sum(lane < split ? r0 : 0.0);
      const double gb = split < 32 ? warp_sum(lane >= split ? g0 : 0.0) : 0.0, rb = split < 32 ? warp_sum(lane >= split ? r0 : 0.0) : 0.0;
      if (lane == 0) {  // a finished neighbour problem's partials in the same slice are never read again
        out[0] = ga;
        out[2] = ra;
        out[3] = gb;
        out[5] = rb;
      }
    }
    grid.sync();
  }
  // NaN systems hand back a NaN step
  for (int slice = gw; slice < P.nslices; slice += GW) {
    const int row = slice * 32 + lane;
    if (row < n && P.S.status[prob_of_row(row, n_per)] == 3) P.x[row] = __longlong_as_double(0x7ff8000000000000LL);
  }
}

// ---- segmented reductions: one block per problem, fixed order ----------------------------------------------------------
// op 0: sum of squares, op 1: max |v| (NaN never wins)
template <int OP>
__global__ void __launch_bounds__(256) seg_reduce_kernel(const double* __restrict__ v, int seg_len, int64_t total, double* __restrict__ out) {
  __shared__ double sm[8];
  const int64_t lo = (int64_t)blockIdx.x * seg_len;
  const int64_t hi = lo + seg_len < total ? lo + seg_len : total;
  double a = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) {
    const double x = v[i];
    if (OP == 0) a = fma(x, x, a);
    else a = fmax(a, fabs(x));
  }
  if (OP == 0) {
    const double s = block_sum_256(a, sm);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
  } else {
    a = warp_max(a);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      double m = 0.0;
      for (int w = 0; w < 8; ++w) m = fmax(m, sm[w]);
      out[blockIdx.x] = m;
    }
  }
}

// ---- the LM controller: one thread per problem --------------------------------------------------------------------------
struct BatchLmState {
  double* lambda;
  double* cachedE;  // energy at prevx (LMSolver.h:155 as cached by dedupe_evaluations)
  double* prevE2;
  double* newE2;
  double* foo;      // ||rhs||_inf of the problem's last own assembly (frozen with the problem)
  double* foo_tmp;  // this iteration's value for every problem
  double* dn2;
  double* energy;
  int* active;    // 1 while the problem is in its loop
  int* accepted;  // this iteration's acceptance (for the masked copies)
  int* exit_path;
  int* ret;
  int* currIter;
  int* factorizations;
  int64_t* lin_iters;
  int* unconverged;
  int* nactive;  // [1]
};
struct BatchLmOpts {
  double funcTolerance, fooTolerance;
  int maxIterations, verbose;
};

// after the assembly: first-order optimality (only with verbose: quirk A), count the factorisation
__global__ void batch_ctl_foo_kernel(BatchLmState S, BatchLmOpts O, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob || !S.active[p]) return;
  S.foo[p] = S.foo_tmp[p];
  if (S.foo[p] < O.fooTolerance && O.verbose) {  // LMSolver.h:123-129 (x = prevx already)
    S.exit_path[p] = 1;
    S.ret[p] = 0;
    S.active[p] = 0;
    return;
  }
  S.factorizations[p]++;
}
// after the solve: failed factorisation analogue, direction small
__global__ void batch_ctl_dir_kernel(BatchLmState S, BatchLmOpts O, int nprob, const int* pcg_status, const int* pcg_iters) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob || !S.active[p]) return;
  S.lin_iters[p] += pcg_iters[p];
  const int st = pcg_status[p];
  if (st == 4) S.unconverged[p]++;
  if (st == 2) {  // LMSolver.h:136-140
    S.exit_path[p] = 2;
    S.ret[p] = 0;
    S.active[p] = 0;
    return;
  }
  const double dn = sqrt(S.dn2[p]);
  if (dn < O.funcTolerance) {  // LMSolver.h:147-152 (x = prevx already)
    S.exit_path[p] = 3;
    S.ret[p] = 1;
    S.active[p] = 0;
    return;
  }
  S.prevE2[p] = S.cachedE[p];
}
// after the trial evaluation: acceptance, function tolerance, lambda rule, iteration bound
__global__ void batch_ctl_accept_kernel(BatchLmState S, BatchLmOpts O, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  S.accepted[p] = 0;
  if (!S.active[p]) return;
  const double pe = S.prevE2[p], ne = S.newE2[p];
  S.energy[p] = ne;  // LMSolver.h:159
  const bool acc = pe > ne;
  S.accepted[p] = acc ? 1 : 0;
  if (acc && fabs(pe - ne) < O.funcTolerance) {  // LMSolver.h:164-168: break, solve() returns false (quirk B)
    S.exit_path[p] = 4;
    S.ret[p] = 0;
    S.active[p] = 2;  // x <- trial is still to be copied; frozen afterwards (batch_ctl_finish_kernel)
    return;
  }
  const double e = acc ? ne : pe;  // LMSolver.h:174-175 after re-linearising at x
  S.energy[p] = e;
  S.cachedE[p] = e;
  // DiagonalDamping.h:65-68
  if ((pe > ne) && (ne != CUDART_INF)) S.lambda[p] /= 10.0;
  else S.lambda[p] *= 10.0;
  const int ci = ++S.currIter[p];
  if (ci > O.maxIterations) {  // LMSolver.h:188 (quirk C); x <- accepted trial is still to be copied
    S.exit_path[p] = 6;
    S.ret[p] = 0;
    S.active[p] = 2;
  }
}
__global__ void batch_ctl_finish_kernel(BatchLmState S, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) *S.nactive = 0;
  if (p >= nprob) return;
  if (S.active[p] == 2) S.active[p] = 0;
}
__global__ void batch_ctl_count_kernel(BatchLmState S, int nprob) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nprob && S.active[p]) atomicAdd(S.nactive, 1);
}
// trial = prevx + dir on active problems (else prevx)
__global__ void batch_trial_kernel(const double* __restrict__ prevx, const double* __restrict__ dir, double* __restrict__ trial, const int* __restrict__ active,
                                   int n_per, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    trial[i] = active[i / n_per] ? prevx[i] + dir[i] : prevx[i];
}
// x = prevx = (accepted ? trial : prevx) on active problems; frozen problems keep x
__global__ void batch_commit_kernel(double* __restrict__ prevx, double* __restrict__ x, const double* __restrict__ trial, const int* __restrict__ active,
                                    const int* __restrict__ accepted, int n_per, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / n_per);
    if (active[p] && accepted[p]) {
      const double t = trial[i];
      x[i] = t;
      prevx[i] = t;
    }
  }
}
__global__ void batch_invert_kernel(const double* __restrict__ a, double* __restrict__ out, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = 1.0 / a[i];
}
// Exact inverse of the 2x2 blocks of a block-diagonal H (h_block2): per row the diagonal and the off-diagonal entry of
// the inverse block, and the block's own off-diagonal entry.  A pair whose block is not positive definite (or not
// finite), or that has no coupling entry, gets the plain Jacobi data (doff = hoff = 0), identically for both rows.
__global__ void batch_invert2_kernel(const double* __restrict__ hdiag, const double* __restrict__ hval, const int64_t* __restrict__ slice_off,
                                     const int* __restrict__ rowlen, int n, double* __restrict__ dinv, double* __restrict__ doff,
                                     double* __restrict__ hoff) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double a = hdiag[i];
    const int o = i ^ 1;
    double ia = 1.0 / a, ib = 0.0, b = 0.0;
    if (o < n && rowlen[i] == 2) {
      // row i holds {i, i+1} (i even: coupling entry second) or {i-1, i} (i odd: coupling entry first)
      const double bb = hval[slice_off[i >> 5] + (int64_t)((i & 1) ? 0 : 1) * 32 + (i & 31)];
      const double c = hdiag[o];
      const double det = a * c - bb * bb;
      if (det > 0.0 && det < 1.79e308 && a > 0.0 && c > 0.0) {
        ia = c / det;
        ib = -bb / det;
        b = bb;
      }
    }
    dinv[i] = ia;
    doff[i] = ib;
    hoff[i] = b;
  }
}
__global__ void batch_fill_kernel(double* p, double v, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

struct BatchWorkspace {
  DevBuf<double> d_prevx, d_x, d_dir, d_trial, d_E, d_Etrial, d_Jv, d_part, d_doff, d_hoff;
  DevBuf<double> d_dbl;  // per-problem doubles: 6 PCG + 8 LM arrays
  DevBuf<int> d_int;     // per-problem ints
  DevBuf<int64_t> d_i64;
  double* hx = nullptr;  // pinned staging for host traits
  double* hE = nullptr;
  double* hJ = nullptr;
  size_t hx_n = 0, hE_n = 0, hJ_n = 0;
  ~BatchWorkspace() {
    if (hx) cudaFreeHost(hx);
    if (hE) cudaFreeHost(hE);
    if (hJ) cudaFreeHost(hJ);
  }
};
void batch_release(spb_context* h) {
  delete (BatchWorkspace*)h->batch_ws;
  h->batch_ws = nullptr;
}

}  // namespace spb

using namespace spb;

extern "C" spb_status spb_lm_solve_batch(spb_handle h, const spb_traits* T, int32_t nproblems, const spb_lm_options* O, spb_lm_result* results,
                                         double* x_out, int x_mem) {
  if (!h || !T || !O || !results || nproblems < 1 || !T->objective_jacobian || !T->initial_solution) return SPB_ERR_ARG;
  try {
    cudaGetLastError();
    SPB_CUDA(cudaSetDevice(h->device));
    auto t_start = std::chrono::steady_clock::now();
    const int n = T->xSize, m = T->ESize, B = nproblems;
    if (n % B != 0 || m % B != 0) throw ArgFail{"batch: xSize and ESize must be multiples of the number of problems (uniform problem sizes)"};
    const int n_per = n / B, m_per = m / B;
    if (n_per < 32) throw ArgFail{"batch: a problem needs at least 32 unknowns"};
    if (h->halo_mode || h->partitioned || h->nranks > 1) throw StateFail{"batch mode runs on a plain single-GPU handle (shard the problems across ranks instead)"};
    if (O->fixed_iterations > 0) throw ArgFail{"batch: fixed_iterations is not supported"};
    // block structure: every entry of a column lies in its problem's residual rows
    for (int j = 0; j < n; ++j) {
      const int p = j / n_per;
      const int q0 = T->colptr[j], q1 = T->colptr[j + 1];
      if (q1 > q0 && (T->rowidx[q0] / m_per != p || T->rowidx[q1 - 1] / m_per != p)) throw ArgFail{"batch: the Jacobian is not block diagonal over the problems"};
    }
    double analyze_s = 0.0;
    {
      int same = 0;
      spb_pattern_matches(h, m, n, T->colptr, T->rowidx, &same);
      if (!same) {
        auto ta = std::chrono::steady_clock::now();
        spb_status s = spb_analyze(h, m, n, T->colptr, T->rowidx);
        if (s != SPB_OK) return s;
        analyze_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - ta).count();
      }
    }
    if (!h->batch_ws) h->batch_ws = new BatchWorkspace;
    BatchWorkspace& W = *(BatchWorkspace*)h->batch_ws;
    cudaStream_t st = h->stream;
    const int64_t nnzJ = h->nnzJ;
    const int nslices = (n + 31) / 32;
    SPB_CUDA(W.d_prevx.alloc(n));
    SPB_CUDA(W.d_x.alloc(n));
    SPB_CUDA(W.d_dir.alloc(n));
    SPB_CUDA(W.d_trial.alloc(n));
    SPB_CUDA(W.d_E.alloc(m));
    SPB_CUDA(W.d_Etrial.alloc(m));
    SPB_CUDA(W.d_Jv.alloc((size_t)nnzJ));
    SPB_CUDA(W.d_part.alloc((size_t)nslices * 6));
    SPB_CUDA(W.d_dbl.alloc((size_t)B * 14));
    SPB_CUDA(W.d_int.alloc((size_t)B * 9 + 4));
    SPB_CUDA(W.d_i64.alloc((size_t)B));
    SPB_CUDA(cudaMemsetAsync(W.d_dbl.p, 0, sizeof(double) * (size_t)B * 14, st));
    SPB_CUDA(cudaMemsetAsync(W.d_int.p, 0, sizeof(int) * ((size_t)B * 9 + 4), st));
    SPB_CUDA(cudaMemsetAsync(W.d_i64.p, 0, sizeof(int64_t) * (size_t)B, st));
    SPB_CUDA(cudaMemsetAsync(W.d_part.p, 0, sizeof(double) * (size_t)nslices * 6, st));
    SPB_CUDA(h->d_u.alloc(n));
    SPB_CUDA(h->d_w.alloc(n));
    SPB_CUDA(h->d_p.alloc(n));
    SPB_CUDA(h->d_q.alloc(n));
    double* D = W.d_dbl.p;
    int* I = W.d_int.p;
    BatchPcgState PS{D, D + B, D + 2 * B, D + 3 * B, D + 4 * B, D + 5 * B, I, I + B};
    BatchLmState LS{D + 6 * B, D + 7 * B, D + 8 * B, D + 9 * B, D + 10 * B, D + 13 * B, D + 11 * B, D + 12 * B, I + 2 * B, I + 3 * B, I + 4 * B,
                    I + 5 * B, I + 6 * B, I + 7 * B, W.d_i64.p, I + 8 * B, I + 9 * B + 2};
    int* pcg_counter = I + 9 * B;  // [2]
    BatchLmOpts LO{O->funcTolerance, O->fooTolerance, O->maxIterations, O->verbose};
    const int gp = (B + 255) / 256;
    const int gv = std::max(1, std::min((n + 255) / 256, h->num_sms * 8));
    // traits evaluation (device traits in place; host traits through pinned staging)
    auto eval = [&](const double* d_x, double* d_E, double* d_Jv) {
      if (T->mem == SPB_DEVICE) {
        stage_begin(h, ST_TRAITS);
        T->objective_jacobian(T->user, d_x, d_E, d_Jv, (void*)st);
        stage_end(h, ST_TRAITS, 0);
        return;
      }
      auto need = [](double*& p, size_t& have, size_t want) {
        if (want <= have && p) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        SPB_CUDA(cudaMallocHost((void**)&p, std::max<size_t>(want, 1) * sizeof(double)));
        have = want;
      };
      need(W.hx, W.hx_n, (size_t)n);
      need(W.hE, W.hE_n, (size_t)m);
      need(W.hJ, W.hJ_n, (size_t)nnzJ);
      SPB_CUDA(cudaMemcpyAsync(W.hx, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
      SPB_CUDA(cudaStreamSynchronize(st));
      T->objective_jacobian(T->user, W.hx, W.hE, d_Jv ? W.hJ : nullptr, nullptr);
      SPB_CUDA(cudaMemcpyAsync(d_E, W.hE, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
      if (d_Jv) SPB_CUDA(cudaMemcpyAsync(d_Jv, W.hJ, (size_t)nnzJ * sizeof(double), cudaMemcpyHostToDevice, st));
      SPB_CUDA(cudaStreamSynchronize(st));
    };
    {
      std::vector<double> x0((size_t)n);
      T->initial_solution(T->user, x0.data());
      SPB_CUDA(cudaMemcpyAsync(W.d_prevx.p, x0.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
      SPB_CUDA(cudaMemcpyAsync(W.d_x.p, W.d_prevx.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      SPB_CUDA(cudaStreamSynchronize(st));
    }
    batch_fill_kernel<<<gp, 256, 0, st>>>(LS.lambda, O->lambda0, B);
    {
      std::vector<int> ones((size_t)B, 1);
      SPB_CUDA(cudaMemcpyAsync(LS.active, ones.data(), sizeof(int) * (size_t)B, cudaMemcpyHostToDevice, st));
      SPB_CUDA(cudaStreamSynchronize(st));
    }
    eval(W.d_prevx.p, W.d_E.p, W.d_Jv.p);  // LMSolver.h:106
    seg_reduce_kernel<0><<<B, 256, 0, st>>>(W.d_E.p, m_per, (int64_t)m, LS.cachedE);  // the energy the first iteration compares against
    h->launches += 2;
    // cooperative grid of the batched PCG
    static int occ_cached = 0;
    if (occ_cached == 0) {
      int occ = 0;
      SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, batch_pcg_kernel, 256, 0));
      occ_cached = std::max(1, occ);
    }
    const int G = std::max(1, std::min((std::max(nslices, B) + 7) / 8, h->num_sms * occ_cached));
    int* h_nactive = (int*)h->h_scalar;  // pinned
    int lm_bodies = 0;
    for (;;) {
      // fused assembly with per-problem lambda (LMSolver.h:117-119,136 + DiagonalDamping rebuild)
      h->d_lambda_vec = LS.lambda;
      h->lambda_n_per = n_per;
      run_assembly(h, W.d_Jv.p, W.d_E.p, 0.0, nullptr);
      h->d_lambda_vec = nullptr;
      stage_begin(h, ST_REDUCE);
      seg_reduce_kernel<1><<<B, 256, 0, st>>>(h->d_rhs.p, n_per, (int64_t)n, LS.foo_tmp);
      batch_ctl_foo_kernel<<<gp, 256, 0, st>>>(LS, LO, B);
      stage_end(h, ST_REDUCE, 2);
      SPB_CUDA(h->d_dinv.alloc(n));
      // block-diagonal H over aligned pairs (extended Rosenbrock batches): the exact inverse blocks precondition, the PCG
      // becomes a direct block solve with one or two refinement iterations instead of ~30 Jacobi iterations on blocks
      // whose condition number is 1e8 far from the optimum
      const bool block2_off = getenv("SPB_BATCH_BLOCK2") && atoi(getenv("SPB_BATCH_BLOCK2")) == 0;  // read per body: the tests compare both
      const bool block2 = h->h_block2 && (n_per % 2 == 0) && !block2_off;
      stage_begin(h, ST_PCGVEC);
      if (block2) {
        SPB_CUDA(W.d_doff.alloc(n));
        SPB_CUDA(W.d_hoff.alloc(n));
        batch_invert2_kernel<<<gv, 256, 0, st>>>(h->d_hdiag.p, h->d_hval.p, h->d_slice_off.p, h->d_rowlen.p, n, h->d_dinv.p, W.d_doff.p, W.d_hoff.p);
      } else {
        batch_invert_kernel<<<gv, 256, 0, st>>>(h->d_hdiag.p, h->d_dinv.p, n);
      }
      stage_end(h, ST_PCGVEC, 1);
      h->launches += 3;
      {
        BatchPcgParams P;
        P.slice_off = h->d_slice_off.p;
        P.col = h->d_sell_col.p;
        P.col16 = h->pcg_col16 ? h->d_sell_col16.p : nullptr;
        P.s16 = h->d_slice16.p;
        P.rowlen = h->d_rowlen.p;
        P.hval = h->d_hval.p;
        P.b = h->d_rhs.p;
        P.dinv = h->d_dinv.p;
        P.doff = block2 ? W.d_doff.p : nullptr;
        P.hoff = block2 ? W.d_hoff.p : nullptr;
        P.hdiag = h->d_hdiag.p;
        P.block2 = block2 ? 1 : 0;
        P.x = W.d_dir.p;
        P.u = h->d_u.p;
        P.p = h->d_p.p;
        P.s = h->d_q.p;
        P.w = h->d_w.p;
        P.part = W.d_part.p;
        P.S = PS;
        P.active = LS.active;
        P.counter = pcg_counter;
        P.n = n;
        P.nslices = nslices;
        P.nprob = B;
        P.n_per = n_per;
        P.maxit = h->pcg_maxit;
        P.rtol = h->pcg_rtol;
        void* args[] = {&P};
        stage_begin(h, ST_PCG);
        SPB_CUDA(cudaLaunchCooperativeKernel((void*)batch_pcg_kernel, dim3(G), dim3(256), args, 0, st));
        stage_end(h, ST_PCG, 1);
      }
      stage_begin(h, ST_REDUCE);
      seg_reduce_kernel<0><<<B, 256, 0, st>>>(W.d_dir.p, n_per, (int64_t)n, LS.dn2);
      batch_ctl_dir_kernel<<<gp, 256, 0, st>>>(LS, LO, B, PS.status, PS.iters);
      stage_end(h, ST_REDUCE, 2);
      stage_begin(h, ST_PCGVEC);
      batch_trial_kernel<<<gv, 256, 0, st>>>(W.d_prevx.p, W.d_dir.p, W.d_trial.p, LS.active, n_per, (int64_t)n);
      stage_end(h, ST_PCGVEC, 1);
      h->launches += 3;
      eval(W.d_trial.p, W.d_Etrial.p, nullptr);  // LMSolver.h:156
      stage_begin(h, ST_REDUCE);
      seg_reduce_kernel<0><<<B, 256, 0, st>>>(W.d_Etrial.p, m_per, (int64_t)m, LS.newE2);
      batch_ctl_accept_kernel<<<gp, 256, 0, st>>>(LS, LO, B);
      stage_end(h, ST_REDUCE, 2);
      stage_begin(h, ST_PCGVEC);
      batch_commit_kernel<<<gv, 256, 0, st>>>(W.d_prevx.p, W.d_x.p, W.d_trial.p, LS.active, LS.accepted, n_per, (int64_t)n);
      stage_end(h, ST_PCGVEC, 1);
      stage_begin(h, ST_REDUCE);
      batch_ctl_finish_kernel<<<gp, 256, 0, st>>>(LS, B);
      batch_ctl_count_kernel<<<gp, 256, 0, st>>>(LS, B);
      stage_end(h, ST_REDUCE, 2);
      h->launches += 5;
      SPB_CUDA(cudaMemcpyAsync(h_nactive, LS.nactive, sizeof(int), cudaMemcpyDeviceToHost, st));
      SPB_CUDA(cudaStreamSynchronize(st));
      ++lm_bodies;
      if (*h_nactive == 0) break;
      // LMSolver.h:174: re-linearise at x (= prevx after the commit); frozen problems re-evaluate to the same values
      eval(W.d_prevx.p, W.d_E.p, W.d_Jv.p);
      if (lm_bodies > O->maxIterations + 2) break;  // cannot happen: the controller stops every problem at the bound
    }
    // results
    std::vector<double> hd((size_t)B * 14);
    std::vector<int> hi((size_t)B * 9 + 4);
    std::vector<int64_t> hl((size_t)B);
    SPB_CUDA(cudaMemcpyAsync(hd.data(), D, sizeof(double) * hd.size(), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaMemcpyAsync(hi.data(), I, sizeof(int) * hi.size(), cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaMemcpyAsync(hl.data(), W.d_i64.p, sizeof(int64_t) * hl.size(), cudaMemcpyDeviceToHost, st));
    if (x_out)
      SPB_CUDA(cudaMemcpyAsync(x_out, W.d_x.p, (size_t)n * sizeof(double), x_mem == SPB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    SPB_CUDA(cudaStreamSynchronize(st));
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    for (int p = 0; p < B; ++p) {
      spb_lm_result& R = results[p];
      memset(&R, 0, sizeof(R));
      R.ret = hi[(size_t)5 * B + p];
      R.exit_path = hi[(size_t)4 * B + p];
      R.currIter = hi[(size_t)6 * B + p];
      R.factorizations = hi[(size_t)7 * B + p];
      R.energy = hd[(size_t)12 * B + p];
      R.fooOptimality = hd[(size_t)10 * B + p];
      R.lambda = hd[(size_t)6 * B + p];
      R.linear_iterations = hl[p];
      R.unconverged_solves = hi[(size_t)8 * B + p];
      R.seconds_total = secs;
      R.seconds_analyze = analyze_s;
    }
    h->h_valid = false;  // the handle's H belongs to the batch's last iteration; not a state callers should build on
    return SPB_OK;
  } catch (const CudaFail& e) {
    h->d_lambda_vec = nullptr;
    h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    return SPB_ERR_CUDA;
  } catch (const ArgFail& e) {
    h->d_lambda_vec = nullptr;
    h->err = e.msg;
    return SPB_ERR_ARG;
  } catch (const StateFail& e) {
    h->d_lambda_vec = nullptr;
    h->err = e.msg;
    return SPB_ERR_STATE;
  } catch (const std::exception& e) {
    h->d_lambda_vec = nullptr;
    h->err = std::string("exception inside the batched LM loop: ") + e.what();
    return SPB_ERR_STATE;
  } catch (...) {
    h->d_lambda_vec = nullptr;
    h->err = "unknown exception inside the batched LM loop";
    return SPB_ERR_STATE;
  }
}
