// cholesky.cu -- numeric phase of the supernodal (multifrontal) LL^T and its triangular
// solves, sm_100a.  Replaces Eigen::SimplicialLLT::compute / solve behind
// EigenSolverWrapper::factorize / solve (EigenSolverWrapper.h:54-78):
//   factorize  -> SPB_NOT_SPD iff some pivot d <= 0 (a NaN pivot passes)
//   solve      -> x = P^T L^-T L^-1 P b
//
// Structure (symbolic phase in chol_symbolic.cpp): every supernode s owns a dense panel
// P_s (r x k column-major: k pivot columns, r = k + |struct(s)| rows) that becomes the
// factor, and a temporary update matrix U_s (u x u, u = r - k).  The assembly tree is
// processed level by level (leaves first), all fronts of a level in the same launches:
//   1. scatter the H values into the panels (one launch for the whole matrix);
//   2. extend-add the children's update matrices into the level's fronts, one launch per
//      child rank so the summation order is fixed (deterministic) and conflict-free;
//   3. blocked right-looking factorisation of the pivot columns, block size 32.  A block step is
//        chol_step_kernel   few CTAs in the group (top separators, latency regime): every CTA factors the
//                           32x32 diagonal block itself in one warp and solves its 128 rows -- one launch;
//                           an extra CTA stores the inverted block for the solves
//        chol_potrf32_warp_kernel + chol_trsm_kernel   thousands of fronts (leaf levels, throughput regime)
//        chol_syrk_kernel   trailing pivot columns of the 256-column panel (K=32) and, once per panel,
//                           everything right of it (K=256); FP64 tensor cores, mma.sync m8n8k4 DMMA
//   4. U_s -= L21 L21^T in one K=k SYRK (DMMA), the bulk of the flops.
// Only the SYRKs are tensor-core work; scatter/extend-add/solves are bandwidth or latency kernels
// (north_star item 3).  tcgen05 has no FP64 path, so DMMA is the FP64 tensor instruction on sm_100a.
// Update matrices live in one arena packed by lifetime (chol_symbolic.cpp), cleared per level.
// Solves: level-ordered.  Groups of many small fronts: one CTA per front, work vector in shared
// memory, children's tails gathered in a fixed order.  Groups with few and/or large fronts: one
// cooperative launch per group runs all its block steps separated by grid barriers (128-pivot tiles
// staged in shared memory); diagonal blocks are applied through their precomputed inverses.
// Debug stopwatches: SPB_CHOL_TIMING=1 (per group), 2 (inside the cooperative forward kernel),
// 3 (fused step kernel), 4 (narrow SYRK).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>

#include "chol_internal.h"
#include "tma.cuh"

namespace cg = cooperative_groups;

namespace spb {

struct CholDevice {
  CholSymbolic sym;
  DevBuf<int> sn_k, sn_r, col0, struct_idx, rel_idx, parent, child_ptr, child_idx, level_sn, rank_children, perm, ea_ptr, ea_idx, small_parents;
  DevBuf<int64_t> L_off, U_off, W_off, struct_ptr, rel_ptr, dinv_off, a_src, a_dst;
  DevBuf<double> Lbuf, Ubuf, Wbuf, Dinv, ywork;
  DevBuf<int> fail;
  DevBuf<long long> dbg;
  // per level: groups of fronts of similar size -> ranges in level_sn (fronts sorted by r inside a level)
  struct Group {
    int begin, count, max_k, max_r;
  };
  std::vector<std::vector<Group>> groups;
  bool factored = false;
  cudaGraphExec_t graph_exec = nullptr;  // the captured launch sequence of one factorisation
  const double* graph_hval = nullptr;    // ... valid for this H value array
  int graph_nl = 0;
  int coop_occ = 0;            // co-resident CTAs per SM of the cooperative solve kernels
  bool smem_attr_set = false;  // per device context, so per handle (not a process-wide static)
  int* h_fail = nullptr;
};

// ---- small device helpers --------------------------------------------------------------
__device__ __forceinline__ unsigned long long chol_global_ns();
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void chol_scatter_a_kernel(const double* __restrict__ hval, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                      int64_t count, double* __restrict__ Lbuf) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) Lbuf[dst[t]] = hval[src[t]];
}

// children[blockIdx.y]: add its update matrix into the parent's panel / update matrix
__global__ void __launch_bounds__(256) chol_extend_add_kernel(const int* __restrict__ children, const int* __restrict__ parent,
                                                              const int* __restrict__ sn_k, const int* __restrict__ sn_r,
                                                              const int64_t* __restrict__ L_off, const int64_t* __restrict__ U_off,
                                                              const int64_t* __restrict__ rel_ptr, const int* __restrict__ rel_idx,
                                                              double* __restrict__ Lbuf, double* __restrict__ Ubuf) {
  const int c = children[blockIdx.y];
  const int p = parent[c];
  const int u = sn_r[c] - sn_k[c];
  const int pk = sn_k[p], pr = sn_r[p], pu = pr - pk;
  const double* Uc = Ubuf + U_off[c];
  double* Pp = Lbuf + L_off[p];
  double* Up = Ubuf + U_off[p];
  const int* rel = rel_idx + rel_ptr[c];
  for (int b = blockIdx.x; b < u; b += gridDim.x) {
    const int lb = rel[b];
    for (int a = b + threadIdx.x; a < u; a += 256) {
      const int la = rel[a];
      const double v = Uc[(int64_t)b * u + a];
      if (lb < pk) Pp[(int64_t)lb * pr + la] += v;
      else Up[(int64_t)(lb - pk) * pu + (la - pk)] += v;
    }
  }
}

// The small children of a parent (update matrices of at most kSmallChildU rows), added by ONE warp per parent in the
// parent's extend-add order, starting at child `first` of its list: same sums in the same order as the rank-wise launches
// would form, without a launch per rank (a parent of the condensed seamless-integration structure has ~100 of them).
__global__ void __launch_bounds__(128) chol_extend_add_small_kernel(const int* __restrict__ parents, int count, int first,
                                                                    const int* __restrict__ ea_ptr, const int* __restrict__ ea_idx,
                                                                    const int* __restrict__ sn_k, const int* __restrict__ sn_r,
                                                                    const int64_t* __restrict__ L_off, const int64_t* __restrict__ U_off,
                                                                    const int64_t* __restrict__ rel_ptr, const int* __restrict__ rel_idx,
                                                                    double* __restrict__ Lbuf, double* __restrict__ Ubuf) {
  const int lane = threadIdx.x & 31;
  const int pi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (pi >= count) return;
  const int p = parents[pi];
  const int pk = sn_k[p], pr = sn_r[p], pu = pr - pk;
  double* Pp = Lbuf + L_off[p];
  double* Up = Ubuf + U_off[p];
  const int c_end = ea_ptr[p + 1];
  for (int ci = ea_ptr[p] + first; ci < c_end; ++ci) {
    const int c = ea_idx[ci];
    const int u = sn_r[c] - sn_k[c];
    const double* Uc = Ubuf + U_off[c];
    const int* rel = rel_idx + rel_ptr[c];
    for (int e = lane; e < u * u; e += 32) {
      const int b = e / u, a = e - b * u;
      if (a < b) continue;
      const int lb = rel[b], la = rel[a];
      const double v = Uc[(int64_t)b * u + a];
      if (lb < pk) Pp[(int64_t)lb * pr + la] += v;
      else Up[(int64_t)(lb - pk) * pu + (la - pk)] += v;
    }
    __syncwarp();  // the next child may add to the same parent entries
  }
}

// A whole front of at most 16 pivots and 32 rows, factored by ONE warp in one launch: Cholesky of the pivot block, the
// rows below it, the update matrix U -= L21 L21^T and the inverted pivot block for the solves.  Lane i keeps row i of the
// r x k panel in registers; columns are eliminated right-looking with the pivot column broadcast by shuffles.  Same
// arithmetic as the block kernels (l_jj = d rsqrt(d), l_ij = a_ij rsqrt(d); failure iff a pivot <= 0, NaN passes).
// With static condensation the first level of the seamless-integration structure is half a million such fronts
// (k = 8, r = 14): three launches of 128-thread CTAs per 32 768 of them before, one launch of warps now.
__global__ void __launch_bounds__(128) chol_tiny_front_kernel(const int* __restrict__ fronts, int count, const int* __restrict__ sn_k,
                                                              const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                              const int64_t* __restrict__ U_off, const int64_t* __restrict__ dinv_off,
                                                              double* __restrict__ Lbuf, double* __restrict__ Ubuf,
                                                              double* __restrict__ Dinv, int* __restrict__ fail) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int fi = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (fi >= count) return;
  const int f = fronts[fi];
  const int k = sn_k[f], r = sn_r[f], u = r - k;
  double* P = Lbuf + L_off[f];
  double a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = (c < k && lane < r && lane >= c) ? P[(int64_t)c * r + lane] : 0.0;
  bool bad = false;
  double myr = 1.0;  // 1 / l_jj of this lane's pivot
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j < k) {  // warp-uniform
      const double d = __shfl_sync(kFull, a[j], j);
      bad = bad || (d <= 0.0);  // NaN passes, like the reference's `if (d <= 0)`
      const double rs = rsqrt(d);
      if (lane == j) {
        a[j] = d * rs;
        myr = rs;
      } else if (lane > j) {
        a[j] = a[j] * rs;
      }
#pragma unroll
      for (int c = j + 1; c < 16; ++c) {
        if (c < k) {
          const double lcj = __shfl_sync(kFull, a[j], c);
          if (lane >= c) a[c] -= a[j] * lcj;
        }
      }
    }
  }
  if (lane == 0 && bad) *fail = 1;
  if (lane < r) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < k && lane >= c) P[(int64_t)c * r + lane] = a[c];
  }
  // update matrix: U(a, b) -= sum_c L(k+a, c) L(k+b, c), lower triangle, column b at a time
  if (u > 0) {
    double* U = Ubuf + U_off[f];
    for (int b = 0; b < u; ++b) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (c < k) s = fma(a[c], __shfl_sync(kFull, a[c], k + b), s);
      const int ai = lane - k;
      if (lane < r && ai >= b) U[(int64_t)b * u + ai] -= s;
    }
  }
  // inverse of the pivot block (identity-padded to 32 x 32, column-major): lane = column, forward substitution
  double x[16];
#pragma unroll
  for (int rw = 0; rw < 16; ++rw) {
    x[rw] = 0.0;
    if (rw < k) {
      double acc = (rw == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int t = 0; t < rw; ++t) acc -= __shfl_sync(kFull, a[t], rw) * x[t];  // x[t] == 0 for t < lane
      const double rd = __shfl_sync(kFull, myr, rw);
      x[rw] = rw >= lane ? acc * rd : 0.0;
    }
  }
  // Only the k x k corner of the front's 32 x 32 block is stored: the tiny solve kernels read nothing else, and the
  // identity padding of half a million fronts was 4 GB of writes per factorisation (8 KB per front for 512 useful bytes).
  // A front is in ONE group per level, and a tiny group is factored and solved by the tiny kernels only.
  if (lane < k) {
    double* out = Dinv + dinv_off[f] + (int64_t)lane * 32;
#pragma unroll
    for (int rw = 0; rw < 16; ++rw)
      if (rw < k) out[rw] = x[rw];
  }
}

// Cholesky of the 32x32 diagonal block at pivot offset kb of every front in the list, plus its
// inverse (for the triangular solves).  1024 threads = one per block entry.  (A single-warp
// register/shuffle variant was measured slower: the leaf levels run tens of thousands of these
// blocks at once and want all 1024 threads per block.)
__global__ void __launch_bounds__(1024) chol_potrf32_kernel(const int* __restrict__ fronts, int kb, const int* __restrict__ sn_k,
                                                            const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                            const int64_t* __restrict__ dinv_off, double* __restrict__ Lbuf,
                                                            double* __restrict__ Dinv, int* __restrict__ fail) {
  __shared__ double S[32][33];
  const int f = fronts[blockIdx.x];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  const int i = threadIdx.x & 31, c = threadIdx.x >> 5;  // row, column
  double* D = Lbuf + L_off[f] + (int64_t)kb * r + kb;
  S[i][c] = (i < nb && c < nb && i >= c) ? D[(int64_t)c * r + i] : (i == c ? 1.0 : 0.0);
  for (int j = 0; j < nb; ++j) {
    __syncthreads();
    if (i == j && c == j) {
      const double d = S[j][j];
      if (d <= 0.0) *fail = 1;  // NaN passes, like the reference's `if (d <= 0)`
      S[j][j] = sqrt(d);
    }
    __syncthreads();
    if (c == j && i > j) S[i][j] = S[i][j] / S[j][j];
    __syncthreads();
    if (c > j && i >= c) S[i][c] -= S[i][j] * S[c][j];
  }
  __syncthreads();
  if (i < nb && c < nb && i >= c) D[(int64_t)c * r + i] = S[i][c];
  // inverse: warp c computes column c of S^-1 by forward substitution, lane t holds x_t
  double x = 0.0;
  for (int row = c; row < 32; ++row) {
    double part = (i >= c && i < row) ? S[row][i] * x : 0.0;
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (i == row) x = ((row == c ? 1.0 : 0.0) - part) / S[row][row];
  }
  Dinv[dinv_off[f] + (int64_t)(kb >> 5) * 1024 + c * 32 + i] = (i >= c) ? x : 0.0;
}

// rows below the diagonal block: X = B * D^-T, one thread per row, column-oriented substitution
// (after x_q is final every later entry takes its update independently: ILP instead of one
// 496-long dependent chain; the accumulation order per entry is unchanged)
__global__ void __launch_bounds__(128) chol_trsm_kernel(const int* __restrict__ fronts, int kb, const int* __restrict__ sn_k,
                                                        const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                        double* __restrict__ Lbuf) {
  __shared__ double Dm[32][33];
  const int f = fronts[blockIdx.y];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  const int rb = kb + nb;
  if ((int64_t)blockIdx.x * 128 >= r - rb) return;
  double* P = Lbuf + L_off[f];
  for (int e = threadIdx.x; e < 1024; e += 128) {
    const int i = e & 31, c = e >> 5;
    Dm[i][c] = (i < nb && c < nb && i >= c) ? P[(int64_t)(kb + c) * r + kb + i] : (i == c ? 1.0 : 0.0);
  }
  __syncthreads();
  const int row = rb + blockIdx.x * 128 + threadIdx.x;
  if (row >= r) return;
  double v[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) v[t] = t < nb ? P[(int64_t)(kb + t) * r + row] : 0.0;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const double xq = v[q] / Dm[q][q];
    v[q] = xq;
#pragma unroll
    for (int t = q + 1; t < 32; ++t) v[t] -= xq * Dm[t][q];
  }
#pragma unroll
  for (int t = 0; t < 32; ++t)
    if (t < nb) P[(int64_t)(kb + t) * r + row] = v[t];
}

// ---- fused block step: Cholesky of the 32x32 diagonal block + the rows below it -------------
// One launch per 32-column step instead of two (potrf32 -> trsm): every CTA of a front loads the
// diagonal block and factors it itself (warp 0, one lane per row, ~4 us; the other warps fetch
// their rows of the panel meanwhile), then solves its 128 rows against it.  All CTAs compute the
// same bits.  The factored block is NOT written back in place -- other CTAs of the launch are
// still reading the unfactored one -- and nothing downstream reads it: the solves apply the
// diagonal blocks through their inverses, which the launch's extra CTA (blockIdx.x == gridDim.x-1)
// computes and stores together with the failure flag.
__device__ __forceinline__ void warp_potrf32(double (*S)[33], int nb, int lane, int* fail_flag, double* rdiag) {
  // lane i keeps row i of the block in registers; columns are eliminated right-looking.  The
  // finished column is published through shared memory (rdiag doubles as the buffer) and read back
  // with broadcast 16-byte loads: a single warp issues shuffles far too slowly for the 496 column
  // entries the updates need.  Rows/columns >= nb are identity padding.
  // One rsqrt per column instead of sqrt + divide (both long software sequences in FP64 on the
  // critical path): l_jj = d * rsqrt(d), l_ij = a_ij * rsqrt(d); rdiag[j] = 1/l_jj afterwards for
  // the row solves, which multiply instead of dividing.
  (void)nb;
  __shared__ __align__(16) double colbuf[4][32];  // one per warp of the CTA
  double* cb = colbuf[(threadIdx.x >> 5) & 3];
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = S[lane][c];
  bool bad = false;
  double myr = 0.0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    bad = bad || (d <= 0.0);  // NaN passes, like the reference's `if (d <= 0)`
    const double rs = rsqrt(d);
    if (lane == j) {
      a[j] = d * rs;
      myr = rs;
    } else if (lane > j) {
      a[j] = a[j] * rs;
    }
    cb[lane] = a[j];
    __syncwarp();
#pragma unroll
    for (int c = (j + 1) & ~1; c < 32; c += 2) {
      const double2 l2 = *reinterpret_cast<const double2*>(cb + c);
      if (c > j && lane >= c) a[c] -= a[j] * l2.x;
      if (lane >= c + 1) a[c + 1] -= a[j] * l2.y;
    }
    __syncwarp();
  }
  if (lane == 0 && bad) *fail_flag = 1;
  rdiag[lane] = myr;
#pragma unroll
  for (int c = 0; c < 32; ++c) S[lane][c] = a[c];
}

// Column `lane` of S^-1 (S lower triangular in shared memory) by forward substitution; the factor
// entries are broadcast reads, the 32 solution entries live in registers.  out: 32x32 column-major.
__device__ __forceinline__ void warp_inverse32(double (*S)[33], const double* rdiag, int lane, double* __restrict__ out) {
  double x[32];
#pragma unroll
  for (int rw = 0; rw < 32; ++rw) {
    double acc = (rw == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int t = 0; t < rw; ++t) acc -= S[rw][t] * x[t];  // x[t] == 0 for t < lane
    x[rw] = rw >= lane ? acc * rdiag[rw] : 0.0;
  }
#pragma unroll
  for (int rw = 0; rw < 32; ++rw) out[lane * 32 + rw] = x[rw];
}

// Throughput variant for levels with many fronts: one WARP per front factors the 32x32 diagonal
// block at pivot offset kb in registers, writes it back and stores its inverse.
__global__ void __launch_bounds__(128) chol_potrf32_warp_kernel(const int* __restrict__ fronts, int count, int kb, const int* __restrict__ sn_k,
                                                                const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                                const int64_t* __restrict__ dinv_off, double* __restrict__ Lbuf,
                                                                double* __restrict__ Dinv, int* __restrict__ fail) {
  __shared__ double Sm[4][32][33];
  __shared__ double Rd[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = blockIdx.x * 4 + warp;
  if (fi >= count) return;
  const int f = fronts[fi];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  double* D = Lbuf + L_off[f] + (int64_t)kb * r + kb;
  double(*S)[33] = Sm[warp];
  for (int c = 0; c < 32; ++c) S[lane][c] = (lane < nb && c < nb && lane >= c) ? D[(int64_t)c * r + lane] : (lane == c ? 1.0 : 0.0);
  __syncwarp();
  int bad = 0;
  warp_potrf32(S, nb, lane, &bad, Rd[warp]);
  if (lane == 0 && bad) *fail = 1;
  __syncwarp();
  for (int c = 0; c < nb; ++c)
    if (lane < nb && lane >= c) D[(int64_t)c * r + lane] = S[lane][c];
  warp_inverse32(S, Rd[warp], lane, Dinv + dinv_off[f] + (int64_t)(kb >> 5) * 1024);
}

__global__ void __launch_bounds__(128) chol_step_kernel(const int* __restrict__ fronts, int kb, const int* __restrict__ sn_k,
                                                        const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                        const int64_t* __restrict__ dinv_off, double* __restrict__ Lbuf,
                                                        double* __restrict__ Dinv, int* __restrict__ fail, long long* dbg) {
  __shared__ double Dm[32][33];
  __shared__ double Rd[32];
  __shared__ int sfail;
  const bool dbg_on = dbg && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
  if (dbg_on) {
    dbg[0] = clock64();
    dbg[5] = (long long)chol_global_ns();
  }
  const int f = fronts[blockIdx.y];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  const int rb = kb + nb;
  const bool inv_cta = blockIdx.x == gridDim.x - 1;
  if (!inv_cta && (int64_t)blockIdx.x * 128 >= r - rb) return;
  double* P = Lbuf + L_off[f];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) sfail = 0;
  // all global loads first (the diagonal block and this thread's row of the panel): one memory round
  // trip; the row stays in flight while warp 0 factors the block
  const int row = rb + blockIdx.x * 128 + tid;
  const bool have_row = !inv_cta && row < r;
  double dl[8], v[32];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = warp + 4 * u;
    dl[u] = (lane < nb && c < nb && lane >= c) ? P[(int64_t)(kb + c) * r + kb + lane] : (lane == c ? 1.0 : 0.0);
  }
#pragma unroll
  for (int t = 0; t < 32; ++t) v[t] = (have_row && t < nb) ? P[(int64_t)(kb + t) * r + row] : 0.0;
  asm volatile("" ::: "memory");
#pragma unroll
  for (int u = 0; u < 8; ++u) Dm[lane][warp + 4 * u] = dl[u];
  __syncthreads();
  if (dbg_on) dbg[1] = clock64();
  if (warp == 0) warp_potrf32(Dm, nb, lane, &sfail, Rd);
  __syncthreads();
  if (dbg_on) dbg[2] = clock64();
  if (inv_cta) {
    if (tid == 0 && sfail) *fail = 1;
    // inverse: warp w computes columns w, w+4, ... of Dm^-1 by forward substitution, lane t holds x_t
    if (warp == 0) warp_inverse32(Dm, Rd, lane, Dinv + dinv_off[f] + (int64_t)(kb >> 5) * 1024);
    return;
  }
  if (!have_row) return;
#pragma unroll
  for (int q = 0; q < 32; ++q) {
    const double xq = v[q] * Rd[q];
    v[q] = xq;
#pragma unroll
    for (int t = q + 1; t < 32; ++t) v[t] -= xq * Dm[t][q];
  }
  if (dbg_on) dbg[3] = clock64();
#pragma unroll
  for (int t = 0; t < 32; ++t)
    if (t < nb) P[(int64_t)(kb + t) * r + row] = v[t];
  if (dbg_on) {
    dbg[4] = clock64();
    dbg[6] = (long long)chol_global_ns();
  }
}

// C -= A_i * A_j^T on 64x64 tiles of the lower triangle, FP64 tensor cores (DMMA m8n8k4).
// mode 0: pivot columns [ke, min(cend, k)) are updated with the factor columns [ka, ke) -- used
//         twice: inside a 256-column panel (ka = kb, ke = kb+32, cend = panel end; K = 32) and once
//         per panel for everything right of it (ka = panel start, ke = panel end, cend = inf; K = 256).
//         Two-level blocking keeps the narrow K=32 updates inside the panel, so the bulk of the
//         trailing matrix is read and written once per 256 columns, not once per 32.
// mode 1: the update matrix U (K = k).
constexpr int kSyrkLd = 68;  // 64 + 4: conflict-free fragment loads
constexpr int kSyrkKC = 16;  // K columns staged per pipeline stage
constexpr int kSyrkSeg = 66; // doubles per bulk copy: a 64-row segment plus the alignment element, even length
template <bool TMA>
__global__ void __launch_bounds__(128) chol_syrk_kernel(const int* __restrict__ fronts, int mode, int ka, int ke, int cend,
                                                        const int* __restrict__ sn_k,
                                                        const int* __restrict__ sn_r, const int64_t* __restrict__ L_off,
                                                        const int64_t* __restrict__ U_off, double* __restrict__ Lbuf,
                                                        double* __restrict__ Ubuf, long long* dbg) {
  __shared__ __align__(16) double As[2 * kSyrkKC * kSyrkLd], Bs[2 * kSyrkKC * kSyrkLd];
  __shared__ uint64_t bar[2];
  const bool dbg_on = dbg && blockIdx.x == 0 && blockIdx.y == gridDim.y - 1 && blockIdx.z == 0 && threadIdx.x == 0;
  if (dbg_on) dbg[0] = clock64();
  const int f = fronts[blockIdx.z];
  const int k = sn_k[f], r = sn_r[f];
  double* P = Lbuf + L_off[f];
  const double* A;
  double* C;
  int64_t ldc;
  int nrows, ncols, K;
  if (mode == 0) {
    if (k <= ke) return;  // no pivot column right of the block
    A = P + (int64_t)ka * r + ke;
    C = P + (int64_t)ke * r + ke;
    ldc = r;
    nrows = r - ke;
    ncols = min(cend, k) - ke;
    K = ke - ka;
  } else {
    const int u = r - k;
    if (u == 0) return;
    A = P + k;
    C = Ubuf + U_off[f];
    ldc = u;
    nrows = ncols = u;
    K = k;
  }
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (ti < tj || ti * 64 >= nrows || tj * 64 >= ncols) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wy = warp >> 1, wx = warp & 1;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  const int lr = lane >> 2, lk = lane & 3;
  if constexpr (TMA) {
    // Two-stage pipeline fed by the TMA unit: every K column of a 64-row operand tile is one contiguous 512-byte
    // segment of the panel, brought in by ONE bulk asynchronous copy (cp.async.bulk, SASS UBLKCP) that completes on the
    // stage's mbarrier -- 32 copy instructions per stage instead of 2 048 eight-byte cp.async (the staging instructions
    // were competing with the DMMA stream for issue slots).  Panel leading dimensions are odd in general, so a segment
    // is only 8-byte aligned: the copy starts at the 16-byte boundary below it (one element early when the segment's
    // address is odd) and moves 66 doubles; the fragment reads add that shift back.  (k0 and kk are even, so the shift
    // of K column k0 + kk + lk depends on lk only: one constant per thread and operand.)  Rows past the end of a column
    // come in as the next column's data and only reach C entries that the store masks; K columns past the end are
    // zero-filled.  The diagonal tile (ti == tj) stages one operand.
    const bool diag = ti == tj;
    const double* Arow = A + ti * 64;
    const double* Brow = A + tj * 64;
    auto odd = [](const double* p) { return (int)((reinterpret_cast<uintptr_t>(p) >> 3) & 1u); };
    const int sha = odd(Arow + (int64_t)lk * r), shb = odd(Brow + (int64_t)lk * r);
    if (tid == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
    }
    __syncthreads();
    auto stage_load = [&](int st, int k0) {
      double* as = As + st * (kSyrkKC * kSyrkLd);
      double* bs = Bs + st * (kSyrkKC * kSyrkLd);
      const int kc = min(kSyrkKC, K - k0);  // K columns of this stage
      if (tid == 0) mbar_expect_tx(&bar[st], (uint32_t)(kc * kSyrkSeg * 8 * (diag ? 1 : 2)));
      if (tid < 2 * kSyrkKC) {
        const int t = tid & (kSyrkKC - 1), which = tid / kSyrkKC;
        if (t < kc && !(which && diag)) {
          const double* src = (which ? Brow : Arow) + (int64_t)(k0 + t) * r;
          tma_bulk_g2s((which ? bs : as) + t * kSyrkLd, src - odd(src), kSyrkSeg * 8, &bar[st]);
        }
      }
      if (kc < kSyrkKC)  // last stage of a K that is not a multiple of 16: the missing columns multiply as zeros
        for (int e = tid; e < (kSyrkKC - kc) * kSyrkLd; e += 128) as[kc * kSyrkLd + e] = bs[kc * kSyrkLd + e] = 0.0;
    };
    const int nchunks = (K + kSyrkKC - 1) / kSyrkKC;
    if (dbg_on) dbg[1] = clock64();
    stage_load(0, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int st = c & 1;
      if (c + 1 < nchunks) stage_load(st ^ 1, (c + 1) * kSyrkKC);  // the other buffer was released by the barrier closing c-1
      mbar_wait(&bar[st], (uint32_t)((c >> 1) & 1));
      if (c + 1 == nchunks) __syncthreads();  // the zero-filled columns of the last stage
      const double* as = As + st * (kSyrkKC * kSyrkLd) + sha;
      const double* bs = diag ? As + st * (kSyrkKC * kSyrkLd) + shb : Bs + st * (kSyrkKC * kSyrkLd) + shb;
  #pragma unroll
      for (int kk = 0; kk < kSyrkKC; kk += 4) {
        double a[4], b[4];
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = as[(kk + lk) * kSyrkLd + wy * 32 + q * 8 + lr];
          b[q] = bs[(kk + lk) * kSyrkLd + wx * 32 + q * 8 + lr];
        }
  #pragma unroll
        for (int p = 0; p < 4; ++p)
  #pragma unroll
          for (int q = 0; q < 4; ++q) dmma_m8n8k4(acc[p][q][0], acc[p][q][1], a[p], b[q]);
      }
      __syncthreads();
    }
  } else {
    // Two-stage software pipeline: the next K-chunk of both row blocks is copied global -> shared
    // asynchronously (cp.async, zero-filled outside the matrix) while the tensor cores work on
    // the current one.
    auto stage_load = [&](int st, int k0) {
      double* as = As + st * (kSyrkKC * kSyrkLd);
      double* bs = Bs + st * (kSyrkKC * kSyrkLd);
      for (int e = tid; e < kSyrkKC * 64; e += 128) {
        const int t = e >> 6, i = e & 63;
        const int gi = ti * 64 + i, gj = tj * 64 + i;
        const bool kin = (k0 + t) < K;
        const bool ia = kin && gi < nrows, ib = kin && gj < nrows;
        const double* sa = ia ? A + (int64_t)(k0 + t) * r + gi : A;
        const double* sb = ib ? A + (int64_t)(k0 + t) * r + gj : A;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(as + t * kSyrkLd + i)), "l"(sa),
                     "r"(ia ? 8 : 0));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(bs + t * kSyrkLd + i)), "l"(sb),
                     "r"(ib ? 8 : 0));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int nchunks = (K + kSyrkKC - 1) / kSyrkKC;
    if (dbg_on) dbg[1] = clock64();
    stage_load(0, 0);
    for (int c = 0; c < nchunks; ++c) {
      const int st = c & 1;
      if (c + 1 < nchunks) {
        stage_load(st ^ 1, (c + 1) * kSyrkKC);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const double* as = As + st * (kSyrkKC * kSyrkLd);
      const double* bs = Bs + st * (kSyrkKC * kSyrkLd);
  #pragma unroll
      for (int kk = 0; kk < kSyrkKC; kk += 4) {
        double a[4], b[4];
  #pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = as[(kk + lk) * kSyrkLd + wy * 32 + q * 8 + lr];
          b[q] = bs[(kk + lk) * kSyrkLd + wx * 32 + q * 8 + lr];
        }
  #pragma unroll
        for (int p = 0; p < 4; ++p)
  #pragma unroll
          for (int q = 0; q < 4; ++q) dmma_m8n8k4(acc[p][q][0], acc[p][q][1], a[p], b[q]);
      }
      __syncthreads();
    }
  }
  if (dbg_on) dbg[2] = clock64();
  // read-modify-write of the C tile in two passes (all 32 loads of a thread in flight, then the
  // stores): written as `C[..] -= acc` the compiler must assume aliasing and serialises 32 memory
  // round trips, which cost more than the K=32 main loop
  double cv[4][4][2];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = ti * 64 + wy * 32 + p * 8 + lr;
        const int j = tj * 64 + wx * 32 + q * 8 + 2 * lk + e;
        cv[p][q][e] = (i < nrows && j < ncols && i >= j) ? C[(int64_t)j * ldc + i] : 0.0;
      }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = ti * 64 + wy * 32 + p * 8 + lr;
        const int j = tj * 64 + wx * 32 + q * 8 + 2 * lk + e;
        if (i < nrows && j < ncols && i >= j) C[(int64_t)j * ldc + i] = cv[p][q][e] - acc[p][q][e];
      }
  if (dbg_on) dbg[3] = clock64();
}

// ---- solves --------------------------------------------------------------------------------
// forward: one CTA per front.  w = [y_pivots ; 0] + sum of the children's tails (fixed order),
// pivots solved block by block through the inverted diagonal blocks, tail -= L21 * y_pivots.
__global__ void __launch_bounds__(256) chol_forward_kernel(const int* __restrict__ fronts, const int* __restrict__ sn_k,
                                                           const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                           const int64_t* __restrict__ L_off, const int64_t* __restrict__ W_off,
                                                           const int64_t* __restrict__ dinv_off, const int* __restrict__ child_ptr,
                                                           const int* __restrict__ child_idx, const int64_t* __restrict__ rel_ptr,
                                                           const int* __restrict__ rel_idx, const double* __restrict__ Lbuf,
                                                           const double* __restrict__ Dinv, double* __restrict__ Wbuf, double* __restrict__ y) {
  extern __shared__ double w[];  // r + 32
  const int f = fronts[blockIdx.x];
  const int k = sn_k[f], r = sn_r[f], c0 = col0[f];
  const int tid = threadIdx.x;
  double* tmp = w + r;
  for (int i = tid; i < r; i += 256) w[i] = i < k ? y[c0 + i] : 0.0;
  __syncthreads();
  for (int ci = child_ptr[f]; ci < child_ptr[f + 1]; ++ci) {
    const int c = child_idx[ci];
    const int ck = sn_k[c], cu = sn_r[c] - ck;
    const double* tail = Wbuf + W_off[c] + ck;
    const int* rel = rel_idx + rel_ptr[c];
    for (int t = tid; t < cu; t += 256) w[rel[t]] += tail[t];
    __syncthreads();
  }
  const double* P = Lbuf + L_off[f];
  const double* Di = Dinv + dinv_off[f];
  for (int kb = 0; kb < k; kb += 32) {
    const int nb = min(32, k - kb);
    // y_blk = Dinv * w_blk  (lower-triangular 32x32, column-major)
    if (tid < 32) {
      double s = 0.0;
      for (int t = 0; t < nb; ++t) s += Di[(int64_t)(kb >> 5) * 1024 + t * 32 + tid] * w[kb + t];
      tmp[tid] = s;
    }
    __syncthreads();
    if (tid < nb) w[kb + tid] = tmp[tid];
    // rows below the block (remaining pivots and the tail): w_i -= sum_t L[i, kb+t] * y_t
    for (int i = kb + nb + tid; i < r; i += 256) {
      double s = w[i];
      for (int t = 0; t < nb; ++t) s -= P[(int64_t)(kb + t) * r + i] * tmp[t];
      w[i] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += 256) y[c0 + i] = w[i];
  double* Wf = Wbuf + W_off[f];
  for (int i = k + tid; i < r; i += 256) Wf[i] = w[i];
}

// backward: x_piv = L11^-T (y_piv - L21^T x_struct); ancestors are already final
__global__ void __launch_bounds__(256) chol_backward_kernel(const int* __restrict__ fronts, const int* __restrict__ sn_k,
                                                            const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                            const int64_t* __restrict__ L_off, const int64_t* __restrict__ dinv_off,
                                                            const int64_t* __restrict__ struct_ptr, const int* __restrict__ struct_idx,
                                                            const double* __restrict__ Lbuf, const double* __restrict__ Dinv,
                                                            double* __restrict__ y) {
  extern __shared__ double w[];  // r + 32
  const int f = fronts[blockIdx.x];
  const int k = sn_k[f], r = sn_r[f], c0 = col0[f];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* tmp = w + r;
  const int* st = struct_idx + struct_ptr[f];
  for (int i = tid; i < r; i += 256) w[i] = i < k ? y[c0 + i] : y[st[i - k]];
  __syncthreads();
  const double* P = Lbuf + L_off[f];
  const double* Di = Dinv + dinv_off[f];
  const int nblk = (k + 31) / 32;
  for (int bi = nblk - 1; bi >= 0; --bi) {
    const int kb = bi * 32;
    const int nb = min(32, k - kb);
    // z_t = w[kb+t] - sum_{i >= kb+nb} L[i, kb+t] * w[i] : one warp per column, strided
    for (int t = warp; t < nb; t += 8) {
      double s = 0.0;
      const double* col = P + (int64_t)(kb + t) * r;
      for (int i = kb + nb + lane; i < r; i += 32) s += col[i] * w[i];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) tmp[t] = w[kb + t] - s;
    }
    __syncthreads();
    // x_blk = Dinv^T * z_blk
    if (tid < nb) {
      double s = 0.0;
      for (int t = tid; t < nb; ++t) s += Di[(int64_t)bi * 1024 + tid * 32 + t] * tmp[t];
      w[kb + tid] = s;
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += 256) y[c0 + i] = w[i];
}

// Tiny fronts (k <= 16 pivots, r <= 32 rows): one WARP per front, the work vector in 32 doubles of shared memory; the same
// sums in the same order as chol_forward_kernel / chol_backward_kernel (a CTA of 256 threads per front was mostly idle
// threads and barriers: the condensed seamless-integration structure has half a million such fronts in its first level).
__global__ void __launch_bounds__(128) chol_forward_tiny_kernel(const int* __restrict__ fronts, int count, const int* __restrict__ sn_k,
                                                                const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                                const int64_t* __restrict__ L_off, const int64_t* __restrict__ W_off,
                                                                const int64_t* __restrict__ dinv_off, const int* __restrict__ child_ptr,
                                                                const int* __restrict__ child_idx, const int64_t* __restrict__ rel_ptr,
                                                                const int* __restrict__ rel_idx, const double* __restrict__ Lbuf,
                                                                const double* __restrict__ Dinv, double* __restrict__ Wbuf,
                                                                double* __restrict__ y) {
  __shared__ double ws[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = blockIdx.x * 4 + warp;
  if (fi >= count) return;
  const int f = fronts[fi];
  const int k = sn_k[f], r = sn_r[f], c0 = col0[f];
  double* w = ws[warp];
  w[lane] = lane < k ? y[c0 + lane] : 0.0;
  __syncwarp();
  for (int ci = child_ptr[f]; ci < child_ptr[f + 1]; ++ci) {
    const int c = child_idx[ci];
    const int ck = sn_k[c], cu = sn_r[c] - ck;
    const double* tail = Wbuf + W_off[c] + ck;
    const int* rel = rel_idx + rel_ptr[c];
    for (int t = lane; t < cu; t += 32) w[rel[t]] += tail[t];
    __syncwarp();
  }
  const double* P = Lbuf + L_off[f];
  const double* Di = Dinv + dinv_off[f];
  double yi = 0.0;
  if (lane < k) {  // only the k x k corner of the inverted block exists for tiny fronts
    double s = 0.0;
    for (int t = 0; t < k; ++t) s += Di[t * 32 + lane] * w[t];
    yi = s;
  }
  __syncwarp();
  if (lane < k) w[lane] = yi;
  __syncwarp();
  if (lane >= k && lane < r) {
    double s = w[lane];
    for (int t = 0; t < k; ++t) s -= P[(int64_t)t * r + lane] * w[t];
    Wbuf[W_off[f] + lane] = s;
  } else if (lane < k) {
    y[c0 + lane] = yi;
  }
}

__global__ void __launch_bounds__(128) chol_backward_tiny_kernel(const int* __restrict__ fronts, int count, const int* __restrict__ sn_k,
                                                                 const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                                 const int64_t* __restrict__ L_off, const int64_t* __restrict__ dinv_off,
                                                                 const int64_t* __restrict__ struct_ptr, const int* __restrict__ struct_idx,
                                                                 const double* __restrict__ Lbuf, const double* __restrict__ Dinv,
                                                                 double* __restrict__ y) {
  __shared__ double zs[4][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = blockIdx.x * 4 + warp;
  if (fi >= count) return;
  const int f = fronts[fi];
  const int k = sn_k[f], r = sn_r[f], c0 = col0[f];
  const int* st = struct_idx + struct_ptr[f];
  const double wi = lane < k ? y[c0 + lane] : (lane < r ? y[st[lane - k]] : 0.0);
  const double* P = Lbuf + L_off[f];
  const double* Di = Dinv + dinv_off[f];
  double* z = zs[warp];
  // z_t = w[t] - sum_{i >= k} L[i, t] * w[i]: lanes hold the rows, the sum runs through the same xor tree as the CTA kernel
  const double wtail = __shfl_sync(0xffffffffu, wi, min(k + lane, 31));  // w[k + lane]
  for (int t = 0; t < k; ++t) {
    double s = (k + lane < r) ? P[(int64_t)t * r + k + lane] * wtail : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const double wt = __shfl_sync(0xffffffffu, wi, t);
    if (lane == 0) z[t] = wt - s;
  }
  __syncwarp();
  if (lane < k) {
    double s = 0.0;
    for (int t = lane; t < k; ++t) s += Di[lane * 32 + t] * z[t];
    y[c0 + lane] = s;
  }
}

// ---- big fronts, launch-per-step variant (fallback: SPB_CHOL_COOP=0) ---------------------------
// A front with thousands of rows cannot be swept by one CTA at memory speed; here the work vector
// lives in global memory (Wbuf) and every 32-column block step is a grid-wide launch: all CTAs
// recompute the 32-entry block solution through the inverted diagonal block, each updates its own
// 256 rows (forward) or 256 earlier pivots (backward).  Superseded by the cooperative kernels below
// (1030 -> 58 launches per solve at C2); kept for A/B measurements.
constexpr int kBigFront = 1536;

__global__ void __launch_bounds__(256) chol_fwd_gather_kernel(const int* __restrict__ fronts, const int* __restrict__ sn_k,
                                                              const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                              const int64_t* __restrict__ W_off, const int* __restrict__ child_ptr,
                                                              const int* __restrict__ child_idx, const int64_t* __restrict__ rel_ptr,
                                                              const int* __restrict__ rel_idx, double* __restrict__ Wbuf,
                                                              const double* __restrict__ y) {
  const int f = fronts[blockIdx.x];
  const int k = sn_k[f], r = sn_r[f], c0 = col0[f];
  double* w = Wbuf + W_off[f];
  for (int i = threadIdx.x; i < r; i += 256) w[i] = i < k ? y[c0 + i] : 0.0;
  __syncthreads();
  for (int ci = child_ptr[f]; ci < child_ptr[f + 1]; ++ci) {  // fixed child order: deterministic
    const int c = child_idx[ci];
    const int ck = sn_k[c], cu = sn_r[c] - ck;
    const double* tail = Wbuf + W_off[c] + ck;
    const int* rel = rel_idx + rel_ptr[c];
    for (int t = threadIdx.x; t < cu; t += 256) w[rel[t]] += tail[t];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) chol_fwd_step_kernel(const int* __restrict__ fronts, int kb, const int* __restrict__ sn_k,
                                                            const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                            const int64_t* __restrict__ L_off, const int64_t* __restrict__ W_off,
                                                            const int64_t* __restrict__ dinv_off, const double* __restrict__ Lbuf,
                                                            const double* __restrict__ Dinv, double* __restrict__ Wbuf, double* __restrict__ y) {
  __shared__ double yb[32];
  const int f = fronts[blockIdx.y];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  const int row0 = kb + nb + blockIdx.x * 256;
  if (row0 >= r && blockIdx.x > 0) return;
  double* w = Wbuf + W_off[f];
  const double* Di = Dinv + dinv_off[f] + (int64_t)(kb >> 5) * 1024;
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int t = 0; t < nb; ++t) s += Di[t * 32 + threadIdx.x] * w[kb + t];
    yb[threadIdx.x] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < nb) y[col0[f] + kb + threadIdx.x] = yb[threadIdx.x];
  const int i = row0 + threadIdx.x;
  if (i < r) {
    const double* P = Lbuf + L_off[f];
    double s = w[i];
    for (int t = 0; t < nb; ++t) s -= P[(int64_t)(kb + t) * r + i] * yb[t];
    w[i] = s;
  }
}

// w_piv = y_piv - L21^T x_struct : one warp per pivot column
__global__ void __launch_bounds__(256) chol_bwd_tail_kernel(const int* __restrict__ fronts, const int* __restrict__ sn_k,
                                                            const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                            const int64_t* __restrict__ L_off, const int64_t* __restrict__ W_off,
                                                            const int64_t* __restrict__ struct_ptr, const int* __restrict__ struct_idx,
                                                            const double* __restrict__ Lbuf, double* __restrict__ Wbuf,
                                                            const double* __restrict__ y) {
  const int f = fronts[blockIdx.y];
  const int k = sn_k[f], r = sn_r[f];
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (j >= k) return;
  const int* st = struct_idx + struct_ptr[f];
  const double* col = Lbuf + L_off[f] + (int64_t)j * r;
  double s = 0.0;
  for (int i = k + lane; i < r; i += 32) s += col[i] * y[st[i - k]];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) Wbuf[W_off[f] + j] = y[col0[f] + j] - s;
}

__global__ void __launch_bounds__(256) chol_bwd_step_kernel(const int* __restrict__ fronts, int kb, const int* __restrict__ sn_k,
                                                            const int* __restrict__ sn_r, const int* __restrict__ col0,
                                                            const int64_t* __restrict__ L_off, const int64_t* __restrict__ W_off,
                                                            const int64_t* __restrict__ dinv_off, const double* __restrict__ Lbuf,
                                                            const double* __restrict__ Dinv, double* __restrict__ Wbuf, double* __restrict__ y) {
  __shared__ double xb[32];
  const int f = fronts[blockIdx.y];
  const int k = sn_k[f];
  if (k <= kb) return;
  const int r = sn_r[f];
  const int nb = min(32, k - kb);
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= kb && blockIdx.x > 0) return;
  double* w = Wbuf + W_off[f];
  const double* Di = Dinv + dinv_off[f] + (int64_t)(kb >> 5) * 1024;
  if (threadIdx.x < 32) {  // x_blk = Dinv^T * w_blk
    double s = 0.0;
    for (int t = threadIdx.x; t < nb; ++t) s += Di[threadIdx.x * 32 + t] * w[kb + t];
    xb[threadIdx.x] = s;
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < nb) y[col0[f] + kb + threadIdx.x] = xb[threadIdx.x];
  if (j < kb) {  // earlier pivots: w_j -= sum_t L[kb+t, j] * x_t  (32 consecutive doubles of column j)
    const double* col = Lbuf + L_off[f] + (int64_t)j * r + kb;
    double s = w[j];
    for (int t = 0; t < nb; ++t) s -= col[t] * xb[t];
    w[j] = s;
  }
}

// ---- cooperative solves: all block steps of a group of fronts in ONE persistent launch -------
// A front with k pivots needs k/32 dependent block steps; as separate launches (above) every
// step pays a launch plus a drain/fill tail (~10 us measured), which made the solve of the top
// separators a 486-launch latency chain.  Here the steps are separated by grid barriers
// (~1.5 us): every CTA recomputes the 32-entry block solution through the inverted diagonal
// block (fixed summation order, so all CTAs hold the same bits) and updates its own 256 rows
// (forward) or 256 earlier pivots (backward).  Used for groups with few and/or large fronts;
// groups of many small fronts keep the one-CTA-per-front shared-memory kernels.
struct CoopSolveArgs {
  const int* fronts;
  int count, max_k, max_r;
  unsigned long long* phase_ns;  // debug stopwatch of CTA 0 (SPB_CHOL_TIMING=2), else null
  int parts;  // CTAs sharing one front (each redoes the tile's small solve, then takes every parts-th row block)
  const int *sn_k, *sn_r, *col0;
  const int64_t *L_off, *W_off, *dinv_off;
  const int *child_ptr, *child_idx;
  const int64_t* rel_ptr;
  const int* rel_idx;
  const int64_t* struct_ptr;
  const int* struct_idx;
  const double *Lbuf, *Dinv;
  double *Wbuf, *y;
};

__device__ __forceinline__ unsigned long long chol_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define CHOL_TICK(k)                                        \
  if (A.phase_ns && blockIdx.x == 0 && threadIdx.x == 0) {  \
    const unsigned long long now_ = chol_global_ns();       \
    A.phase_ns[k] += now_ - tick_;                          \
    tick_ = now_;                                           \
  }
constexpr int kCoopTile = 128;  // pivots per grid step: the tile's own triangular solve is redone by every CTA
constexpr int kBlkPitch = 33;   // shared-memory pitch of a 32x32 block: conflict-free by rows and by columns
constexpr int kBlkSize = kBlkPitch * 32;
// shared memory of the cooperative solves: 4 inverted diagonal blocks + the 6 blocks below them
constexpr size_t kCoopSmem = (size_t)(10 * kBlkSize) * sizeof(double);

// Stage the tile's factor data (read-only) in shared memory with every load in flight at once:
// Dinv blocks b = 0..3 at sm[b], block (bi, bj), bi > bj, of the tile's lower triangle at
// sm[4 + bi(bi-1)/2 + bj]; element (i, t) of a block at t * kBlkPitch + i.
__device__ __forceinline__ void coop_stage_tile(const double* __restrict__ P, const double* __restrict__ Dinv_f, int r, int kb, int nbt,
                                                double* sm) {
  const int tid = threadIdx.x, i = tid & 31, tq = tid >> 5;
  const int nblk = (nbt + 31) >> 5;
  // all (up to 40) loads of a thread are issued before the first store: one memory round trip, not ten
  double dv[4][4], lv[6][4];
#pragma unroll
  for (int bq = 0; bq < 4; ++bq) {
    const double* Di = Dinv_f + (int64_t)((kb >> 5) + bq) * 1024;
#pragma unroll
    for (int u = 0; u < 4; ++u) dv[bq][u] = bq < nblk ? Di[(tq * 4 + u) * 32 + i] : 0.0;
  }
#pragma unroll
  for (int bi = 1; bi < 4; ++bi)
#pragma unroll
    for (int bj = 0; bj < bi; ++bj) {
      const int row = bi * 32 + i;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        lv[bi * (bi - 1) / 2 + bj][u] = row < nbt ? P[(int64_t)(kb + bj * 32 + tq * 4 + u) * r + kb + row] : 0.0;
    }
  asm volatile("" ::: "memory");  // keep the loads above the stores (issue them back to back)
#pragma unroll
  for (int bq = 0; bq < 4; ++bq)
#pragma unroll
    for (int u = 0; u < 4; ++u) sm[bq * kBlkSize + (tq * 4 + u) * kBlkPitch + i] = dv[bq][u];
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int u = 0; u < 4; ++u) sm[(4 + q) * kBlkSize + (tq * 4 + u) * kBlkPitch + i] = lv[q][u];
}

// L2 prefetch of the factor data of the tile at kb (read-only), issued before a grid barrier so the
// staging loads after it are L2 hits
__device__ __forceinline__ void coop_prefetch_tile(const double* __restrict__ P, const double* __restrict__ Dinv_f, int r, int kb, int nbt) {
  const int tid = threadIdx.x;
  const int nblk = (nbt + 31) >> 5;
  // one 128-byte line per prefetch: Dinv blocks are 8 KB contiguous each, L columns nbt*8 bytes each
  for (int q = tid; q < nblk * 64; q += 256) {
    const double* a = Dinv_f + (int64_t)(kb >> 5) * 1024 + (int64_t)q * 16;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
  }
  const int lines = (nbt * 8 + 127) / 128 + 1;
  for (int q = tid; q < nbt * lines; q += 256) {
    const int t = q / lines, ln = q - t * lines;
    const double* a = P + (int64_t)(kb + t) * r + kb + ln * 16;
    if (ln * 16 < nbt + 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
  }
}

// out[0..32) = Dinv_blk * v  (TRANS: Dinv_blk^T * v) from the staged block, all 256 threads, fixed order
template <bool TRANS>
__device__ __forceinline__ void coop_diag_apply(const double* Dsm, const double* v, int nb, double (*part)[32], double* out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double sacc = 0.0;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int t = warp * 4 + u;
    if (!TRANS) {
      if (t < nb) sacc += Dsm[t * kBlkPitch + lane] * v[t];  // lower-triangular block: zeros above the diagonal
    } else {
      if (t < nb && t >= lane) sacc += Dsm[lane * kBlkPitch + t] * v[t];
    }
  }
  part[warp][lane] = sacc;
  __syncthreads();
  if (tid < 32) {
    double r = part[0][tid];
#pragma unroll
    for (int q = 1; q < 8; ++q) r += part[q][tid];
    out[tid] = r;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256, 2) chol_fwd_coop_kernel(CoopSolveArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double tile_sm[];
  __shared__ double part[8][32];
  __shared__ double wloc[kCoopTile];
  __shared__ double ys[kCoopTile];
  __shared__ double red[4][64];
  const int tid = threadIdx.x;
  // gather: w = [y_pivots ; 0] + children's tails, children in a fixed order
  for (int fi = blockIdx.x; fi < A.count; fi += gridDim.x) {
    const int f = A.fronts[fi];
    const int k = A.sn_k[f], r = A.sn_r[f], c0 = A.col0[f];
    double* w = A.Wbuf + A.W_off[f];
    for (int i = tid; i < r; i += 256) w[i] = i < k ? A.y[c0 + i] : 0.0;
    __syncthreads();
    for (int ci = A.child_ptr[f]; ci < A.child_ptr[f + 1]; ++ci) {
      const int c = A.child_idx[ci];
      const int ck = A.sn_k[c], cu = A.sn_r[c] - ck;
      const double* tail = A.Wbuf + A.W_off[c] + ck;
      const int* rel = A.rel_idx + A.rel_ptr[c];
      for (int t = tid; t < cu; t += 256) w[rel[t]] += tail[t];
      __syncthreads();
    }
  }
  grid.sync();
  const int items = A.count * A.parts;
  unsigned long long tick_ = A.phase_ns ? chol_global_ns() : 0ull;
  for (int kb = 0; kb < A.max_k; kb += kCoopTile) {
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int fi = item / A.parts, pt = item - fi * A.parts;
      const int f = A.fronts[fi];
      const int k = A.sn_k[f];
      if (k <= kb) continue;
      const int r = A.sn_r[f];
      const int nbt = min(kCoopTile, k - kb);
      if (pt > 0 && kb + nbt + pt * 64 >= r) continue;  // no rows for this part
      double* w = A.Wbuf + A.W_off[f];
      const double* P = A.Lbuf + A.L_off[f];
      coop_stage_tile(P, A.Dinv + A.dinv_off[f], r, kb, nbt, tile_sm);
      if (tid < nbt) wloc[tid] = w[kb + tid];
      __syncthreads();
      CHOL_TICK(0)
      for (int sb = 0, bq = 0; sb < nbt; sb += 32, ++bq) {  // the tile's triangular solve, 32 pivots at a time
        const int nb = min(32, nbt - sb);
        coop_diag_apply<false>(tile_sm + bq * kBlkSize, wloc + sb, nb, part, ys + sb);
        const int i = sb + 32 + tid;  // remaining pivots of the tile (the blocks below this one)
        if (i < nbt) {
          const int bi = i >> 5;
          const double* blk = tile_sm + (4 + bi * (bi - 1) / 2 + bq) * kBlkSize + (i & 31);
          double sv = wloc[i];
#pragma unroll 8
          for (int t = 0; t < 32; ++t) sv -= blk[t * kBlkPitch] * ys[sb + t];
          wloc[i] = sv;
        }
        __syncthreads();
      }
      if (pt == 0 && tid < nbt) A.y[A.col0[f] + kb + tid] = ys[tid];
      if (tid >= nbt && tid < kCoopTile) ys[tid] = 0.0;  // partial tile: the row update multiplies zeros there
      __syncthreads();
      CHOL_TICK(1)
      // rows below the tile (later pivots and the tail): 64 rows at a time, four threads per row, each
      // takes 32 of the tile's columns with all its loads in flight; partial sums meet in a fixed order
      for (int row0 = kb + nbt + pt * 64; row0 < r; row0 += A.parts * 64) {
        const int i = row0 + (tid & 63), cgp = tid >> 6;
        double acc = 0.0;
        const double wi = (tid < 64 && i < r) ? w[i] : 0.0;
        if (i < r) {
          const double* col = P + (int64_t)(kb + cgp * 32) * r + i;
          const int tn = min(32, nbt - cgp * 32);
          double lv[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) lv[t] = t < tn ? col[(int64_t)t * r] : 0.0;
#pragma unroll
          for (int t = 0; t < 32; ++t) acc += lv[t] * ys[cgp * 32 + t];
        }
        red[cgp][tid & 63] = acc;
        __syncthreads();
        if (tid < 64 && i < r) w[i] = wi - (((red[0][tid] + red[1][tid]) + red[2][tid]) + red[3][tid]);
        __syncthreads();
      }
      CHOL_TICK(2)
      if (kb + kCoopTile < k) coop_prefetch_tile(P, A.Dinv + A.dinv_off[f], r, kb + kCoopTile, min(kCoopTile, k - kb - kCoopTile));
    }
    grid.sync();
    CHOL_TICK(3)
  }
}

__global__ void __launch_bounds__(256, 2) chol_bwd_coop_kernel(CoopSolveArgs A) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ double tile_sm[];
  __shared__ double part[8][32];
  __shared__ double wloc[kCoopTile];
  __shared__ double xs[kCoopTile];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // w_piv = y_piv - L21^T x_struct : one warp per pivot column
  {
    const int ngroups = (A.max_k + 7) / 8;
    for (int item = blockIdx.x; item < A.count * ngroups; item += gridDim.x) {
      const int fi = item / ngroups, gq = item - fi * ngroups;
      const int f = A.fronts[fi];
      const int k = A.sn_k[f], r = A.sn_r[f];
      const int j = gq * 8 + warp;
      if (j >= k) continue;
      const int* st = A.struct_idx + A.struct_ptr[f];
      const double* col = A.Lbuf + A.L_off[f] + (int64_t)j * r;
      double sacc = 0.0;
      int i = k + lane;
      for (; i + 96 < r; i += 128) {  // four independent loads (and gathers) in flight per lane
        const double a0 = col[i], a1 = col[i + 32], a2 = col[i + 64], a3 = col[i + 96];
        const double b0 = A.y[st[i - k]], b1 = A.y[st[i + 32 - k]], b2 = A.y[st[i + 64 - k]], b3 = A.y[st[i + 96 - k]];
        sacc += a0 * b0;
        sacc += a1 * b1;
        sacc += a2 * b2;
        sacc += a3 * b3;
      }
      for (; i < r; i += 32) sacc += col[i] * A.y[st[i - k]];
      for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
      if (lane == 0) A.Wbuf[A.W_off[f] + j] = A.y[A.col0[f] + j] - sacc;
    }
  }
  grid.sync();
  const int items = A.count * A.parts;
  for (int kb = ((A.max_k - 1) / kCoopTile) * kCoopTile; kb >= 0; kb -= kCoopTile) {
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int fi = item / A.parts, pt = item - fi * A.parts;
      const int f = A.fronts[fi];
      const int k = A.sn_k[f];
      if (k <= kb) continue;
      if (pt > 0 && pt * 32 >= kb) continue;  // no earlier pivot columns for this part
      const int r = A.sn_r[f];
      const int nbt = min(kCoopTile, k - kb);
      double* w = A.Wbuf + A.W_off[f];
      const double* P = A.Lbuf + A.L_off[f];
      coop_stage_tile(P, A.Dinv + A.dinv_off[f], r, kb, nbt, tile_sm);
      if (tid < nbt) wloc[tid] = w[kb + tid];
      __syncthreads();
      for (int bq = (nbt - 1) >> 5; bq >= 0; --bq) {  // the tile's own back substitution
        const int sb = bq * 32;
        const int nb = min(32, nbt - sb);
        coop_diag_apply<true>(tile_sm + bq * kBlkSize, wloc + sb, nb, part, xs + sb);
        if (tid < sb) {  // earlier pivots of the tile: w_j -= sum_t L[sb+t, j] x_t
          const double* blk = tile_sm + (4 + bq * (bq - 1) / 2 + (tid >> 5)) * kBlkSize + (tid & 31) * kBlkPitch;
          double sv = wloc[tid];
#pragma unroll 8
          for (int t = 0; t < 32; ++t)
            if (t < nb) sv -= blk[t] * xs[sb + t];
          wloc[tid] = sv;
        }
        __syncthreads();
      }
      if (pt == 0 && tid < nbt) A.y[A.col0[f] + kb + tid] = xs[tid];
      // earlier pivots: w_j -= sum_t L[kb+t, j] * x_t -- 32 columns at a time, four per warp, the lanes
      // across the tile's (consecutive) rows so every load is a full line and 16 are in flight per lane
      for (int c = pt; c * 32 < kb; c += A.parts) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = c * 32 + warp * 4 + u;
          if (j < kb) {
            const double* col = P + (int64_t)j * r + kb;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < kCoopTile / 32; ++q) {
              const int t = lane + 32 * q;
              if (t < nbt) acc += col[t] * xs[t];
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) w[j] -= acc;
          }
        }
      }
      __syncthreads();
      if (kb > 0) coop_prefetch_tile(P, A.Dinv + A.dinv_off[f], r, kb - kCoopTile, kCoopTile);
    }
    grid.sync();
  }
}

__global__ void chol_permute_kernel(const double* __restrict__ in, const int* __restrict__ perm, double* __restrict__ out, int n, int dir) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dir == 0) out[i] = in[perm[i]];  // y = P b
  else out[perm[i]] = in[i];           // x = P^T y
}

// ---- host orchestration ---------------------------------------------------------------------
static CholDevice* ensure_symbolic(spb_context* h) {
  if (h->chol) return (CholDevice*)h->chol;
  CholDevice* D = new CholDevice;
  h->chol = D;
  cudaStream_t st = h->stream;
  chol_symbolic(h->H, h->S, D->sym);
  CholSymbolic& C = D->sym;
  const int ns = C.nsuper;
  std::vector<int> k(ns), r(ns), col0(ns);
  for (int s = 0; s < ns; ++s) {
    k[s] = C.sn_col0[s + 1] - C.sn_col0[s];
    r[s] = k[s] + (int)(C.struct_ptr[(size_t)s + 1] - C.struct_ptr[s]);
    col0[s] = C.sn_col0[s];
  }
  // sort fronts inside each level by front size and cut groups of similar size
  D->groups.assign(C.nlevels, {});
  for (int l = 0; l < C.nlevels; ++l) {
    int b = C.level_ptr[l], e = C.level_ptr[(size_t)l + 1];
    std::sort(C.level_sn.begin() + b, C.level_sn.begin() + e, [&](int a, int c) { return r[a] != r[c] ? r[a] < r[c] : a < c; });
    int g0 = b;
    while (g0 < e) {
      int g1 = g0;
      int lim = std::max(64, 2 * r[C.level_sn[g0]]);
      // a level with few fronts is one group whatever their sizes: its fronts are independent, and two groups would be
      // two dependent-looking launch sequences on one stream where one runs them side by side (the upper levels of a
      // dissection tree are in the launch-latency regime)
      static const int merge_below = getenv("SPB_CHOL_MERGE_LEVEL") ? atoi(getenv("SPB_CHOL_MERGE_LEVEL")) : 4096;
      if (e - b <= merge_below) lim = 0x7fffffff;
      int mk = 0, mr = 0;
      while (g1 < e && r[C.level_sn[g1]] <= lim && g1 - g0 < 32768) {
        mk = std::max(mk, k[C.level_sn[g1]]);
        mr = std::max(mr, r[C.level_sn[g1]]);
        ++g1;
      }
      D->groups[l].push_back({g0, g1 - g0, mk, mr});
      g0 = g1;
    }
  }
  SPB_CUDA(D->sn_k.upload(k, st));
  SPB_CUDA(D->sn_r.upload(r, st));
  SPB_CUDA(D->col0.upload(col0, st));
  SPB_CUDA(D->struct_idx.upload(C.struct_idx, st));
  SPB_CUDA(D->struct_ptr.upload(C.struct_ptr, st));
  SPB_CUDA(D->rel_idx.upload(C.rel_idx, st));
  SPB_CUDA(D->rel_ptr.upload(C.rel_ptr, st));
  SPB_CUDA(D->parent.upload(C.sn_parent, st));
  SPB_CUDA(D->child_ptr.upload(C.child_ptr, st));
  SPB_CUDA(D->child_idx.upload(C.child_idx, st));
  SPB_CUDA(D->level_sn.upload(C.level_sn, st));
  SPB_CUDA(D->rank_children.upload(C.rank_children, st));
  SPB_CUDA(D->ea_ptr.upload(C.ea_ptr, st));
  SPB_CUDA(D->ea_idx.upload(C.ea_idx, st));
  SPB_CUDA(D->small_parents.upload(C.small_parents, st));
  SPB_CUDA(D->perm.upload(C.perm, st));
  SPB_CUDA(D->L_off.upload(C.L_off, st));
  SPB_CUDA(D->U_off.upload(C.U_off, st));
  SPB_CUDA(D->W_off.upload(C.W_off, st));
  SPB_CUDA(D->dinv_off.upload(C.dinv_off, st));
  SPB_CUDA(D->a_src.upload(C.a_src, st));
  SPB_CUDA(D->a_dst.upload(C.a_dst, st));
  // + slack: the SYRK staging copies 66-double windows that may run past the last column of the last panel
  SPB_CUDA(D->Lbuf.alloc((size_t)C.L_off[ns] + 128));
  SPB_CUDA(cudaMemsetAsync(D->Lbuf.p + C.L_off[ns], 0, 128 * sizeof(double), h->stream));
  SPB_CUDA(D->Ubuf.alloc((size_t)std::max<int64_t>(1, C.U_total)));
  SPB_CUDA(D->Wbuf.alloc((size_t)C.W_off[ns]));
  SPB_CUDA(D->Dinv.alloc((size_t)C.dinv_off[ns]));
  SPB_CUDA(D->ywork.alloc((size_t)h->n));
  SPB_CUDA(D->fail.alloc(1));
  SPB_CUDA(cudaMallocHost((void**)&D->h_fail, sizeof(int)));
  SPB_CUDA(cudaStreamSynchronize(st));
  // the big host maps are on the device now
  std::vector<int64_t>().swap(C.a_src);
  std::vector<int64_t>().swap(C.a_dst);
  return D;
}

spb_status chol_factorize(spb_context* h) {
  if (h->n == 0) return SPB_OK;
  CholDevice* D = ensure_symbolic(h);
  CholSymbolic& C = D->sym;
  cudaStream_t st = h->stream;
  const int ns = C.nsuper;
  if ((size_t)(C.max_r + 32) * sizeof(double) > 200 * 1024) throw ArgFail{"front too large for the triangular-solve kernels"};
  static const bool fused_step = !(getenv("SPB_CHOL_FUSED") && atoi(getenv("SPB_CHOL_FUSED")) == 0);
  stage_begin(h, ST_FACTOR);
  int nl = 0;
  // The launch sequence of a factorisation depends on the symbolic state only (no host decision inside it), and on the
  // upper levels of the tree it is hundreds of dependent 10-20 us kernels: it is captured into a CUDA graph the first time
  // and replayed afterwards (one graph launch instead of ~600-1200 kernel launches and memsets per factorisation).
  auto enqueue = [&]() {
  SPB_CUDA(cudaMemsetAsync(D->Lbuf.p, 0, (size_t)C.L_off[ns] * sizeof(double), st));
  SPB_CUDA(cudaMemsetAsync(D->fail.p, 0, sizeof(int), st));
  {
    const int64_t cnt = (int64_t)D->a_src.n;
    chol_scatter_a_kernel<<<h->num_sms * 8, 256, 0, st>>>(h->d_hval.p, D->a_src.p, D->a_dst.p, cnt, D->Lbuf.p);
    ++nl;
  }
  // SYRK/GEMM operand staging: cp.async (default) or TMA bulk copies (SPB_CHOL_TMA=1).  Measured on one B200: the
  // TMA form is 2-3 % slower (48.9 vs 47.5 ms at C2, 808 vs 791 ms per LM iteration at the 500k-face size): a 64-row
  // operand segment is 512 bytes and the TMA unit is issue-bound on copies that small; the one-instruction-per-tile 2-D
  // tensor-map form needs 16-byte column strides, i.e. even leading dimensions, which the packed panels do not have.
  const bool syrk_tma = getenv("SPB_CHOL_TMA") && atoi(getenv("SPB_CHOL_TMA")) > 0;
  auto syrk = [&](dim3 grid, const int* fl, int mode, int ka, int ke, int cend, const int* snk, const int* snr, const int64_t* lo, const int64_t* uo,
                  double* Lb, double* Ub, long long* dbgp) {
    if (syrk_tma) chol_syrk_kernel<true><<<grid, 128, 0, st>>>(fl, mode, ka, ke, cend, snk, snr, lo, uo, Lb, Ub, dbgp);
    else chol_syrk_kernel<false><<<grid, 128, 0, st>>>(fl, mode, ka, ke, cend, snk, snr, lo, uo, Lb, Ub, dbgp);
  };
  for (int l = 0; l < C.nlevels; ++l) {
    // the update matrices born at this level start from zero (their arena blocks are reused across levels)
    for (int z = C.u_zero_ptr[l]; z < C.u_zero_ptr[(size_t)l + 1]; ++z)
      SPB_CUDA(cudaMemsetAsync(D->Ubuf.p + C.u_zero_range[(size_t)2 * z], 0, (size_t)C.u_zero_range[(size_t)2 * z + 1] * sizeof(double), st));
    // extend-add, one launch per child rank
    for (int rk = C.level_rank_ptr[l]; rk < C.level_rank_ptr[(size_t)l + 1]; ++rk) {
      const int b = C.rank_ptr[rk], cnt = C.rank_ptr[(size_t)rk + 1] - b;
      for (int off = 0; off < cnt; off += 32768) {
        const int c = std::min(32768, cnt - off);
        // enough CTAs to fill the GPU even when a rank has one or two (large) children
        const int gx = std::max(16, std::min(2048, (h->num_sms * 16 + c - 1) / c));
        chol_extend_add_kernel<<<dim3(gx, c), 256, 0, st>>>(D->rank_children.p + b + off, D->parent.p, D->sn_k.p, D->sn_r.p, D->L_off.p,
                                                            D->U_off.p, D->rel_ptr.p, D->rel_idx.p, D->Lbuf.p, D->Ubuf.p);
        ++nl;
      }
    }
    {
      const int sp0 = C.small_parents_ptr[l], spc = C.small_parents_ptr[(size_t)l + 1] - sp0;
      if (spc > 0) {
        chol_extend_add_small_kernel<<<(spc + 3) / 4, 128, 0, st>>>(D->small_parents.p + sp0, spc, C.level_big_ranks[l], D->ea_ptr.p, D->ea_idx.p,
                                                                    D->sn_k.p, D->sn_r.p, D->L_off.p, D->U_off.p, D->rel_ptr.p, D->rel_idx.p,
                                                                    D->Lbuf.p, D->Ubuf.p);
        ++nl;
      }
    }
    for (const CholDevice::Group& G : D->groups[l]) {
      const int* list = D->level_sn.p + G.begin;
      if (G.max_k <= 16 && G.max_r <= 32) {  // tiny fronts: the whole step in one launch, a warp per front
        chol_tiny_front_kernel<<<(G.count + 3) / 4, 128, 0, st>>>(list, G.count, D->sn_k.p, D->sn_r.p, D->L_off.p, D->U_off.p, D->dinv_off.p,
                                                                  D->Lbuf.p, D->Ubuf.p, D->Dinv.p, D->fail.p);
        ++nl;
        continue;
      }
      constexpr int kPanel = 256;
      for (int ob = 0; ob < G.max_k; ob += kPanel) {
        const int oe = ob + kPanel;
        for (int kb = ob; kb < std::min(oe, G.max_k); kb += 32) {
          const int rows_below = std::max(0, G.max_r - kb - 1);
          if (fused_step && (int64_t)((rows_below + 127) / 128 + 1) * G.count <= 2 * h->num_sms) {  // latency regime
            const int rows = rows_below;
            static const bool dbg3 = getenv("SPB_CHOL_TIMING") && atoi(getenv("SPB_CHOL_TIMING")) == 3;
            long long* dbgp = nullptr;
            if (dbg3 && G.count == 1 && kb == 0 && l == C.nlevels - 1) {
              SPB_CUDA(D->dbg.alloc(8));
              dbgp = D->dbg.p;
            }
            chol_step_kernel<<<dim3((rows + 127) / 128 + 1, G.count), 128, 0, st>>>(list, kb, D->sn_k.p, D->sn_r.p, D->L_off.p, D->dinv_off.p,
                                                                                    D->Lbuf.p, D->Dinv.p, D->fail.p, dbgp);
            if (dbgp) {
              long long t[8];
              cudaStreamSynchronize(st);
              cudaMemcpy(t, dbgp, sizeof(t), cudaMemcpyDeviceToHost);
              fprintf(stderr, "[spb chol timing] step kernel root kb=0: load %lld potrf %lld trsm %lld store %lld cycles; total %lld cycles in %lld ns\n",
                      t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[4] - t[0], t[6] - t[5]);
            }
            ++nl;
          } else {
            if (fused_step)
              chol_potrf32_warp_kernel<<<(G.count + 3) / 4, 128, 0, st>>>(list, G.count, kb, D->sn_k.p, D->sn_r.p, D->L_off.p, D->dinv_off.p,
                                                                          D->Lbuf.p, D->Dinv.p, D->fail.p);
            else
              chol_potrf32_kernel<<<G.count, 1024, 0, st>>>(list, kb, D->sn_k.p, D->sn_r.p, D->L_off.p, D->dinv_off.p, D->Lbuf.p, D->Dinv.p,
                                                            D->fail.p);
            ++nl;
            const int rows = G.max_r - kb - 1;
            if (rows > 0) {
              chol_trsm_kernel<<<dim3((rows + 127) / 128, G.count), 128, 0, st>>>(list, kb, D->sn_k.p, D->sn_r.p, D->L_off.p, D->Lbuf.p);
              ++nl;
            }
          }
          if (kb + 32 < std::min(oe, G.max_k)) {  // narrow update, confined to the panel's columns
            const int nti = (G.max_r - kb - 32 + 63) / 64, ntj = (std::min(oe, G.max_k) - kb - 32 + 63) / 64;
            static const bool dbg4 = getenv("SPB_CHOL_TIMING") && atoi(getenv("SPB_CHOL_TIMING")) == 4;
            long long* dbgp = nullptr;
            if (dbg4 && G.count == 1 && kb == 0 && l == C.nlevels - 1) {
              SPB_CUDA(D->dbg.alloc(8));
              dbgp = D->dbg.p;
            }
            syrk(dim3(ntj, nti, G.count), list, 0, kb, kb + 32, oe, D->sn_k.p, D->sn_r.p, D->L_off.p, D->U_off.p,
                                                                    D->Lbuf.p, D->Ubuf.p, dbgp);
            if (dbgp) {
              long long t[8];
              cudaStreamSynchronize(st);
              cudaMemcpy(t, dbgp, sizeof(t), cudaMemcpyDeviceToHost);
              fprintf(stderr, "[spb chol timing] narrow syrk root kb=0 grid (%d,%d): setup %lld mainloop %lld epilogue %lld cycles\n", ntj, nti,
                      t[1] - t[0], t[2] - t[1], t[3] - t[2]);
            }
            ++nl;
          }
        }
        if (oe < G.max_k) {  // everything right of the panel, once, with K = 256
          const int nti = (G.max_r - oe + 63) / 64, ntj = (G.max_k - oe + 63) / 64;
          syrk(dim3(ntj, nti, G.count), list, 0, ob, oe, 0x7fffffff, D->sn_k.p, D->sn_r.p, D->L_off.p,
                                                                  D->U_off.p, D->Lbuf.p, D->Ubuf.p, nullptr);
          ++nl;
        }
      }
      const int umax = G.max_r;  // upper bound on u inside the group
      const int nt = (umax + 63) / 64;
      syrk(dim3(nt, nt, G.count), list, 1, 0, 0, 0, D->sn_k.p, D->sn_r.p, D->L_off.p, D->U_off.p, D->Lbuf.p,
                                                             D->Ubuf.p, nullptr);
      ++nl;
    }
  }
  };
  static const bool graph_on = !(getenv("SPB_CHOL_GRAPH") && atoi(getenv("SPB_CHOL_GRAPH")) == 0);
  static const bool dbg_any = getenv("SPB_CHOL_TIMING") && atoi(getenv("SPB_CHOL_TIMING")) > 0;
  if (graph_on && !dbg_any && D->graph_exec && D->graph_hval == h->d_hval.p) {
    SPB_CUDA(cudaGraphLaunch(D->graph_exec, st));
    nl = D->graph_nl;
  } else if (graph_on && !dbg_any && cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
    if (D->graph_exec) {
      cudaGraphExecDestroy(D->graph_exec);
      D->graph_exec = nullptr;
    }
    cudaGraph_t g = nullptr;
    try {
      enqueue();
    } catch (...) {
      cudaStreamEndCapture(st, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    SPB_CUDA(cudaStreamEndCapture(st, &g));
    cudaGraphExec_t ex = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&ex, g, 0);
    cudaGraphDestroy(g);
    if (ie == cudaSuccess) {
      D->graph_exec = ex;
      D->graph_hval = h->d_hval.p;
      D->graph_nl = nl;
      SPB_CUDA(cudaGraphLaunch(ex, st));
    } else {
      cudaGetLastError();
      nl = 0;
      enqueue();  // no graph on this system: plain launches
    }
  } else {
    cudaGetLastError();
    enqueue();
  }
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_FACTOR, nl);
  SPB_CUDA(cudaMemcpyAsync(D->h_fail, D->fail.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SPB_CUDA(cudaStreamSynchronize(st));
  D->factored = (*D->h_fail == 0);
  return D->factored ? SPB_OK : SPB_NOT_SPD;
}

// cooperative solve of a group: large fronts (the shared-memory kernels cannot hold them) and
// groups with too few fronts to fill the GPU with one CTA per front
static bool use_coop(spb_context* h, const CholDevice::Group& G) {
  static const int mode = getenv("SPB_CHOL_COOP") ? atoi(getenv("SPB_CHOL_COOP")) : 1;
  if (mode == 0) return false;
  return G.max_r > kBigFront || (G.count < 2 * h->num_sms && G.max_k > 64);
}
static CoopSolveArgs coop_args(spb_context* h, CholDevice* D, const CholDevice::Group& G) {
  CoopSolveArgs A;
  A.phase_ns = nullptr;
  A.parts = std::max(1, std::min((2 * h->num_sms + G.count - 1) / G.count, (G.max_r + 63) / 64));
  A.fronts = D->level_sn.p + G.begin;
  A.count = G.count;
  A.max_k = G.max_k;
  A.max_r = G.max_r;
  A.sn_k = D->sn_k.p;
  A.sn_r = D->sn_r.p;
  A.col0 = D->col0.p;
  A.L_off = D->L_off.p;
  A.W_off = D->W_off.p;
  A.dinv_off = D->dinv_off.p;
  A.child_ptr = D->child_ptr.p;
  A.child_idx = D->child_idx.p;
  A.rel_ptr = D->rel_ptr.p;
  A.rel_idx = D->rel_idx.p;
  A.struct_ptr = D->struct_ptr.p;
  A.struct_idx = D->struct_idx.p;
  A.Lbuf = D->Lbuf.p;
  A.Dinv = D->Dinv.p;
  A.Wbuf = D->Wbuf.p;
  A.y = D->ywork.p;
  return A;
}
static int coop_grid(spb_context* h, CholDevice* D, int items) {
  if (D->coop_occ == 0) {
    int a = 0, b = 0;
    SPB_CUDA(cudaFuncSetAttribute(chol_fwd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    SPB_CUDA(cudaFuncSetAttribute(chol_bwd_coop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmem));
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, chol_fwd_coop_kernel, 256, kCoopSmem));
    SPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, chol_bwd_coop_kernel, 256, kCoopSmem));
    D->coop_occ = std::max(1, std::min(a, b));
  }
  // barriers get slower with more CTAs: no more than the work can use, at most two per SM
  return std::max(1, std::min(items, h->num_sms * std::min(2, D->coop_occ)));
}

spb_status chol_solve(spb_context* h, const double* d_b, double* d_x) {
  if (h->n == 0) return SPB_OK;
  CholDevice* D = (CholDevice*)h->chol;
  if (!D || !D->factored) throw StateFail{"Cholesky solve without a successful factorisation"};
  CholSymbolic& C = D->sym;
  cudaStream_t st = h->stream;
  const int n = h->n;
  stage_begin(h, ST_TRISOLVE);
  int nl = 0;
  chol_permute_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_b, D->perm.p, D->ywork.p, n, 0);
  ++nl;
  const size_t smem = (size_t)(C.max_r + 32) * sizeof(double);
  if (!D->smem_attr_set) {
    SPB_CUDA(cudaFuncSetAttribute(chol_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SPB_CUDA(cudaFuncSetAttribute(chol_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    D->smem_attr_set = true;
  }
  static const bool dbg = getenv("SPB_CHOL_TIMING") && atoi(getenv("SPB_CHOL_TIMING")) > 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (dbg) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
  }
  auto dbg_begin = [&]() {
    if (dbg) cudaEventRecord(e0, st);
  };
  auto dbg_end = [&](const char* what, int l, const CholDevice::Group& G) {
    if (!dbg) return;
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(stderr, "[spb chol timing] %s level %d fronts %d max_k %d max_r %d coop %d: %.1f us\n", what, l, G.count, G.max_k, G.max_r,
            (int)use_coop(h, G), ms * 1e3);
  };
  for (int l = 0; l < C.nlevels; ++l) {
    for (const CholDevice::Group& G : D->groups[l]) {
      const int* list = D->level_sn.p + G.begin;
      dbg_begin();
      struct Scope {
        decltype(dbg_end)& f;
        int l;
        const CholDevice::Group& G;
        ~Scope() { f("fwd", l, G); }
      } scope{dbg_end, l, G};
      if (use_coop(h, G)) {
        CoopSolveArgs A = coop_args(h, D, G);
        static const bool dbg2 = getenv("SPB_CHOL_TIMING") && atoi(getenv("SPB_CHOL_TIMING")) > 1;
        DevBuf<unsigned long long> d_ns;
        if (dbg2) {
          SPB_CUDA(d_ns.alloc(4));
          SPB_CUDA(cudaMemsetAsync(d_ns.p, 0, 32, st));
          A.phase_ns = d_ns.p;
        }
        struct Pr {
          bool on; DevBuf<unsigned long long>& b; cudaStream_t st; int steps;
          ~Pr() {
            if (!on) return;
            unsigned long long t[4];
            cudaStreamSynchronize(st);
            cudaMemcpy(t, b.p, 32, cudaMemcpyDeviceToHost);
            fprintf(stderr, "[spb chol timing]   fwd CTA0 per tile step: stage %.2f solve %.2f rows %.2f sync %.2f us (%d steps)\n", t[0] * 1e-3 / steps,
                    t[1] * 1e-3 / steps, t[2] * 1e-3 / steps, t[3] * 1e-3 / steps, steps);
          }
        } pr{dbg2, d_ns, st, (G.max_k + kCoopTile - 1) / kCoopTile};
        void* args[] = {&A};
        SPB_CUDA(cudaLaunchCooperativeKernel((void*)chol_fwd_coop_kernel, dim3(coop_grid(h, D, G.count * A.parts)), dim3(256), args, kCoopSmem, st));
        ++nl;
        continue;
      }
      if (G.max_r > kBigFront) {
        chol_fwd_gather_kernel<<<G.count, 256, 0, st>>>(list, D->sn_k.p, D->sn_r.p, D->col0.p, D->W_off.p, D->child_ptr.p, D->child_idx.p,
                                                        D->rel_ptr.p, D->rel_idx.p, D->Wbuf.p, D->ywork.p);
        ++nl;
        for (int kb = 0; kb < G.max_k; kb += 32) {
          const int rows = std::max(1, G.max_r - kb - 1);
          chol_fwd_step_kernel<<<dim3((rows + 255) / 256, G.count), 256, 0, st>>>(list, kb, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p,
                                                                                D->W_off.p, D->dinv_off.p, D->Lbuf.p, D->Dinv.p, D->Wbuf.p,
                                                                                D->ywork.p);
          ++nl;
        }
        continue;
      }
      if (G.max_k <= 16 && G.max_r <= 32) {
        chol_forward_tiny_kernel<<<(G.count + 3) / 4, 128, 0, st>>>(list, G.count, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p, D->W_off.p,
                                                                    D->dinv_off.p, D->child_ptr.p, D->child_idx.p, D->rel_ptr.p, D->rel_idx.p,
                                                                    D->Lbuf.p, D->Dinv.p, D->Wbuf.p, D->ywork.p);
        ++nl;
        continue;
      }
      const size_t sm = (size_t)(G.max_r + 32) * sizeof(double);
      chol_forward_kernel<<<G.count, 256, sm, st>>>(D->level_sn.p + G.begin, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p, D->W_off.p,
                                                    D->dinv_off.p, D->child_ptr.p, D->child_idx.p, D->rel_ptr.p, D->rel_idx.p, D->Lbuf.p,
                                                    D->Dinv.p, D->Wbuf.p, D->ywork.p);
      ++nl;
    }
  }
  for (int l = C.nlevels - 1; l >= 0; --l) {
    for (const CholDevice::Group& G : D->groups[l]) {
      const int* list = D->level_sn.p + G.begin;
      dbg_begin();
      struct Scope {
        decltype(dbg_end)& f;
        int l;
        const CholDevice::Group& G;
        ~Scope() { f("bwd", l, G); }
      } scope{dbg_end, l, G};
      if (use_coop(h, G)) {
        CoopSolveArgs A = coop_args(h, D, G);
        void* args[] = {&A};
        SPB_CUDA(cudaLaunchCooperativeKernel((void*)chol_bwd_coop_kernel, dim3(coop_grid(h, D, G.count * std::max(A.parts, (G.max_k + 7) / 8))), dim3(256), args, kCoopSmem, st));
        ++nl;
        continue;
      }
      if (G.max_r > kBigFront) {
        chol_bwd_tail_kernel<<<dim3((G.max_k + 7) / 8, G.count), 256, 0, st>>>(list, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p, D->W_off.p,
                                                                              D->struct_ptr.p, D->struct_idx.p, D->Lbuf.p, D->Wbuf.p,
                                                                              D->ywork.p);
        ++nl;
        for (int kb = ((G.max_k - 1) / 32) * 32; kb >= 0; kb -= 32) {
          chol_bwd_step_kernel<<<dim3(std::max(1, (kb + 255) / 256), G.count), 256, 0, st>>>(list, kb, D->sn_k.p, D->sn_r.p, D->col0.p,
                                                                                           D->L_off.p, D->W_off.p, D->dinv_off.p, D->Lbuf.p,
                                                                                           D->Dinv.p, D->Wbuf.p, D->ywork.p);
          ++nl;
        }
        continue;
      }
      if (G.max_k <= 16 && G.max_r <= 32) {
        chol_backward_tiny_kernel<<<(G.count + 3) / 4, 128, 0, st>>>(list, G.count, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p, D->dinv_off.p,
                                                                     D->struct_ptr.p, D->struct_idx.p, D->Lbuf.p, D->Dinv.p, D->ywork.p);
        ++nl;
        continue;
      }
      const size_t sm = (size_t)(G.max_r + 32) * sizeof(double);
      chol_backward_kernel<<<G.count, 256, sm, st>>>(D->level_sn.p + G.begin, D->sn_k.p, D->sn_r.p, D->col0.p, D->L_off.p, D->dinv_off.p,
                                                     D->struct_ptr.p, D->struct_idx.p, D->Lbuf.p, D->Dinv.p, D->ywork.p);
      ++nl;
    }
  }
  (void)smem;
  chol_permute_kernel<<<(n + 255) / 256, 256, 0, st>>>(D->ywork.p, D->perm.p, d_x, n, 1);
  ++nl;
  SPB_CUDA(cudaGetLastError());
  stage_end(h, ST_TRISOLVE, nl);
  return SPB_OK;
}

void chol_release(spb_context* h) {
  CholDevice* D = (CholDevice*)h->chol;
  if (D) {
    if (D->h_fail) cudaFreeHost(D->h_fail);
    if (D->graph_exec) cudaGraphExecDestroy(D->graph_exec);
    delete D;
  }
  h->chol = nullptr;
}

// exported for tests / bench: nnz(L), flops, supernodes, levels of the current symbolic phase
bool chol_info(spb_context* h, int64_t* nnzL, double* flops, int* nsuper, int* nlevels) {
  CholDevice* D = (CholDevice*)h->chol;
  if (!D) return false;
  if (nnzL) *nnzL = D->sym.nnzL;
  if (flops) *flops = D->sym.flops;
  if (nsuper) *nsuper = D->sym.nsuper;
  if (nlevels) *nlevels = D->sym.nlevels;
  return true;
}

}  // namespace spb
