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
// Two kernels, both "coalesced stage -> sequential segment sum":
//  * assemble_pairs: one CTA per chunk of strictly-lower H slots.  The chunk's pair records
//    ([first:1][po][qo], 2 bytes for J columns <=128 entries) are streamed coalesced; each
//    lane gathers its two J values (read-only path, L1/L2 resident: a slot's partners sit
//    in one or two 128-byte lines), multiplies, and parks the product in shared memory.
//    The slot of a record is recovered without per-record indices from a warp ballot of
//    the `first` flags plus one uint16 per 32 records.  Then one thread per slot adds its
//    products in order and writes H(i,j) and its mirror H(j,i) straight into the SpMV
//    storage (sliced ELL), so no second pass over H is needed.
//  * assemble_cols: one CTA per chunk of J columns.  Values, row indices and the gathered
//    residuals are staged coalesced in (bank-skewed) shared memory, then one thread per
//    column walks its column in order producing sum v*f, sum (lambda*v)*v and sum v*v, and
//    emits rhs, d, sqrt(d)^2 onto the diagonal, the Jacobi diagonal, and a block max of
//    |rhs| folded with atomicMax on the (non-negative) bit pattern.
#include <algorithm>

#include "spb_internal.h"

namespace spb {

template <typename RecT, int POBITS>
__global__ void __launch_bounds__(256) assemble_pairs_kernel(
    const RecT* __restrict__ recs, const uint16_t* __restrict__ grp_slot, const int* __restrict__ chunk_slot,
    const int64_t* __restrict__ chunk_rec, const int* __restrict__ chunk_npairs, const int* __restrict__ chunk_col0,
    const int* __restrict__ low_colptr, const int* __restrict__ low_row, const uint16_t* __restrict__ low_rank,
    const uint16_t* __restrict__ low_rank_up, const int* __restrict__ Jcolptr, const int64_t* __restrict__ slice_off,
    const double* __restrict__ Jv, double* __restrict__ hval) {
  __shared__ double prod[kPairTile];
  __shared__ int s_pbase[kSlotCap], s_qbase[kSlotCap], s_col[kSlotCap], s_start[kSlotCap + 1];
  const int c = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const int s0 = chunk_slot[c], ns = chunk_slot[c + 1] - s0;
  const int64_t rec0 = chunk_rec[c];
  const int np = chunk_npairs[c];
  constexpr uint32_t MASK = (1u << POBITS) - 1u;

  // phase 0: per-slot bases.  The column of a slot is found by bisection over the chunk's
  // column range (columns without strictly-lower entries make the range loose, not wrong).
  {
    const int jlo = chunk_col0[c], jhi = chunk_col0[c + 1];
    for (int t = tid; t < ns; t += 256) {
      const int s = s0 + t;
      int lo = jlo, hi = jhi;  // find j with low_colptr[j] <= s < low_colptr[j+1]
      while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (low_colptr[mid] <= s) lo = mid; else hi = mid - 1;
      }
      s_col[t] = lo;
      s_qbase[t] = Jcolptr[lo];
      s_pbase[t] = Jcolptr[low_row[s]];
    }
  }
  __syncthreads();

  const bool multi = np > kPairTile;  // only possible for a single-slot chunk
  double carry = 0.0;
  bool have = false;
  for (int a = 0; a < np; a += kPairTile) {
    const int b = min(np, a + kPairTile);
    const int bround = (b + 31) & ~31;
    // phase A: products
    for (int t = a + tid; t < bround; t += 256) {
      const uint32_t rec = (uint32_t)recs[rec0 + t];  // chunks are padded to 32 records with zeros
      const uint32_t first = rec >> (2 * POBITS);
      const unsigned ballot = __ballot_sync(0xffffffffu, first != 0);
      const int g = (int)((rec0 + (t - lane)) >> 5);
      const int sl = (int)grp_slot[g] + __popc(ballot & ((2u << lane) - 1u) & ~1u);
      if (t < b) {
        const int po = (int)((rec >> POBITS) & MASK), qo = (int)(rec & MASK);
        const double vp = __ldg(Jv + s_pbase[sl] + po);
        const double vq = __ldg(Jv + s_qbase[sl] + qo);
        prod[t - a] = __dmul_rn(vp, vq);
        if (first) s_start[sl] = t - a;
      }
    }
    if (tid == 0) {
      s_start[ns] = b - a;
      if (a > 0) s_start[0] = 0;  // continuing slot of a multi-tile chunk
    }
    __syncthreads();
    // phase B: ordered sums
    if (!multi) {
      for (int t = tid; t < ns; t += 256) {
        const int beg = s_start[t], end = s_start[t + 1];
        double acc = prod[beg];
        for (int k = beg + 1; k < end; ++k) acc = __dadd_rn(acc, prod[k]);
        const int s = s0 + t;
        const int i = low_row[s], j = s_col[t];
        const int64_t plo = slice_off[i >> 5] + (int64_t)low_rank[s] * 32 + (i & 31);
        const int64_t pup = slice_off[j >> 5] + (int64_t)low_rank_up[s] * 32 + (j & 31);
        hval[plo] = acc;
        hval[pup] = acc;
      }
    } else if (tid == 0) {
      int k = 0;
      if (!have) {
        carry = prod[0];
        have = true;
        k = 1;
      }
      for (; k < b - a; ++k) carry = __dadd_rn(carry, prod[k]);
    }
    __syncthreads();
  }
  if (multi && tid == 0) {
    const int s = s0;
    const int i = low_row[s], j = s_col[0];
    const int64_t plo = slice_off[i >> 5] + (int64_t)low_rank[s] * 32 + (i & 31);
    const int64_t pup = slice_off[j >> 5] + (int64_t)low_rank_up[s] * 32 + (j & 31);
    hval[plo] = carry;
    hval[pup] = carry;
  }
}

__device__ __forceinline__ int skew(int i) { return i + (i >> 4); }

__global__ void __launch_bounds__(256) assemble_cols_kernel(
    const int* __restrict__ colchunk, const int* __restrict__ Jcolptr, const int* __restrict__ Jrow,
    const double* __restrict__ Jv, const double* __restrict__ f, double lambda, const int* __restrict__ nup1,
    const int64_t* __restrict__ slice_off, double* __restrict__ hval, double* __restrict__ hdiag,
    double* __restrict__ rhs, double* __restrict__ dvec, unsigned long long* __restrict__ foo_bits, int defer_damping) {
  __shared__ double sv[kEntTile + kEntTile / 16 + 1], sf[kEntTile + kEntTile / 16 + 1];
  __shared__ double s_max[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int j0 = colchunk[c], j1 = colchunk[c + 1];
  const int e0 = Jcolptr[j0], e1 = Jcolptr[j1];
  constexpr int CPT = kColCap / 256;  // columns per thread
  double a_vf[CPT], a_d[CPT], a_vv[CPT];
  bool started[CPT];
  int cs[CPT], ce[CPT];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int j = j0 + tid + k * 256;
    a_vf[k] = 0.0;
    a_d[k] = 0.0;
    a_vv[k] = 0.0;
    started[k] = false;
    cs[k] = j < j1 ? Jcolptr[j] : 0;
    ce[k] = j < j1 ? Jcolptr[j + 1] : 0;
  }
  for (int a = e0; a < e1 || a == e0; a += kEntTile) {
    const int b = min(e1, a + kEntTile);
    for (int e = a + tid; e < b; e += 256) {
      const double v = Jv[e];
      sv[skew(e - a)] = v;
      sf[skew(e - a)] = f ? __ldg(f + Jrow[e]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int lo = max(cs[k], a), hi = min(ce[k], b);
      for (int e = lo; e < hi; ++e) {
        const double v = sv[skew(e - a)], fr = sf[skew(e - a)];
        a_vf[k] = __dadd_rn(a_vf[k], __dmul_rn(v, fr));              // tmp += v*f   (LMSolver.h:117)
        a_d[k] = __dadd_rn(a_d[k], __dmul_rn(__dmul_rn(lambda, v), v));  // (lambda*v)*v (DiagonalDamping.h:35)
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
  SPB_CUDA(h->d_rowlen.upload(h->S.rowlen, st));
  SPB_CUDA(h->d_slice_off.upload(h->S.slice_off, st));
  SPB_CUDA(h->d_nup1.upload(h->H.nup1, st));
  SPB_CUDA(h->d_hval.alloc((size_t)h->S.padded));
  SPB_CUDA(cudaMemsetAsync(h->d_hval.p, 0, (size_t)h->S.padded * sizeof(double), st));
  SPB_CUDA(h->d_hdiag.alloc(h->n));
  SPB_CUDA(h->d_rhs.alloc(h->n));
  SPB_CUDA(h->d_dvec.alloc(h->n));
  SPB_CUDA(cudaStreamSynchronize(st));
  std::vector<int>().swap(h->S.col);  // the big host copy is no longer needed
  h->d_pos_of_slot.release();
}

void upload_analysis(spb_context* h, const int* colptr, const int* rowidx, AssemblyPlan& P) {
  cudaStream_t st = h->stream;
  std::vector<int> cp(colptr, colptr + h->n + 1), ri(rowidx, rowidx + colptr[h->n]);
  SPB_CUDA(h->d_Jcolptr.upload(cp, st));
  SPB_CUDA(h->d_Jrow.upload(ri, st));
  h->rec_bits = P.rec_bits;
  if (P.rec_bits == 16) SPB_CUDA(h->d_rec16.upload(P.rec16, st)); else SPB_CUDA(h->d_rec32.upload(P.rec32, st));
  SPB_CUDA(h->d_grp_slot.upload(P.grp_slot, st));
  SPB_CUDA(h->d_low_rank.upload(P.low_rank, st));
  SPB_CUDA(h->d_low_rank_up.upload(P.low_rank_up, st));
  SPB_CUDA(h->d_chunk_slot.upload(P.chunk_slot, st));
  SPB_CUDA(h->d_chunk_npairs.upload(P.chunk_npairs, st));
  SPB_CUDA(h->d_chunk_rec.upload(P.chunk_rec, st));
  SPB_CUDA(h->d_low_colptr.upload(P.low_colptr, st));
  SPB_CUDA(h->d_low_row.upload(P.low_row, st));
  SPB_CUDA(h->d_colchunk.upload(P.colchunk, st));
  // first column of every chunk (+ sentinel): bisection bounds for the slot->column search
  h->nchunks = (int)P.chunk_npairs.size();
  std::vector<int> col0((size_t)h->nchunks + 1, h->n > 0 ? h->n - 1 : 0);
  {
    int j = 0;
    for (int c = 0; c < h->nchunks; ++c) {
      int s = P.chunk_slot[c];
      while (P.low_colptr[(size_t)j + 1] <= s) ++j;
      col0[c] = j;
    }
  }
  SPB_CUDA(h->d_chunk_col0.upload(col0, st));
  h->ncolchunks = (int)P.colchunk.size() - 1;
  h->nslots = P.nslots;
  h->npairs = P.npairs;
  SPB_CUDA(h->d_foo_bits.alloc(1));
  SPB_CUDA(cudaStreamSynchronize(st));
}

void run_assembly(spb_context* h, const double* d_Jv, const double* d_f, double lambda, double* foo_host) {
  cudaStream_t st = h->stream;
  stage_begin(h, ST_ASSEMBLY);
  int nl = 0;
  if (d_f) SPB_CUDA(cudaMemsetAsync(h->d_foo_bits.p, 0, sizeof(unsigned long long), st));
  const bool part = h->partitioned;
  double* hv = part ? h->d_hval_part.p : h->d_hval.p;
  if (h->ncolchunks > 0) {
    assemble_cols_kernel<<<h->ncolchunks, 256, 0, st>>>(h->d_colchunk.p, h->d_Jcolptr.p, h->d_Jrow.p, d_Jv, d_f, lambda,
                                                        h->d_nup1.p, h->d_slice_off.p, hv, h->d_hdiag.p,
                                                        part ? h->d_rhs_part.p : h->d_rhs.p, part ? h->d_dvec_part.p : h->d_dvec.p,
                                                        h->d_foo_bits.p, part ? 1 : 0);
    ++nl;
  }
  if (h->nchunks > 0) {
    if (h->rec_bits == 16)
      assemble_pairs_kernel<uint16_t, 7><<<h->nchunks, 256, 0, st>>>(
          h->d_rec16.p, h->d_grp_slot.p, h->d_chunk_slot.p, h->d_chunk_rec.p, h->d_chunk_npairs.p, h->d_chunk_col0.p,
          h->d_low_colptr.p, h->d_low_row.p, h->d_low_rank.p, h->d_low_rank_up.p, h->d_Jcolptr.p, h->d_slice_off.p, d_Jv,
          hv);
    else
      assemble_pairs_kernel<uint32_t, 15><<<h->nchunks, 256, 0, st>>>(
          h->d_rec32.p, h->d_grp_slot.p, h->d_chunk_slot.p, h->d_chunk_rec.p, h->d_chunk_npairs.p, h->d_chunk_col0.p,
          h->d_low_colptr.p, h->d_low_row.p, h->d_low_rank.p, h->d_low_rank_up.p, h->d_Jcolptr.p, h->d_slice_off.p, d_Jv,
          hv);
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
