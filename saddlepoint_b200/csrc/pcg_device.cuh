// pcg_device.cuh -- device helpers shared by the PCG kernels (pcg.cu, pcg_peer.cu): warp/block reductions,
// the sliced-ELL FP64 SpMV inner loops, and the redundant per-block finalisation of grid-wide sums.
#pragma once
#include <cstdint>

#include <cooperative_groups.h>

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

// ---- SpMV ----------------------------------------------------------------------------
// y = H x over the slice range [s_begin, s_end); lane == row inside a 32-row slice; the 8
// warps of a block stride over the range.  Returns this thread's share of x.y.
// NC=true: x is immutable for the kernel's lifetime, gathers may use the non-coherent path.
// Loads of the H stream.  PIN=false: evict-first streaming loads (ld.global.cs), so the stream
// never displaces the PCG vectors from L2.  PIN=true: L2 evict_last policy -- the persistent PCG
// kernel re-reads the same H every iteration, so a fixed leading part of every block's slice
// range is kept resident in the 126 MB L2 across iterations and only the rest comes from HBM.
#ifndef SPB_SPMV_UNROLL8
#define SPB_SPMV_UNROLL8 1
#endif
template <bool PIN>
__device__ __forceinline__ int ld_h_i32(const int* p, uint64_t pol) {
  if (!PIN) return __ldcs(p);
  int v;
  asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
template <bool PIN>
__device__ __forceinline__ double ld_h_f64(const double* p, uint64_t pol) {
  if (!PIN) return __ldcs(p);
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
// the FP32 copy of the value stream (relaxed-precision products of the late PCG iterations): streaming loads only
template <bool PIN>
__device__ __forceinline__ double ld_h_f64(const float* p, uint64_t) {
  return (double)__ldcs(p);
}
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// C16: the slice's column indices are read from the 16-bit stream (column = row + delta - 32768),
// 10 instead of 12 bytes per entry; slices with a delta that does not fit keep the 32-bit stream.
// row sum of one lane's row of a slice (no store): acc = sum_k H(row, c_k) x[c_k], ascending k
// VT: double (the assembled values) or float (their FP32 copy, see pcg.cu "relaxed-precision products").  W32: also
// write the FP32 copy of every value read (VT = double only): the conversion rides on the solve's first product.
template <bool NC, bool PIN, bool C16, typename VT = double, bool W32 = false>
__device__ __forceinline__ double spmv_slice_acc(const int64_t* __restrict__ slice_off, const int* __restrict__ col,
                                                 const int* __restrict__ rowlen, const VT* __restrict__ hval, const double* x,
                                                 int n, int slice, int lane, uint64_t pol, const uint16_t* __restrict__ col16,
                                                 float* __restrict__ h32out = nullptr) {
  const int row = slice * 32 + lane;
  const int64_t off = slice_off[slice];
  const int width = (int)((slice_off[slice + 1] - off) >> 5);
  const int len = row < n ? rowlen[row] : 0;
  const int* cp = col + off + lane;
  const uint16_t* cp16 = col16 + off + lane;
  const VT* vp = hval + off + lane;
  float* wp = W32 ? h32out + off + lane : nullptr;
  const int rbase = row - 32768;
  double acc = 0.0;
  int k = 0;
#if SPB_SPMV_UNROLL8
  for (; k + 8 <= width; k += 8) {  // eight entries in flight per lane
    int c[8];
    double v[8], xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      c[u] = C16 ? rbase + (int)__ldcs(cp16 + (k + u) * 32) : ld_h_i32<PIN>(cp + (k + u) * 32, pol);
      v[u] = ld_h_f64<PIN>(vp + (k + u) * 32, pol);
    }
    if (W32) {
#pragma unroll
      for (int u = 0; u < 8; ++u) __stcs(wp + (k + u) * 32, (float)v[u]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) xv[u] = NC ? __ldg(x + c[u]) : x[c[u]];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (k + u < len) acc = fma(v[u], xv[u], acc);
  }
#endif
  for (; k + 4 <= width; k += 4) {
    int c0, c1, c2, c3;
    if (C16) {
      c0 = rbase + (int)__ldcs(cp16 + (k + 0) * 32);
      c1 = rbase + (int)__ldcs(cp16 + (k + 1) * 32);
      c2 = rbase + (int)__ldcs(cp16 + (k + 2) * 32);
      c3 = rbase + (int)__ldcs(cp16 + (k + 3) * 32);
    } else {
      c0 = ld_h_i32<PIN>(cp + (k + 0) * 32, pol);
      c1 = ld_h_i32<PIN>(cp + (k + 1) * 32, pol);
      c2 = ld_h_i32<PIN>(cp + (k + 2) * 32, pol);
      c3 = ld_h_i32<PIN>(cp + (k + 3) * 32, pol);
    }
    const double v0 = ld_h_f64<PIN>(vp + (k + 0) * 32, pol), v1 = ld_h_f64<PIN>(vp + (k + 1) * 32, pol),
                 v2 = ld_h_f64<PIN>(vp + (k + 2) * 32, pol), v3 = ld_h_f64<PIN>(vp + (k + 3) * 32, pol);
    if (W32) {
      __stcs(wp + (k + 0) * 32, (float)v0);
      __stcs(wp + (k + 1) * 32, (float)v1);
      __stcs(wp + (k + 2) * 32, (float)v2);
      __stcs(wp + (k + 3) * 32, (float)v3);
    }
    const double x0 = NC ? __ldg(x + c0) : x[c0], x1 = NC ? __ldg(x + c1) : x[c1], x2 = NC ? __ldg(x + c2) : x[c2], x3 = NC ? __ldg(x + c3) : x[c3];
    if (k + 0 < len) acc = fma(v0, x0, acc);
    if (k + 1 < len) acc = fma(v1, x1, acc);
    if (k + 2 < len) acc = fma(v2, x2, acc);
    if (k + 3 < len) acc = fma(v3, x3, acc);
  }
  for (; k < width; ++k) {
    const int c0 = C16 ? rbase + (int)__ldcs(cp16 + k * 32) : ld_h_i32<PIN>(cp + k * 32, pol);
    const double v0 = ld_h_f64<PIN>(vp + k * 32, pol);
    if (W32) __stcs(wp + k * 32, (float)v0);
    const double x0 = NC ? __ldg(x + c0) : x[c0];
    if (k < len) acc = fma(v0, x0, acc);
  }
  return acc;
}

template <bool NC, bool PIN, bool C16, typename VT = double, bool W32 = false>
__device__ __forceinline__ double spmv_slice(const int64_t* __restrict__ slice_off, const int* __restrict__ col,
                                             const int* __restrict__ rowlen, const VT* __restrict__ hval, const double* x,
                                             double* y, int n, int slice, int lane, uint64_t pol, const uint16_t* __restrict__ col16,
                                             float* __restrict__ h32out = nullptr) {
  const double acc = spmv_slice_acc<NC, PIN, C16, VT, W32>(slice_off, col, rowlen, hval, x, n, slice, lane, pol, col16, h32out);
  const int row = slice * 32 + lane;
  if (row < n) {
    y[row] = acc;
    return (NC ? __ldg(x + row) : x[row]) * acc;
  }
  return 0.0;
}

// y = H x over the slice range [s_begin, s_end); lane == row inside a 32-row slice; the 8
// warps of a block stride over the range.  Returns this thread's share of x.y.
// NC=true: x is immutable for the kernel's lifetime, gathers may use the non-coherent path.
// Slices below s_pin_end are loaded with the L2 evict_last policy (see ld_h_*).
template <bool NC, typename VT = double, bool W32 = false>
__device__ __forceinline__ double spmv_range(const int64_t* __restrict__ slice_off, const int* __restrict__ col,
                                             const int* __restrict__ rowlen, const VT* __restrict__ hval, const double* x,
                                             double* y, int n, int s_begin, int s_end, int s_pin_end = 0,
                                             const uint16_t* __restrict__ col16 = nullptr, const uint8_t* __restrict__ s16 = nullptr,
                                             int first = -1, int step = 8, float* __restrict__ h32out = nullptr) {
  // default: the 8 warps of a block stride over the block's contiguous range [s_begin, s_end); first/step override
  // the walk (sweep order: first = global warp index, step = warps in the grid, s_end = all slices)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double loc = 0.0;
  uint64_t pol = 0;
  if (s_pin_end > s_begin) pol = l2_evict_last_policy();
  for (int slice = first >= 0 ? first : s_begin + warp; slice < s_end; slice += step) {
    if (col16 && s16[slice]) loc += spmv_slice<NC, false, true, VT, W32>(slice_off, col, rowlen, hval, x, y, n, slice, lane, pol, col16, h32out);
    else if (slice < s_pin_end) loc += spmv_slice<NC, true, false, VT, W32>(slice_off, col, rowlen, hval, x, y, n, slice, lane, pol, col16, h32out);
    else loc += spmv_slice<NC, false, false, VT, W32>(slice_off, col, rowlen, hval, x, y, n, slice, lane, pol, col16, h32out);
  }
  return loc;
}

// Reductions finalised redundantly by every block of a cooperative grid from per-block partials, in one fixed order.
__device__ __forceinline__ double block_total(const double* part, int count, double* sm, double* bc) {
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) s += __ldcg(part + i);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = sm[0];
    for (int w = 1; w < 8; ++w) r += sm[w];
    *bc = r;
  }
  __syncthreads();
  const double r = *bc;
  __syncthreads();
  return r;
}

// two reductions finalised in one pass (one L2 round trip, one set of block barriers)
__device__ __forceinline__ void block_total2(const double* partA, const double* partB, int count, double* sm2, double* bc2,
                                             double& outA, double& outB) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < count; i += 256) {
    a += __ldcg(partA + i);
    b += __ldcg(partB + i);
  }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    sm2[threadIdx.x >> 5] = a;
    sm2[8 + (threadIdx.x >> 5)] = b;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    const double* src = sm2 + 8 * threadIdx.x;
    double r = src[0];
    for (int w = 1; w < 8; ++w) r += src[w];
    bc2[threadIdx.x] = r;
  }
  __syncthreads();
  outA = bc2[0];
  outB = bc2[1];
  __syncthreads();
}

}  // namespace spb
