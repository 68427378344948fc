// problems.cu -- device-resident SolverTraits for the benchmark workloads.
//
// These are the *inputs* of the hot path, evaluated where the hot path runs so that the
// 134 MB of Jacobian values per iteration never cross PCIe:
//  * "LF": sparse analogue of benchmark/linear_function/main.cpp (f = A x - 1, J = A, x0 = 1)
//    on a g x g torus with 8 entries per residual row;
//  * extended Rosenbrock: benchmark/rosenbrock/main.cpp:47-62 replicated over independent
//    2-variable blocks, including its explicit structural zero JVals(3)=0.0 (config C1/C4).
// Both evaluate residuals in the same operation order as a plain host loop (ascending
// column, separate multiply and add), so host and device traits produce identical bits.
#include <algorithm>
#include <numeric>

#include "spb_internal.h"

using namespace spb;

struct spb_problem {
  spb_context* h = nullptr;
  int kind = 0;  // 0 LF, 1 Rosenbrock
  int n = 0, m = 0;           // m: residual rows evaluated by this object (the owned rows when partitioned)
  int m_global = 0, row_begin = 0, row_end = 0;
  int64_t nnzJ = 0;
  std::vector<int> colptr, rowidx;  // CSC pattern of J (host)
  std::vector<double> x0;
  // LF
  DevBuf<int> d_ell_col;      // [8][m] column-major ELL of A's rows (ascending column)
  DevBuf<double> d_ell_val;
  DevBuf<double> d_csc_val;   // constant Jacobian values, CSC order
  // Rosenbrock
  int nblocks = 0;
};

namespace {

inline uint64_t splitmix64_at(uint64_t seed, uint64_t ctr) {
  uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline double unit_pm1(uint64_t bits) { return 2.0 * ((double)(bits >> 11) * (1.0 / 9007199254740992.0)) - 1.0; }

__global__ void lf_residual_kernel(const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                                   double* __restrict__ E, int m) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = col[(size_t)k * m + r];
    const double a = val[(size_t)k * m + r];
    s = __dadd_rn(s, __dmul_rn(a, __ldg(x + c)));
  }
  E[r] = s - 1.0;
}

// CSC order of the 2x2 block b: (2b,2b) (2b+1,2b) | (2b,2b+1) (2b+1,2b+1)
__global__ void rosenbrock_kernel(const double* __restrict__ x, double* __restrict__ E, double* __restrict__ Jv, int nblocks) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const double VC = 25.0;
  const double xa = x[2 * b], xb = x[2 * b + 1];
  E[2 * b] = __dmul_rn(VC, __dadd_rn(xb, -__dmul_rn(xa, xa)));
  E[2 * b + 1] = 1.0 - xa;
  if (Jv) {
    Jv[4 * b + 0] = __dmul_rn(__dmul_rn(-2.0, VC), xa);
    Jv[4 * b + 1] = -1.0;
    Jv[4 * b + 2] = VC;
    Jv[4 * b + 3] = 0.0;
  }
}

void lf_objjac(void* user, const double* x, double* E, double* Jv, void* stream) {
  spb_problem* P = (spb_problem*)user;
  cudaStream_t st = (cudaStream_t)stream;
  lf_residual_kernel<<<(P->m + 255) / 256, 256, 0, st>>>(P->d_ell_col.p, P->d_ell_val.p, x, E, P->m);
  P->h->launches++;
  P->h->timing.count[ST_TRAITS]++;
  if (Jv) cudaMemcpyAsync(Jv, P->d_csc_val.p, (size_t)P->nnzJ * sizeof(double), cudaMemcpyDeviceToDevice, st);
}
void rb_objjac(void* user, const double* x, double* E, double* Jv, void* stream) {
  spb_problem* P = (spb_problem*)user;
  cudaStream_t st = (cudaStream_t)stream;
  rosenbrock_kernel<<<(P->nblocks + 255) / 256, 256, 0, st>>>(x, E, Jv, P->nblocks);
  P->h->launches++;
  P->h->timing.count[ST_TRAITS]++;
}
void any_x0(void* user, double* x0) {
  spb_problem* P = (spb_problem*)user;
  std::copy(P->x0.begin(), P->x0.end(), x0);
}

}  // namespace

extern "C" {

static spb_status build_lf(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, int32_t rb, int32_t re, bool part, spb_problem_t* out) {
  if (!h || !out || g < 3 || rows_mult < 1) return SPB_ERR_ARG;
  if ((int64_t)g * g * rows_mult * 8 > 0x7fffffffLL) return SPB_ERR_ARG;
  *out = nullptr;
  spb_problem* P = new spb_problem;
  try {
    SPB_CUDA(cudaSetDevice(h->device));
    P->h = h;
    P->kind = 0;
    const int n = g * g, mg = rows_mult * n;
    if (!part) {
      rb = 0;
      re = mg;
    }
    if (rb < 0 || re > mg || re < rb) {
      delete P;
      return SPB_ERR_ARG;
    }
    const int ml = re - rb;
    P->n = n;
    P->m = ml;
    P->m_global = mg;
    P->row_begin = part ? rb : 0;
    P->row_end = part ? re : 0;
    P->x0.assign(n, 1.0);
    // column indices of every row (the global pattern is needed on every rank); values of owned rows
    std::vector<int> rc((size_t)mg * 8);
    std::vector<double> rv((size_t)ml * 8);
    static const int dy[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
    static const int dx[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    for (int r = 0; r < mg; ++r) {
      const int c0 = r % n, y0 = c0 / g, x0 = c0 % g;
      const uint64_t h0 = splitmix64_at(seed, (uint64_t)r * 8);
      const int drop = (int)(h0 & 7);
      const bool own = r >= rb && r < re;
      int* c = &rc[(size_t)r * 8];
      double vbuf[8];
      double* v = own ? &rv[(size_t)(r - rb) * 8] : vbuf;
      c[0] = c0;
      v[0] = own ? 2.0 + unit_pm1(h0) : 0.0;
      int s = 1;
      for (int k = 0; k < 8; ++k) {
        if (k == drop) continue;
        c[s] = ((y0 + dy[k] + g) % g) * g + (x0 + dx[k] + g) % g;
        v[s] = own ? unit_pm1(splitmix64_at(seed, (uint64_t)r * 8 + s)) : 0.0;
        ++s;
      }
      // ascending column inside the row (insertion sort of 8)
      for (int a = 1; a < 8; ++a) {
        int cc = c[a];
        double vv = v[a];
        int b = a - 1;
        while (b >= 0 && c[b] > cc) {
          c[b + 1] = c[b];
          v[b + 1] = v[b];
          --b;
        }
        c[b + 1] = cc;
        v[b + 1] = vv;
      }
    }
    // global CSC pattern by counting sort over columns (rows visited ascending => sorted columns);
    // owned values in the same order restricted to the owned rows
    P->colptr.assign((size_t)n + 1, 0);
    for (size_t e = 0; e < rc.size(); ++e) P->colptr[(size_t)rc[e] + 1]++;
    for (int j = 0; j < n; ++j) P->colptr[(size_t)j + 1] += P->colptr[j];
    P->rowidx.resize(rc.size());
    {
      std::vector<int> cur(P->colptr.begin(), P->colptr.end() - 1);
      for (int r = 0; r < mg; ++r)
        for (int k = 0; k < 8; ++k) P->rowidx[cur[rc[(size_t)r * 8 + k]]++] = r;
    }
    std::vector<int> lptr((size_t)n + 1, 0);
    for (int r = rb; r < re; ++r)
      for (int k = 0; k < 8; ++k) lptr[(size_t)rc[(size_t)r * 8 + k] + 1]++;
    for (int j = 0; j < n; ++j) lptr[(size_t)j + 1] += lptr[j];
    std::vector<double> cv((size_t)ml * 8);
    {
      std::vector<int> cur(lptr.begin(), lptr.end() - 1);
      for (int r = rb; r < re; ++r)
        for (int k = 0; k < 8; ++k) cv[cur[rc[(size_t)r * 8 + k]]++] = rv[(size_t)(r - rb) * 8 + k];
    }
    P->nnzJ = (int64_t)ml * 8;
    // ELL (k-major) copies of the owned rows for the residual kernel
    std::vector<int> ec((size_t)ml * 8);
    std::vector<double> evv((size_t)ml * 8);
    for (int r = 0; r < ml; ++r)
      for (int k = 0; k < 8; ++k) {
        ec[(size_t)k * ml + r] = rc[(size_t)(r + rb) * 8 + k];
        evv[(size_t)k * ml + r] = rv[(size_t)r * 8 + k];
      }
    SPB_CUDA(P->d_ell_col.upload(ec, h->stream));
    SPB_CUDA(P->d_ell_val.upload(evv, h->stream));
    SPB_CUDA(P->d_csc_val.upload(cv, h->stream));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
  } catch (const CudaFail& e) {
    h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    delete P;
    return SPB_ERR_CUDA;
  }
  *out = P;
  return SPB_OK;
}

spb_status spb_problem_lf(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, spb_problem_t* out) {
  return build_lf(h, g, rows_mult, seed, 0, 0, false, out);
}

spb_status spb_problem_lf_rows(spb_handle h, int32_t g, int32_t rows_mult, uint64_t seed, int32_t row_begin, int32_t row_end,
                               spb_problem_t* out) {
  return build_lf(h, g, rows_mult, seed, row_begin, row_end, true, out);
}

spb_status spb_problem_rosenbrock(spb_handle h, int32_t nblocks, const double* x0, spb_problem_t* out) {
  if (!h || !out || nblocks < 1) return SPB_ERR_ARG;
  spb_problem* P = new spb_problem;
  P->h = h;
  P->kind = 1;
  P->nblocks = nblocks;
  P->n = 2 * nblocks;
  P->m = 2 * nblocks;
  P->nnzJ = 4LL * nblocks;
  P->x0.resize(P->n);
  for (int b = 0; b < nblocks; ++b) {
    P->x0[2 * b] = x0 ? x0[2 * b] : -100.0;
    P->x0[2 * b + 1] = x0 ? x0[2 * b + 1] : 100.0;
  }
  P->colptr.resize((size_t)P->n + 1);
  P->rowidx.resize((size_t)P->nnzJ);
  for (int j = 0; j <= P->n; ++j) P->colptr[j] = 2 * j;
  for (int b = 0; b < nblocks; ++b) {
    P->rowidx[4 * b + 0] = 2 * b;
    P->rowidx[4 * b + 1] = 2 * b + 1;
    P->rowidx[4 * b + 2] = 2 * b;
    P->rowidx[4 * b + 3] = 2 * b + 1;
  }
  *out = P;
  return SPB_OK;
}

spb_status spb_problem_traits(spb_problem_t P, spb_traits* T) {
  if (!P || !T) return SPB_ERR_ARG;
  T->xSize = P->n;
  T->ESize = P->row_end > P->row_begin ? P->m_global : P->m;
  T->row_begin = P->row_begin;
  T->row_end = P->row_end;
  T->colptr = P->colptr.data();
  T->rowidx = P->rowidx.data();
  T->pattern_version = (uint64_t)(uintptr_t)P | 1u;  // pattern is immutable for the problem's lifetime
  T->mem = SPB_DEVICE;
  T->user = P;
  T->initial_solution = any_x0;
  T->objective_jacobian = P->kind == 0 ? lf_objjac : rb_objjac;
  T->pre_iteration = nullptr;
  T->post_iteration = nullptr;
  return SPB_OK;
}

spb_status spb_problem_destroy(spb_problem_t P) {
  if (P) {
    cudaSetDevice(P->h->device);
    delete P;
  }
  return SPB_OK;
}

}  // extern "C"
