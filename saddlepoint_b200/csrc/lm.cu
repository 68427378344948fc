// lm.cu -- the Levenberg-Marquardt loop of LMSolver::solve (LMSolver.h:86-191) with
// DiagonalDamping (DiagonalDamping.h:24-91), state resident on the device.
//
// Control flow, stop tests and their quirks are the reference's:
//   A  the first-order-optimality break only fires when verbose (LMSolver.h:123-129)
//   B  the function-tolerance break falls through to `return false` (LMSolver.h:161-169,190)
//   C  the loop runs while currIter<=maxIterations (LMSolver.h:188)
//   E  lambda /= 10 iff prev>new && new!=inf, else *= 10 (DiagonalDamping.h:65-68)
//   F,G damping enters H as fl(sqrt(d))^2 from the same J as rhs (assembly.cu)
// What changes is where the arithmetic runs: gradient, ||.||_inf, damping and the normal
// matrix are one fused assembly (assembly.cu); factorize/solve are the GPU LinearSolver;
// energies are deterministic device reductions; only a handful of scalars cross PCIe per
// iteration.  With dedupe_evaluations the three redundant objective evaluations per
// iteration (LMSolver.h:154 and DiagonalDamping.h:58-61 re-evaluate points whose residuals
// are already known bit-for-bit) are skipped; the trajectory is unchanged.
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

#include "spb_internal.h"

using namespace spb;

namespace {
struct HostPinned {
  double* p = nullptr;
  size_t n = 0;
  ~HostPinned() {
    if (p) cudaFreeHost(p);
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    SPB_CUDA(cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(double)));
    n = count;
  }
};

}  // namespace

namespace spb {
// Buffers of the LM loop, cached on the handle so repeated solves (the seamless-integration
// driver re-inits and re-solves in a loop, benchmark/seamless_integration/main.cpp:69-80)
// pay for allocation once.
struct LmWorkspace {
  DevBuf<double> d_prevx, d_x, d_dir, d_trial, d_E, d_Etrial, d_Jv;
  HostPinned hx, hE, hJ, hx0;
  uint64_t pattern_version = 0;
  const void* pattern_ptr = nullptr;
};
void lm_release(spb_context* h) {
  delete (LmWorkspace*)h->lm_ws;
  h->lm_ws = nullptr;
}
}  // namespace spb

namespace {
// Evaluates the traits at a device-resident x; leaves EVec (and J values) on the device.
struct Evaluator {
  spb_context* h;
  const spb_traits* T;
  int n, m;
  int64_t nnzJ;
  HostPinned &hx, &hE, &hJ;
  void init() {
    if (T->mem == SPB_HOST) {
      hx.alloc(n);
      hE.alloc(m);
      hJ.alloc((size_t)nnzJ);
    }
  }
  void eval(const double* d_x, double* d_E, double* d_Jv /*nullptr: no Jacobian*/) {
    cudaStream_t st = h->stream;
    if (T->mem == SPB_DEVICE) {
      stage_begin(h, ST_TRAITS);
      T->objective_jacobian(T->user, d_x, d_E, d_Jv, (void*)st);
      stage_end(h, ST_TRAITS, 0);
    } else {
      SPB_CUDA(cudaMemcpyAsync(hx.p, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
      SPB_CUDA(cudaStreamSynchronize(st));
      T->objective_jacobian(T->user, hx.p, hE.p, d_Jv ? hJ.p : nullptr, nullptr);
      SPB_CUDA(cudaMemcpyAsync(d_E, hE.p, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, st));
      if (d_Jv) SPB_CUDA(cudaMemcpyAsync(d_Jv, hJ.p, (size_t)nnzJ * sizeof(double), cudaMemcpyHostToDevice, st));
      SPB_CUDA(cudaStreamSynchronize(st));
    }
  }
  // pre/post iteration hooks always see what the traits' memory space expects
  const double* view(const double* d_x) {
    if (T->mem == SPB_DEVICE) return d_x;
    SPB_CUDA(cudaMemcpyAsync(hx.p, d_x, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    SPB_CUDA(cudaStreamSynchronize(h->stream));
    return hx.p;
  }
};
}  // namespace

extern "C" {

void spb_lm_default_options(spb_lm_options* o) {
  if (!o) return;
  o->maxIterations = 100;
  o->funcTolerance = 10e-10;
  o->fooTolerance = 10e-10;
  o->lambda0 = 0.01;
  o->verbose = 0;
  o->dedupe_evaluations = 0;
  o->fixed_iterations = 0;
}

spb_status spb_lm_solve(spb_handle h, const spb_traits* T, const spb_lm_options* O, spb_lm_result* R, double* x_out, int x_mem,
                        double* trace, int32_t trace_cap, int32_t* trace_len) {
  if (!h || !T || !O || !R || !T->objective_jacobian || !T->initial_solution) return SPB_ERR_ARG;
  memset(R, 0, sizeof(*R));
  if (trace_len) *trace_len = 0;
  try {
    cudaGetLastError();
    SPB_CUDA(cudaSetDevice(h->device));
    auto t_start = std::chrono::steady_clock::now();
    // Row-partitioned traits (multi-GPU, one big problem): ESize/colptr/rowidx describe the GLOBAL
    // Jacobian, this rank evaluates and owns residual rows [row_begin,row_end) only.
    const bool part = T->row_end > T->row_begin;
    const int n = T->xSize, m = part ? (T->row_end - T->row_begin) : T->ESize;
    // pattern: analyse once, re-analyse only when the fingerprint changes
    if (!h->lm_ws) h->lm_ws = new LmWorkspace;
    LmWorkspace& W = *(LmWorkspace*)h->lm_ws;
    {
      // a caller-provided pattern_version lets repeated solves skip re-hashing the arrays
      int same = 0;
      if (T->pattern_version != 0 && h->analyzed && W.pattern_version == T->pattern_version && W.pattern_ptr == (const void*)T->colptr &&
          h->m == m && h->n == n && h->partitioned == part && (!part || (h->row_begin == T->row_begin && h->row_end == T->row_end))) {
        same = 1;
      } else if (!part) {
        spb_pattern_matches(h, m, n, T->colptr, T->rowidx, &same);
        if (h->partitioned) same = 0;
      }
      W.pattern_version = T->pattern_version;
      W.pattern_ptr = (const void*)T->colptr;
      if (!same) {
        auto ta = std::chrono::steady_clock::now();
        spb_status s = part ? spb_analyze_partitioned(h, T->ESize, n, T->colptr, T->rowidx, T->row_begin, T->row_end)
                            : spb_analyze(h, m, n, T->colptr, T->rowidx);
        if (s != SPB_OK) return s;
        R->seconds_analyze = std::chrono::duration<double>(std::chrono::steady_clock::now() - ta).count();
      }
    }
    cudaStream_t st = h->stream;
    const int64_t nnzJ = h->nnzJ;
    DevBuf<double>&d_prevx = W.d_prevx, &d_x = W.d_x, &d_dir = W.d_dir, &d_trial = W.d_trial, &d_E = W.d_E, &d_Etrial = W.d_Etrial,
                  &d_Jv = W.d_Jv;
    SPB_CUDA(d_prevx.alloc(n));
    SPB_CUDA(d_x.alloc(n));
    SPB_CUDA(d_dir.alloc(n));
    SPB_CUDA(d_trial.alloc(n));
    SPB_CUDA(d_E.alloc(m));
    SPB_CUDA(d_Etrial.alloc(m));
    SPB_CUDA(d_Jv.alloc((size_t)nnzJ));
    Evaluator ev{h, T, n, m, nnzJ, W.hx, W.hE, W.hJ};
    ev.init();
    {
      HostPinned& x0 = W.hx0;
      x0.alloc(n);
      T->initial_solution(T->user, x0.p);
      SPB_CUDA(cudaMemcpyAsync(d_prevx.p, x0.p, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
      SPB_CUDA(cudaMemcpyAsync(d_x.p, d_prevx.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      SPB_CUDA(cudaStreamSynchronize(st));
    }
    const bool verbose = O->verbose != 0;
    const bool dedupe = O->dedupe_evaluations != 0;
    const bool fixed = O->fixed_iterations > 0;
    double lambda = O->lambda0;
    int currIter = 0, ntrace = 0;
    double energy = 0.0, foo = 0.0;
    double cachedPrevEnergy = 0.0;
    bool havePrevEnergy = false;
    int exit_path = 6, ret = 0;
    auto copy_vec = [&](double* dst, const double* src) {
      SPB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    };

    ev.eval(d_prevx.p, d_E.p, d_Jv.p);  // LMSolver.h:106 ; DT->init is folded into the assembly (lambda)
    int bodies = 0;
    do {
      if (T->pre_iteration) T->pre_iteration(T->user, ev.view(d_prevx.p));
      // LMSolver.h:117-119 + DiagonalDamping rebuild + LMSolver.h:136's product, fused
      run_assembly(h, d_Jv.p, d_E.p, lambda, &foo);
      if (!fixed && foo < O->fooTolerance) {
        copy_vec(d_x.p, d_prevx.p);
        if (verbose) {
          exit_path = 1;
          break;
        }
      }
      spb_status fs = spb_factorize(h);
      R->factorizations++;
      if (fs == SPB_NOT_SPD) {
        exit_path = 2;
        ret = 0;
        goto done;
      }
      if (fs != SPB_OK) return fs;
      int lin_it = 0;
      spb_status ss = spb_solve(h, nullptr, d_dir.p, 1, SPB_DEVICE, &lin_it);
      R->linear_iterations += lin_it;
      if (ss == SPB_NOT_SPD) {  // PCG met p.Hp<=0: the iterative analogue of a failed factorisation
        exit_path = 2;
        ret = 0;
        goto done;
      }
      if (ss != SPB_OK && ss != SPB_NOT_CONVERGED) return ss;
      double dn2 = 0.0;
      dev_squared_norm(h, d_dir.p, n, &dn2);
      const double dn = std::sqrt(dn2);
      if (!fixed && dn < O->funcTolerance) {
        copy_vec(d_x.p, d_prevx.p);
        ret = 1;
        exit_path = 3;
        goto done;
      }
      double prevEnergy2, newEnergy2;
      if (dedupe && havePrevEnergy) {
        prevEnergy2 = cachedPrevEnergy;
      } else {
        if (!dedupe) ev.eval(d_prevx.p, d_E.p, nullptr);  // LMSolver.h:154 (same values as already held)
        dev_squared_norm_global(h, d_E.p, m, &prevEnergy2);
      }
      dev_add(h, d_prevx.p, d_dir.p, d_trial.p, n);
      ev.eval(d_trial.p, d_Etrial.p, nullptr);  // LMSolver.h:156
      dev_squared_norm_global(h, d_Etrial.p, m, &newEnergy2);
      energy = newEnergy2;
      const bool accepted = prevEnergy2 > newEnergy2;
      if (trace && ntrace < trace_cap) {
        double* t = trace + (size_t)ntrace * 8;
        t[0] = prevEnergy2; t[1] = newEnergy2; t[2] = lambda; t[3] = foo; t[4] = dn; t[5] = accepted ? 1.0 : 0.0;
        t[6] = (double)lin_it; t[7] = 0.0;
        ++ntrace;
      }
      if (accepted) {
        copy_vec(d_x.p, d_trial.p);
        if (!fixed && std::fabs(prevEnergy2 - newEnergy2) < O->funcTolerance) {
          exit_path = 4;
          break;
        }
      } else {
        copy_vec(d_x.p, d_prevx.p);
      }
      // LMSolver.h:174-175: re-linearise at x
      if (!dedupe) {
        ev.eval(d_x.p, d_E.p, d_Jv.p);
        dev_squared_norm_global(h, d_E.p, m, &energy);
      } else if (accepted) {
        ev.eval(d_x.p, d_E.p, d_Jv.p);  // same residual as the trial evaluation, plus the new Jacobian
        energy = newEnergy2;
      } else {
        energy = prevEnergy2;  // x==prevx: residual and Jacobian already current
      }
      cachedPrevEnergy = energy;
      havePrevEnergy = true;
      // DT->update (DiagonalDamping.h:58-68); the dampJ rebuild is the next assembly
      {
        double pe = prevEnergy2, ne = newEnergy2;
        if (!dedupe) {
          ev.eval(d_prevx.p, d_Etrial.p, nullptr);
          dev_squared_norm_global(h, d_Etrial.p, m, &pe);
          dev_add(h, d_prevx.p, d_dir.p, d_trial.p, n);
          ev.eval(d_trial.p, d_Etrial.p, nullptr);
          dev_squared_norm_global(h, d_Etrial.p, m, &ne);
        }
        if ((pe > ne) && (ne != std::numeric_limits<double>::infinity())) lambda /= 10.0; else lambda *= 10.0;
      }
      if (T->post_iteration && T->post_iteration(T->user, ev.view(d_x.p))) {
        ret = 1;
        exit_path = 5;
        goto done;
      }
      currIter++;
      copy_vec(d_prevx.p, d_x.p);
      ++bodies;
      if (fixed && bodies >= O->fixed_iterations) break;
    } while (fixed || currIter <= O->maxIterations);
  done:
    if (x_out) {
      SPB_CUDA(cudaMemcpyAsync(x_out, d_x.p, (size_t)n * sizeof(double),
                               x_mem == SPB_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    SPB_CUDA(cudaStreamSynchronize(st));
    R->ret = ret;
    R->exit_path = exit_path;
    R->currIter = currIter;
    R->energy = energy;
    R->fooOptimality = foo;
    R->lambda = lambda;
    R->seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    if (trace_len) *trace_len = ntrace;
    return SPB_OK;
  } catch (const CudaFail& e) {
    h->err = std::string("CUDA error: ") + cudaGetErrorString(e.e) + " at " + e.where;
    return SPB_ERR_CUDA;
  } catch (const NcclFail& e) {
    h->err = std::string("NCCL error ") + std::to_string(e.code) + " at " + e.where;
    return SPB_ERR_NCCL;
  } catch (const ArgFail& e) {
    h->err = e.msg;
    return SPB_ERR_ARG;
  } catch (const StateFail& e) {
    h->err = e.msg;
    return SPB_ERR_STATE;
  }
}

}  // extern "C"
