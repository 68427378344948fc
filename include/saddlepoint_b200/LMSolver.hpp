// LMSolver.hpp -- SaddlePoint::LMSolver<LinearSolver, SolverTraits, DampingTraits> on the GPU.
//
// Keeps the concept API of reference include/SaddlePoint/LMSolver.h:22-193:
//   init(LS*, ST*, DT*, maxIterations=100, funcTolerance=10e-10, fooTolerance=10e-10)
//   bool solve(verbose)          -- same return convention, incl. `false` on a normal
//                                   function-tolerance stop (LMSolver.h:161-169,190)
//   public x, x0, prevx, energy, fooOptimality, currIter, LS, ST, DT, tolerances
// and the SolverTraits concept it consumes (xSize, ESize, initial_solution,
// objective_jacobian(x, EVec, J, computeJacobian), pre_iteration, post_iteration).
// The loop itself runs inside libsaddlepoint_b200.so (spb_lm_solve): the user's traits stay on
// the host and are called through trampolines, everything else (gradient, damping, normal
// matrix, factorize/solve, energies) is on the device.
//
// DampingTraits: the damping arithmetic (DiagonalDamping.h:31-43,58-68,73-85) is fused into the GPU assembly, so
// only DT->currLambda is read (at solve entry) and written back (at exit); a user-supplied DampingTraits::update
// with a different rule is NOT consulted -- use SaddlePoint::DiagonalDamping (or the reference's own LMSolver.h
// with B200SolverWrapper as its LinearSolver) for custom damping policies.
//
// Type parameters VectorT / SparseT default to Eigen's when SPB_USE_EIGEN is defined, else to
// the shims in SparseTypes.hpp; any type with Eigen's raw accessors works.
#ifndef SADDLEPOINT_B200_LM_SOLVER_HPP
#define SADDLEPOINT_B200_LM_SOLVER_HPP
#include <cstring>
#include <exception>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "../saddlepoint_b200.h"
#include "B200SolverWrapper.hpp"
#include "DiagonalDamping.hpp"
#include "SparseTypes.hpp"
#ifdef SPB_USE_EIGEN
#include <Eigen/Core>
#include <Eigen/Sparse>
#endif

namespace SaddlePoint {

#ifdef SPB_USE_EIGEN
typedef Eigen::VectorXd DefaultVector;
typedef Eigen::SparseMatrix<double> DefaultSparse;
#else
typedef shim::Vector DefaultVector;
typedef shim::SparseMatrix DefaultSparse;
#endif

template <class LinearSolver, class SolverTraits, class DampingTraits, class VectorT = DefaultVector, class SparseT = DefaultSparse>
class LMSolver {
 public:
  VectorT x;      // current solution; always updated
  VectorT prevx;  // kept for API compatibility (the iterate history lives on the device)
  VectorT x0;     // the initial solution
  VectorT d;

  LinearSolver* LS;
  SolverTraits* ST;
  DampingTraits* DT;

  int maxIterations;
  double funcTolerance;
  double fooTolerance;

  double energy;
  double fooOptimality;
  int currIter;

  // extras (not in the reference)
  int factorizations = 0;
  long long linearIterations = 0;
  int exitPath = 0;
  bool dedupeEvaluations = false;  // true: skip the three bit-identical re-evaluations per iteration

  LMSolver() : LS(nullptr), ST(nullptr), DT(nullptr), maxIterations(100), funcTolerance(10e-10), fooTolerance(10e-10), energy(0), fooOptimality(0), currIter(0) {}

  void init(LinearSolver* _LS, SolverTraits* _ST, DampingTraits* _DT, int _maxIterations = 100, double _funcTolerance = 10e-10,
            double _fooTolerance = 10e-10) {
    LS = _LS;
    ST = _ST;
    DT = _DT;
    maxIterations = _maxIterations;
    funcTolerance = _funcTolerance;
    fooTolerance = _fooTolerance;
    d.resize(ST->xSize);
    x.resize(ST->xSize);
    x0.resize(ST->xSize);
    prevx.resize(ST->xSize);
  }

  bool solve(const bool verbose) {
    if (verbose) std::cout << "******Beginning Optimization******" << std::endl;
    Bridge br;
    br.self = this;
    ST->initial_solution(x0);
    // The pattern has to be known before the library can analyse it: evaluate once here and
    // hand this very evaluation to the loop's first objective_jacobian call (LMSolver.h:106), so
    // the user's traits see exactly the reference's call sequence.
    ST->objective_jacobian(x0, br.E0, br.J0, true);
    br.first_pending = true;
    spb_traits T;
    std::memset(&T, 0, sizeof(T));
    T.xSize = (int32_t)ST->xSize;
    T.ESize = (int32_t)ST->ESize;
    T.colptr = br.J0.outerIndexPtr();
    T.rowidx = br.J0.innerIndexPtr();
    T.pattern_version = 0;
    T.mem = SPB_HOST;
    T.user = &br;
    T.initial_solution = &Bridge::initial_solution;
    T.objective_jacobian = &Bridge::objective_jacobian;
    T.pre_iteration = &Bridge::pre_iteration;
    T.post_iteration = &Bridge::post_iteration;
    spb_lm_options O;
    spb_lm_default_options(&O);
    O.maxIterations = maxIterations;
    O.funcTolerance = funcTolerance;
    O.fooTolerance = fooTolerance;
    O.lambda0 = DT->currLambda;
    O.verbose = verbose ? 1 : 0;
    O.dedupe_evaluations = dedupeEvaluations ? 1 : 0;
    spb_lm_result R;
    x.resize(ST->xSize);
    spb_status s = spb_lm_solve(LS->handle(), &T, &O, &R, x.data(), SPB_HOST, nullptr, 0, nullptr);
    // Nothing is thrown across the C ABI: a failure inside a trampoline (the user's traits threw, or the Jacobian
    // pattern changed mid-solve) is parked in the bridge, the loop is told to stop at its next post_iteration, and
    // the exception continues from here, after spb_lm_solve has returned normally.
    if (br.error) std::rethrow_exception(br.error);
    if (s != SPB_OK) throw std::runtime_error(std::string("spb_lm_solve failed: ") + spb_last_error(LS->handle()));
    energy = R.energy;
    fooOptimality = R.fooOptimality;
    currIter = R.currIter;
    DT->currLambda = R.lambda;
    factorizations = R.factorizations;
    linearIterations = R.linear_iterations;
    exitPath = R.exit_path;
    if (verbose) {
      static const char* why[] = {"", "First-order optimality has been reached", "Solver Failed to factorize! ",
                                  "Stopping since direction magnitude small.", "Stopping sincefunction didn't change above tolerance.",
                                  "ST->Post_iteration() gave a stop", "maximum iterations reached"};
      std::cout << why[R.exit_path] << std::endl;
      std::cout << "iterations: " << currIter << "  energy: " << energy << "  firstOrderOptimality: " << fooOptimality << std::endl;
    } else if (R.exit_path == 2) {
      std::cout << "Solver Failed to factorize! " << std::endl;  // printed unconditionally by the reference (LMSolver.h:138)
    }
    return R.ret != 0;
  }

 private:
  struct Bridge {
    LMSolver* self = nullptr;
    VectorT E0;
    SparseT J0;
    bool first_pending = false;
    std::exception_ptr error;  // first failure inside a trampoline; rethrown by solve() after the C call returns
    template <class F>
    void guarded(F&& f) {
      if (error) return;
      try {
        f();
      } catch (...) {
        error = std::current_exception();
      }
    }
    // same compressed pattern: sizes, nnz, outer and inner index arrays (not just the nonzero count)
    static bool same_pattern(const SparseT& A, const SparseT& B) {
      if (A.rows() != B.rows() || A.cols() != B.cols() || A.nonZeros() != B.nonZeros()) return false;
      return std::memcmp(A.outerIndexPtr(), B.outerIndexPtr(), sizeof(int) * (size_t)(A.cols() + 1)) == 0 &&
             std::memcmp(A.innerIndexPtr(), B.innerIndexPtr(), sizeof(int) * (size_t)A.nonZeros()) == 0;
    }
    static void copy_in(VectorT& v, const double* p, long n) {
      v.resize(n);
      std::memcpy(v.data(), p, sizeof(double) * (size_t)n);
    }
    static void initial_solution(void* user, double* out) {
      Bridge* b = (Bridge*)user;
      std::memcpy(out, b->self->x0.data(), sizeof(double) * (size_t)b->self->ST->xSize);
    }
    static void objective_jacobian(void* user, const double* xp, double* Ep, double* Jp, void*) {
      Bridge* b = (Bridge*)user;
      SolverTraits* ST = b->self->ST;
      const long n = ST->xSize, m = ST->ESize;
      if (b->first_pending && Jp) {  // LMSolver.h:106: already evaluated in solve()
        b->first_pending = false;
        std::memcpy(Ep, b->E0.data(), sizeof(double) * (size_t)m);
        std::memcpy(Jp, b->J0.valuePtr(), sizeof(double) * (size_t)b->J0.nonZeros());
        return;
      }
      b->guarded([&] {
        VectorT xv, E;
        SparseT J;
        copy_in(xv, xp, n);
        ST->objective_jacobian(xv, E, J, Jp != nullptr);
        if ((long)E.size() != m) throw std::runtime_error("objective_jacobian returned an EVec of the wrong size");
        std::memcpy(Ep, E.data(), sizeof(double) * (size_t)m);
        if (Jp) {
          // the assembly plan belongs to the pattern analysed at solve() entry (the reference re-derives it every
          // iteration, LMSolver.h:136; within one solve its benchmarks never change it)
          if (!same_pattern(J, b->J0)) throw std::runtime_error("Jacobian pattern changed inside LMSolver::solve");
          std::memcpy(Jp, J.valuePtr(), sizeof(double) * (size_t)J.nonZeros());
        }
      });
      if (b->error) {  // keep the device work defined until the loop stops at post_iteration
        std::memset(Ep, 0, sizeof(double) * (size_t)m);
        if (Jp) std::memset(Jp, 0, sizeof(double) * (size_t)b->J0.nonZeros());
      }
    }
    static void pre_iteration(void* user, const double* xp) {
      Bridge* b = (Bridge*)user;
      b->guarded([&] {
        VectorT xv;
        copy_in(xv, xp, b->self->ST->xSize);
        b->self->ST->pre_iteration(xv);
      });
    }
    static int post_iteration(void* user, const double* xp) {
      Bridge* b = (Bridge*)user;
      int stop = 0;
      b->guarded([&] {
        VectorT xv;
        copy_in(xv, xp, b->self->ST->xSize);
        stop = b->self->ST->post_iteration(xv) ? 1 : 0;
      });
      return b->error ? 1 : stop;  // a parked failure ends the loop here
    }
  };
};

}  // namespace SaddlePoint
#endif
