// The reference's Rosenbrock benchmark (benchmark/rosenbrock/main.cpp) on the GPU solver:
// same problem (VALLEY_COEFF 25, x0 = (-100,100), lambda0 = 0.01, maxIter 1000, verbose),
// traits written against the current objective_jacobian concept (LMSolver.h:106), built from
// (JRows, JCols, JVals) triplets with setFromTriplets semantics like check_traits.h:35-43.
//
//   g++ -O2 -std=c++17 -Iinclude examples/rosenbrock_main.cpp -Lsaddlepoint_b200/lib
//     -lsaddlepoint_b200 -Wl,-rpath,$PWD/saddlepoint_b200/lib -o rosenbrock_gpu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "saddlepoint_b200/LMSolver.hpp"
#include "saddlepoint_b200/check_traits.hpp"

typedef SaddlePoint::B200SolverWrapper LinearSolver;
typedef SaddlePoint::shim::Vector Vec;
typedef SaddlePoint::shim::SparseMatrix SpMat;

#define VALLEY_COEFF 25.0

class RosenbrockTraits {
 public:
  int xSize, ESize;
  int nblocks = 1;
  void init(int blocks) {
    nblocks = blocks;
    xSize = 2 * blocks;
    ESize = 2 * blocks;
  }
  void initial_solution(Vec& x0) {
    x0.resize(xSize);
    for (int b = 0; b < nblocks; ++b) {
      x0(2 * b) = -100;
      x0(2 * b + 1) = 100;
    }
  }
  void pre_iteration(const Vec&) {}
  bool post_iteration(const Vec&) { return false; }
  void objective_jacobian(const Vec& x, Vec& EVec, SpMat& J, bool computeJacobian) {
    EVec.resize(ESize);
    for (int b = 0; b < nblocks; ++b) {
      EVec(2 * b) = VALLEY_COEFF * (x(2 * b + 1) - x(2 * b) * x(2 * b));
      EVec(2 * b + 1) = 1.0 - x(2 * b);
    }
    if (!computeJacobian) return;
    std::vector<SaddlePoint::shim::Triplet> tris;
    for (int b = 0; b < nblocks; ++b) {
      tris.emplace_back(2 * b, 2 * b, -2.0 * VALLEY_COEFF * x(2 * b));
      tris.emplace_back(2 * b, 2 * b + 1, VALLEY_COEFF);
      tris.emplace_back(2 * b + 1, 2 * b, -1.0);
      tris.emplace_back(2 * b + 1, 2 * b + 1, 0.0);  // the reference keeps this structural zero
    }
    J.resize(ESize, xSize);
    J.setFromTriplets(tris.begin(), tris.end());
  }
};

int main(int argc, char* argv[]) {
  int blocks = argc > 1 ? atoi(argv[1]) : 1;
  int solver = argc > 2 ? atoi(argv[2]) : SPB_SOLVER_CHOLESKY;
  RosenbrockTraits slTraits;
  LinearSolver lSolver(solver);
  SaddlePoint::DiagonalDamping<RosenbrockTraits> dTraits(0.01);
  SaddlePoint::LMSolver<LinearSolver, RosenbrockTraits, SaddlePoint::DiagonalDamping<RosenbrockTraits> > lmSolver;
  slTraits.init(blocks);
  if (argc > 3) {  // gradient check first, as one does after coding a traits class (check_traits.h:17)
    Vec xc;
    slTraits.initial_solution(xc);
    spb_check_result chk = SaddlePoint::check_traits(slTraits, xc, lSolver);
    std::printf("CHECK max=%.17g discrepancies=%lld colors=%d evaluations=%d\n", chk.max_diff, (long long)chk.discrepancies, chk.ncolors,
                chk.evaluations);
  }
  lmSolver.init(&lSolver, &slTraits, &dTraits, 1000);
  bool ret = lmSolver.solve(true);
  std::printf("RESULT ret=%d currIter=%d factorizations=%d energy=%.17g lambda=%.17g x0=%.17g x1=%.17g\n", (int)ret, lmSolver.currIter,
              lmSolver.factorizations, lmSolver.energy, dTraits.currLambda, lmSolver.x(0), lmSolver.x(1));
  return 0;
}
