// The LinearSolver concept surface of B200SolverWrapper, exercised the way the reference's callers use
// EigenSolverWrapper (include/SaddlePoint/EigenSolverWrapper.h:26-78):
//   factorize(A)                         :54   (explicit symmetric matrix)
//   solve(const VectorXd&, VectorXd&)    :72
//   solve(const MatrixXd&, MatrixXd&)    :62   (one solve per column)
//   public members `solver`, `A`         :26,28
// plus the failure behaviour of the drop-in: an asymmetric pattern is refused, a Jacobian whose pattern
// changes in the middle of LMSolver::solve surfaces as a C++ exception AFTER the C call has returned
// (nothing is thrown across the C ABI), and a user exception thrown inside the traits does the same.
//
//   g++ -O2 -std=c++17 -Iinclude examples/wrapper_concept.cpp -Lsaddlepoint_b200/lib
//     -lsaddlepoint_b200 -Wl,-rpath,$PWD/saddlepoint_b200/lib -o wrapper_concept
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "saddlepoint_b200/LMSolver.hpp"

typedef SaddlePoint::B200SolverWrapper LinearSolver;
typedef SaddlePoint::shim::Vector Vec;
typedef SaddlePoint::shim::Matrix Mat;
typedef SaddlePoint::shim::SparseMatrix SpMat;
typedef SaddlePoint::shim::Triplet Trip;

// 1-D Laplacian + 3 I: symmetric positive definite, tridiagonal
static SpMat laplacian(int n) {
  std::vector<Trip> t;
  for (int i = 0; i < n; ++i) {
    t.emplace_back(i, i, 5.0);
    if (i > 0) t.emplace_back(i, i - 1, -1.0);
    if (i + 1 < n) t.emplace_back(i, i + 1, -1.0);
  }
  SpMat A(n, n);
  A.setFromTriplets(t.begin(), t.end());
  return A;
}
static double residual(int n, const double* x, const double* b) {
  double worst = 0.0;
  for (int i = 0; i < n; ++i) {
    double r = 5.0 * x[i] - (i > 0 ? x[i - 1] : 0.0) - (i + 1 < n ? x[i + 1] : 0.0) - b[i];
    worst = std::fmax(worst, std::fabs(r));
  }
  return worst;
}

// f_i = x_i - 1 (i < n), plus one coupling row whose column changes when `drift` is set: same nnz, other pattern
struct DriftingTraits {
  int xSize = 8, ESize = 9;
  int calls = 0;
  bool drift = false, thrower = false;
  void initial_solution(Vec& x0) {
    x0.resize(xSize);
    x0.setConstant(3.0);
  }
  void pre_iteration(const Vec&) {}
  bool post_iteration(const Vec&) { return false; }
  void objective_jacobian(const Vec& x, Vec& E, SpMat& J, bool computeJacobian) {
    E.resize(ESize);
    for (int i = 0; i < xSize; ++i) E(i) = x(i) - 1.0;
    E(xSize) = 0.5 * (x(0) - x(1));
    if (!computeJacobian) return;
    ++calls;
    if (thrower && calls >= 2) throw std::logic_error("traits failure");
    const int other = (drift && calls >= 2) ? 2 : 1;  // second Jacobian: entry (8,1) moves to (8,2)
    std::vector<Trip> t;
    for (int i = 0; i < xSize; ++i) t.emplace_back(i, i, 1.0);
    t.emplace_back(xSize, 0, 0.5);
    t.emplace_back(xSize, other, -0.5);
    J.resize(ESize, xSize);
    J.setFromTriplets(t.begin(), t.end());
  }
};

int main(int argc, char* argv[]) {
  const int solver = argc > 1 ? atoi(argv[1]) : SPB_SOLVER_CHOLESKY;
  const int n = 300;
  int failures = 0;
  LinearSolver ls(solver);
  SpMat A = laplacian(n);
  if (!ls.factorize(A)) ++failures;
  if (ls.A.rows() != n || ls.A.cols() != n || ls.A.nonZeros() != A.nonZeros() || ls.solver != ls.handle()) ++failures;
  Vec b(n), x;
  for (int i = 0; i < n; ++i) b(i) = std::sin(0.1 * i);
  if (!ls.solve(b, x) || x.size() != n) ++failures;
  const double r1 = residual(n, x.data(), b.data());
  Mat B(n, 3), X;
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < n; ++i) B(i, c) = std::cos(0.05 * i * (c + 1));
  if (!ls.solve(B, X) || X.rows() != n || X.cols() != 3) ++failures;
  double r2 = 0.0;
  for (int c = 0; c < 3; ++c) r2 = std::fmax(r2, residual(n, X.data() + (size_t)c * n, B.data() + (size_t)c * n));
  std::printf("SOLVE vec_residual=%.3e mat_residual=%.3e status=%d\n", r1, r2, (int)ls.last_status());
  if (!(r1 < 1e-12) || !(r2 < 1e-12)) ++failures;

  // an asymmetric pattern is refused (Eigen would silently read the lower triangle only)
  {
    std::vector<Trip> t;
    for (int i = 0; i < 4; ++i) t.emplace_back(i, i, 2.0);
    t.emplace_back(2, 0, 1.0);  // no (0,2)
    SpMat U(4, 4);
    U.setFromTriplets(t.begin(), t.end());
    const bool ok = ls.factorize(U);
    std::printf("ASYMMETRIC refused=%d status=%d msg=%s\n", (int)!ok, (int)ls.last_status(), ls.last_error().c_str());
    if (ok || ls.last_status() != SPB_ERR_ARG) ++failures;
  }

  // pattern drift and a throwing traits class inside LMSolver::solve
  for (int mode = 0; mode < 3; ++mode) {
    DriftingTraits tr;
    tr.drift = mode == 1;
    tr.thrower = mode == 2;
    LinearSolver lsol(solver);
    SaddlePoint::DiagonalDamping<DriftingTraits> damp(0.01);
    SaddlePoint::LMSolver<LinearSolver, DriftingTraits, SaddlePoint::DiagonalDamping<DriftingTraits> > lm;
    lm.init(&lsol, &tr, &damp, 50);
    const char* what = "";
    bool threw = false;
    try {
      lm.solve(false);
    } catch (const std::exception& e) {
      threw = true;
      what = e.what();
      std::printf("LM mode=%d threw=1 what=%s\n", mode, what);
    }
    if (!threw) std::printf("LM mode=%d threw=0 x0=%.12f iterations=%d\n", mode, lm.x(0), lm.currIter);
    if (threw != (mode != 0)) ++failures;
    if (mode == 0 && std::fabs(lm.x(0) - 1.0) > 1e-8) ++failures;
    // the handle is still usable after a parked failure
    if (mode != 0 && !lsol.factorize(A)) ++failures;
  }
  std::printf("RESULT failures=%d\n", failures);
  return failures == 0 ? 0 : 1;
}
