// A Gauss-Newton loop written against the reference's LEGACY LinearSolver concept -- the one
// GNSolver.h:128-175 and SLSolver.h:121-166 use: the normal matrix is handed over as triplets,
//   LS->analyze(HRows, HCols, true)   once,   LS->factorize(HVals, true) + LS->solve(rhs, d)   per step,
// upper triangle only, one triplet per (residual row, column pair) product, duplicates summed by
// the linear solver.  Runs SaddlePoint::B200SolverWrapper through exactly those calls.
//
// Problem: Broyden tridiagonal system as least squares, f_i = (3 - 2 x_i) x_i - x_{i-1} - 2 x_{i+1} + 1,
// x0 = -1 (More, Garbow, Hillstrom #30); the optimum has zero residual.
//
//   g++ -O2 -std=c++17 -Iinclude examples/gauss_newton_legacy.cpp -Lsaddlepoint_b200/lib
//     -lsaddlepoint_b200 -Wl,-rpath,$PWD/saddlepoint_b200/lib -o gn_legacy
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "saddlepoint_b200/B200SolverWrapper.hpp"

struct BroydenTraits {
  int xSize = 0;
  std::vector<int> JRows, JCols;  // sorted by row, as the legacy MatrixPattern requires
  std::vector<double> JVals, EVec;
  void init(int n) {
    xSize = n;
    for (int i = 0; i < n; ++i)
      for (int j = i - 1; j <= i + 1; ++j)
        if (j >= 0 && j < n) {
          JRows.push_back(i);
          JCols.push_back(j);
        }
    JVals.resize(JRows.size());
    EVec.resize(n);
  }
  void update(const std::vector<double>& x) {
    size_t t = 0;
    for (int i = 0; i < xSize; ++i) {
      const double lo = i > 0 ? x[i - 1] : 0.0, hi = i + 1 < xSize ? x[i + 1] : 0.0;
      EVec[i] = (3.0 - 2.0 * x[i]) * x[i] - lo - 2.0 * hi + 1.0;
      if (i > 0) JVals[t++] = -1.0;
      JVals[t++] = 3.0 - 4.0 * x[i];
      if (i + 1 < xSize) JVals[t++] = -2.0;
    }
  }
};

int main(int argc, char* argv[]) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000;
  const int solver = argc > 2 ? atoi(argv[2]) : SPB_SOLVER_CHOLESKY;
  BroydenTraits ST;
  ST.init(n);
  // pattern of J^T J, upper triangle, one triplet per product; pair list (a, b) of J entries
  std::vector<int> HRows, HCols, pa, pb;
  for (size_t r0 = 0; r0 < ST.JRows.size();) {
    size_t r1 = r0;
    while (r1 < ST.JRows.size() && ST.JRows[r1] == ST.JRows[r0]) ++r1;
    for (size_t a = r0; a < r1; ++a)
      for (size_t b = r0; b < r1; ++b)
        if (ST.JCols[b] >= ST.JCols[a]) {
          HRows.push_back(ST.JCols[a]);
          HCols.push_back(ST.JCols[b]);
          pa.push_back((int)a);
          pb.push_back((int)b);
        }
    r0 = r1;
  }
  SaddlePoint::B200SolverWrapper LS(solver);
  if (!LS.analyze(HRows, HCols, true)) {
    std::printf("analyze failed: %s\n", LS.last_error().c_str());
    return 1;
  }
  std::vector<double> x(n, -1.0), HVals(HRows.size()), rhs(n), d;
  int it = 0;
  double energy = 0.0;
  for (; it < 50; ++it) {
    ST.update(x);
    energy = 0.0;
    for (double e : ST.EVec) energy += e * e;
    if (energy < 1e-26) break;
    for (size_t t = 0; t < HVals.size(); ++t) HVals[t] = ST.JVals[pa[t]] * ST.JVals[pb[t]];
    for (double& v : rhs) v = 0.0;
    for (size_t t = 0; t < ST.JRows.size(); ++t) rhs[ST.JCols[t]] += ST.JVals[t] * -ST.EVec[ST.JRows[t]];
    if (!LS.factorize(HVals, true)) {
      std::printf("Solver Failed to factorize!\n");
      return 2;
    }
    LS.solve(rhs, d);
    for (int i = 0; i < n; ++i) x[i] += d[i];
  }
  std::printf("RESULT iterations=%d energy=%.17g x0=%.17g xmid=%.17g xlast=%.17g\n", it, energy, x[0], x[n / 2], x[n - 1]);
  return 0;
}
