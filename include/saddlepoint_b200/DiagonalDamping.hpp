// DiagonalDamping.hpp -- the DampingTraits concept class.
//
// Same public surface as reference include/SaddlePoint/DiagonalDamping.h:20-96: public
// currLambda (callers poke it, benchmark/seamless_integration/main.cpp:73), init(), update(),
// ctor default 0.01.  The Marquardt scaling d_j = lambda * sum_r J(r,j)^2 and the augmented
// rows sqrt(d_j) are produced inside the fused GPU assembly (spb_assemble), so dampJ is never
// materialised: init()/update() leave it untouched.  update() still applies the reference's
// lambda rule (DiagonalDamping.h:58-68) for callers that drive the loop themselves.
#ifndef SADDLEPOINT_B200_DIAGONAL_DAMPING_HPP
#define SADDLEPOINT_B200_DIAGONAL_DAMPING_HPP
#include <iostream>
#include <limits>

namespace SaddlePoint {

template <class SolverTraits>
class DiagonalDamping {
 public:
  double currLambda;

  template <class SparseMat, class Vec>
  void init(const SparseMat& /*J*/, const Vec& /*initSolution*/, const bool verbose, SparseMat& /*dampJ*/) {
    if (verbose) std::cout << "Initial Lambda: " << currLambda << std::endl;
  }

  template <class SparseMat, class Vec>
  bool update(SolverTraits& ST, const SparseMat& /*J*/, const Vec& currSolution, const Vec& direction, const bool verbose,
              SparseMat& /*dampJ*/) {
    Vec EVec, trial;
    SparseMat stubJ;
    ST.objective_jacobian(currSolution, EVec, stubJ, false);
    double prevEnergy2 = 0.0;
    for (long i = 0; i < (long)EVec.size(); ++i) prevEnergy2 += EVec.data()[i] * EVec.data()[i];
    trial.resize(currSolution.size());
    for (long i = 0; i < (long)trial.size(); ++i) trial.data()[i] = currSolution.data()[i] + direction.data()[i];
    ST.objective_jacobian(trial, EVec, stubJ, false);
    double newEnergy2 = 0.0;
    for (long i = 0; i < (long)EVec.size(); ++i) newEnergy2 += EVec.data()[i] * EVec.data()[i];
    if ((prevEnergy2 > newEnergy2) && (newEnergy2 != std::numeric_limits<double>::infinity())) currLambda /= 10.0;
    else currLambda *= 10.0;
    if (verbose) std::cout << "Current Lambda: " << currLambda << std::endl;
    return prevEnergy2 > newEnergy2;
  }

  DiagonalDamping(double _currLambda = 0.01) : currLambda(_currLambda) {}
  ~DiagonalDamping() {}
};

}  // namespace SaddlePoint
#endif
