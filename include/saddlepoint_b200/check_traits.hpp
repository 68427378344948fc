// check_traits.hpp -- SaddlePoint::check_traits(Traits, CurrSolution) on the GPU solver library.
//
// Same job and the same printed report as reference include/SaddlePoint/check_traits.h:19-83
// (central differences with step 10e-5, FE entries below 10e-7 dropped, every discrepancy above
// 10e-7 listed, then "Maximum gradient difference" and its location), but columns of the Jacobian
// that share no residual row are perturbed together: 2*ncolors objective evaluations instead of
// 2*xSize.  Needs a handle: pass the LinearSolver the problem will be solved with.
// Extra over the reference: a finite-difference response in a row that has no entry in any of the
// perturbed columns means the declared pattern misses a derivative; it is reported as well.
#ifndef SADDLEPOINT_B200_CHECK_TRAITS_HPP
#define SADDLEPOINT_B200_CHECK_TRAITS_HPP
#include <cmath>
#include <cstring>
#include <iostream>
#include <stdexcept>
#include <vector>

#include "LMSolver.hpp"

namespace SaddlePoint {

template <class SolverTraits, class LinearSolver, class VectorT = DefaultVector, class SparseT = DefaultSparse>
spb_check_result check_traits(SolverTraits& Traits, const VectorT& CurrSolution, LinearSolver& LS) {
  using namespace std;
  cout << "WARNING: FE gradient checking, reducing performance!!!" << endl;
  cout << "******************************************************" << endl;
  cout << "Solution Size: " << Traits.xSize << endl;
  struct Ctx {
    SolverTraits* ST;
  } ctx{&Traits};
  VectorT EVec;
  SparseT TraitGradient;
  Traits.objective_jacobian(CurrSolution, EVec, TraitGradient, true);  // for the pattern and the report
  spb_traits T;
  memset(&T, 0, sizeof(T));
  T.xSize = (int32_t)Traits.xSize;
  T.ESize = (int32_t)TraitGradient.rows();
  T.colptr = TraitGradient.outerIndexPtr();
  T.rowidx = TraitGradient.innerIndexPtr();
  T.mem = SPB_HOST;
  T.user = &ctx;
  T.initial_solution = [](void*, double*) {};
  T.objective_jacobian = [](void* user, const double* xp, double* Ep, double* Jp, void*) {
    SolverTraits* ST = ((Ctx*)user)->ST;
    VectorT xv, E;
    SparseT J;
    xv.resize(ST->xSize);
    memcpy(xv.data(), xp, sizeof(double) * (size_t)ST->xSize);
    ST->objective_jacobian(xv, E, J, Jp != nullptr);
    memcpy(Ep, E.data(), sizeof(double) * (size_t)E.size());
    if (Jp) memcpy(Jp, J.valuePtr(), sizeof(double) * (size_t)J.nonZeros());
  };
  spb_check_result R;
  vector<double> fe((size_t)TraitGradient.nonZeros());
  spb_status s = spb_check_traits(LS.handle(), &T, CurrSolution.data(), SPB_HOST, 0.0, 0.0, &R, fe.data());
  if (s != SPB_OK) throw runtime_error(string("spb_check_traits failed: ") + spb_last_error(LS.handle()));
  const double* ours = TraitGradient.valuePtr();
  for (long j = 0; j < (long)Traits.xSize; ++j)
    for (long q = T.colptr[j]; q < T.colptr[j + 1]; ++q) {
      const double f = fabs(fe[q]) > 10e-7 ? fe[q] : 0.0;
      if (fabs(f - ours[q]) > 10e-7) {
        cout << "Gradient Discrepancy at: (" << T.rowidx[q] << "," << j << ") of " << f - ours[q] << endl;
        cout << "Our Value: " << ours[q] << endl;
        cout << "Computed Value: " << f << endl;
      }
    }
  if (R.outside_count > 0)
    cout << "Derivatives outside the declared Jacobian pattern: " << R.outside_count << " responses, largest " << R.outside_max << endl;
  cout << "Maximum gradient difference: " << R.max_diff << endl;
  cout << "At Location: (" << R.max_row << "," << R.max_col << ")" << endl;
  return R;
}

}  // namespace SaddlePoint
#endif
