// B200SolverWrapper.hpp -- the GPU LinearSolver concept class.
//
// Drop-in for SaddlePoint::EigenSolverWrapper<Eigen::SimplicialLLT<...>>
// (reference include/SaddlePoint/EigenSolverWrapper.h:23-79): same member functions, same
// bool results (factorize() == false  <=>  the matrix is not positive definite).
//   analyze(JPattern)    EigenSolverWrapper.h:30   (takes J's pattern; also analyze_pattern)
//   factorize(A)         EigenSolverWrapper.h:54   (explicit symmetric matrix, both triangles)
//   solve(rhs, x)        EigenSolverWrapper.h:72   (vector) / :62 (columns of a matrix)
// and the legacy triplet overloads the GN / SL solvers and EigenSingleSolveWrapper call:
//   analyze(HRows, HCols[, symmetric])   GNSolver.h:131, SLSolver.h:123, EigenSolverWrapper.h:118
//   factorize(HVals[, symmetric])        GNSolver.h:169, SLSolver.h:159, EigenSolverWrapper.h:119
// A caller that keeps the reference's own LMSolver.h can plug this class in unchanged: the
// product dampJ^T*dampJ it passes to factorize() is uploaded and factorised on the GPU.
// SaddlePoint::LMSolver in this directory goes further and never forms that product on the
// host (fused GPU assembly), using handle() directly.
#ifndef SADDLEPOINT_B200_SOLVER_WRAPPER_HPP
#define SADDLEPOINT_B200_SOLVER_WRAPPER_HPP
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

#include "../saddlepoint_b200.h"

namespace SaddlePoint {

namespace detail {
template <class T, class = void>
struct has_outer_index : std::false_type {};
template <class T>
struct has_outer_index<T, std::void_t<decltype(std::declval<const T&>().outerIndexPtr())>> : std::true_type {};
// dense arguments: a matrix type has rows()/cols() (Eigen::MatrixXd, and Eigen::VectorXd with cols()==1)
template <class T, class = void>
struct has_cols : std::false_type {};
template <class T>
struct has_cols<T, std::void_t<decltype(std::declval<const T&>().cols()), decltype(std::declval<const T&>().rows())>> : std::true_type {};
}  // namespace detail

// What the reference keeps in its public member `A` (EigenSolverWrapper.h:28; a host copy of the last matrix given
// to factorize, used by solve() to size x): here the matrix lives on the device, `A` keeps its shape.
struct FactorizedMatrixInfo {
  long rows_ = 0, cols_ = 0, nnz_ = 0;
  long rows() const { return rows_; }
  long cols() const { return cols_; }
  long nonZeros() const { return nnz_; }
};

class B200SolverWrapper {
 public:
  // the reference's public members (EigenSolverWrapper.h:26,28): `solver` is the GPU solver object standing in
  // for the Eigen solver (the C-ABI handle; spb_* calls may be made on it directly), `A` the factorised matrix
  spb_handle solver = nullptr;
  FactorizedMatrixInfo A;

  // default: the direct solver, the SimplicialLLT replacement (same default in the Python mirror)
  explicit B200SolverWrapper(int solver_kind = SPB_SOLVER_CHOLESKY, int device = 0, double pcg_rtol = 1e-13, int pcg_maxit = 10000) {
    if (spb_create(device, nullptr, &h_) != SPB_OK)
      throw std::runtime_error("saddlepoint_b200: no CUDA device (there is no CPU fallback)");
    solver = h_;
    spb_set_solver(h_, solver_kind, pcg_rtol, pcg_maxit);
  }
  ~B200SolverWrapper() { spb_destroy(h_); }
  B200SolverWrapper(const B200SolverWrapper&) = delete;
  B200SolverWrapper& operator=(const B200SolverWrapper&) = delete;

  spb_handle handle() const { return h_; }
  int last_iterations() const { return last_iters_; }
  // status of the last factorize / solve call (solve() itself always returns true, like the reference's)
  spb_status last_status() const { return last_status_; }
  // which pair-assembly kernel the last fused assembly on this handle ran (3 = stream form, 2 / 1 = general kernels)
  int assembly_form() const { return spb_assembly_form(h_); }

  template <class SparsePattern>
  bool analyze(const SparsePattern& JPattern) {
    return spb_analyze(h_, (int32_t)JPattern.rows(), (int32_t)JPattern.cols(), JPattern.outerIndexPtr(), JPattern.innerIndexPtr()) == SPB_OK;
  }
  template <class SparsePattern>
  bool analyze_pattern(const SparsePattern& JPattern) {
    return analyze(JPattern);
  }

  // legacy: the matrix itself as triplets; symmetric == true lists row <= col only (mirrored here)
  template <class IndexVec>
  bool analyze(const IndexVec& HRows, const IndexVec& HCols, bool symmetric = false) {
    if (HRows.size() != HCols.size()) return false;
    int32_t n = 0;
    for (long t = 0; t < (long)HRows.size(); ++t) {
      if (HRows.data()[t] + 1 > n) n = HRows.data()[t] + 1;
      if (HCols.data()[t] + 1 > n) n = HCols.data()[t] + 1;
    }
    return spb_analyze_triplets(h_, n, (int64_t)HRows.size(), HRows.data(), HCols.data(), symmetric ? 1 : 0) == SPB_OK;
  }

  // A sparse matrix (anything with outerIndexPtr()): the explicit symmetric matrix, both triangles.
  // A dense vector: legacy factorize(HVals[, symmetric]), one value per analysed triplet.
  template <class MatOrVals>
  bool factorize(const MatOrVals& A, bool symmetric = false) {
    (void)symmetric;  // fixed at analyze(HRows, HCols, symmetric), as in the legacy wrapper's callers
    if constexpr (detail::has_outer_index<MatOrVals>::value) {
      last_status_ = spb_set_matrix(h_, (int32_t)A.cols(), A.outerIndexPtr(), A.innerIndexPtr(), A.valuePtr(), SPB_HOST);
      if (last_status_ != SPB_OK) return false;
      this->A.rows_ = (long)A.rows();
      this->A.cols_ = (long)A.cols();
      this->A.nnz_ = (long)A.nonZeros();
      last_status_ = spb_factorize(h_);
    } else {
      last_status_ = spb_factorize_triplets(h_, A.data(), SPB_HOST);
    }
    return last_status_ == SPB_OK;
  }

  // bool solve(const VectorXd& rhs, VectorXd& x) (EigenSolverWrapper.h:72-78) and
  // bool solve(const MatrixXd& rhs, MatrixXd& x) (EigenSolverWrapper.h:62-70: one solve per column):
  // a dense argument with rows()/cols() is taken column by column (column-major data()), anything else as
  // one vector.  Always returns true like the reference; on a failed solve x is zero-filled and
  // last_status() says why.
  template <class Dense>
  bool solve(const Dense& rhs, Dense& x) {
    int nrhs = 1;
    long len = 0;
    if constexpr (detail::has_cols<Dense>::value) {
      nrhs = (int)rhs.cols();
      len = (long)rhs.rows();
      x.resize(rhs.rows(), rhs.cols());
    } else {
      len = (long)rhs.size();
      x.resize(rhs.size());
    }
    last_status_ = spb_solve(h_, rhs.data(), x.data(), nrhs, SPB_HOST, &last_iters_);
    if (last_status_ != SPB_OK && last_status_ != SPB_NOT_CONVERGED)
      for (long i = 0; i < len * nrhs; ++i) x.data()[i] = 0.0;
    return true;  // EigenSolverWrapper.h:69,77
  }
  // the same on raw column-major storage
  bool solve_columns(const double* rhs, int nrhs, double* x) {
    last_status_ = spb_solve(h_, rhs, x, nrhs, SPB_HOST, &last_iters_);
    return true;
  }
  std::string last_error() const { return spb_last_error(h_); }

 private:
  spb_handle h_ = nullptr;
  int last_iters_ = 0;
  spb_status last_status_ = SPB_OK;
};

}  // namespace SaddlePoint
#endif
