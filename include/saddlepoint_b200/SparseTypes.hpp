// SparseTypes.hpp -- minimal stand-ins for Eigen::VectorXd / Eigen::SparseMatrix<double>.
//
// The concept classes in this directory are matrix-type agnostic: they only use Eigen's raw
// compressed-storage accessors (rows(), cols(), nonZeros(), outerIndexPtr(), innerIndexPtr(),
// valuePtr()) and data()/size()/resize() on vectors, so they compile unchanged against real
// Eigen types (define SPB_USE_EIGEN before including LMSolver.hpp) or against these shims
// where Eigen is not installed (it is not in this image).
#ifndef SADDLEPOINT_B200_SPARSE_TYPES_HPP
#define SADDLEPOINT_B200_SPARSE_TYPES_HPP
#include <algorithm>
#include <cstdint>
#include <vector>

namespace SaddlePoint {
namespace shim {

class Vector {
 public:
  Vector() = default;
  explicit Vector(long n) : v_((size_t)n, 0.0) {}
  long size() const { return (long)v_.size(); }
  void resize(long n) { v_.resize((size_t)n); }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(long i) { return v_[(size_t)i]; }
  double operator()(long i) const { return v_[(size_t)i]; }
  double& operator[](long i) { return v_[(size_t)i]; }
  double operator[](long i) const { return v_[(size_t)i]; }
  void setZero() { std::fill(v_.begin(), v_.end(), 0.0); }
  void setConstant(double c) { std::fill(v_.begin(), v_.end(), c); }

 private:
  std::vector<double> v_;
};

// Column-major dense matrix (stand-in for Eigen::MatrixXd in the matrix right-hand-side solve overload).
class Matrix {
 public:
  Matrix() = default;
  Matrix(long r, long c) : r_(r), c_(c), v_((size_t)(r * c), 0.0) {}
  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  void resize(long r, long c) {
    r_ = r;
    c_ = c;
    v_.resize((size_t)(r * c));
  }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(long i, long j) { return v_[(size_t)(j * r_ + i)]; }
  double operator()(long i, long j) const { return v_[(size_t)(j * r_ + i)]; }

 private:
  long r_ = 0, c_ = 0;
  std::vector<double> v_;
};

struct Triplet {
  int r, c;
  double v;
  Triplet(int r_, int c_, double v_) : r(r_), c(c_), v(v_) {}
  int row() const { return r; }
  int col() const { return c; }
  double value() const { return v; }
};

// Column-major compressed sparse matrix, int32 indices (Eigen's default StorageIndex).
class SparseMatrix {
 public:
  SparseMatrix() = default;
  SparseMatrix(long rows, long cols) : rows_(rows), cols_(cols), outer_((size_t)cols + 1, 0) {}
  long rows() const { return rows_; }
  long cols() const { return cols_; }
  long nonZeros() const { return outer_.empty() ? 0 : outer_.back(); }
  long outerSize() const { return cols_; }
  const int* outerIndexPtr() const { return outer_.data(); }
  const int* innerIndexPtr() const { return inner_.data(); }
  const double* valuePtr() const { return val_.data(); }
  double* valuePtr() { return val_.data(); }
  void resize(long rows, long cols) {
    rows_ = rows;
    cols_ = cols;
    outer_.assign((size_t)cols + 1, 0);
    inner_.clear();
    val_.clear();
  }
  // Eigen setFromTriplets semantics: sorted inner indices, duplicates summed in insertion
  // order, explicit zeros kept.
  template <class It>
  void setFromTriplets(It first, It last) {
    std::vector<std::vector<std::pair<int, double>>> byrow((size_t)rows_);
    for (It t = first; t != last; ++t) {
      auto& row = byrow[(size_t)t->row()];
      bool dup = false;
      for (auto& e : row)
        if (e.first == t->col()) {
          e.second += t->value();
          dup = true;
          break;
        }
      if (!dup) row.emplace_back(t->col(), t->value());
    }
    outer_.assign((size_t)cols_ + 1, 0);
    for (auto& row : byrow)
      for (auto& e : row) outer_[(size_t)e.first + 1]++;
    for (long c = 0; c < cols_; ++c) outer_[(size_t)c + 1] += outer_[(size_t)c];
    inner_.resize((size_t)outer_.back());
    val_.resize((size_t)outer_.back());
    std::vector<int> cur(outer_.begin(), outer_.end() - 1);
    for (long r = 0; r < rows_; ++r)
      for (auto& e : byrow[(size_t)r]) {
        int d = cur[(size_t)e.first]++;
        inner_[(size_t)d] = (int)r;
        val_[(size_t)d] = e.second;
      }
  }
  void setCompressed(long rows, long cols, const int* outer, const int* inner, const double* val) {
    rows_ = rows;
    cols_ = cols;
    outer_.assign(outer, outer + cols + 1);
    inner_.assign(inner, inner + outer[cols]);
    if (val) val_.assign(val, val + outer[cols]); else val_.assign((size_t)outer[cols], 0.0);
  }

 private:
  long rows_ = 0, cols_ = 0;
  std::vector<int> outer_, inner_;
  std::vector<double> val_;
};

}  // namespace shim
}  // namespace SaddlePoint
#endif
