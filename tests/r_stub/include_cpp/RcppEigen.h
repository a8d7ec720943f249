// Miniature of Eigen::MappedSparseMatrix<double> as RcppEigen hands a dgCMatrix to C++: the compressed-column arrays
// of the R object, viewed in place, with the InnerIterator the reference's loops use (tests only, see Rcpp.h here).
#ifndef RSTUB_RCPPEIGEN_H
#define RSTUB_RCPPEIGEN_H
#include "Rcpp.h"

namespace Eigen {

template <typename T>
class MappedSparseMatrix {
  public:
    MappedSparseMatrix(long nr, long nc, const int* p, const int* i, const T* x) : nr_(nr), nc_(nc), p_(p), i_(i), x_(x) {}
    long rows() const { return nr_; }
    long cols() const { return nc_; }

    class InnerIterator {
      public:
        InnerIterator(const MappedSparseMatrix& m, long outer) : m_(m), k_(m.p_[outer]), end_(m.p_[outer + 1]) {}
        InnerIterator& operator++() {
            ++k_;
            return *this;
        }
        operator bool() const { return k_ < end_; }
        long index() const { return m_.i_[k_]; }
        T value() const { return m_.x_[k_]; }

      private:
        const MappedSparseMatrix& m_;
        long k_, end_;
    };

  private:
    long nr_, nc_;
    const int* p_;
    const int* i_;
    const T* x_;
};

}  // namespace Eigen
#endif
