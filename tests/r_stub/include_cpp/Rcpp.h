// Miniature of the Rcpp containers the reference's C++ natives use (tests only; see tests/r_stub/include/Rinternals.h
// for the C side): vectors and column-major matrices that VIEW caller memory or own zero-filled storage, copies that
// share storage like Rcpp's (so `NumericMatrix x = px;` updates px in place, as in the reference), min / max, stop().
// Only containers and iteration live here -- every arithmetic statement that is executed comes from the reference's
// own source file, compiled unmodified where it lies (oracle/Makefile target `ref`).
#ifndef RSTUB_RCPP_H
#define RSTUB_RCPP_H
#include <cmath>
#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace Rcpp {

struct exception : std::runtime_error {
    explicit exception(const std::string& m) : std::runtime_error(m) {}
};
inline void stop(const std::string& m) { throw exception(m); }

template <typename T>
class Vector {
  public:
    Vector() = default;
    explicit Vector(long n) : own_(std::make_shared<std::vector<T>>((size_t)n, T())), p_(own_->data()), n_(n) {}
    Vector(T* p, long n) : p_(p), n_(n) {}  // a view of caller memory (an R vector)
    long size() const { return n_; }
    long length() const { return n_; }
    T& operator[](long i) { return p_[i]; }
    const T& operator[](long i) const { return p_[i]; }
    T* begin() { return p_; }
    T* end() { return p_ + n_; }
    const T* begin() const { return p_; }
    const T* end() const { return p_ + n_; }

  private:
    std::shared_ptr<std::vector<T>> own_;
    T* p_ = nullptr;
    long n_ = 0;
};

namespace internal {
struct NamedPlaceHolder {};
}  // namespace internal
static const internal::NamedPlaceHolder _ = internal::NamedPlaceHolder();
template <typename T>
class RowView;

template <typename T>
class Matrix {
  public:
    Matrix() = default;
    Matrix(int nr, int nc) : own_(std::make_shared<std::vector<T>>((size_t)nr * nc, T())), p_(own_->data()), nr_(nr), nc_(nc) {}
    Matrix(T* p, int nr, int nc) : p_(p), nr_(nr), nc_(nc) {}
    int nrow() const { return nr_; }
    int ncol() const { return nc_; }
    int rows() const { return nr_; }
    int cols() const { return nc_; }
    long size() const { return (long)nr_ * nc_; }
    long length() const { return size(); }
    T& operator()(long i, long j) { return p_[(size_t)j * nr_ + i]; }
    const T& operator()(long i, long j) const { return p_[(size_t)j * nr_ + i]; }
    T& operator[](long i) { return p_[i]; }
    const T& operator[](long i) const { return p_[i]; }
    T* begin() { return p_; }
    const T* begin() const { return p_; }
    RowView<T> operator()(long i, const internal::NamedPlaceHolder&) const;  // m(i, _)

  private:
    std::shared_ptr<std::vector<T>> own_;
    T* p_ = nullptr;
    int nr_ = 0, nc_ = 0;
};

// ---- the little sugar src/cG_calcGibbsProbY.cpp's cross-check variant (cG_calcGibbsProbY_Simple) is written in: row
// views `m(i, _)`, eager element-wise +, -, *, lgamma and a left-to-right sum.  The production native of that file
// (cG_CalcGibbsProbY) uses none of it -- only indexing and the C library's scalar lgamma.
template <typename T>
class RowView {
  public:
    RowView(T* first, long stride, long n) : p_(first), stride_(stride), n_(n) {}
    long size() const { return n_; }
    T operator[](long j) const { return p_[(size_t)j * stride_]; }
    RowView& operator=(const Vector<T>& v) {
        for (long j = 0; j < n_; ++j) p_[(size_t)j * stride_] = v[j];
        return *this;
    }

  private:
    T* p_;
    long stride_, n_;
};
template <typename T>
inline RowView<T> Matrix<T>::operator()(long i, const internal::NamedPlaceHolder&) const {
    return RowView<T>(const_cast<T*>(p_) + i, nr_, nc_);
}
template <typename T>
inline Vector<T> operator-(const RowView<T>& a, const RowView<T>& b) {
    Vector<T> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] - b[j];
    return r;
}
template <typename T>
inline Vector<T> operator+(const RowView<T>& a, const RowView<T>& b) {
    Vector<T> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] + b[j];
    return r;
}
inline Vector<double> operator+(const Vector<int>& a, double s) {
    Vector<double> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] + s;
    return r;
}
inline Vector<double> operator*(const Vector<int>& a, double s) {
    Vector<double> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] * s;
    return r;
}
inline Vector<double> operator+(const Vector<int>& a, const Vector<double>& b) {
    Vector<double> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] + b[j];
    return r;
}
inline Vector<double> operator+(const Matrix<int>& a, double s) {
    Vector<double> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = a[j] + s;
    return r;
}
inline Vector<double> lgamma(const Vector<double>& a) {
    Vector<double> r(a.size());
    for (long j = 0; j < a.size(); ++j) r[j] = ::lgamma(a[j]);
    return r;
}
inline double sum(const Vector<double>& a) {
    double t = 0;
    for (long j = 0; j < a.size(); ++j) t += a[j];
    return t;
}

typedef Vector<int> IntegerVector;
typedef Vector<double> NumericVector;
typedef Matrix<int> IntegerMatrix;
typedef Matrix<double> NumericMatrix;

template <typename T>
inline T min(const Vector<T>& v) {
    T m = v[0];
    for (long i = 1; i < v.size(); ++i) m = v[i] < m ? v[i] : m;
    return m;
}
template <typename T>
inline T max(const Vector<T>& v) {
    T m = v[0];
    for (long i = 1; i < v.size(); ++i) m = v[i] > m ? v[i] : m;
    return m;
}

}  // namespace Rcpp
#endif
