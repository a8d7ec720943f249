// extern "C" doors into the reference's C++ natives (src/matrixSumsSparse.cpp, compiled unmodified as its own
// translation unit against the container miniatures in include_cpp/): plain pointers in, status + message out, so that
// tests/test_reference_natives.py can call them through ctypes.  Tests only.
#include <cstring>
#include <string>

#include "RcppEigen.h"

using Rcpp::IntegerVector;
using Rcpp::NumericMatrix;
typedef Eigen::MappedSparseMatrix<double> Sparse;

// the reference's exported functions (src/matrixSumsSparse.cpp:9-187)
NumericMatrix colSumByGroupSparse(const Sparse& counts, const IntegerVector& group, const int& K);
NumericMatrix rowSumByGroupSparse(const Sparse& counts, const IntegerVector& group, const int& L);
NumericMatrix colSumByGroupChangeSparse(const Sparse& counts, const NumericMatrix& px, const IntegerVector& group,
                                        const IntegerVector& pgroup, const int& K);
NumericMatrix rowSumByGroupChangeSparse(const Sparse& counts, const NumericMatrix& px, const IntegerVector& group,
                                        const IntegerVector& pgroup, const int& L);

static std::string g_msg;

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_msg = e.what();
        return 1;
    }
}

extern "C" {

const char* ref_last_error() { return g_msg.c_str(); }

int ref_colSumByGroupSparse(int nr, int nc, const int* p, const int* i, const double* x, int* group, long ngroup, int K,
                            double* out /* nr x K */) {
    return guarded([&] {
        NumericMatrix r = colSumByGroupSparse(Sparse(nr, nc, p, i, x), IntegerVector(group, ngroup), K);
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)r.size());
    });
}
int ref_rowSumByGroupSparse(int nr, int nc, const int* p, const int* i, const double* x, int* group, long ngroup, int L,
                            double* out /* L x nc */) {
    return guarded([&] {
        NumericMatrix r = rowSumByGroupSparse(Sparse(nr, nc, p, i, x), IntegerVector(group, ngroup), L);
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)r.size());
    });
}
// px is the caller's matrix: the reference updates it in place and returns it
int ref_colSumByGroupChangeSparse(int nr, int nc, const int* p, const int* i, const double* x, double* px, int px_nr,
                                  int px_nc, int* group, long ngroup, int* pgroup, long npgroup, int K) {
    return guarded([&] {
        colSumByGroupChangeSparse(Sparse(nr, nc, p, i, x), NumericMatrix(px, px_nr, px_nc), IntegerVector(group, ngroup),
                                  IntegerVector(pgroup, npgroup), K);
    });
}
int ref_rowSumByGroupChangeSparse(int nr, int nc, const int* p, const int* i, const double* x, double* px, int px_nr,
                                  int px_nc, int* group, long ngroup, int* pgroup, long npgroup, int L) {
    return guarded([&] {
        rowSumByGroupChangeSparse(Sparse(nr, nc, p, i, x), NumericMatrix(px, px_nr, px_nc), IntegerVector(group, ngroup),
                                  IntegerVector(pgroup, npgroup), L);
    });
}
}
