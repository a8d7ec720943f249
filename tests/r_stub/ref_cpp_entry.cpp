// extern "C" doors into the reference's C++ natives (src/matrixSumsSparse.cpp and src/cG_calcGibbsProbY.cpp, each
// compiled unmodified as its own translation unit against the container miniatures in include_cpp/): plain pointers in, status + message out, so that
// tests/test_reference_natives.py can call them through ctypes.  Tests only.
#include <cstring>
#include <string>

#include "RcppEigen.h"

using Rcpp::IntegerMatrix;
using Rcpp::IntegerVector;
using Rcpp::NumericMatrix;
using Rcpp::NumericVector;
typedef Eigen::MappedSparseMatrix<double> Sparse;

// the reference's exported functions (src/matrixSumsSparse.cpp:9-187)
NumericMatrix colSumByGroupSparse(const Sparse& counts, const IntegerVector& group, const int& K);
NumericMatrix rowSumByGroupSparse(const Sparse& counts, const IntegerVector& group, const int& L);
NumericMatrix colSumByGroupChangeSparse(const Sparse& counts, const NumericMatrix& px, const IntegerVector& group,
                                        const IntegerVector& pgroup, const int& K);
NumericMatrix rowSumByGroupChangeSparse(const Sparse& counts, const NumericMatrix& px, const IntegerVector& group,
                                        const IntegerVector& pgroup, const int& L);

// src/cG_calcGibbsProbY.cpp: the production native (:186-249) and the reference's own cross-check variants (:7-183)
NumericVector cG_CalcGibbsProbY(const int index, const NumericVector& counts, const NumericMatrix& nTSbyC,
                                const NumericVector& nbyTS, const IntegerVector& nGbyTS, const NumericVector& nbyG,
                                const IntegerVector& y, const int L, const int nG, const NumericVector& lg_beta,
                                const NumericVector& lg_gamma, const NumericVector& lg_delta, const double delta);
NumericVector cG_CalcGibbsProbY_fastRow(const int index, const IntegerMatrix& counts, const IntegerMatrix& nTSbyC,
                                        const IntegerVector& nbyTS, const IntegerVector& nGbyTS, const IntegerVector& nbyG,
                                        const IntegerVector& y, const int L, const int nG, const NumericVector& lg_beta,
                                        const NumericVector& lg_gamma, const NumericVector& lg_delta, const double delta);
NumericVector cG_CalcGibbsProbY_ori(const int index, const IntegerMatrix& counts, const IntegerMatrix& nTSbyC,
                                    const IntegerVector& nbyTS, const IntegerVector& nGbyTS, const IntegerVector& nbyG,
                                    const IntegerVector& y, const int L, const int nG, const NumericVector& lg_beta,
                                    const NumericVector& lg_gamma, const NumericVector& lg_delta, const double delta);
NumericVector cG_calcGibbsProbY_Simple(const IntegerMatrix counts, IntegerVector nGbyTS, IntegerMatrix nTSbyC,
                                       IntegerVector nbyTS, IntegerVector nbyG, const IntegerVector y, const int L,
                                       const int index, const double gamma, const double beta, const double delta);

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

// counts_row: the gene's row of the nG x ncol count matrix (what R passes: counts[index, ]); tables as R builds them
// (lg_beta[k] = lgamma(k + beta), lg_gamma[k] = lgamma(k + gamma), lg_delta = c(NA, lgamma(seq(nG + L) * delta)))
int ref_cG_CalcGibbsProbY(int index, double* counts_row, int ncol, double* nTSbyC, double* nbyTS, int* nGbyTS, double* nbyG,
                          int* y, int L, int nG, double* lg_beta, long nlb, double* lg_gamma, long nlg, double* lg_delta,
                          long nld, double delta, double* probs) {
    return guarded([&] {
        NumericVector r = cG_CalcGibbsProbY(index, NumericVector(counts_row, ncol), NumericMatrix(nTSbyC, L, ncol),
                                            NumericVector(nbyTS, L), IntegerVector(nGbyTS, L), NumericVector(nbyG, nG),
                                            IntegerVector(y, nG), L, nG, NumericVector(lg_beta, nlb),
                                            NumericVector(lg_gamma, nlg), NumericVector(lg_delta, nld), delta);
        std::memcpy(probs, r.begin(), sizeof(double) * (size_t)L);
    });
}
// which: 0 = _fastRow, 1 = _ori (integer matrices, the whole nG x ncol count matrix)
int ref_cG_CalcGibbsProbY_int(int which, int index, int* counts, int ncol, int* nTSbyC, int* nbyTS, int* nGbyTS, int* nbyG,
                              int* y, int L, int nG, double* lg_beta, long nlb, double* lg_gamma, long nlg,
                              double* lg_delta, long nld, double delta, double* probs) {
    return guarded([&] {
        auto f = which == 0 ? cG_CalcGibbsProbY_fastRow : cG_CalcGibbsProbY_ori;
        NumericVector r = f(index, IntegerMatrix(counts, nG, ncol), IntegerMatrix(nTSbyC, L, ncol), IntegerVector(nbyTS, L),
                            IntegerVector(nGbyTS, L), IntegerVector(nbyG, nG), IntegerVector(y, nG), L, nG,
                            NumericVector(lg_beta, nlb), NumericVector(lg_gamma, nlg), NumericVector(lg_delta, nld), delta);
        std::memcpy(probs, r.begin(), sizeof(double) * (size_t)L);
    });
}
// the slow variant "that matches the equations more directly" (src/cG_calcGibbsProbY.cpp:4-6): it evaluates the full
// expression per candidate module, so its values differ from the others by a constant
int ref_cG_calcGibbsProbY_Simple(int* counts, int ncol, int* nGbyTS, int* nTSbyC, int* nbyTS, int* nbyG, int* y, int L,
                                 int nG, int index, double gamma, double beta, double delta, double* probs) {
    return guarded([&] {
        NumericVector r = cG_calcGibbsProbY_Simple(IntegerMatrix(counts, nG, ncol), IntegerVector(nGbyTS, L),
                                                   IntegerMatrix(nTSbyC, L, ncol), IntegerVector(nbyTS, L),
                                                   IntegerVector(nbyG, nG), IntegerVector(y, nG), L, index, gamma, beta,
                                                   delta);
        std::memcpy(probs, r.begin(), sizeof(double) * (size_t)L);
    });
}
}
