/* CPU stand-in for the closure-level entry points of include/celda_cuda.h, answering through the oracle
 * (oracle/libcelda_oracle.so).  Test infrastructure ONLY (tests/test_r_shim.py): linked IN FRONT of libcelda_cuda.so
 * under the R .Call shim, it lets the shim's marshalling of the R closures' arguments and result lists (argument
 * order, copies, dims, names) be checked numerically on a box without a GPU.  Never shipped, never on the product
 * path: the product library has no CPU fallback. */
#include <stddef.h>

#include "celda_cuda.h"

int o_cCCalcGibbsProbZ(const double *counts, int R, long nM, int *mCPByS, int nS, double *nGByCP, const double *nByC,
                       double *nCP, int *z, const int *s, int K, double alpha, double beta, int doSample,
                       const int *ix, const double *u, long first, long count, double *probs, double *margin);
int o_cGCalcGibbsProbY(const double *counts, int nG, int ncol, double *nTSByC, double *nByTS, int *nGByTS,
                       const double *nByG, int *y, int L, double beta, double delta, double gamma,
                       const double *lgbeta, const double *lggamma, const double *lgdelta, int doSample,
                       const int *ix, const double *u, long first, long count, double *probs, double *margin);
int o_cCCalcEMProbZ(int sparse, int R, long nM, const int *p, const int *ri, const double *x, int *mCPByS, int nS,
                    double *nGByCP, double *nCP, int *z, const int *s, int K, double alpha, double beta, int doSample,
                    double *probs);
double o_cCCalcLL(const int *mCPByS, const double *nGByCP, int K, int nS, int nG, double alpha, double beta);
double o_cGCalcLL(const double *nTSByC, const double *nByTS, const double *nByG, int nGenes, const int *nGByTS,
                  long nM, int L, double beta, double delta, double gamma);
const char *o_last_error(void);

static int mock_calls = 0;
int mock_celda_cuda_calls(void) { return mock_calls; }

const char *celda_cuda_last_error(void) { return o_last_error(); }

int celda_cCCalcGibbsProbZ(const double *counts, int *mCPByS, int nS, double *nGByCP, const double *nByC, double *nCP,
                           int *z, const int *s, int K, int nG, long long nM, double alpha, double beta, int doSample,
                           const int *ix, const double *u, double *probs) {
    ++mock_calls;
    return o_cCCalcGibbsProbZ(counts, nG, (long)nM, mCPByS, nS, nGByCP, nByC, nCP, z, s, K, alpha, beta, doSample, ix, u,
                              0, (long)nM, probs, NULL);
}

int celda_cGCalcGibbsProbY(const double *counts, int ncol, double *nTSByC, double *nByTS, int *nGByTS,
                           const double *nByG, int *y, int L, int nG, double beta, double delta, double gamma,
                           int doSample, const int *ix, const double *u, double *probs) {
    ++mock_calls;
    return o_cGCalcGibbsProbY(counts, nG, ncol, nTSByC, nByTS, nGByTS, nByG, y, L, beta, delta, gamma, NULL, NULL, NULL,
                              doSample, ix, u, 0, nG, probs, NULL);
}

int celda_cCCalcEMProbZ(int nG, long long nM, const int *p, const int *i, const double *x, int *mCPByS, int nS,
                        double *nGByCP, double *nCP, int *z, const int *s, int K, double alpha, double beta,
                        int doSample, double *probs) {
    ++mock_calls;
    return o_cCCalcEMProbZ(p != NULL, nG, (long)nM, p, i, x, mCPByS, nS, nGByCP, nCP, z, s, K, alpha, beta, doSample,
                           probs);
}

int celda_cCCalcLL(const int *mCPByS, const double *nGByCP, int K, int nS, int nG, double alpha, double beta,
                   double *ll) {
    ++mock_calls;
    *ll = o_cCCalcLL(mCPByS, nGByCP, K, nS, nG, alpha, beta);
    return 0;
}

int celda_cGCalcLL(const double *nTSByC, const double *nByTS, const double *nByG, int nGenes, const int *nGByTS,
                   long long nM, int L, double beta, double delta, double gamma, double *ll) {
    ++mock_calls;
    *ll = o_cGCalcLL(nTSByC, nByTS, nByG, nGenes, nGByTS, (long)nM, L, beta, delta, gamma);
    return 0;
}
