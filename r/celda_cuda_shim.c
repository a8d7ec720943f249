/*
 * celda_cuda_shim.c -- the .Call shims a celda maintainer adds to celda/src/ to bind libcelda_cuda
 * (include/celda_cuda.h).  R.h / Rinternals.h are not present in the build image, so this
 * file is compiled (-Wall -Wextra -Werror) and driven through `.Call` against a miniature of the R C API,
 * tests/r_stub/ (tests/test_r_shim.py); the same C ABI is exercised through ctypes by celda_b200/ and the rest of
 * the test-suite.
 *
 * Conventions: every shim unpacks R objects into plain pointers, calls the C ABI, and raises an R error with
 * celda_cuda_last_error() only AFTER the call has returned (the library has released its temporaries by then).
 * `*Change` shims return the `px` they were given, modified in place, like the reference natives
 * (src/matrixSums.c:136-142, src/matrixSumsSparse.cpp:117,173).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "celda_cuda.h"

#define CHECK(call) do { if ((call)) Rf_error("%s", celda_cuda_last_error()); } while (0)

static SEXP slot(SEXP x, const char *n) { return R_do_slot(x, Rf_install(n)); }
static int nlev(SEXP f) { return Rf_isFactor(f) ? Rf_nlevels(f) : -1; } /* -1 == "not a factor" in the C ABI */

/* ---- src/matrixSumsSparse.cpp ---------------------------------------------------------------- */
SEXP _celda_cuda_colSumByGroupSparse(SEXP counts, SEXP group, SEXP K) {
    int *dim = INTEGER(slot(counts, "Dim")), k = Rf_asInteger(K);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, dim[0], k));
    int rc = celda_colSumByGroupSparse(dim[0], dim[1], INTEGER(slot(counts, "p")), INTEGER(slot(counts, "i")),
                                       REAL(slot(counts, "x")), INTEGER(group), XLENGTH(group), k, REAL(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}

SEXP _celda_cuda_rowSumByGroupSparse(SEXP counts, SEXP group, SEXP L) {
    int *dim = INTEGER(slot(counts, "Dim")), l = Rf_asInteger(L);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, l, dim[1]));
    int rc = celda_rowSumByGroupSparse(dim[0], dim[1], INTEGER(slot(counts, "p")), INTEGER(slot(counts, "i")),
                                       REAL(slot(counts, "x")), INTEGER(group), XLENGTH(group), l, REAL(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}

SEXP _celda_cuda_colSumByGroupChangeSparse(SEXP counts, SEXP px, SEXP group, SEXP pgroup, SEXP K) {
    int *dim = INTEGER(slot(counts, "Dim"));
    CHECK(celda_colSumByGroupChangeSparse(dim[0], dim[1], INTEGER(slot(counts, "p")), INTEGER(slot(counts, "i")),
                                          REAL(slot(counts, "x")), REAL(px), Rf_nrows(px), INTEGER(group),
                                          XLENGTH(group), INTEGER(pgroup), XLENGTH(pgroup), Rf_asInteger(K)));
    return px;
}

SEXP _celda_cuda_rowSumByGroupChangeSparse(SEXP counts, SEXP px, SEXP group, SEXP pgroup, SEXP L) {
    int *dim = INTEGER(slot(counts, "Dim"));
    CHECK(celda_rowSumByGroupChangeSparse(dim[0], dim[1], INTEGER(slot(counts, "p")), INTEGER(slot(counts, "i")),
                                          REAL(slot(counts, "x")), REAL(px), Rf_ncols(px), INTEGER(group),
                                          XLENGTH(group), INTEGER(pgroup), XLENGTH(pgroup), Rf_asInteger(L)));
    return px;
}

/* ---- src/matrixSums.c (dense twins; same .Call names as the reference, NAMESPACE:162-170) ----- */
SEXP _rowSumByGroup(SEXP x, SEXP group) {
    int nr = Rf_nrows(x), nc = Rf_ncols(x), nl = nlev(group);
    SEXP ans = PROTECT(Rf_allocMatrix(INTSXP, nl < 0 ? 0 : nl, nc));
    int rc = celda_rowSumByGroup_int(INTEGER(x), nr, nc, INTEGER(group), XLENGTH(group), nl, INTEGER(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}
SEXP _colSumByGroup(SEXP x, SEXP group) {
    int nr = Rf_nrows(x), nc = Rf_ncols(x), nl = nlev(group);
    SEXP ans = PROTECT(Rf_allocMatrix(INTSXP, nr, nl < 0 ? 0 : nl));
    int rc = celda_colSumByGroup_int(INTEGER(x), nr, nc, INTEGER(group), XLENGTH(group), nl, INTEGER(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}
SEXP _rowSumByGroupChange(SEXP x, SEXP px, SEXP group, SEXP pgroup) {
    CHECK(celda_rowSumByGroupChange_int(INTEGER(x), Rf_nrows(x), Rf_ncols(x), INTEGER(px), Rf_nrows(px), Rf_ncols(px),
                                        INTEGER(group), XLENGTH(group), nlev(group), INTEGER(pgroup), XLENGTH(pgroup),
                                        nlev(pgroup)));
    return px;
}
SEXP _colSumByGroupChange(SEXP x, SEXP px, SEXP group, SEXP pgroup) {
    CHECK(celda_colSumByGroupChange_int(INTEGER(x), Rf_nrows(x), Rf_ncols(x), INTEGER(px), Rf_nrows(px), Rf_ncols(px),
                                        INTEGER(group), XLENGTH(group), nlev(group), INTEGER(pgroup), XLENGTH(pgroup),
                                        nlev(pgroup)));
    return px;
}
SEXP _rowSumByGroup_numeric(SEXP x, SEXP group) {
    int nr = Rf_nrows(x), nc = Rf_ncols(x), nl = nlev(group);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, nl < 0 ? 0 : nl, nc));
    int rc = celda_rowSumByGroup_numeric(REAL(x), nr, nc, INTEGER(group), XLENGTH(group), nl, REAL(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}
SEXP _colSumByGroup_numeric(SEXP x, SEXP group) {
    int nr = Rf_nrows(x), nc = Rf_ncols(x), nl = nlev(group);
    SEXP ans = PROTECT(Rf_allocMatrix(REALSXP, nr, nl < 0 ? 0 : nl));
    int rc = celda_colSumByGroup_numeric(REAL(x), nr, nc, INTEGER(group), XLENGTH(group), nl, REAL(ans));
    UNPROTECT(1);
    CHECK(rc);
    return ans;
}
SEXP _rowSumByGroupChange_numeric(SEXP x, SEXP px, SEXP group, SEXP pgroup) {
    CHECK(celda_rowSumByGroupChange_numeric(REAL(x), Rf_nrows(x), Rf_ncols(x), REAL(px), Rf_nrows(px), Rf_ncols(px),
                                            INTEGER(group), XLENGTH(group), nlev(group), INTEGER(pgroup),
                                            XLENGTH(pgroup), nlev(pgroup)));
    return px;
}
SEXP _colSumByGroupChange_numeric(SEXP x, SEXP px, SEXP group, SEXP pgroup) {
    CHECK(celda_colSumByGroupChange_numeric(REAL(x), Rf_nrows(x), Rf_ncols(x), REAL(px), Rf_nrows(px), Rf_ncols(px),
                                            INTEGER(group), XLENGTH(group), nlev(group), INTEGER(pgroup),
                                            XLENGTH(pgroup), nlev(pgroup)));
    return px;
}

/* ---- src/perplexity.c, src/matrixNorm.cpp, src/eigenMatMultInt.cpp, src/cG_calcGibbsProbY.cpp -- */
SEXP _perplexityG(SEXP x, SEXP phi, SEXP psi, SEXP group) {
    double ans = 0;
    CHECK(celda_perplexityG(INTEGER(x), Rf_nrows(x), Rf_ncols(x), REAL(phi), Rf_nrows(phi), Rf_ncols(phi), REAL(psi),
                            Rf_nrows(psi), Rf_ncols(psi), INTEGER(group), XLENGTH(group), nlev(group), &ans));
    return Rf_ScalarReal(ans);
}

SEXP _celda_cuda_fastNormPropLog(SEXP counts, SEXP alpha) {
    SEXP res = PROTECT(Rf_allocMatrix(REALSXP, Rf_nrows(counts), Rf_ncols(counts)));
    int rc = celda_fastNormPropLog(REAL(counts), Rf_nrows(counts), Rf_ncols(counts), Rf_asReal(alpha), REAL(res));
    UNPROTECT(1);
    CHECK(rc);
    return res;
}

SEXP _celda_cuda_eigenMatMultInt(SEXP A, SEXP B) {
    SEXP C = PROTECT(Rf_allocMatrix(REALSXP, Rf_ncols(A), Rf_ncols(B)));
    int rc = celda_eigenMatMultInt(REAL(A), Rf_nrows(A), Rf_ncols(A), INTEGER(B), Rf_ncols(B), REAL(C));
    UNPROTECT(1);
    CHECK(rc);
    return C;
}

SEXP _celda_cuda_eigenMatMultNumeric(SEXP A, SEXP B) {
    SEXP C = PROTECT(Rf_allocMatrix(REALSXP, Rf_ncols(A), Rf_ncols(B)));
    int rc = celda_eigenMatMultNumeric(REAL(A), Rf_nrows(A), Rf_ncols(A), REAL(B), Rf_ncols(B), REAL(C));
    UNPROTECT(1);
    CHECK(rc);
    return C;
}

/* cG_CalcGibbsProbY(index, counts, nTSbyC, nbyTS, nGbyTS, nbyG, y, L, nG, lg_beta, lg_gamma, lg_delta, delta):
 * the three tables are not needed (value-identical direct evaluation); beta and gamma are passed instead. */
SEXP _celda_cuda_cG_CalcGibbsProbY(SEXP index, SEXP counts, SEXP nTSbyC, SEXP nbyTS, SEXP nGbyTS, SEXP nbyG, SEXP y,
                                   SEXP L, SEXP nG, SEXP beta, SEXP gamma, SEXP delta) {
    int l = Rf_asInteger(L);
    SEXP probs = PROTECT(Rf_allocVector(REALSXP, l));
    SEXP g = PROTECT(Rf_coerceVector(nGbyTS, INTSXP)); /* quirk Q4: a double vector in R */
    int rc = celda_cG_CalcGibbsProbY(Rf_asInteger(index), REAL(counts), (int)XLENGTH(counts), REAL(nTSbyC), REAL(nbyTS),
                                     INTEGER(g), REAL(nbyG), INTEGER(y), l, Rf_asInteger(nG), Rf_asReal(beta),
                                     Rf_asReal(gamma), Rf_asReal(delta), REAL(probs));
    UNPROTECT(2);
    CHECK(rc);
    return probs;
}

/* ---- log-likelihoods (R closures .cCGCalcLL / .cCCalcLL / .cGCalcLL) --------------------------- */
SEXP _celda_cuda_cCGCalcLL(SEXP K, SEXP L, SEXP mCPByS, SEXP nTSByCP, SEXP nByG, SEXP nByTS, SEXP nGByTS, SEXP nS,
                           SEXP hyper /* alpha, beta, delta, gamma */) {
    double ll = 0, *h = REAL(hyper);
    SEXP g = PROTECT(Rf_coerceVector(nGByTS, INTSXP));
    int rc = celda_cCGCalcLL(Rf_asInteger(K), Rf_asInteger(L), INTEGER(mCPByS), REAL(nTSByCP), REAL(nByG),
                             (int)XLENGTH(nByG), REAL(nByTS), INTEGER(g), Rf_asInteger(nS), h[0], h[1], h[2], h[3], &ll);
    UNPROTECT(1);
    CHECK(rc);
    return Rf_ScalarReal(ll);
}

/* ---- device-resident chain --------------------------------------------------------------------- */
static void chain_finalizer(SEXP p) {
    celda_chain *ch = (celda_chain *)R_ExternalPtrAddr(p);
    if (ch) celda_chain_destroy(ch);
    R_ClearExternalPtr(p);
}
static celda_chain *chain_of(SEXP p) {
    celda_chain *ch = (celda_chain *)R_ExternalPtrAddr(p);
    if (!ch) Rf_error("celda chain handle has been released");
    return ch;
}

SEXP _celda_cuda_chain_create(SEXP model, SEXP counts, SEXP s, SEXP nS, SEXP K, SEXP L, SEXP hyper, SEXP device) {
    int *dim = INTEGER(slot(counts, "Dim"));
    double *h = REAL(hyper);
    celda_chain *ch = NULL;
    CHECK(celda_chain_create(&ch, Rf_asInteger(device), Rf_asInteger(model), dim[0], dim[1],
                             INTEGER(slot(counts, "p")), INTEGER(slot(counts, "i")), REAL(slot(counts, "x")),
                             Rf_isNull(s) ? NULL : INTEGER(s), Rf_asInteger(nS), Rf_asInteger(K), Rf_asInteger(L), h[0],
                             h[1], h[2], h[3]));
    SEXP p = PROTECT(R_MakeExternalPtr(ch, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(p, chain_finalizer, TRUE);
    UNPROTECT(1);
    return p;
}

SEXP _celda_cuda_chain_set_state(SEXP p, SEXP z, SEXP y) {
    CHECK(celda_chain_set_state(chain_of(p), Rf_isNull(z) ? NULL : INTEGER(z), Rf_isNull(y) ? NULL : INTEGER(y)));
    return R_NilValue;
}

/* one iteration of R/celda_CG.R:543-602; ixy/uy/ixz/uz drawn by R under the caller's with_seed, or all NULL for
 * the device Philox generator (after _celda_cuda_chain_seed) */
SEXP _celda_cuda_chain_iterate(SEXP p, SEXP ixy, SEXP uy, SEXP ixz, SEXP uz) {
    double ll = 0;
    CHECK(celda_chain_iterate(chain_of(p), Rf_isNull(ixy) ? NULL : INTEGER(ixy), Rf_isNull(uy) ? NULL : REAL(uy),
                              Rf_isNull(ixz) ? NULL : INTEGER(ixz), Rf_isNull(uz) ? NULL : REAL(uz), &ll));
    return Rf_ScalarReal(ll);
}

SEXP _celda_cuda_chain_iterate_em(SEXP p, SEXP ixy, SEXP uy) {
    double ll = 0;
    CHECK(celda_chain_iterate_em(chain_of(p), Rf_isNull(ixy) ? NULL : INTEGER(ixy), Rf_isNull(uy) ? NULL : REAL(uy), &ll));
    return Rf_ScalarReal(ll);
}

SEXP _celda_cuda_chain_seed(SEXP p, SEXP seed, SEXP stream) {
    CHECK(celda_chain_seed(chain_of(p), (unsigned long long)Rf_asReal(seed), (unsigned long long)Rf_asReal(stream)));
    return R_NilValue;
}

SEXP _celda_cuda_chain_get_state(SEXP p, SEXP nM, SEXP nG) {
    SEXP z = PROTECT(Rf_allocVector(INTSXP, Rf_asInteger(nM))), y = PROTECT(Rf_allocVector(INTSXP, Rf_asInteger(nG)));
    int rc = celda_chain_get_state(chain_of(p), INTEGER(z), INTEGER(y));
    SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(out, 0, z);
    SET_VECTOR_ELT(out, 1, y);
    UNPROTECT(3);
    CHECK(rc);
    return out;
}

SEXP _celda_cuda_chain_loglik(SEXP p) {
    double ll = 0;
    CHECK(celda_chain_loglik(chain_of(p), &ll));
    return Rf_ScalarReal(ll);
}

SEXP _celda_cuda_chain_perplexity(SEXP p) {
    double v = 0;
    CHECK(celda_chain_perplexity(chain_of(p), &v));
    return Rf_ScalarReal(v);
}

/* celdaGridSearch (R/celdaGridSearch.R:290-391): the next (K, L) run on the same resident counts */
SEXP _celda_cuda_chain_reconfigure(SEXP p, SEXP K, SEXP L) {
    CHECK(celda_chain_reconfigure((celda_chain *)R_ExternalPtrAddr(p), Rf_asInteger(K), Rf_asInteger(L)));
    return R_NilValue;
}

/* split heuristics (R/split_clusters.R): log-likelihood of alternative labels, chain state untouched; NULL keeps a side */
SEXP _celda_cuda_chain_loglik_alt(SEXP p, SEXP z, SEXP y) {
    double ll = 0;
    CHECK(celda_chain_loglik_alt((celda_chain *)R_ExternalPtrAddr(p), Rf_isNull(z) ? NULL : INTEGER(z),
                                 Rf_isNull(y) ? NULL : INTEGER(y), &ll));
    return Rf_ScalarReal(ll);
}

/* perplexity(x, newCounts = ...) (R/perplexity.R:233-325): newCounts as a dgCMatrix */
SEXP _celda_cuda_chain_perplexity_new(SEXP p, SEXP counts) {
    double v = 0;
    CHECK(celda_chain_perplexity_new((celda_chain *)R_ExternalPtrAddr(p), INTEGER(slot(counts, "p")),
                                     INTEGER(slot(counts, "i")), REAL(slot(counts, "x")), &v));
    return Rf_ScalarReal(v);
}

/* resamplePerplexity(doResampling = TRUE) (R/perplexity.R:449-498, 764-775): numResample perplexities */
SEXP _celda_cuda_chain_perplexity_resampled(SEXP p, SEXP seed, SEXP n) {
    int k = Rf_asInteger(n);
    SEXP ans = PROTECT(Rf_allocVector(REALSXP, k));
    CHECK(celda_chain_perplexity_resampled((celda_chain *)R_ExternalPtrAddr(p), (unsigned long long)Rf_asReal(seed), k, REAL(ans)));
    UNPROTECT(1);
    return ans;
}

/* celda_G: opt-in scan of the count columns only, and its per-sweep report (include/celda_cuda.h) */
SEXP _celda_cuda_chain_set_scan_mode(SEXP p, SEXP mode) {
    CHECK(celda_chain_set_scan_mode((celda_chain *)R_ExternalPtrAddr(p), Rf_asInteger(mode)));
    return R_NilValue;
}

SEXP _celda_cuda_chain_scan_margin(SEXP p) {
    SEXP ans = PROTECT(Rf_allocVector(REALSXP, 2));
    CHECK(celda_chain_scan_margin((celda_chain *)R_ExternalPtrAddr(p), REAL(ans)));
    UNPROTECT(1);
    return ans;
}

static const R_CallMethodDef CudaEntries[] = {
    {"_celda_cuda_colSumByGroupSparse", (DL_FUNC)&_celda_cuda_colSumByGroupSparse, 3},
    {"_celda_cuda_rowSumByGroupSparse", (DL_FUNC)&_celda_cuda_rowSumByGroupSparse, 3},
    {"_celda_cuda_colSumByGroupChangeSparse", (DL_FUNC)&_celda_cuda_colSumByGroupChangeSparse, 5},
    {"_celda_cuda_rowSumByGroupChangeSparse", (DL_FUNC)&_celda_cuda_rowSumByGroupChangeSparse, 5},
    {"_rowSumByGroup", (DL_FUNC)&_rowSumByGroup, 2},
    {"_colSumByGroup", (DL_FUNC)&_colSumByGroup, 2},
    {"_rowSumByGroupChange", (DL_FUNC)&_rowSumByGroupChange, 4},
    {"_colSumByGroupChange", (DL_FUNC)&_colSumByGroupChange, 4},
    {"_rowSumByGroup_numeric", (DL_FUNC)&_rowSumByGroup_numeric, 2},
    {"_colSumByGroup_numeric", (DL_FUNC)&_colSumByGroup_numeric, 2},
    {"_rowSumByGroupChange_numeric", (DL_FUNC)&_rowSumByGroupChange_numeric, 4},
    {"_colSumByGroupChange_numeric", (DL_FUNC)&_colSumByGroupChange_numeric, 4},
    {"_perplexityG", (DL_FUNC)&_perplexityG, 4},
    {"_celda_cuda_fastNormPropLog", (DL_FUNC)&_celda_cuda_fastNormPropLog, 2},
    {"_celda_cuda_eigenMatMultInt", (DL_FUNC)&_celda_cuda_eigenMatMultInt, 2},
    {"_celda_cuda_eigenMatMultNumeric", (DL_FUNC)&_celda_cuda_eigenMatMultNumeric, 2},
    {"_celda_cuda_cG_CalcGibbsProbY", (DL_FUNC)&_celda_cuda_cG_CalcGibbsProbY, 12},
    {"_celda_cuda_cCGCalcLL", (DL_FUNC)&_celda_cuda_cCGCalcLL, 9},
    {"_celda_cuda_chain_create", (DL_FUNC)&_celda_cuda_chain_create, 8},
    {"_celda_cuda_chain_set_state", (DL_FUNC)&_celda_cuda_chain_set_state, 3},
    {"_celda_cuda_chain_iterate", (DL_FUNC)&_celda_cuda_chain_iterate, 5},
    {"_celda_cuda_chain_iterate_em", (DL_FUNC)&_celda_cuda_chain_iterate_em, 3},
    {"_celda_cuda_chain_seed", (DL_FUNC)&_celda_cuda_chain_seed, 3},
    {"_celda_cuda_chain_get_state", (DL_FUNC)&_celda_cuda_chain_get_state, 3},
    {"_celda_cuda_chain_loglik", (DL_FUNC)&_celda_cuda_chain_loglik, 1},
    {"_celda_cuda_chain_perplexity", (DL_FUNC)&_celda_cuda_chain_perplexity, 1},
    {"_celda_cuda_chain_reconfigure", (DL_FUNC)&_celda_cuda_chain_reconfigure, 3},
    {"_celda_cuda_chain_loglik_alt", (DL_FUNC)&_celda_cuda_chain_loglik_alt, 3},
    {"_celda_cuda_chain_perplexity_new", (DL_FUNC)&_celda_cuda_chain_perplexity_new, 2},
    {"_celda_cuda_chain_perplexity_resampled", (DL_FUNC)&_celda_cuda_chain_perplexity_resampled, 3},
    {"_celda_cuda_chain_set_scan_mode", (DL_FUNC)&_celda_cuda_chain_set_scan_mode, 2},
    {"_celda_cuda_chain_scan_margin", (DL_FUNC)&_celda_cuda_chain_scan_margin, 1},
    {NULL, NULL, 0}};

/* called from R_init_celda (src/RcppExports.cpp:339-342) next to the Rcpp-generated table */
void celda_cuda_register(DllInfo *dll) { R_registerRoutines(dll, NULL, CudaEntries, NULL, NULL); }
