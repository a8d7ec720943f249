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

/* a modifiable copy of x as `type`, attributes (dim) kept: the closures below return updated copies, R arguments are
 * never modified.  The caller PROTECTs the result. */
static SEXP fresh(SEXP x, SEXPTYPE type) {
    SEXP y = PROTECT(Rf_coerceVector(x, type));
    if (y == x) y = Rf_duplicate(x);
    UNPROTECT(1);
    return y;
}
static SEXP named_list(int n, const char *const *names, SEXP *vals) {
    SEXP out = PROTECT(Rf_allocVector(VECSXP, n)), nm = PROTECT(Rf_allocVector(STRSXP, n));
    for (int k = 0; k < n; ++k) {
        SET_VECTOR_ELT(out, k, vals[k]);
        SET_STRING_ELT(nm, k, Rf_mkChar(names[k]));
    }
    Rf_setAttrib(out, R_NamesSymbol, nm);
    UNPROTECT(2);
    return out;
}

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

/* ---- the sweeps at closure granularity (R/celda_C.R:620-756, R/celda_G.R:602-660) ---------------
 * Same arguments and the same named list as the R closures; the draws the closure would take from R's generator are
 * passed in: ix = sample(seq_len(n)), u = runif(n), drawn by the caller under its with_seed.
 * `counts` is the dense matrix the closure indexes (nTSByC / nGByCP in celda_CG, as.matrix(counts) in celda_C / G). */
SEXP _celda_cuda_cCCalcGibbsProbZ(SEXP counts, SEXP mCPByS, SEXP nGByCP, SEXP nByC, SEXP nCP, SEXP z, SEXP s, SEXP K,
                                  SEXP nG, SEXP nM, SEXP alpha, SEXP beta, SEXP doSample, SEXP ix, SEXP u) {
    int k = Rf_asInteger(K), nm = Rf_asInteger(nM);
    SEXP cnt = PROTECT(Rf_coerceVector(counts, REALSXP)), byc = PROTECT(Rf_coerceVector(nByC, REALSXP));
    SEXP ss = PROTECT(Rf_coerceVector(s, INTSXP)), ord = PROTECT(Rf_coerceVector(ix, INTSXP));
    SEXP m = PROTECT(fresh(mCPByS, INTSXP)), n = PROTECT(fresh(nGByCP, REALSXP)), cp = PROTECT(fresh(nCP, REALSXP));
    SEXP zz = PROTECT(fresh(z, INTSXP)), probs = PROTECT(Rf_allocMatrix(REALSXP, k, nm));
    int rc = celda_cCCalcGibbsProbZ(REAL(cnt), INTEGER(m), Rf_ncols(m), REAL(n), REAL(byc), REAL(cp), INTEGER(zz),
                                    INTEGER(ss), k, Rf_asInteger(nG), nm, Rf_asReal(alpha), Rf_asReal(beta),
                                    Rf_asLogical(doSample), INTEGER(ord), REAL(u), REAL(probs));
    static const char *const names[] = {"mCPByS", "nGByCP", "nCP", "z", "probs"};
    SEXP vals[] = {m, n, cp, zz, probs};
    SEXP out = named_list(5, names, vals);
    UNPROTECT(9);
    CHECK(rc);
    return out;
}

SEXP _celda_cuda_cGCalcGibbsProbY(SEXP counts, SEXP nTSByC, SEXP nByTS, SEXP nGByTS, SEXP nByG, SEXP y, SEXP L, SEXP nG,
                                  SEXP beta, SEXP delta, SEXP gamma, SEXP doSample, SEXP ix, SEXP u) {
    int l = Rf_asInteger(L), ng = Rf_asInteger(nG);
    SEXP cnt = PROTECT(Rf_coerceVector(counts, REALSXP)), byg = PROTECT(Rf_coerceVector(nByG, REALSXP));
    SEXP ord = PROTECT(Rf_coerceVector(ix, INTSXP));
    SEXP t = PROTECT(fresh(nTSByC, REALSXP)), ts = PROTECT(fresh(nByTS, REALSXP)), gt = PROTECT(fresh(nGByTS, INTSXP));
    SEXP yy = PROTECT(fresh(y, INTSXP)), probs = PROTECT(Rf_allocMatrix(REALSXP, l, ng));
    int rc = celda_cGCalcGibbsProbY(REAL(cnt), Rf_ncols(cnt), REAL(t), REAL(ts), INTEGER(gt), REAL(byg), INTEGER(yy), l, ng,
                                    Rf_asReal(beta), Rf_asReal(delta), Rf_asReal(gamma), Rf_asLogical(doSample),
                                    INTEGER(ord), REAL(u), REAL(probs));
    static const char *const names[] = {"nTSByC", "nGByTS", "nByTS", "y", "probs"};
    SEXP vals[] = {t, gt, ts, yy, probs};
    SEXP out = named_list(5, names, vals);
    UNPROTECT(8);
    CHECK(rc);
    return out;
}

/* .cCCalcEMProbZ: counts as a dgCMatrix (S4) or a dense matrix; nByC is not needed by the EM step (kept for the signature) */
SEXP _celda_cuda_cCCalcEMProbZ(SEXP counts, SEXP mCPByS, SEXP nGByCP, SEXP nByC, SEXP nCP, SEXP z, SEXP s, SEXP K, SEXP nG,
                               SEXP nM, SEXP alpha, SEXP beta, SEXP doSample) {
    (void)nByC;
    int k = Rf_asInteger(K), nm = Rf_asInteger(nM), sparse = IS_S4_OBJECT(counts);
    SEXP cnt = PROTECT(sparse ? slot(counts, "x") : Rf_coerceVector(counts, REALSXP));
    SEXP ss = PROTECT(Rf_coerceVector(s, INTSXP));
    SEXP m = PROTECT(fresh(mCPByS, INTSXP)), n = PROTECT(fresh(nGByCP, REALSXP)), cp = PROTECT(fresh(nCP, REALSXP));
    SEXP zz = PROTECT(fresh(z, INTSXP)), probs = PROTECT(Rf_allocMatrix(REALSXP, k, nm));
    int rc = celda_cCCalcEMProbZ(Rf_asInteger(nG), nm, sparse ? INTEGER(slot(counts, "p")) : NULL,
                                 sparse ? INTEGER(slot(counts, "i")) : NULL, REAL(cnt), INTEGER(m), Rf_ncols(m), REAL(n),
                                 REAL(cp), INTEGER(zz), INTEGER(ss), k, Rf_asReal(alpha), Rf_asReal(beta),
                                 Rf_asLogical(doSample), REAL(probs));
    static const char *const names[] = {"mCPByS", "nGByCP", "nCP", "z", "probs"};
    SEXP vals[] = {m, n, cp, zz, probs};
    SEXP out = named_list(5, names, vals);
    UNPROTECT(7);
    CHECK(rc);
    return out;
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

/* .cCCalcLL(mCPByS, nGByCP, s, z, K, nS, nG, alpha, beta) R/celda_C.R:760-788 (s and z are unused by the formula) */
SEXP _celda_cuda_cCCalcLL(SEXP mCPByS, SEXP nGByCP, SEXP s, SEXP z, SEXP K, SEXP nS, SEXP nG, SEXP alpha, SEXP beta) {
    (void)s;
    (void)z;
    double ll = 0;
    SEXP m = PROTECT(Rf_coerceVector(mCPByS, INTSXP)), n = PROTECT(Rf_coerceVector(nGByCP, REALSXP));
    int rc = celda_cCCalcLL(INTEGER(m), REAL(n), Rf_asInteger(K), Rf_asInteger(nS), Rf_asInteger(nG), Rf_asReal(alpha),
                            Rf_asReal(beta), &ll);
    UNPROTECT(2);
    CHECK(rc);
    return Rf_ScalarReal(ll);
}

/* .cGCalcLL(nTSByC, nByTS, nByG, nGByTS, nM, nG, L, beta, delta, gamma) R/celda_G.R:664-702 */
SEXP _celda_cuda_cGCalcLL(SEXP nTSByC, SEXP nByTS, SEXP nByG, SEXP nGByTS, SEXP nM, SEXP nG, SEXP L, SEXP beta, SEXP delta,
                          SEXP gamma) {
    double ll = 0;
    SEXP t = PROTECT(Rf_coerceVector(nTSByC, REALSXP)), ts = PROTECT(Rf_coerceVector(nByTS, REALSXP));
    SEXP g = PROTECT(Rf_coerceVector(nByG, REALSXP)), gt = PROTECT(Rf_coerceVector(nGByTS, INTSXP));
    int rc = celda_cGCalcLL(REAL(t), REAL(ts), REAL(g), Rf_asInteger(nG), INTEGER(gt), (long long)Rf_asReal(nM),
                            Rf_asInteger(L), Rf_asReal(beta), Rf_asReal(delta), Rf_asReal(gamma), &ll);
    UNPROTECT(4);
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

/* single sweeps on the handle: .cGCalcGibbsProbY / .cCCalcGibbsProbZ / .cCCalcEMProbZ with the tables resident;
 * doSample = FALSE leaves the labels alone and only fills the probabilities (clusterProbability, the second-best
 * labels of the 'split' initialisation) */
SEXP _celda_cuda_chain_sweep_y(SEXP p, SEXP ix, SEXP u, SEXP doSample) {
    CHECK(celda_chain_sweep_y(chain_of(p), INTEGER(ix), REAL(u), Rf_asLogical(doSample)));
    return R_NilValue;
}
SEXP _celda_cuda_chain_sweep_z_gibbs(SEXP p, SEXP ix, SEXP u, SEXP doSample) {
    CHECK(celda_chain_sweep_z_gibbs(chain_of(p), INTEGER(ix), REAL(u), Rf_asLogical(doSample)));
    return R_NilValue;
}
SEXP _celda_cuda_chain_sweep_z_em(SEXP p, SEXP doSample) {
    CHECK(celda_chain_sweep_z_em(chain_of(p), Rf_asLogical(doSample)));
    return R_NilValue;
}

/* list(z = K x nM, y = L x nG) of the last sweeps' probabilities; a dimension of 0 skips that side (celda_C / celda_G) */
SEXP _celda_cuda_chain_get_probs(SEXP p, SEXP K, SEXP nM, SEXP L, SEXP nG) {
    int k = Rf_asInteger(K), nm = Rf_asInteger(nM), l = Rf_asInteger(L), ng = Rf_asInteger(nG);
    SEXP pz = PROTECT(Rf_allocMatrix(REALSXP, k, nm)), py = PROTECT(Rf_allocMatrix(REALSXP, l, ng));
    int rc = celda_chain_get_probs(chain_of(p), (k && nm) ? REAL(pz) : NULL, (l && ng) ? REAL(py) : NULL);
    static const char *const names[] = {"z", "y"};
    SEXP vals[] = {pz, py};
    SEXP out = named_list(2, names, vals);
    UNPROTECT(2);
    CHECK(rc);
    return out;
}

/* the count tables of the handle as .cCGDecomposeCounts / .cCDecomposeCounts / .cGDecomposeCounts return them
 * (R/celda_CG.R:902-934); dimensions that do not apply to the handle's model are passed as 0 and come back empty */
SEXP _celda_cuda_chain_get_tables(SEXP p, SEXP nG, SEXP nM, SEXP nS, SEXP K, SEXP L) {
    int ng = Rf_asInteger(nG), nm = Rf_asInteger(nM), ns = Rf_asInteger(nS), k = Rf_asInteger(K), l = Rf_asInteger(L);
    SEXP v[9];
    v[0] = PROTECT(Rf_allocMatrix(INTSXP, k, ns));   /* mCPByS  */
    v[1] = PROTECT(Rf_allocMatrix(REALSXP, l, nm));  /* nTSByC  */
    v[2] = PROTECT(Rf_allocMatrix(REALSXP, l, k));   /* nTSByCP */
    v[3] = PROTECT(Rf_allocMatrix(REALSXP, ng, k));  /* nGByCP  */
    v[4] = PROTECT(Rf_allocVector(REALSXP, k));      /* nCP     */
    v[5] = PROTECT(Rf_allocVector(REALSXP, nm));     /* nByC    */
    v[6] = PROTECT(Rf_allocVector(REALSXP, ng));     /* nByG    */
    v[7] = PROTECT(Rf_allocVector(REALSXP, l));      /* nByTS   */
    v[8] = PROTECT(Rf_allocVector(INTSXP, l));       /* nGByTS  */
    int rc = celda_chain_get_tables(chain_of(p), INTEGER(v[0]), REAL(v[1]), REAL(v[2]), REAL(v[3]), REAL(v[4]), REAL(v[5]),
                                    REAL(v[6]), REAL(v[7]), INTEGER(v[8]));
    static const char *const names[] = {"mCPByS", "nTSByC", "nTSByCP", "nGByCP", "nCP", "nByC", "nByG", "nByTS", "nGByTS"};
    SEXP out = named_list(9, names, v);
    UNPROTECT(9);
    CHECK(rc);
    return out;
}

/* which GPU the stateless entry points use; celdaGridSearch workers call it once with their share of the devices */
SEXP _celda_cuda_set_device(SEXP device) {
    CHECK(celda_cuda_set_device(Rf_asInteger(device)));
    return R_NilValue;
}
SEXP _celda_cuda_device_count(void) {
    int n = 0;
    CHECK(celda_cuda_device_count(&n));
    return Rf_ScalarInteger(n);
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
    {"_celda_cuda_cCCalcGibbsProbZ", (DL_FUNC)&_celda_cuda_cCCalcGibbsProbZ, 15},
    {"_celda_cuda_cGCalcGibbsProbY", (DL_FUNC)&_celda_cuda_cGCalcGibbsProbY, 14},
    {"_celda_cuda_cCCalcEMProbZ", (DL_FUNC)&_celda_cuda_cCCalcEMProbZ, 13},
    {"_celda_cuda_cCCalcLL", (DL_FUNC)&_celda_cuda_cCCalcLL, 9},
    {"_celda_cuda_cGCalcLL", (DL_FUNC)&_celda_cuda_cGCalcLL, 10},
    {"_celda_cuda_chain_sweep_y", (DL_FUNC)&_celda_cuda_chain_sweep_y, 4},
    {"_celda_cuda_chain_sweep_z_gibbs", (DL_FUNC)&_celda_cuda_chain_sweep_z_gibbs, 4},
    {"_celda_cuda_chain_sweep_z_em", (DL_FUNC)&_celda_cuda_chain_sweep_z_em, 2},
    {"_celda_cuda_chain_get_probs", (DL_FUNC)&_celda_cuda_chain_get_probs, 5},
    {"_celda_cuda_chain_get_tables", (DL_FUNC)&_celda_cuda_chain_get_tables, 6},
    {"_celda_cuda_set_device", (DL_FUNC)&_celda_cuda_set_device, 1},
    {"_celda_cuda_device_count", (DL_FUNC)&_celda_cuda_device_count, 0},
    {NULL, NULL, 0}};

/* called from R_init_celda (src/RcppExports.cpp:339-342) next to the Rcpp-generated table */
void celda_cuda_register(DllInfo *dll) { R_registerRoutines(dll, NULL, CudaEntries, NULL, NULL); }
