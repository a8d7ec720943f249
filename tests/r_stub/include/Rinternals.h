/* Miniature stand-in for R's C API: just enough of Rinternals.h to compile and run r/celda_cuda_shim.c without an
 * R installation (tests/test_r_shim.py).  Test infrastructure, not part of the product: the semantics follow
 * "Writing R Extensions" for the handful of entry points the shim uses; objects live in an arena that
 * rstub_reset() frees.  A real build uses R's own headers. */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rstub_sexp *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;

#define NILSXP 0
#define SYMSXP 1
#define CHARSXP 9
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
#define EXTPTRSXP 22
#define S4SXP 25

extern SEXP R_NilValue;
extern SEXP R_NamesSymbol;

int *INTEGER(SEXP x);
double *REAL(SEXP x);
R_xlen_t XLENGTH(SEXP x);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
const char *R_CHAR(SEXP x);
#define CHAR(x) R_CHAR(x)
int *LOGICAL(SEXP x);

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type);
SEXP Rf_ScalarReal(double v);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_mkChar(const char *s);
SEXP Rf_duplicate(SEXP x);
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value);
SEXP Rf_getAttrib(SEXP x, SEXP name);
int Rf_asLogical(SEXP x);
Rboolean Rf_isMatrix(SEXP x);
int IS_S4_OBJECT(SEXP x);
SEXP Rf_install(const char *name);
SEXP R_do_slot(SEXP obj, SEXP name);
int Rf_asInteger(SEXP x);
double Rf_asReal(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
int Rf_nlevels(SEXP x);
Rboolean Rf_isFactor(SEXP x);
Rboolean Rf_isNull(SEXP x);
void Rf_error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));

SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)

typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

int Rf_length(SEXP x);
#define LENGTH(x) Rf_length(x)
extern int R_NaInt;
#define NA_INTEGER R_NaInt

/* R's short names (Rinternals.h without R_NO_REMAP): what the reference's own C sources use */
#ifndef R_NO_REMAP
#define error(...) Rf_error(__VA_ARGS__)
#define allocMatrix(t, r, c) Rf_allocMatrix(t, r, c)
#define allocVector(t, n) Rf_allocVector(t, n)
#define isFactor(x) Rf_isFactor(x)
#define nlevels(x) Rf_nlevels(x)
#define nrows(x) Rf_nrows(x)
#define ncols(x) Rf_ncols(x)
#define length(x) Rf_length(x)
#endif

#ifdef __cplusplus
}
#endif
#endif
