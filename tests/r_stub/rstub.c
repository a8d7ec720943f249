/* Miniature R runtime behind tests/r_stub/include: vectors with dim / levels / S4 slots, external pointers with
 * finalizers, Rf_error as a longjmp back to rstub_call, and the .Call routine table R_registerRoutines receives.
 * Test infrastructure (tests/test_r_shim.py): it lets r/celda_cuda_shim.c be compiled and driven the way R's .Call
 * would drive it, without an R installation.  Not part of the product. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

struct rstub_slot {
    char name[32];
    SEXP value;
    struct rstub_slot *next;
};

struct rstub_sexp {
    SEXPTYPE type;
    R_xlen_t length;
    void *data;          /* int / double / SEXP payload, or the external pointer's address */
    int has_dim, nrow, ncol;
    int nlevels;         /* >= 0: an R factor with that many levels */
    char sym[32];        /* SYMSXP */
    struct rstub_slot *slots;
    SEXP names;          /* the "names" attribute (a STRSXP) */
    R_CFinalizer_t fin;
    struct rstub_sexp *next_alloc;
};

static struct rstub_sexp nil_value = {NILSXP, 0, NULL, 0, 0, 0, -1, "", NULL, NULL, NULL, NULL};
static struct rstub_sexp names_symbol = {SYMSXP, 0, NULL, 0, 0, 0, -1, "names", NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_value;
SEXP R_NamesSymbol = &names_symbol;
int R_NaInt = (-2147483647 - 1);

static struct rstub_sexp *arena = NULL;
static int protect_depth = 0, protect_max = 0, protect_underflow = 0;
static jmp_buf *err_jmp = NULL;
static char err_msg[1024];
static const R_CallMethodDef *routines = NULL;

static SEXP node(SEXPTYPE type) {
    SEXP s = (SEXP)calloc(1, sizeof(*s));
    s->type = type;
    s->nlevels = -1;
    s->next_alloc = arena;
    arena = s;
    return s;
}

static void need(SEXP x, SEXPTYPE type, const char *what) {
    if (!x || x->type != type) Rf_error("%s: wrong SEXP type %d", what, x ? (int)x->type : -1);
}

/* ---------------------------------------------------------------- the subset of R's API the shim uses */
int *INTEGER(SEXP x) { need(x, INTSXP, "INTEGER()"); return (int *)x->data; }
double *REAL(SEXP x) { need(x, REALSXP, "REAL()"); return (double *)x->data; }
R_xlen_t XLENGTH(SEXP x) { return x->length; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) {
    need(x, VECSXP, "SET_VECTOR_ELT()");
    if (i < 0 || i >= x->length) Rf_error("SET_VECTOR_ELT(): index out of range");
    ((SEXP *)x->data)[i] = v;
    return v;
}
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) {
    need(x, VECSXP, "VECTOR_ELT()");
    if (i < 0 || i >= x->length) Rf_error("VECTOR_ELT(): index out of range");
    return ((SEXP *)x->data)[i];
}

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) {
    size_t w = (type == INTSXP || type == LGLSXP) ? sizeof(int) : type == REALSXP ? sizeof(double)
               : (type == VECSXP || type == STRSXP) ? sizeof(SEXP) : 0;
    if (!w || n < 0) Rf_error("allocVector: unsupported type %u or negative length", type);
    SEXP s = node(type);
    s->length = n;
    s->data = calloc((size_t)n + 1, w);
    if (type == VECSXP || type == STRSXP)
        for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)s->data)[i] = R_NilValue;
    return s;
}
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
    if (nrow < 0 || ncol < 0) Rf_error("negative extents to matrix");
    SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
    s->has_dim = 1;
    s->nrow = nrow;
    s->ncol = ncol;
    return s;
}
SEXP Rf_coerceVector(SEXP x, SEXPTYPE type) {
    if (x->type == type) return x;
    if (x->type == REALSXP && type == INTSXP) {
        SEXP s = Rf_allocVector(INTSXP, x->length);
        for (R_xlen_t i = 0; i < x->length; ++i) ((int *)s->data)[i] = (int)((double *)x->data)[i];
        s->has_dim = x->has_dim, s->nrow = x->nrow, s->ncol = x->ncol; /* attributes are kept */
        return s;
    }
    if ((x->type == INTSXP || x->type == LGLSXP) && type == REALSXP) {
        SEXP s = Rf_allocVector(REALSXP, x->length);
        for (R_xlen_t i = 0; i < x->length; ++i) ((double *)s->data)[i] = (double)((int *)x->data)[i];
        s->has_dim = x->has_dim, s->nrow = x->nrow, s->ncol = x->ncol;
        return s;
    }
    Rf_error("coerceVector: unsupported conversion %u -> %u", x->type, type);
}
SEXP Rf_ScalarReal(double v) {
    SEXP s = Rf_allocVector(REALSXP, 1);
    ((double *)s->data)[0] = v;
    return s;
}
SEXP Rf_ScalarInteger(int v) {
    SEXP s = Rf_allocVector(INTSXP, 1);
    ((int *)s->data)[0] = v;
    return s;
}
SEXP Rf_mkChar(const char *str) {
    SEXP s = node(CHARSXP);
    s->length = (R_xlen_t)strlen(str);
    s->data = malloc(strlen(str) + 1);
    strcpy((char *)s->data, str);
    return s;
}
const char *R_CHAR(SEXP x) { need(x, CHARSXP, "CHAR()"); return (const char *)x->data; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) {
    need(x, STRSXP, "SET_STRING_ELT()");
    need(v, CHARSXP, "SET_STRING_ELT() value");
    if (i < 0 || i >= x->length) Rf_error("SET_STRING_ELT(): index out of range");
    ((SEXP *)x->data)[i] = v;
}
SEXP STRING_ELT(SEXP x, R_xlen_t i) {
    need(x, STRSXP, "STRING_ELT()");
    if (i < 0 || i >= x->length) Rf_error("STRING_ELT(): index out of range");
    return ((SEXP *)x->data)[i];
}
int *LOGICAL(SEXP x) { need(x, LGLSXP, "LOGICAL()"); return (int *)x->data; }
SEXP Rf_duplicate(SEXP x) {
    if (x->type != INTSXP && x->type != REALSXP && x->type != LGLSXP) Rf_error("duplicate: unsupported type %u", x->type);
    SEXP s = Rf_allocVector(x->type, x->length);
    memcpy(s->data, x->data, (size_t)x->length * (x->type == REALSXP ? sizeof(double) : sizeof(int)));
    s->has_dim = x->has_dim, s->nrow = x->nrow, s->ncol = x->ncol, s->nlevels = x->nlevels;
    return s;
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP value) {
    need(name, SYMSXP, "setAttrib()");
    if (strcmp(name->sym, "names")) Rf_error("setAttrib: only \"names\" is supported");
    need(value, STRSXP, "setAttrib(names)");
    if (value->length != x->length) Rf_error("'names' attribute must be the same length as the vector");
    x->names = value;
    return value;
}
SEXP Rf_getAttrib(SEXP x, SEXP name) {
    need(name, SYMSXP, "getAttrib()");
    return (!strcmp(name->sym, "names") && x->names) ? x->names : R_NilValue;
}
int Rf_asLogical(SEXP x) {
    if (x->length < 1) Rf_error("asLogical: zero-length argument");
    if (x->type == LGLSXP || x->type == INTSXP) return ((int *)x->data)[0] != 0;
    if (x->type == REALSXP) return ((double *)x->data)[0] != 0.0;
    Rf_error("asLogical: unsupported type %u", x->type);
}
Rboolean Rf_isMatrix(SEXP x) { return x->has_dim ? TRUE : FALSE; }
int IS_S4_OBJECT(SEXP x) { return x->type == S4SXP; }
SEXP Rf_install(const char *name) {
    SEXP s = node(SYMSXP);
    snprintf(s->sym, sizeof(s->sym), "%s", name);
    return s;
}
SEXP R_do_slot(SEXP obj, SEXP name) {
    need(name, SYMSXP, "R_do_slot()");
    for (struct rstub_slot *p = obj->slots; p; p = p->next)
        if (!strcmp(p->name, name->sym)) return p->value;
    Rf_error("no slot of name \"%s\" for this object", name->sym);
}
int Rf_asInteger(SEXP x) {
    if (x->length < 1) Rf_error("asInteger: zero-length argument");
    if (x->type == INTSXP || x->type == LGLSXP) return ((int *)x->data)[0];
    if (x->type == REALSXP) return (int)((double *)x->data)[0];
    Rf_error("asInteger: unsupported type %u", x->type);
}
double Rf_asReal(SEXP x) {
    if (x->length < 1) Rf_error("asReal: zero-length argument");
    if (x->type == REALSXP) return ((double *)x->data)[0];
    if (x->type == INTSXP) return (double)((int *)x->data)[0];
    Rf_error("asReal: unsupported type %u", x->type);
}
int Rf_length(SEXP x) { return (int)x->length; }
int Rf_nrows(SEXP x) { return x->has_dim ? x->nrow : (int)x->length; }
int Rf_ncols(SEXP x) { return x->has_dim ? x->ncol : 1; }
int Rf_nlevels(SEXP x) { return x->nlevels < 0 ? 0 : x->nlevels; }
Rboolean Rf_isFactor(SEXP x) { return (x->type == INTSXP && x->nlevels >= 0) ? TRUE : FALSE; }
Rboolean Rf_isNull(SEXP x) { return x->type == NILSXP ? TRUE : FALSE; }

void Rf_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_msg, sizeof(err_msg), fmt, ap);
    va_end(ap);
    if (!err_jmp) {
        fprintf(stderr, "rstub: Rf_error outside rstub_call: %s\n", err_msg);
        abort();
    }
    longjmp(*err_jmp, 1);
}

SEXP Rf_protect(SEXP x) {
    if (++protect_depth > protect_max) protect_max = protect_depth;
    return x;
}
void Rf_unprotect(int n) {
    protect_depth -= n;
    if (protect_depth < 0) {
        protect_underflow = 1;
        protect_depth = 0;
    }
}

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
    (void)tag;
    (void)prot;
    SEXP s = node(EXTPTRSXP);
    s->data = p;
    return s;
}
void *R_ExternalPtrAddr(SEXP s) { need(s, EXTPTRSXP, "R_ExternalPtrAddr()"); return s->data; }
void R_ClearExternalPtr(SEXP s) { need(s, EXTPTRSXP, "R_ClearExternalPtr()"); s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) {
    (void)onexit;
    need(s, EXTPTRSXP, "R_RegisterCFinalizerEx()");
    s->fin = fun;
}

int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
    (void)info;
    (void)c;
    (void)f;
    (void)e;
    routines = call;
    return 1;
}

/* ---------------------------------------------------------------- what the Python test drives (ctypes) */
SEXP rstub_nil(void) { return R_NilValue; }
SEXP rstub_int(R_xlen_t n, const int *v) {
    SEXP s = Rf_allocVector(INTSXP, n);
    if (n) memcpy(s->data, v, sizeof(int) * (size_t)n);
    return s;
}
SEXP rstub_real(R_xlen_t n, const double *v) {
    SEXP s = Rf_allocVector(REALSXP, n);
    if (n) memcpy(s->data, v, sizeof(double) * (size_t)n);
    return s;
}
SEXP rstub_logical(int v) {
    SEXP s = Rf_allocVector(LGLSXP, 1);
    ((int *)s->data)[0] = v;
    return s;
}
const char *rstub_name(SEXP x, long long i) { return x->names ? R_CHAR(((SEXP *)x->names->data)[i]) : ""; }
void rstub_set_dim(SEXP x, int nrow, int ncol) { x->has_dim = 1; x->nrow = nrow; x->ncol = ncol; }
void rstub_set_levels(SEXP x, int nlevels) { x->nlevels = nlevels; }
SEXP rstub_s4(void) { return node(S4SXP); }
void rstub_set_slot(SEXP obj, const char *name, SEXP v) {
    struct rstub_slot *p = (struct rstub_slot *)calloc(1, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "%s", name);
    p->value = v;
    p->next = obj->slots;
    obj->slots = p;
}
int rstub_type(SEXP x) { return (int)x->type; }
long long rstub_length(SEXP x) { return (long long)x->length; }
void *rstub_data(SEXP x) { return x->data; }
int rstub_nrow(SEXP x) { return Rf_nrows(x); }
int rstub_ncol(SEXP x) { return Rf_ncols(x); }
SEXP rstub_vector_elt(SEXP x, long long i) { return ((SEXP *)x->data)[i]; }
int rstub_protect_balance(void) { return protect_underflow ? -1 : protect_depth; }

int rstub_n_routines(void) {
    int n = 0;
    while (routines && routines[n].name) ++n;
    return n;
}
const char *rstub_routine_name(int i) { return routines[i].name; }
int rstub_routine_nargs(int i) { return routines[i].numArgs; }

typedef SEXP (*F0)(void);
typedef SEXP (*F1)(SEXP);
typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F11)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F12)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F13)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F14)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F15)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F16)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

/* .Call(name, ...): looks the routine up in the registered table, checks the arity like R does, runs it under a
 * setjmp that Rf_error returns to.  NULL + message in `err` on an R error. */
SEXP rstub_call(const char *name, int nargs, SEXP *a, char *err, int errlen) {
    const R_CallMethodDef *volatile r = routines; /* read again after the longjmp */
    err[0] = 0;
    while (r && r->name && strcmp(r->name, name)) r = r + 1;
    if (!r || !r->name) {
        snprintf(err, errlen, "\"%s\" not available for .Call()", name);
        return NULL;
    }
    if (r->numArgs != nargs) {
        snprintf(err, errlen, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, r->numArgs, name);
        return NULL;
    }
    jmp_buf jb;
    jmp_buf *outer = err_jmp;
    const int depth = protect_depth;
    SEXP out = NULL;
    err_jmp = &jb;
    if (setjmp(jb) == 0) {
        DL_FUNC f = r->fun;
        switch (nargs) {
        case 0: out = ((F0)f)(); break;
        case 1: out = ((F1)f)(a[0]); break;
        case 2: out = ((F2)f)(a[0], a[1]); break;
        case 3: out = ((F3)f)(a[0], a[1], a[2]); break;
        case 4: out = ((F4)f)(a[0], a[1], a[2], a[3]); break;
        case 5: out = ((F5)f)(a[0], a[1], a[2], a[3], a[4]); break;
        case 6: out = ((F6)f)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
        case 7: out = ((F7)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
        case 8: out = ((F8)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]); break;
        case 9: out = ((F9)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
        case 10: out = ((F10)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]); break;
        case 11: out = ((F11)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10]); break;
        case 12: out = ((F12)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11]); break;
        case 13: out = ((F13)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12]); break;
        case 14: out = ((F14)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13]); break;
        case 15: out = ((F15)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14]); break;
        case 16: out = ((F16)f)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]); break;
        default: snprintf(err, errlen, "rstub_call: arity %d not wired", nargs); break;
        }
        if (out && protect_depth != depth) {
            snprintf(err, errlen, "stack imbalance in '%s': %d then %d", name, depth, protect_depth);
            out = NULL;
        }
    } else {
        snprintf(err, errlen, "%s", err_msg);
        protect_depth = depth; /* R unwinds the protect stack on error */
        out = NULL;
    }
    err_jmp = outer;
    return out;
}

/* the same for a routine given by address (a library that does not register its routines) */
static R_CallMethodDef one_routine[2];
SEXP rstub_call_ptr(void *fn, int nargs, SEXP *a, char *err, int errlen) {
    const R_CallMethodDef *saved = routines;
    union { void *p; DL_FUNC f; } cv; /* data pointer -> function pointer */
    cv.p = fn;
    one_routine[0].name = "<by address>";
    one_routine[0].fun = cv.f;
    one_routine[0].numArgs = nargs;
    one_routine[1].name = NULL;
    routines = one_routine;
    SEXP out = rstub_call("<by address>", nargs, a, err, errlen);
    routines = saved;
    return out;
}

/* gc() at the end of a session: run the finalizers of live external pointers, then free everything */
int rstub_reset(void) {
    int finalized = 0;
    for (struct rstub_sexp *s = arena; s; s = s->next_alloc)
        if (s->type == EXTPTRSXP && s->fin) {
            s->fin(s);
            s->fin = NULL;
            ++finalized;
        }
    while (arena) {
        struct rstub_sexp *s = arena;
        arena = s->next_alloc;
        if (s->type != EXTPTRSXP) free(s->data);
        while (s->slots) {
            struct rstub_slot *p = s->slots;
            s->slots = p->next;
            free(p);
        }
        free(s);
    }
    protect_depth = protect_max = protect_underflow = 0;
    return finalized;
}
