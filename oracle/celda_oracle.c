/*
 * celda_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A plain-C restatement of the reference's collapsed-Gibbs hot path (campbio/celda 1.18.2),
 * written in the reference's own loop order.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (celda_b200/csrc, libcelda_cuda.so) never links, imports or calls it.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 *
 * PARITY PIN (1), the reference's own natives: there is no R, Rcpp or Eigen in the build container, but
 * the reference's native sources compile UNMODIFIED against miniatures of the R C API and of the Rcpp /
 * Eigen containers (tests/r_stub, my own code): `make ref` builds oracle/_ref/ from
 * /root/reference/src/{matrixSums.c, perplexity.c, matrixSumsSparse.cpp, cG_calcGibbsProbY.cpp}, and
 * tests/test_reference_natives.py holds this file against them -- the twelve aggregation natives and
 * _perplexityG bit for bit and message for message, the y kernel cG_CalcGibbsProbY to 2-3 ulp of its
 * largest lgamma term (libm lgamma there, R's lgammafn here) with identical argmax and draws.
 * PARITY PIN (2), the R closures (z sweep, EM step, log-likelihoods, decompositions, sampler) need an R
 * interpreter and cannot run here; they are pinned against the reference's own
 * fixtures: tests/test_oracle_golden.py checks the three log-likelihood functions
 * against data/celda{CG,C,G}Mod.rda finalLogLik (bit-identical or <= 4e-16 relative), the 9
 * grid-search models of data/celdaCGGridSearchRes.rda, the aggregation known answers of
 * tests/testthat/test-matrixSums.R, and the Delta-probs == Delta-LL identity that ties the
 * sweep probabilities to the pinned log-likelihood.  The sweep draws themselves have no
 * reference test or fixture (SURVEY.md section 8c "what stays unpinned"): parity of the
 * draws is "unpinned by the reference", pinned only through those identities.
 *
 * Third-party arithmetic that is NOT under /root/reference (base R, unpinned version,
 * DESCRIPTION:17 "R (>= 4.0)") is restated from its published algorithm:
 *   - lgamma:  R nmath lgammafn/gammafn/lgammacor/chebyshev_eval (SLATEC-derived series).
 *   - log/exp: the platform libm in R; restated here as the fdlibm (Sun, 1993) algorithms
 *              e_log.c / e_exp.c so that results do not depend on the host libm
 *              (< 1 ulp; checked against glibc in tests/test_oracle_math.py).
 *   - sum():   R accumulates in LDOUBLE, which is 80-bit on x86 and plain double on
 *              aarch64 / --disable-long-double builds.  Default here: plain double, serial,
 *              left to right (define CELDA_ORACLE_LDOUBLE to get the x86 behaviour; the test
 *              suite checks both modes give identical draws on the fixtures).
 *   - sample.int(n, 1, TRUE, prob): do_sample -> FixupProb -> ProbSampleReplace -> revsort
 *              (R src/main/random.c, sort.c), restated in o_sample_ll().
 * All floating point here is evaluated in the written order with contraction disabled
 * (compile with -ffp-contract=off, see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef CELDA_ORACLE_LDOUBLE
typedef long double acc_t;
#else
typedef double acc_t;
#endif

static char g_err[512];
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return 1; } while (0)
const char *o_last_error(void) { return g_err; }
int o_uses_long_double(void) { return sizeof(acc_t) > sizeof(double); }

/* ------------------------------------------------------------------ deterministic libm */

static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* natural log, fdlibm e_log.c algorithm (argument reduction x = 2^k (1+f), s = f/(2+f),
 * degree-14 even polynomial in s). */
double o_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
        two54 = 1.80143985094819840000e+16,
        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
        Lg7 = 1.479819860511658591e-01;
    uint64_t ux = d2u(x);
    int32_t hx = (int32_t)(ux >> 32);
    uint32_t lx = (uint32_t)ux;
    int32_t k = 0, i, j;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -HUGE_VAL;
        if (hx < 0) return NAN;
        k -= 54; x *= two54; ux = d2u(x); hx = (int32_t)(ux >> 32);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    ux = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ux & 0xffffffffu);
    x = u2d(ux);
    k += (i >> 20);
    double f = x - 1.0, dk, R, s, z, w, t1, t2, hfsq;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; dk = (double)k; return dk * ln2_hi + dk * ln2_lo; }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

/* exp, fdlibm e_exp.c algorithm (x = k ln2 + r, |r| <= 0.5 ln2, Remez degree-5 for
 * r coth(r/2)-style rational form). */
double o_exp(double x)
{
    static const double halF[2] = {0.5, -0.5}, huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
        o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
        ln2HI[2] = {6.93147180369123816490e-01, -6.93147180369123816490e-01},
        ln2LO[2] = {1.90821492927058770002e-10, -1.90821492927058770002e-10},
        invln2 = 1.44269504088896338700e+00,
        P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
        P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    double y, hi = 0.0, lo = 0.0, c, t;
    int32_t k = 0, xsb;
    uint64_t ux = d2u(x);
    uint32_t hx = (uint32_t)(ux >> 32);
    xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | (uint32_t)ux) != 0) return x + x;
            return (xsb == 0) ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) { hi = x - ln2HI[xsb]; lo = ln2LO[xsb]; k = 1 - xsb - xsb; }
        else {
            k = (int32_t)(invln2 * x + halF[xsb]);
            t = k;
            hi = x - t * ln2HI[0];
            lo = t * ln2LO[0];
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) {
        if (huge + x > 1.0) return 1.0 + x;
    } else k = 0;
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) { return u2d(d2u(y) + ((uint64_t)(int64_t)k << 52)); }
    return u2d(d2u(y) + ((uint64_t)(int64_t)(k + 1000) << 52)) * twom1000;
}

/* R nmath chebyshev_eval (Clenshaw recurrence, coefficients a[0..n-1]). */
static double o_cheb(double x, const double *a, int n)
{
    double b0 = 0, b1 = 0, b2 = 0, twox = x * 2;
    for (int i = 1; i <= n; i++) { b2 = b1; b1 = b0; b0 = twox * b1 - b2 + a[n - i]; }
    return (b0 - b2) * 0.5;
}

/* R nmath gammafn restricted to 0 < x <= 10 (the only range lgammafn sends it). */
static double o_gamma_small(double x)
{
    static const double gamcs[22] = {
        +.8571195590989331421920062399942e-2, +.4415381324841006757191315771652e-2,
        +.5685043681599363378632664588789e-1, -.4219835396418560501012500186624e-2,
        +.1326808181212460220584006796352e-2, -.1893024529798880432523947023886e-3,
        +.3606925327441245256578082217225e-4, -.6056761904460864218485548290365e-5,
        +.1055829546302283344731823509093e-5, -.1811967365542384048291855891166e-6,
        +.3117724964715322277790254593169e-7, -.5354219639019687140874081024347e-8,
        +.9193275519859588946887786825940e-9, -.1577941280288339761767423273953e-9,
        +.2707980622934954543266540433089e-10, -.4646818653825730144081661058933e-11,
        +.7973350192007419656460767175359e-12, -.1368078209830916025799499172309e-12,
        +.2347319486563800657233471771688e-13, -.4027432614949066932766570534699e-14,
        +.6910051747372100912138336975257e-15, -.1185584500221992907052387126192e-15};
    int n = (int)x;
    double y = x - n;
    --n;
    double value = o_cheb(y * 2 - 1, gamcs, 22) + .9375;
    if (n == 0) return value;
    if (n < 0) { /* 0 < x < 1 : one step of the downward recursion */
        return value / x;
    }
    for (int i = 1; i <= n; i++) value *= (y + i);
    return value;
}

/* R nmath lgammacor for x >= 10. */
static double o_lgammacor(double x)
{
    static const double algmcs[5] = {
        +.1666389480451863247205729650822e+0, -.1384948176067563840732986059135e-4,
        +.9810825646924729426157171547487e-8, -.1809129475572494194263306266719e-10,
        +.6221098041892605227126015543416e-13};
    if (x < 94906265.62425156) { double t = 10 / x; return o_cheb(t * t * 2 - 1, algmcs, 5) / x; }
    return 1 / (x * 12);
}

/* R nmath lgammafn, positive arguments (every argument on the path is count + prior > 0). */
double o_lgamma(double x)
{
    static const double M_LN_SQRT_2PI_ = 0.918938533204672741780329736406;
    if (!(x > 0)) return (x == 0) ? HUGE_VAL : NAN; /* pole / NA propagate */
    if (x < 1e-306) return -o_log(x);
    if (x <= 10) return o_log(fabs(o_gamma_small(x)));
    if (x > 1e17) return x * (o_log(x) - 1.);
    if (x > 4934720.) return M_LN_SQRT_2PI_ + (x - 0.5) * o_log(x) - x;
    return M_LN_SQRT_2PI_ + (x - 0.5) * o_log(x) - x + o_lgammacor(x);
}

void o_log_vec(const double *x, long n, double *out) { for (long i = 0; i < n; i++) out[i] = o_log(x[i]); }
void o_exp_vec(const double *x, long n, double *out) { for (long i = 0; i < n; i++) out[i] = o_exp(x[i]); }
void o_lgamma_vec(const double *x, long n, double *out) { for (long i = 0; i < n; i++) out[i] = o_lgamma(x[i]); }

/* ------------------------------------------------------------------ aggregation */

static int check_labels(const int *g, long n, int K, const char *what, const char *kname)
{
    for (long i = 0; i < n; i++)
        if (g[i] < 1 || g[i] > K) FAIL("The entries in '%s' need to be between 1 and '%s'.", what, kname);
    return 0;
}

/* src/matrixSumsSparse.cpp:9-38.  out: nr x K column-major, zero-filled here. */
int o_colSumByGroupSparse(int nr, int nc, const int *p, const int *ri, const double *x, const int *group,
                          long ngroup, int K, double *out)
{
    if (nc != ngroup) FAIL("Length of 'group' must be equal to the number of columns in 'counts'.");
    for (long i = 0; i < ngroup; i++) /* :21 -- this one message of the file has no full stop */
        if (group[i] < 1 || group[i] > K) FAIL("The entries in 'group' need to be between 1 and 'K'");
    if (K > nc) FAIL("'K' cannot be bigger than the number of columns in 'counts'.");
    memset(out, 0, sizeof(double) * (size_t)nr * K);
    for (int j = 0; j < nc; ++j)
        for (int t = p[j]; t < p[j + 1]; ++t) out[(size_t)(group[j] - 1) * nr + ri[t]] += x[t];
    return 0;
}

/* src/matrixSumsSparse.cpp:44-73.  out: L x nc column-major. */
int o_rowSumByGroupSparse(int nr, int nc, const int *p, const int *ri, const double *x, const int *group,
                          long ngroup, int L, double *out)
{
    if (nr != ngroup) FAIL("Length of 'group' must be equal to the number of rows in 'counts'.");
    if (check_labels(group, ngroup, L, "group", "L")) return 1;
    if (L > nr) FAIL("'L' cannot be bigger than the number of rows in 'counts'.");
    memset(out, 0, sizeof(double) * (size_t)L * nc);
    for (int j = 0; j < nc; ++j)
        for (int t = p[j]; t < p[j + 1]; ++t) out[(size_t)j * L + group[ri[t]] - 1] += x[t];
    return 0;
}

/* src/matrixSumsSparse.cpp:85-131.  px (nr x K) is updated IN PLACE (quirk Q3). */
int o_colSumByGroupChangeSparse(int nr, int nc, const int *p, const int *ri, const double *x, double *px,
                                int px_rows, const int *group, long ngroup, const int *pgroup, long npgroup, int K)
{
    if (nc != ngroup) FAIL("Length of 'group' must be equal to the number of columns in 'counts'.");
    if (ngroup != npgroup) FAIL("Length of 'group' must equal 'pgroup'.");
    if (check_labels(group, ngroup, K, "group", "K")) return 1;
    if (check_labels(pgroup, npgroup, K, "pgroup", "K")) return 1;
    if (px_rows != nr) FAIL("'px' and 'counts' must have the same number of rows.");
    if (K > nc) FAIL("'K' cannot be bigger than the number of columns in 'counts'.");
    for (int j = 0; j < nc; ++j)
        if (group[j] != pgroup[j])
            for (int t = p[j]; t < p[j + 1]; ++t) {
                px[(size_t)(group[j] - 1) * nr + ri[t]] += x[t];
                px[(size_t)(pgroup[j] - 1) * nr + ri[t]] -= x[t];
            }
    return 0;
}

/* src/matrixSumsSparse.cpp:141-187.  px (L x nc) updated in place; full nnz scan like the reference. */
int o_rowSumByGroupChangeSparse(int nr, int nc, const int *p, const int *ri, const double *x, double *px,
                                int px_cols, const int *group, long ngroup, const int *pgroup, long npgroup, int L)
{
    if (nr != ngroup) FAIL("Length of 'group' must be equal to the number of rows in 'counts'.");
    if (ngroup != npgroup) FAIL("Length of 'group' must equal 'pgroup'.");
    if (check_labels(group, ngroup, L, "group", "L")) return 1;
    if (check_labels(pgroup, npgroup, L, "pgroup", "L")) return 1;
    if (px_cols != nc) FAIL("'px' and 'counts' must have the same number of rows.");
    if (L > nr) FAIL("'L' cannot be bigger than the number of rows in 'counts'.");
    for (int j = 0; j < nc; ++j)
        for (int t = p[j]; t < p[j + 1]; ++t) {
            int r = ri[t];
            if (group[r] != pgroup[r]) {
                px[(size_t)j * L + group[r] - 1] += x[t];
                px[(size_t)j * L + pgroup[r] - 1] -= x[t];
            }
        }
    return 0;
}

#define NA_INT INT32_MIN

/* Dense twins, src/matrixSums.c.  `nl` stands for nlevels(group); a negative nl stands for
 * "group is not a factor" (the R-level precondition the C code tests with isFactor).
 * int32 variants wrap silently like the reference (:42). */
#define DENSE_AGG(NAME, T)                                                                                         \
    /* src/matrixSums.c:5-48 (int) / :198-236 (numeric) */                                                         \
    int o_rowSumByGroup_##NAME(const T *x, int nr, int nc, const int *group, long ngroup, int nl, T *ans)          \
    {                                                                                                              \
        if (nl < 0) FAIL("The grouping argument must be a factor");                                                \
        if (ngroup != nr) FAIL("The length of the grouping argument must match the number of rows in the matrix"); \
        for (int i = 0; i < nr; i++) if (group[i] == NA_INT) FAIL("Labels in group and pgroup must not be NA.");  \
        memset(ans, 0, sizeof(T) * (size_t)nl * nc);                                                               \
        for (int j = 0; j < nc; j++)                                                                               \
            for (int i = 0; i < nr; i++) ans[(size_t)j * nl + group[i] - 1] += x[(size_t)j * nr + i];             \
        return 0;                                                                                                  \
    }                                                                                                              \
    /* src/matrixSums.c:50-93 / :240-278 */                                                                        \
    int o_colSumByGroup_##NAME(const T *x, int nr, int nc, const int *group, long ngroup, int nl, T *ans)          \
    {                                                                                                              \
        if (nl < 0) FAIL("The grouping argument must be a factor");                                                \
        if (ngroup != nc)                                                                                          \
            FAIL("The length of the grouping argument must match the number of columns in the matrix");            \
        for (int i = 0; i < nc; i++) if (group[i] == NA_INT) FAIL("Labels in group and pgroup must not be NA.");  \
        memset(ans, 0, sizeof(T) * (size_t)nr * nl);                                                               \
        for (int j = 0; j < nc; j++)                                                                               \
            for (int i = 0; i < nr; i++) ans[(size_t)(group[j] - 1) * nr + i] += x[(size_t)j * nr + i];           \
        return 0;                                                                                                  \
    }                                                                                                              \
    /* src/matrixSums.c:97-143 / :282-327 ; px (nl x nc) in place */                                               \
    int o_rowSumByGroupChange_##NAME(const T *x, int nr, int nc, T *px, int px_nr, int px_nc, const int *group,    \
                                     long ngroup, int nl, const int *pgroup, long npgroup, int nlp)                \
    {                                                                                                              \
        if (nl < 0 || nlp < 0) FAIL("The grouping arguments must be factors");                                     \
        if (nl != nlp || nl != px_nr)                                                                              \
            FAIL("group and pgroup must have the same number of levels equal to row number of px");                \
        if (nc != px_nc) FAIL("x and the previously summed matrix, px, must have the same number of columns.");    \
        if (ngroup != npgroup || ngroup != nr)                                                                     \
            FAIL("group label and previous group label must be the same length as the number of rows in x.");      \
        for (int i = 0; i < nr; i++)                                                                               \
            if (group[i] == NA_INT || pgroup[i] == NA_INT) FAIL("Labels in group and pgroup must not be NA.");     \
        for (int i = 0; i < nr; i++)                                                                               \
            if (pgroup[i] != group[i])                                                                             \
                for (int j = 0; j < nc; j++) {                                                                     \
                    px[(size_t)j * nl + pgroup[i] - 1] -= x[(size_t)j * nr + i];                                   \
                    px[(size_t)j * nl + group[i] - 1] += x[(size_t)j * nr + i];                                    \
                }                                                                                                  \
        return 0;                                                                                                  \
    }                                                                                                              \
    /* src/matrixSums.c:147-194 / :331-379 ; px (nr x nl) in place */                                              \
    int o_colSumByGroupChange_##NAME(const T *x, int nr, int nc, T *px, int px_nr, int px_nc, const int *group,    \
                                     long ngroup, int nl, const int *pgroup, long npgroup, int nlp)                \
    {                                                                                                              \
        if (nl < 0 || nlp < 0) FAIL("The grouping arguments must be factors");                                     \
        if (nl != nlp || nl != px_nc)                                                                              \
            FAIL("group and pgroup must have the same number of levels equal to column number of px");             \
        if (nr != px_nr) FAIL("x and the previously summed matrix, pxc must have the same number of rows");        \
        if (ngroup != npgroup || ngroup != nc)                                                                     \
            FAIL("group label and previous group label must be the same length as the number of columns in x.");   \
        for (int i = 0; i < nc; i++)                                                                               \
            if (group[i] == NA_INT || pgroup[i] == NA_INT) FAIL("Labels in group and pgroup must not be NA.");     \
        for (int j = 0; j < nc; j++)                                                                               \
            if (group[j] != pgroup[j])                                                                             \
                for (int i = 0; i < nr; i++) {                                                                     \
                    px[(size_t)(group[j] - 1) * nr + i] += x[(size_t)j * nr + i];                                  \
                    px[(size_t)(pgroup[j] - 1) * nr + i] -= x[(size_t)j * nr + i];                                 \
                }                                                                                                  \
        return 0;                                                                                                  \
    }

/* int32 accumulation must wrap like the reference; do it in unsigned to keep it defined in C */
typedef int32_t i32w;
DENSE_AGG(numeric, double)
#pragma GCC diagnostic push
DENSE_AGG(int, i32w)
#pragma GCC diagnostic pop

/* table(factor(z, levels = 1:K), s)  (R/celda_CG.R:904, R/celda_C.R:805): K x nS, column-major, int. */
void o_table_zs(const int *z, const int *s, long nM, int K, int nS, int *m)
{
    memset(m, 0, sizeof(int) * (size_t)K * nS);
    for (long c = 0; c < nM; c++) m[(size_t)(s[c] - 1) * K + z[c] - 1] += 1;
}

/* ------------------------------------------------------------------ sampler */

/* R sort.c revsort: descending heapsort of a[] carrying ib[] (NOT stable). */
static void o_revsort(double *a, int *ib, int n)
{
    int l, j, ir, i, ii;
    double ra;
    if (n <= 1) return;
    a--; ib--;
    l = (n >> 1) + 1;
    ir = n;
    for (;;) {
        if (l > 1) { l = l - 1; ra = a[l]; ii = ib[l]; }
        else {
            ra = a[ir]; ii = ib[ir];
            a[ir] = a[1]; ib[ir] = ib[1];
            if (--ir == 1) { a[1] = ra; ib[1] = ii; return; }
        }
        i = l; j = l << 1;
        while (j <= ir) {
            if (j < ir && a[j] > a[j + 1]) ++j;
            if (ra > a[j]) { a[i] = a[j]; ib[i] = ib[j]; j += (i = j); }
            else j = ir + 1;
        }
        a[i] = ra; ib[i] = ii;
    }
}

/* .sampleLl (R/celda_functions.R:1-11) with the uniform INJECTED:
 * exp(ll - max) ; / sum ; sample.int(n, 1, TRUE, prob) = FixupProb (renormalise by a plain-double sum of
 * the positive entries) ; Walker alias if > 200 entries have n*p > 0.1 ;
 * ProbSampleReplace (revsort descending, cumulative sum, first j < n-1 with u <= cum[j]).
 * Returns the 1-based label.  *margin (optional) = min_j |u - cum[j]| over the scanned prefix. */
int o_sample_ll(const double *ll, int n, double u, double *work_p, int *work_perm, double *margin)
{
    double mx = ll[0];
    for (int i = 1; i < n; i++) if (ll[i] > mx) mx = ll[i];
    acc_t sum = 0;
    for (int i = 0; i < n; i++) { work_p[i] = o_exp(ll[i] - mx); sum += work_p[i]; }
    double dsum = (double)sum;
    for (int i = 0; i < n; i++) work_p[i] = work_p[i] / dsum;
    /* FixupProb */
    double s2 = 0.0;
    int npos = 0;
    for (int i = 0; i < n; i++) {
        if (!isfinite(work_p[i]) || work_p[i] < 0) return -2; /* "NA in probability vector" */
        if (work_p[i] > 0) { npos++; s2 += work_p[i]; }
    }
    if (npos == 0) return -3; /* "too few positive probabilities" */
    for (int i = 0; i < n; i++) work_p[i] /= s2;
    int nc = 0;
    for (int i = 0; i < n; i++) if (n * work_p[i] > 0.1) nc++;
    if (nc > 200) {
        /* Walker's alias method (R random.c walker_ProbSampleReplace, restated): q = n p; entries with q < 1 are listed
         * from the left of HL, the others from the right; each small entry takes the current large one as its alias,
         * which gives up 1 - q of its mass; q[i] += i; one uniform rU = u n picks cell k = floor(rU) and returns k if
         * rU < q[k], else its alias. */
        int *HL = work_perm, H = -1, Lp = n;
        int *al = (int *)malloc(sizeof(int) * (size_t)n);
        if (!al) return -4;
        for (int i = 0; i < n; i++) {
            work_p[i] = work_p[i] * n;
            if (work_p[i] < 1.) HL[++H] = i; else HL[--Lp] = i;
        }
        for (int i = 0; i < n; i++) al[i] = i;
        if (H >= 0 && Lp < n) {
            for (int k = 0; k < n - 1; k++) {
                int i = HL[k], j = HL[Lp];
                al[i] = j;
                work_p[j] += work_p[i] - 1;
                if (work_p[j] < 1.) Lp++;
                if (Lp >= n) break;
            }
        }
        for (int i = 0; i < n; i++) work_p[i] += i;
        double rU = u * n;
        int k = (int)rU;
        if (k > n - 1) k = n - 1;
        int ans = (rU < work_p[k]) ? k + 1 : al[k] + 1;
        free(al);
        if (margin) *margin = fabs(rU - work_p[k]) / n;
        return ans;
    }
    for (int i = 0; i < n; i++) work_perm[i] = i + 1;
    o_revsort(work_p, work_perm, n);
    for (int i = 1; i < n; i++) work_p[i] += work_p[i - 1];
    int j, nm1 = n - 1;
    double mg = HUGE_VAL;
    for (j = 0; j < nm1; j++) {
        double d = fabs(u - work_p[j]);
        if (d < mg) mg = d;
        if (u <= work_p[j]) break;
    }
    if (margin) *margin = mg;
    return work_perm[j];
}

/* ------------------------------------------------------------------ z sweeps */

static double sum_lgamma_col_plus(const double *ncol, const double *x, int R, double sign, double beta)
{
    acc_t s = 0;
    if (x == NULL) for (int r = 0; r < R; r++) s += o_lgamma(ncol[r] + beta);
    else if (sign > 0) for (int r = 0; r < R; r++) s += o_lgamma(ncol[r] + x[r] + beta);
    else for (int r = 0; r < R; r++) s += o_lgamma(ncol[r] - x[r] + beta);
    return (double)s;
}

/* .cCCalcGibbsProbZ (R/celda_C.R:620-714) on a dense column-major counts matrix (R x nM); in celda_CG this is
 * nTSByC with R = L (R/celda_CG.R:567-580).  ix (1-based visiting order) and u (one uniform per visited cell,
 * in visiting order) are injected.  Tables are updated in place.  probs: K x nM.
 * first/count select a sub-range of the visiting order (bounded CPU-baseline samples); margin optional [nM]. */
int o_cCCalcGibbsProbZ(const double *counts, int R, long nM, int *mCPByS, int nS, double *nGByCP, const double *nByC,
                       double *nCP, int *z, const int *s, int K, double alpha, double beta, int doSample,
                       const int *ix, const double *u, long first, long count, double *probs, double *margin)
{
    double *wp = malloc(sizeof(double) * K), *pr = malloc(sizeof(double) * K);
    int *wperm = malloc(sizeof(int) * K);
    double Rbeta = R * beta;
    int rc = 0;
    for (long t = first; t < first + count; t++) {
        long i = ix[t] - 1;
        const double *x = counts + (size_t)i * R;
        int zi = z[i] - 1, si = s[i] - 1;
        mCPByS[(size_t)si * K + zi] -= 1;                                            /* :659 */
        for (int j = 0; j < K; j++) {
            const double *ncol = nGByCP + (size_t)j * R;
            double p = o_log(mCPByS[(size_t)si * K + j] + alpha);
            if (j != zi) {                                                           /* :664-676 */
                p = p + sum_lgamma_col_plus(ncol, x, R, +1, beta);
                p = p - o_lgamma(nCP[j] + nByC[i] + Rbeta);
                p = p - sum_lgamma_col_plus(ncol, NULL, R, 0, beta);
                p = p + o_lgamma(nCP[j] + Rbeta);
            } else {                                                                 /* :681-687 */
                p = p + sum_lgamma_col_plus(ncol, NULL, R, 0, beta);
                p = p - o_lgamma(nCP[j] + Rbeta);
                p = p - sum_lgamma_col_plus(ncol, x, R, -1, beta);
                p = p + o_lgamma(nCP[j] - nByC[i] + Rbeta);
            }
            pr[j] = p;
            if (probs) probs[(size_t)i * K + j] = p;
        }
        int newz = zi;
        if (doSample) {                                                              /* :693-695 */
            int lab = o_sample_ll(pr, K, u[t], wp, wperm, margin ? &margin[t] : NULL);
            if (lab < 0) { snprintf(g_err, sizeof g_err, "sampler failure %d at cell %ld", lab, i + 1); rc = 1; break; }
            newz = lab - 1;
        }
        if (newz != zi) {                                                            /* :697-703 */
            double *a = nGByCP + (size_t)zi * R, *b = nGByCP + (size_t)newz * R;
            for (int r = 0; r < R; r++) { a[r] -= x[r]; b[r] += x[r]; }
            nCP[zi] -= nByC[i]; nCP[newz] += nByC[i];
            z[i] = newz + 1;
        }
        mCPByS[(size_t)si * K + newz] += 1;                                          /* :704 */
    }
    free(wp); free(pr); free(wperm);
    return rc;
}

/* fastNormPropLog (src/matrixNorm.cpp:36-52): log((x + a) / (colSum + nrow a)), column-major nr x nc. */
int o_fastNormPropLog(const double *x, int nr, int nc, double a, double *res)
{
    double atot = nr * a;
    for (int j = 0; j < nc; j++) {
        acc_t cs = 0; /* Rcpp sugar colSums accumulates in long double (LDOUBLE) like R */
        for (int i = 0; i < nr; i++) cs += x[(size_t)j * nr + i];
        double c = (double)cs;
        if (c + atot == 0)
            FAIL("Division by 0. Make sure colSums of counts does not contain 0 after rounding counts to integers.");
        for (int i = 0; i < nr; i++) res[(size_t)j * nr + i] = o_log((x[(size_t)j * nr + i] + a) / (c + atot));
    }
    return 0;
}

/* eigenMatMultInt / eigenMatMultNumeric (src/eigenMatMultInt.cpp:11-26): C = A^T B, A: n x k, B: n x m, C: k x m.
 * Eigen's blocked summation order is unspecified; restated as a serial dot product over n (tests use 1e-12 rel). */
void o_matMultT(const double *A, int n, int k, const double *B, int m, double *C)
{
    for (int c = 0; c < m; c++)
        for (int a = 0; a < k; a++) {
            double acc = 0;
            for (int r = 0; r < n; r++) acc += A[(size_t)a * n + r] * B[(size_t)c * n + r];
            C[(size_t)c * k + a] = acc;
        }
}

/* t(phi) %*% counts for a dgCMatrix (R/celda_C.R:974): probs[k, c] = sum over stored entries of column c, in
 * storage order, of phi[row, k] * x.  out: K x nc. */
void o_countsTimesProbsSparse(int nr, int nc, const int *p, const int *ri, const double *x, const double *phi, int K,
                              double *out)
{
    for (int c = 0; c < nc; c++)
        for (int k = 0; k < K; k++) {
            double acc = 0;
            for (int t = p[c]; t < p[c + 1]; t++) acc += phi[(size_t)k * nr + ri[t]] * x[t];
            out[(size_t)c * K + k] = acc;
        }
}

/* .cCCalcEMProbZ (R/celda_C.R:717-756) for dgCMatrix counts (sparse != 0) or dense double counts (R x nM).
 * nGByCP updated in place through colSumByGroupChange; nCP = colSums; mCPByS = table. */
int o_cCCalcEMProbZ(int sparse, int R, long nM, const int *p, const int *ri, const double *x, int *mCPByS, int nS,
                    double *nGByCP, double *nCP, int *z, const int *s, int K, double alpha, double beta, int doSample,
                    double *probs)
{
    double *md = malloc(sizeof(double) * K * nS), *theta = malloc(sizeof(double) * K * nS);
    double *phi = malloc(sizeof(double) * (size_t)R * K);
    for (int i = 0; i < K * nS; i++) md[i] = mCPByS[i];
    int rc = o_fastNormPropLog(md, K, nS, alpha, theta);                              /* :732 */
    if (!rc) rc = o_fastNormPropLog(nGByCP, R, K, beta, phi);                         /* :733 */
    if (rc) { free(md); free(theta); free(phi); return rc; }
    if (sparse) o_countsTimesProbsSparse(R, (int)nM, p, ri, x, phi, K, probs);        /* :736 */
    else o_matMultT(phi, R, K, x, (int)nM, probs);
    for (long c = 0; c < nM; c++)
        for (int k = 0; k < K; k++) probs[(size_t)c * K + k] += theta[(size_t)(s[c] - 1) * K + k];
    if (doSample) {
        int *zprev = malloc(sizeof(int) * nM);
        memcpy(zprev, z, sizeof(int) * nM);
        for (long c = 0; c < nM; c++) {                                               /* :740 which.max, first max */
            int best = 0;
            for (int k = 1; k < K; k++) if (probs[(size_t)c * K + k] > probs[(size_t)c * K + best]) best = k;
            z[c] = best + 1;
        }
        /* .cCReDecomposeCounts R/celda_C.R:825-839 */
        if (sparse) rc = o_colSumByGroupChangeSparse(R, (int)nM, p, ri, x, nGByCP, R, z, nM, zprev, nM, K);
        else rc = o_colSumByGroupChange_numeric(x, R, (int)nM, nGByCP, R, K, z, nM, K, zprev, nM, K);
        for (int k = 0; k < K; k++) {
            acc_t cs = 0;
            for (int r = 0; r < R; r++) cs += nGByCP[(size_t)k * R + r];
            nCP[k] = (double)cs;
        }
        o_table_zs(z, s, nM, K, nS, mCPByS);
        free(zprev);
    }
    free(md); free(theta); free(phi);
    return rc;
}

/* ------------------------------------------------------------------ y sweep */

static inline double tab_or_eval(const double *tab, double idx, double shift, double scale)
{
    /* lg_beta[k] = lgamma(k + beta), lg_gamma[k] = lgamma(k + gamma), lg_delta[k] = lgamma(k * delta)
     * (R/celda_CG.R:428-429,543).  tab == NULL evaluates the entry directly: same value, no table. */
    long k = (long)idx;
    if (tab) return tab[k];
    /* lgdelta <- c(NA, lgamma(seq(nG + L) * delta)): entry 0 is NA (R/celda_CG.R:429, R/celda_G.R:332) */
    if (scale != 0 && k == 0) return NAN;
    return o_lgamma(scale != 0 ? (double)k * scale : (double)k + shift);
}

/* cG_CalcGibbsProbY (src/cG_calcGibbsProbY.cpp:187-249).  counts: length ncol; nTSbyC: L x ncol column-major.
 * The scalar lgamma calls at :238-245 are libm lgamma in the reference; here o_lgamma like everything else. */
void o_cG_CalcGibbsProbY(int index, const double *counts, int ncol, const double *nTSbyC, const double *nbyTS,
                         const int *nGbyTS, const double *nbyG, const int *y, int L, const double *lg_beta,
                         const double *lg_gamma, const double *lg_delta, double beta, double gamma, double delta,
                         double *probs)
{
    int index0 = index - 1, cur = y[index0] - 1;
    for (int i = 0; i < L; i++) probs[i] = 0;
    for (int col = 0; col < ncol; col++)
        for (int i = 0; i < L; i++) {
            size_t j = (size_t)col * L + i;
            if (i == cur) {
                probs[i] += tab_or_eval(lg_beta, nTSbyC[j], beta, 0);
                probs[i] -= tab_or_eval(lg_beta, nTSbyC[j] - counts[col], beta, 0);
            } else {
                probs[i] += tab_or_eval(lg_beta, nTSbyC[j] + counts[col], beta, 0);
                probs[i] -= tab_or_eval(lg_beta, nTSbyC[j], beta, 0);
            }
        }
    for (int i = 0; i < L; i++) {
        if (i == cur) {
            probs[i] += tab_or_eval(lg_gamma, nGbyTS[i], gamma, 0);
            probs[i] -= tab_or_eval(lg_gamma, nGbyTS[i] - 1, gamma, 0);
            probs[i] += tab_or_eval(lg_delta, nGbyTS[i], 0, delta);
            probs[i] -= tab_or_eval(lg_delta, nGbyTS[i] - 1, 0, delta);
            probs[i] += o_lgamma(nbyTS[i] - nbyG[index0] + (nGbyTS[i] - 1) * delta);
            probs[i] -= o_lgamma(nbyTS[i] + nGbyTS[i] * delta);
        } else {
            probs[i] += tab_or_eval(lg_gamma, nGbyTS[i] + 1, gamma, 0);
            probs[i] -= tab_or_eval(lg_gamma, nGbyTS[i], gamma, 0);
            probs[i] += tab_or_eval(lg_delta, nGbyTS[i] + 1, 0, delta);
            probs[i] -= tab_or_eval(lg_delta, nGbyTS[i], 0, delta);
            probs[i] += o_lgamma(nbyTS[i] + nGbyTS[i] * delta);
            probs[i] -= o_lgamma(nbyTS[i] + nbyG[index0] + (nGbyTS[i] + 1) * delta);
        }
    }
}

/* cG_calcGibbsProbY_Simple (src/cG_calcGibbsProbY.cpp:8-48): brute-force definition, equal to the fast form up
 * to one additive constant per gene.  counts: nG x ncol column-major (whole matrix). */
void o_cG_calcGibbsProbY_Simple(const double *counts, int nG, int ncol, int *nGbyTS, double *nTSbyC, double *nbyTS,
                                const double *nbyG, const int *y, int L, int index, double gamma, double beta,
                                double delta, double *probs)
{
    int index0 = index - 1, cur = y[index0] - 1;
    nGbyTS[cur] -= 1; nbyTS[cur] -= nbyG[index0];
    for (int c = 0; c < ncol; c++) nTSbyC[(size_t)c * L + cur] -= counts[(size_t)c * nG + index0];
    for (int i = 0; i < L; i++) {
        nGbyTS[i] += 1; nbyTS[i] += nbyG[index0];
        for (int c = 0; c < ncol; c++) nTSbyC[(size_t)c * L + i] += counts[(size_t)c * nG + index0];
        double p = 0, a = 0;
        for (int l = 0; l < L; l++) a += o_lgamma(nGbyTS[l] + gamma);
        p += a; a = 0;
        for (size_t q = 0; q < (size_t)L * ncol; q++) a += o_lgamma(nTSbyC[q] + beta);
        p += a; a = 0;
        for (int l = 0; l < L; l++) a += o_lgamma(nGbyTS[l] * delta);
        p += a; a = 0;
        for (int l = 0; l < L; l++) a += o_lgamma(nbyTS[l] + (nGbyTS[l] * delta));
        p -= a;
        probs[i] = p;
        nGbyTS[i] -= 1; nbyTS[i] -= nbyG[index0];
        for (int c = 0; c < ncol; c++) nTSbyC[(size_t)c * L + i] -= counts[(size_t)c * nG + index0];
    }
    nGbyTS[cur] += 1; nbyTS[cur] += nbyG[index0];
    for (int c = 0; c < ncol; c++) nTSbyC[(size_t)c * L + cur] += counts[(size_t)c * nG + index0];
}

/* .cGCalcGibbsProbY (R/celda_G.R:602-660).  counts: nG x ncol column-major dense (nGByCP in celda_CG,
 * R/celda_CG.R:544-558; the densified raw matrix in celda_G).  Tables may be NULL (direct evaluation). */
int o_cGCalcGibbsProbY(const double *counts, int nG, int ncol, double *nTSByC, double *nByTS, int *nGByTS,
                       const double *nByG, int *y, int L, double beta, double delta, double gamma,
                       const double *lgbeta, const double *lggamma, const double *lgdelta, int doSample,
                       const int *ix, const double *u, long first, long count, double *probs, double *margin)
{
    double *row = malloc(sizeof(double) * ncol), *pr = malloc(sizeof(double) * L), *wp = malloc(sizeof(double) * L);
    int *wperm = malloc(sizeof(int) * L);
    int rc = 0;
    for (long t = first; t < first + count; t++) {
        int i = ix[t] - 1;
        for (int c = 0; c < ncol; c++) row[c] = counts[(size_t)c * nG + i];           /* as.numeric(counts[i, ]) :623 */
        o_cG_CalcGibbsProbY(i + 1, row, ncol, nTSByC, nByTS, nGByTS, nByG, y, L, lgbeta, lggamma, lgdelta, beta, gamma,
                            delta, pr);
        if (probs) memcpy(probs + (size_t)i * L, pr, sizeof(double) * L);
        if (doSample) {
            int prev = y[i] - 1;
            int lab = o_sample_ll(pr, L, u[t], wp, wperm, margin ? &margin[t] : NULL);
            if (lab < 0) { snprintf(g_err, sizeof g_err, "sampler failure %d at gene %d", lab, i + 1); rc = 1; break; }
            int ny = lab - 1;
            if (ny != prev) {                                                         /* :641-649 */
                for (int c = 0; c < ncol; c++) {
                    nTSByC[(size_t)c * L + prev] -= row[c];
                    nTSByC[(size_t)c * L + ny] += row[c];
                }
                nGByTS[prev] -= 1; nByTS[prev] -= nByG[i];
                nGByTS[ny] += 1; nByTS[ny] += nByG[i];
                y[i] = ny + 1;
            }
        }
    }
    free(row); free(pr); free(wp); free(wperm);
    return rc;
}

/* ------------------------------------------------------------------ log-likelihoods */

static double ll_theta(const int *m, int K, int nS, double alpha)
{
    /* R/celda_CG.R:854-860 == R/celda_C.R:771-777 */
    double a = nS * o_lgamma(K * alpha);
    acc_t b = 0;
    for (int i = 0; i < K * nS; i++) b += o_lgamma(m[i] + alpha);
    double c = -nS * K * o_lgamma(alpha);
    acc_t d = 0;
    for (int s = 0; s < nS; s++) {
        acc_t cs = 0;
        for (int k = 0; k < K; k++) cs += m[(size_t)s * K + k] + alpha;
        d += o_lgamma((double)cs);
    }
    return a + (double)b + c + -(double)d;
}

static double ll_phi(const double *n, int R, long ncols, double beta)
{
    /* R/celda_CG.R:862-868 (nTSByCP, L x K), R/celda_C.R:779-785 (nGByCP, G x K), R/celda_G.R:677-683 (nTSByC) */
    double a = ncols * o_lgamma(R * beta);
    acc_t b = 0;
    for (size_t i = 0; i < (size_t)R * ncols; i++) b += o_lgamma(n[i] + beta);
    double c = -(double)ncols * R * o_lgamma(beta);
    acc_t d = 0;
    for (long k = 0; k < ncols; k++) {
        acc_t cs = 0;
        for (int r = 0; r < R; r++) cs += n[(size_t)k * R + r] + beta;
        d += o_lgamma((double)cs);
    }
    return a + (double)b + c + -(double)d;
}

static double ll_psi(const double *nByG, int nGenes, const double *nByTS, const int *nGByTS, int L, double delta)
{
    /* R/celda_CG.R:852,870-876 == R/celda_G.R:674,685-691 ; nG <- sum(nGByTS) (G + L, pseudo-genes included) */
    acc_t nGs = 0;
    for (int l = 0; l < L; l++) nGs += nGByTS[l];
    double nG = (double)nGs;
    acc_t a = 0, b = 0, d = 0;
    for (int l = 0; l < L; l++) a += o_lgamma(nGByTS[l] * delta);
    for (int g = 0; g < nGenes; g++) b += o_lgamma(nByG[g] + delta);
    double c = -nG * o_lgamma(delta);
    for (int l = 0; l < L; l++) d += o_lgamma(nByTS[l] + (nGByTS[l] * delta));
    return (double)a + (double)b + c + -(double)d;
}

static double ll_eta(const int *nGByTS, int L, double gamma)
{
    /* R/celda_CG.R:878-884 == R/celda_G.R:693-699 (sum(lgamma(sum(..))) there is a sum of one value) */
    double a = o_lgamma(L * gamma);
    acc_t b = 0, es = 0;
    for (int l = 0; l < L; l++) b += o_lgamma(nGByTS[l] + gamma);
    double c = -L * o_lgamma(gamma);
    for (int l = 0; l < L; l++) es += nGByTS[l] + gamma;
    double d = -o_lgamma((double)es);
    return a + (double)b + c + d;
}

/* .cCGCalcLL (R/celda_CG.R:839-888): final <- thetaLl + phiLl + psiLl + etaLl. */
double o_cCGCalcLL(int K, int L, const int *mCPByS, const double *nTSByCP, const double *nByG, int nGenes,
                   const double *nByTS, const int *nGByTS, int nS, double alpha, double beta, double delta,
                   double gamma)
{
    double th = ll_theta(mCPByS, K, nS, alpha), ph = ll_phi(nTSByCP, L, K, beta);
    double ps = ll_psi(nByG, nGenes, nByTS, nGByTS, L, delta), et = ll_eta(nGByTS, L, gamma);
    return th + ph + ps + et;
}

/* .cCCalcLL (R/celda_C.R:760-788): theta + phi over nGByCP (nG x K). */
double o_cCCalcLL(const int *mCPByS, const double *nGByCP, int K, int nS, int nG, double alpha, double beta)
{
    return ll_theta(mCPByS, K, nS, alpha) + ll_phi(nGByCP, nG, K, beta);
}

/* .cGCalcLL (R/celda_G.R:664-702): final <- phiLl + psiLl + etaLl, phi over nTSByC (L x nM). */
double o_cGCalcLL(const double *nTSByC, const double *nByTS, const double *nByG, int nGenes, const int *nGByTS,
                  long nM, int L, double beta, double delta, double gamma)
{
    double ph = ll_phi(nTSByC, L, nM, beta);
    double ps = ll_psi(nByG, nGenes, nByTS, nGByTS, L, delta), et = ll_eta(nGByTS, L, gamma);
    return ph + ps + et;
}

/* ------------------------------------------------------------------ perplexity */

static void colnorm(const double *x, int nr, int nc, double add, double *out)
{
    /* normalizeCounts(x + add, "proportion") : (x + add) / colSums(x + add)  (R/factorizeMatrix.R:284-296) */
    for (int j = 0; j < nc; j++) {
        acc_t cs = 0;
        for (int i = 0; i < nr; i++) cs += x[(size_t)j * nr + i] + add;
        for (int i = 0; i < nr; i++) out[(size_t)j * nr + i] = (x[(size_t)j * nr + i] + add) / (double)cs;
    }
}

static double logsumexp(const double *v, int n)
{
    /* matrixStats::logSumExp: max + log(sum(exp(v - max))) with the max term counted as exactly 1 (log1p form is
     * not used by matrixStats; it computes log(sum) with sum accumulated in long double). */
    int im = 0;
    for (int i = 1; i < n; i++) if (v[i] > v[im]) im = i;
    acc_t s = 0;
    for (int i = 0; i < n; i++) if (i != im) s += o_exp(v[i] - v[im]);
    return v[im] + log1p((double)s);
}

/* .perplexityCelda_C (R/perplexity.R:233-259) / .perplexityCelda_CG (:262-293) shared tail.
 * M: nr x K column-major log-probabilities; theta: log(sample posterior) K x nS. */
static double perplexity_tail(int nr, long nc, const int *p, const int *ri, const double *x, const double *M, int K,
                              const double *theta, const int *s)
{
    double *col = malloc(sizeof(double) * K);
    acc_t logPx = 0, tot = 0;
    for (long c = 0; c < nc; c++) {
        for (int k = 0; k < K; k++) {
            double acc = 0;
            for (int t = p[c]; t < p[c + 1]; t++) acc += M[(size_t)k * nr + ri[t]] * x[t];
            col[k] = acc + theta[(size_t)(s[c] - 1) * K + k];
        }
        logPx += logsumexp(col, K);
        for (int t = p[c]; t < p[c + 1]; t++) tot += x[t];
    }
    free(col);
    return o_exp(-((double)logPx / (double)tot));
}

double o_perplexity_C(int nr, long nc, const int *p, const int *ri, const double *x, const int *mCPByS, int nS,
                      const double *nGByCP, int K, const int *s, double alpha, double beta)
{
    double *md = malloc(sizeof(double) * K * nS), *th = malloc(sizeof(double) * K * nS);
    double *M = malloc(sizeof(double) * (size_t)nr * K);
    for (int i = 0; i < K * nS; i++) md[i] = mCPByS[i];
    colnorm(md, K, nS, alpha, th);
    for (int i = 0; i < K * nS; i++) th[i] = o_log(th[i]);
    colnorm(nGByCP, nr, K, beta, M);
    for (size_t i = 0; i < (size_t)nr * K; i++) M[i] = o_log(M[i]);
    double r = perplexity_tail(nr, nc, p, ri, x, M, K, th, s);
    free(md); free(th); free(M);
    return r;
}

/* psi (module posterior) per gene: (nByG[g] + delta) / sum over genes of the same module (R/factorizeMatrix.R:281-283) */
static void psi_per_gene(const double *nByG, const int *y, int nG, int L, double delta, double *psi)
{
    acc_t *den = calloc(L, sizeof(acc_t));
    for (int g = 0; g < nG; g++) den[y[g] - 1] += nByG[g] + delta;
    for (int g = 0; g < nG; g++) psi[g] = (nByG[g] + delta) / (double)den[y[g] - 1];
    free(den);
}

double o_perplexity_CG(int nr, long nc, const int *p, const int *ri, const double *x, const int *mCPByS, int nS,
                       const double *nTSByCP, const double *nByG, const int *y, int K, int L, const int *s,
                       double alpha, double beta, double delta)
{
    double *md = malloc(sizeof(double) * K * nS), *th = malloc(sizeof(double) * K * nS);
    double *phi = malloc(sizeof(double) * (size_t)L * K), *psi = malloc(sizeof(double) * nr);
    double *M = malloc(sizeof(double) * (size_t)nr * K);
    for (int i = 0; i < K * nS; i++) md[i] = mCPByS[i];
    colnorm(md, K, nS, alpha, th);
    for (int i = 0; i < K * nS; i++) th[i] = o_log(th[i]);
    colnorm(nTSByCP, L, K, beta, phi);
    psi_per_gene(nByG, y, nr, L, delta, psi);
    /* log(psi %*% phi): psi has one non-zero per row, so the product is a single multiply (R/perplexity.R:285) */
    for (int k = 0; k < K; k++)
        for (int g = 0; g < nr; g++) M[(size_t)k * nr + g] = o_log(psi[g] * phi[(size_t)k * L + y[g] - 1]);
    double r = perplexity_tail(nr, nc, p, ri, x, M, K, th, s);
    free(md); free(th); free(phi); free(psi); free(M);
    return r;
}

/* .perplexityCelda_G (R/perplexity.R:296-325): exp(-sum_gc x log(psi[g] phi_cell[y_g, c]) / sum x). */
double o_perplexity_G(int nr, long nc, const int *p, const int *ri, const double *x, const double *nTSByC,
                      const double *nByG, const int *y, int L, double beta, double delta)
{
    double *psi = malloc(sizeof(double) * nr), *phic = malloc(sizeof(double) * L);
    psi_per_gene(nByG, y, nr, L, delta, psi);
    acc_t logPx = 0, tot = 0;
    for (long c = 0; c < nc; c++) {
        colnorm(nTSByC + (size_t)c * L, L, 1, beta, phic);
        for (int t = p[c]; t < p[c + 1]; t++) {
            int g = ri[t];
            logPx += o_log(psi[g] * phic[y[g] - 1]) * x[t];
            tot += x[t];
        }
    }
    free(psi); free(phic);
    return o_exp(-((double)logPx / (double)tot));
}

/* _perplexityG (src/perplexity.c:5-62): dense int x (nr x nc), phi (nl x nc), psi (nr x nl), factor group. */
int o_perplexityG(const int *x, int nr, int nc, const double *phi, int phi_nr, int phi_nc, const double *psi,
                  int psi_nr, int psi_nc, const int *group, long ngroup, int nl, double *ans)
{
    if (nl < 0) FAIL("The grouping argument must be a factor");
    if (ngroup != nr) FAIL("The length of the grouping argument must match the number of rows in the matrix.");
    if (phi_nc != nc) FAIL("The R_phi and R_x must have the same number of colums.");
    if (phi_nr != nl) FAIL("R_phi must have the same number of rows as the number of levels in R_group.");
    if (psi_nr != nr) FAIL("The R_psi and R_x must have the same number of rows.");
    if (psi_nc != nl) FAIL("R_phi must have the same number of columns as the number of levels in R_group.");
    for (int i = 0; i < nr; i++)
        if (group[i] == NA_INT || group[i] < 0 || group[i] > nr)
            FAIL("Labels in group and pgroup must not be NA and must less than or equal to the number of rows in the matrix.");
    double a = 0;
    for (int j = 0; j < nc; j++)
        for (int i = 0; i < nr; i++)
            a += x[(size_t)j * nr + i] * o_log(phi[(size_t)j * nl + (group[i] - 1)] * psi[(size_t)nr * (group[i] - 1) + i]);
    *ans = a;
    return 0;
}
