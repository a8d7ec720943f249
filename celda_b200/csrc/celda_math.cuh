// Device arithmetic for the celda hot path (sm_100a).
//
// Every probability on the sweep path is built from log, exp and lgamma of "count + prior" arguments.  The
// reference takes these from base R / libm (R nmath lgammafn, platform log/exp: R/celda_C.R:664-689,
// R/celda_functions.R:1-11, src/cG_calcGibbsProbY.cpp:238-245).  To make draws reproducible bit for bit against a
// CPU restatement, the device evaluates them with IEEE basic operations only, in a fixed order, with FMA
// contraction disabled for this translation unit (nvcc -fmad=false):
//   dlog   : argument reduction x = 2^k (1+f), s = f/(2+f), even polynomial (the classic fdlibm scheme)
//   dexp   : x = k ln2 + r, rational form in r with a degree-5 polynomial (fdlibm scheme)
//   dlgamma: R's algorithm -- Chebyshev series for gamma on (0,1) + recursion for x <= 10, Stirling + Chebyshev
//            correction above.
#pragma once
#include <cstdint>

namespace celda {

__device__ __forceinline__ double cheb_eval(double x, const double* a, int n) {
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    const double twox = x * 2.0;
#pragma unroll
    for (int i = 1; i <= n; ++i) {
        b2 = b1;
        b1 = b0;
        b0 = twox * b1 - b2 + a[n - i];
    }
    return (b0 - b2) * 0.5;
}

__device__ __forceinline__ double dlog(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    unsigned lx = (unsigned)__double2loint(x);
    int k = 0;
    if (hx < 0x00100000) {  // zero, negative or subnormal
        if (((hx & 0x7fffffff) | lx) == 0) return -__longlong_as_double(0x7ff0000000000000LL);
        if (hx < 0) return __longlong_as_double(0x7ff8000000000000LL);
        k -= 54;
        x *= 1.80143985094819840000e+16;
        hx = __double2hiint(x);
        lx = (unsigned)__double2loint(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), (int)lx);
    k += (i >> 20);
    const double f = x - 1.0;
    const double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {  // |f| < 2^-20
        if (f == 0.0) return (k == 0) ? 0.0 : dk * ln2_hi + dk * ln2_lo;
        const double R = f * f * (0.5 - 0.33333333333333333 * f);
        return (k == 0) ? f - R : dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    i = (hx - 0x6147a) | (0x6b851 - hx);
    if (i > 0) {
        const double hfsq = 0.5 * f * f;
        return (k == 0) ? f - (hfsq - s * (hfsq + R)) : dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    return (k == 0) ? f - s * (f - R) : dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

__device__ __forceinline__ double dexp(double x) {
    const double ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00, twom1000 = 9.33263618503218878990e-302;
    const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    unsigned hx = (unsigned)__double2hiint(x);
    const int xsb = (hx >> 31) & 1;
    hx &= 0x7fffffff;
    if (hx >= 0x40862E42) {
        if (hx >= 0x7ff00000) {
            if (((hx & 0xfffff) | (unsigned)__double2loint(x)) != 0) return x + x;
            return xsb ? 0.0 : x;
        }
        if (x > 7.09782712893383973096e+02) return __longlong_as_double(0x7ff0000000000000LL);
        if (x < -7.45133219101941108420e+02) return 0.0;
    }
    double hi = 0.0, lo = 0.0;
    int k = 0;
    if (hx > 0x3fd62e42) {
        if (hx < 0x3FF0A2B2) {
            hi = xsb ? x + ln2HI : x - ln2HI;
            lo = xsb ? -ln2LO : ln2LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int)(invln2 * x + (xsb ? -0.5 : 0.5));
            const double t = (double)k;
            hi = x - t * ln2HI;
            lo = t * ln2LO;
        }
        x = hi - lo;
    } else if (hx < 0x3e300000) {
        return 1.0 + x;
    }
    const double t = x * x;
    const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    const double y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return __longlong_as_double(__double_as_longlong(y) + ((long long)k << 52));
    return __longlong_as_double(__double_as_longlong(y) + ((long long)(k + 1000) << 52)) * twom1000;
}

__device__ __constant__ double kGamCs[22] = {
    +.8571195590989331421920062399942e-2,  +.4415381324841006757191315771652e-2,
    +.5685043681599363378632664588789e-1,  -.4219835396418560501012500186624e-2,
    +.1326808181212460220584006796352e-2,  -.1893024529798880432523947023886e-3,
    +.3606925327441245256578082217225e-4,  -.6056761904460864218485548290365e-5,
    +.1055829546302283344731823509093e-5,  -.1811967365542384048291855891166e-6,
    +.3117724964715322277790254593169e-7,  -.5354219639019687140874081024347e-8,
    +.9193275519859588946887786825940e-9,  -.1577941280288339761767423273953e-9,
    +.2707980622934954543266540433089e-10, -.4646818653825730144081661058933e-11,
    +.7973350192007419656460767175359e-12, -.1368078209830916025799499172309e-12,
    +.2347319486563800657233471771688e-13, -.4027432614949066932766570534699e-14,
    +.6910051747372100912138336975257e-15, -.1185584500221992907052387126192e-15};

// gamma(x) for 0 < x <= 10 (cold path: only counts below ten land here).
__device__ __noinline__ double dgamma_small(double x) {
    int n = (int)x;
    const double y = x - (double)n;
    --n;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    const double twox = (y * 2.0 - 1.0) * 2.0;
    for (int i = 1; i <= 22; ++i) {
        b2 = b1;
        b1 = b0;
        b0 = twox * b1 - b2 + kGamCs[22 - i];
    }
    double value = (b0 - b2) * 0.5 + .9375;
    if (n == 0) return value;
    if (n < 0) return value / x;
    for (int i = 1; i <= n; ++i) value *= (y + (double)i);
    return value;
}

// Stirling correction for x >= 10.
__device__ __forceinline__ double dlgammacor(double x) {
    const double algmcs[5] = {+.1666389480451863247205729650822e+0, -.1384948176067563840732986059135e-4,
                              +.9810825646924729426157171547487e-8, -.1809129475572494194263306266719e-10,
                              +.6221098041892605227126015543416e-13};
    if (x < 94906265.62425156) {
        const double t = 10.0 / x;
        return cheb_eval(t * t * 2.0 - 1.0, algmcs, 5) / x;
    }
    return 1.0 / (x * 12.0);
}

// lgamma for x > 10 (the hot branch).
__device__ __forceinline__ double dlgamma_large(double x) {
    const double ln_sqrt_2pi = 0.918938533204672741780329736406;
    if (x > 1e17) return x * (dlog(x) - 1.0);
    if (x > 4934720.0) return ln_sqrt_2pi + (x - 0.5) * dlog(x) - x;
    return ln_sqrt_2pi + (x - 0.5) * dlog(x) - x + dlgammacor(x);
}

// Generic lgamma kept out of line: the sweeps call it only off the hot range.
__device__ __noinline__ double dlgamma_generic(double x) {
    if (!(x > 0.0))
        return (x == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff8000000000000LL);
    if (x <= 10.0) {
        if (x < 1e-306) return -dlog(x);
        return dlog(fabs(dgamma_small(x)));
    }
    return dlgamma_large(x);
}

// Hot-range lgamma: 10 < x <= 1e17 (a count plus a prior), straight-line code.  Performs exactly the
// operations of dlgamma_large()/dlog() for such x (x is normal, exponent k >= 3; the Stirling correction is
// evaluated always and dropped above lgammafn's 4934720 cut-off, so table totals in the millions stay converged);
// the two data-dependent forms of the log tail are both evaluated and selected, which costs a few flops but keeps
// the warp converged.  Arguments within 2^-20 of a power of two (e.g. 63 + 1) take the
// generic routine.
__device__ __constant__ double kHot[20] = {
    6.93147180369123816490e-01, 1.90821492927058770002e-10,                                   // 0 ln2_hi, 1 ln2_lo
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,             // 2.. Lg1..Lg7
    2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01,
    +.6221098041892605227126015543416e-13, -.1809129475572494194263306266719e-10,             // 9.. algmcs[4..0]
    +.9810825646924729426157171547487e-8, -.1384948176067563840732986059135e-4, +.1666389480451863247205729650822e+0,
    0.918938533204672741780329736406, 0.0, 0.0, 0.0, 0.0, 0.0};

__device__ __forceinline__ double dlgamma_hot(double x) {
    if (!(x > 10.0 && x <= 1e17)) return dlgamma_generic(x);
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    if ((0x000fffff & (2 + hx)) < 3) return dlgamma_generic(x);
    const int i = (hx + 0x95f64) & 0x100000;
    const double xm = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    const double f = xm - 1.0;
    const double dk = (double)k;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (kHot[3] + w * (kHot[5] + w * kHot[7]));
    const double t2 = z * (kHot[2] + w * (kHot[4] + w * (kHot[6] + w * kHot[8])));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dkhi = dk * kHot[0], dklo = dk * kHot[1];
    const double la = dkhi - ((hfsq - (s * (hfsq + R) + dklo)) - f);
    const double lb = dkhi - ((s * (f - R) - dklo) - f);
    const double lg = (((hx - 0x6147a) | (0x6b851 - hx)) > 0) ? la : lb;
    // Stirling correction, x < 94906265.6
    const double t = 10.0 / x;
    const double tx = (t * t * 2.0 - 1.0) * 2.0;
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    b2 = b1; b1 = b0; b0 = tx * b1 - b2 + kHot[9];
    b2 = b1; b1 = b0; b0 = tx * b1 - b2 + kHot[10];
    b2 = b1; b1 = b0; b0 = tx * b1 - b2 + kHot[11];
    b2 = b1; b1 = b0; b0 = tx * b1 - b2 + kHot[12];
    b2 = b1; b1 = b0; b0 = tx * b1 - b2 + kHot[13];
    const double cor = ((b0 - b2) * 0.5) / x;
    const double base = kHot[14] + (x - 0.5) * lg - x;
    return (x > 4934720.0) ? base : base + cor;  // above the cut-off lgammafn drops the correction term
}

__device__ __forceinline__ double dlgamma(double x) {
    if (!(x > 0.0))
        return (x == 0.0) ? __longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff8000000000000LL);
    if (x <= 10.0) {
        if (x < 1e-306) return -dlog(x);
        return dlog(fabs(dgamma_small(x)));
    }
    return dlgamma_large(x);
}

// lg_delta[k] of the reference: c(NA, lgamma(seq(nG + L) * delta)) (R/celda_CG.R:429, R/celda_G.R:332): index 0 is NA,
// so a module count of zero makes the probability vector NA ("NA in probability vector" from sample.int).
__device__ __forceinline__ double dlgdelta(int k, double delta) {
    if (k == 0) return __longlong_as_double(0x7ff8000000000000LL);
    return dlgamma_hot((double)k * delta);
}

}  // namespace celda
