// libm-parity helpers of the device code.
//
// The reference's generated C evaluates I_Dsqrt / I_DDsqrt as `.5*pow(x,-.5)` / `-.25*pow(x,-1.5)`
// (lib/csparse/instructionsTableUTHash.c:1414-1426) with the HOST libm: glibc's pow errs by < 0.52 ULP, i.e. it
// returns the correctly rounded power on all but ~8e-4 of arguments (measured: tools/bench/pow_parity.c).
// CUDA's pow() is specified to 2 ULP and differs from it on ~1 argument in 4.  tc_pow_m05 / tc_pow_m15 below
// return the CORRECTLY ROUNDED x^-1/2 and x^-3/2 (double-double Newton correction of 1/sqrt(x), exact
// residuals through FMA), which is what glibc returns except on those ~8e-4 -- the closest a device
// function can get to the oracle's rounding sequence without carrying glibc's tables.
#pragma once
#include <cfloat>
#include <cmath>

// x^-1/2 = hi + lo to ~2^-104 relative, for x in the range where no intermediate under/overflows
__device__ __forceinline__ void tc_rsqrt_dd(double x, double &hi, double &lo) {
    const double r = 1.0 / sqrt(x);                            // IEEE sqrt and division: |r sqrt(x) - 1| < 2^-51
    const double t = __dmul_rn(r, r), tl = __fma_rn(r, r, -t); // r*r = t + tl exactly
    const double s = __dmul_rn(x, t);
    const double sl = __dadd_rn(__fma_rn(x, t, -s), __dmul_rn(x, tl));   // x*r*r = s + sl (to 2^-106)
    const double e = __dadd_rn(__dadd_rn(1.0, -s), -sl);       // 1 - x r^2: (1-s) is exact (Sterbenz)
    hi = r;
    lo = __dmul_rn(__dmul_rn(0.5, r), e);                      // r (1 + e/2) = x^-1/2 (1 + O(e^2))
}

__device__ __forceinline__ double tc_pow_m05(double x) {       // pow(x, -.5)
    if (!(x > 1e-290 && x < 1e290)) return pow(x, -.5);        // zeros, negatives, NaN, inf, extreme exponents
    double hi, lo;
    tc_rsqrt_dd(x, hi, lo);
    return __dadd_rn(hi, lo);
}

__device__ __forceinline__ double tc_pow_m15(double x) {       // pow(x, -1.5) = x^-1/2 / x
    if (!(x > 1e-190 && x < 1e190)) return pow(x, -1.5);
    double hi, lo;
    tc_rsqrt_dd(x, hi, lo);
    const double q0 = hi / x;
    const double rem = __fma_rn(-q0, x, hi);                   // hi - q0*x exactly
    return __dadd_rn(q0, __dadd_rn(rem, lo) / x);
}
