// fp64 elementary functions for the assembly kernels: log, exp, reciprocal and square root of POSITIVE NORMAL arguments,
// accurate to about one unit in the last place, branch-free, with their polynomial coefficients in the constant bank.
//
// Why not the CUDA math library: its log()/exp()/pow() are exact-rounding-grade general-purpose routines; inlined into the
// assembly kernel they spend about as many issue slots on materialising fp64 immediates (UMOV/IMAD.MOV pairs) and on
// special-case branches (BSSY/BSYNC/BRA) as on fp64 arithmetic (profiles/r2a: 263 fp64 of 704 instructions per row).  Every
// argument on this path is a positive, finite, normal number well inside the exponent range (heights, radii, wind speeds,
// temperatures in Kelvin), so none of that is needed.  The parity bar on assembled coefficients is 1e-12 relative; these
// functions are held to <= 4e-16 against libm on the argument ranges of the path by tests/test_fastmath.py (host build of this
// same header) and by pbsm3d_debug_math on the device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define PBSM3D_HD __host__ __device__ __forceinline__
#else
#define PBSM3D_HD inline
#endif

namespace pbsm3d {

PBSM3D_HD double bits_to_double(long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double d;
    std::memcpy(&d, &b, sizeof(d));
    return d;
#endif
}
PBSM3D_HD long long double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(d);
#else
    long long b;
    std::memcpy(&b, &d, sizeof(b));
    return b;
#endif
}

// 1/d: hardware seed (MUFU.RCP64H, ~2^-20) + two Newton steps.  |d| normal and 1/d normal.
PBSM3D_HD double frcp(double d) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / d;
#endif
}

// sqrt(x), x > 0 normal: hardware seed of 1/sqrt(x) (MUFU.RSQ64H) + two coupled Newton steps + one residual correction.
PBSM3D_HD double fsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
#else
    return std::sqrt(x);
#endif
}

// Polynomial coefficients live in ONE constant-bank array so that the device code reads them as uniform-register operands
// (LDCU.128: two doubles per instruction) instead of building every 64-bit immediate from two 32-bit moves.
#define PBSM3D_MATH_TABLE                                                                                                         \
    {6.93147180369123816490e-01 /* ln2_hi */, 1.90821492927058770002e-10 /* ln2_lo */,                                           \
     6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,                     \
     1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01 /* Lg1..Lg7 */, 1.4426950408889634074,          \
     1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880, 1.0 / 3628800, 1.0 / 39916800,      \
     1.0 / 479001600, 1.0 / 6227020800.0 /* 1/n!, n = 2..13 */}
static const double kMathHost[22] = PBSM3D_MATH_TABLE;
#if defined(__CUDACC__)
__constant__ double kMathDev[22] = PBSM3D_MATH_TABLE;
#endif
#if defined(__CUDA_ARCH__)
#define PBSM3D_KM(i) kMathDev[i]
#else
#define PBSM3D_KM(i) kMathHost[i]
#endif

// ln(x), x > 0 normal.  Argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh(s), s = (m-1)/(m+1), with the
// classic 7-term minimax polynomial in s^2 (Remez coefficients as tabulated in FreeBSD msun e_log.c) and the hi/lo split of
// ln 2.
PBSM3D_HD double flog(double x) {
    const double ln2_hi = PBSM3D_KM(0), ln2_lo = PBSM3D_KM(1);
    const double Lg1 = PBSM3D_KM(2), Lg2 = PBSM3D_KM(3), Lg3 = PBSM3D_KM(4), Lg4 = PBSM3D_KM(5), Lg5 = PBSM3D_KM(6), Lg6 = PBSM3D_KM(7),
                 Lg7 = PBSM3D_KM(8);
    const long long ix = double_to_bits(x);
    int hi = (int)(ix >> 32);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int adj = (hi + 0x95f64) & 0x100000;  // mantissa >= sqrt(2): halve it, bump the exponent
    k += adj >> 20;
    hi |= adj ^ 0x3ff00000;
    const double m = bits_to_double(((long long)hi << 32) | (ix & 0xffffffffLL));
    const double f = m - 1.0;
    const double s = f * frcp(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

// exp(x), |x| < 700.  x = k ln2 + r, |r| <= ln2/2; e^r by its Taylor polynomial of degree 13 (truncation < 4e-18 on the reduced
// range) evaluated in Estrin form (five dependent steps instead of thirteen); scaling by 2^k through the exponent field.
PBSM3D_HD double fexp(double x) {
    const double ln2_hi = PBSM3D_KM(0), ln2_lo = PBSM3D_KM(1);
    const double kf = rint(x * PBSM3D_KM(9));
    double r = fma(-kf, ln2_hi, x);
    r = fma(-kf, ln2_lo, r);
    const double c2 = PBSM3D_KM(10), c3 = PBSM3D_KM(11), c4 = PBSM3D_KM(12), c5 = PBSM3D_KM(13), c6 = PBSM3D_KM(14), c7 = PBSM3D_KM(15),
                 c8 = PBSM3D_KM(16), c9 = PBSM3D_KM(17), c10 = PBSM3D_KM(18), c11 = PBSM3D_KM(19), c12 = PBSM3D_KM(20),
                 c13 = PBSM3D_KM(21);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double a0 = 1.0 + r, a1 = fma(c3, r, c2), a2 = fma(c5, r, c4), a3 = fma(c7, r, c6), a4 = fma(c9, r, c8),
                 a5 = fma(c11, r, c10), a6 = fma(c13, r, c12);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
    const double p = fma(d1, r8, d0);
    return bits_to_double(double_to_bits(p) + ((long long)(int)kf << 52));
}

// cbrt(x) for x in [1, 1.5]: quadratic seed (|error| < 2e-4) + two Halley steps (cubic convergence).
PBSM3D_HD double fcbrt_1_15(double x) {
    double y = 0.59528904 + x * (0.48254113 - 0.07759479 * x);  // least-squares fit on [1, 1.5]
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const double y3 = y * y * y;
        y = y * (y3 + 2.0 * x) * frcp(2.0 * y3 + x);
    }
    return y;
}

// x^y, x > 0: exp(y ln x).  Relative error ~ |y ln x| ulp: <= 2e-15 for every use on this path (|y ln x| < 20).
PBSM3D_HD double fpow(double x, double y) { return fexp(y * flog(x)); }

}  // namespace pbsm3d
