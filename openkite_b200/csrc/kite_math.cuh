// =====================================================================================
// kite_math.cuh -- lean FP64 special functions for the kite kernels (sm_100a).
//
// More than half of the FP64 instructions of one RHS evaluation were CUDA libm special functions
// (profiles/r1a_*: 1210 FP64 instr per RK4 step, ~160 per RHS in sqrt/div/asin/atan2/exp incl. their
// special-case paths).  These replacements keep ~1 ulp accuracy on the value ranges the model can
// produce, drop the denormal/NaN slow paths (non-finite trajectories are flagged by the kernels
// instead) and share work between related quantities:
//   rcp / rsqrt : MUFU seed (rcp.approx / rsqrt.approx .ftz.f64) + ONE cubically convergent step
//   asin        : odd minimax polynomial on |x| <= 0.7072 (scripts/fit_math_polys.py), complement identity
//                 asin(s) = sign(s) (pi/2 - asin(cos)) beyond, so the whole range is branch free
//   logistic    : exp by Cody-Waite reduction + degree-11 polynomial, then rcp
// Host builds (tests/cpu_shim) emulate the MUFU seeds with float precision so the same source is
// checked on the CPU.
// =====================================================================================
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

namespace kite {

__device__ __forceinline__ double rcp_seed(double a) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
#else
    return (double)(1.0f / (float)a);
#endif
}
__device__ __forceinline__ double rsqrt_seed(double a) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
#else
    return (double)(1.0f / sqrtf((float)a));
#endif
}

// High word of a double as a signed integer.  Sign and magnitude tests on it run on the integer pipe (ISETP) and leave the
// FP64 pipe to the arithmetic: for finite a, hi(a) > 0 <=> a >= 2^-1022 (a positive normal), hi(a) < 0 <=> sign bit set.
__device__ __forceinline__ int hi_word(double a) {
#ifdef __CUDA_ARCH__
    return __double2hiint(a);
#else
    int64_t b; memcpy(&b, &a, sizeof b); return (int)(b >> 32);
#endif
}
#ifndef KITE_INT_CMP
#define KITE_INT_CMP 1      // measured (profiles/r2b_sweep.log): 82.51 -> 81.91 ms per config-2 pass, 7 DSETP per RHS off the FP64 pipe
#endif
__device__ __forceinline__ bool is_pos(double a) {            // a > 0 (denormals count as 0 in the integer form)
#if KITE_INT_CMP
    return hi_word(a) > 0;
#else
    return a > 0.0;
#endif
}
__device__ __forceinline__ bool is_neg(double a) {            // a < 0
#if KITE_INT_CMP
    return hi_word(a) < 0;
#else
    return a < 0.0;
#endif
}

// 1/a: seed error e0 <= 2^-19  ->  y0 (1 + e + e^2), e = 1 - a y0, error e0^3 <= 2^-57.   MUFU + 3 DFMA.
__device__ __forceinline__ double fast_rcp(double a) {
    const double y0 = rcp_seed(a);
    const double e = fma(-a, y0, 1.0);
    const double e2 = fma(e, e, e);
    return fma(y0, e2, y0);
}
// 1/sqrt(a): y0 (1 + e/2 + 3 e^2/8), e = 1 - a y0^2, error ~ (5/16) e0^3.   MUFU + 5 FP64.
__device__ __forceinline__ double fast_rsqrt(double a) {
    const double y0 = rsqrt_seed(a);
    const double t = a * y0;
    const double e = fma(-t, y0, 1.0);
    const double p = fma(0.375, e, 0.5) * e;
    return fma(y0, p, y0);
}

// NOTE: the coefficients are written as literals on purpose.  FP64 instructions on sm_100a take constants only through
// uniform registers; a coefficient TABLE (__constant__ or constexpr array) is hoisted out of the stage loop into
// ~40 uniform registers, which overflow into vector registers and from there into local memory (228 B of spills
// in k_rk4_rollout).  Literals are re-materialised next to their use (2 UMOV each) and cost no registers.
#ifndef KITE_POLY_SPLIT
#define KITE_POLY_SPLIT 0
#endif

// asin(x) = x + x u P(u), u = x^2, |x| <= 0.7072; degree 16, max relative error 4.4e-16 (scripts/fit_math_polys.py)
#define KITE_ASIN_POLY_MAX 0.7072
__device__ __forceinline__ double asin_poly(double x) {
    const double u = x * x;
#if KITE_POLY_SPLIT
    // even/odd split in u: two independent Horner chains of half the depth
    const double u2 = u * u;
    double pe = 5.27314318387392955e-01, po = -1.67855479930664897e+00;
    pe = fma(pe, u2, 2.55375757459231778e+00);
    po = fma(po, u2, -2.35809246022577668e+00);
    pe = fma(pe, u2, 1.48923001382088138e+00);
    po = fma(po, u2, -6.56023493249017098e-01);
    pe = fma(pe, u2, 2.23097241478172448e-01);
    po = fma(po, u2, -4.34132617562938208e-02);
    pe = fma(pe, u2, 1.89243203271123074e-02);
    po = fma(po, u2, 1.03696475274575421e-02);
    pe = fma(pe, u2, 1.40738595019858897e-02);
    po = fma(po, u2, 1.73458138976341353e-02);
    pe = fma(pe, u2, 2.23724499523806075e-02);
    po = fma(po, u2, 3.03819370896512217e-02);
    pe = fma(pe, u2, 4.46428572403844703e-02);
    po = fma(po, u2, 7.49999999994898220e-02);
    pe = fma(pe, u2, 1.66666666666667102e-01);
    const double p = fma(po, u, pe);
#else
    double p = 5.27314318387392955e-01;
    p = fma(p, u, -1.67855479930664897e+00);
    p = fma(p, u, 2.55375757459231778e+00);
    p = fma(p, u, -2.35809246022577668e+00);
    p = fma(p, u, 1.48923001382088138e+00);
    p = fma(p, u, -6.56023493249017098e-01);
    p = fma(p, u, 2.23097241478172448e-01);
    p = fma(p, u, -4.34132617562938208e-02);
    p = fma(p, u, 1.89243203271123074e-02);
    p = fma(p, u, 1.03696475274575421e-02);
    p = fma(p, u, 1.40738595019858897e-02);
    p = fma(p, u, 1.73458138976341353e-02);
    p = fma(p, u, 2.23724499523806075e-02);
    p = fma(p, u, 3.03819370896512217e-02);
    p = fma(p, u, 4.46428572403844703e-02);
    p = fma(p, u, 7.49999999994898220e-02);
    p = fma(p, u, 1.66666666666667102e-01);
#endif
    return fma(x * u, p, x);
}
// Angle in [-pi/2, pi/2] from its sine s and cosine c >= 0 (s^2 + c^2 = 1), branch free over the whole range:
//   |s| <= 1/sqrt2 : asin(s)            |s| > 1/sqrt2 : sign(s) (pi/2 - asin(c)),  c < 1/sqrt2
// so the polynomial argument never leaves |x| <= 0.7072 and a warp never diverges into libm (random-control
// rollouts sit at |sideslip| > 37 deg for ~40% of the horizon: profiles/r1f sweep).
__device__ __forceinline__ double asin_sc(double s, double c) {
#if KITE_INT_CMP
    // |s| > 0.70710678 on the high words (0x3FE6A09E = hi(1/sqrt 2); the polynomial is valid up to 0.7072 on either side)
    const bool big = (hi_word(s) & 0x7fffffff) > 0x3FE6A09E;
#else
    const bool big = fabs(s) > 0.70710678118654752;
#endif
    const double r = asin_poly(big ? c : s);
    const double t = (1.5707963267948966 - r) + 6.123233995736766e-17;
    return big ? copysign(t, s) : r;
}
// atan2(y, x) from the normalised pair s = y/hypot, c = x/hypot, any quadrant, branch free:
//   c >= 0 : asin_sc(s, c)              c < 0 : sign(s) pi - asin_sc(s, -c)
__device__ __forceinline__ double atan2_sc(double s, double c) {
    const double r = asin_sc(s, fabs(c));
    const double t = (copysign(3.141592653589793, s) - r) + copysign(1.2246467991473532e-16, s);
    return is_neg(c) ? t : r;
}

// logistic(x) = 1 / (1 + exp(-x)); argument clamped to +-700 (result 0 / 1 to within 1e-304 beyond).
__device__ __forceinline__ double fast_logistic(double x) {
    double a = -x;
    a = fmin(fmax(a, -700.0), 700.0);
    // n = rint(a * log2(e)) by the 1.5 * 2^52 trick; r = a - n ln2 (Cody-Waite, hi part has 32 trailing zero bits)
    const double magic = 6755399441055744.0;
    const double tn = fma(a, 1.44269504088896339e+00, magic);
    const double nf = tn - magic;
    double r = fma(nf, -6.93146705627441406e-01, a);
    r = fma(nf, -4.74932503903167256e-07, r);
    // exp(r), |r| <= ln2/2, degree 11 (scripts/fit_math_polys.py)
#if KITE_POLY_SPLIT
    const double r2 = r * r;
    double po = 2.51100492048186583e-08, pe = 2.76326547225277896e-07;
    po = fma(po, r2, 2.75572408872298695e-06);
    pe = fma(pe, r2, 2.48014854415613131e-05);
    po = fma(po, r2, 1.98412698900764028e-04);
    pe = fma(pe, r2, 1.38888889523528631e-03);
    po = fma(po, r2, 8.33333333331958900e-03);
    pe = fma(pe, r2, 4.16666666664879531e-02);
    po = fma(po, r2, 1.66666666666666796e-01);
    pe = fma(pe, r2, 5.00000000000001887e-01);
    po = fma(po, r2, 1.00000000000000000e+00);
    pe = fma(pe, r2, 1.00000000000000000e+00);
    const double p = fma(po, r, pe);
#else
    double p = 2.51100492048186583e-08;
    p = fma(p, r, 2.76326547225277896e-07);
    p = fma(p, r, 2.75572408872298695e-06);
    p = fma(p, r, 2.48014854415613131e-05);
    p = fma(p, r, 1.98412698900764028e-04);
    p = fma(p, r, 1.38888889523528631e-03);
    p = fma(p, r, 8.33333333331958900e-03);
    p = fma(p, r, 4.16666666664879531e-02);
    p = fma(p, r, 1.66666666666666796e-01);
    p = fma(p, r, 5.00000000000001887e-01);
    p = fma(p, r, 1.00000000000000000e+00);
    p = fma(p, r, 1.00000000000000000e+00);
#endif
    // scale by 2^n: n is in the low word of tn (|n| <= 1010, so the biased exponent stays normal)
    int64_t bits;
    memcpy(&bits, &tn, sizeof bits);
    const int64_t n = (int64_t)(int32_t)(bits & 0xFFFFFFFF);
    const int64_t sb = (n + 1023) << 52;
    double scale;
    memcpy(&scale, &sb, sizeof scale);
    const double ex = p * scale;                  // exp(-x)
    return fast_rcp(1.0 + ex);
}

}  // namespace kite
