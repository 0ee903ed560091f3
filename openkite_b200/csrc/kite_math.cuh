// =====================================================================================
// kite_math.cuh -- lean FP64 special functions for the kite kernels (sm_100a).
//
// More than half of the FP64 instructions of one RHS evaluation were CUDA libm special functions
// (profiles/r1a_*: 1210 FP64 instr per RK4 step, ~160 per RHS in sqrt/div/asin/atan2/exp incl. their
// special-case paths).  These replacements keep ~1 ulp accuracy on the value ranges the model can
// produce, drop the denormal/NaN slow paths (non-finite trajectories are flagged by the kernels
// instead) and share work between related quantities:
//   rcp / rsqrt : MUFU seed (rcp.approx / rsqrt.approx .ftz.f64) + ONE cubically convergent step
//   asin        : odd minimax polynomial on |x| <= 0.6 (scripts/fit_math_polys.py), libm fallback outside
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

// asin(x) = x + x u P(u), u = x^2, |x| <= ASIN_FAST_MAX; max relative error 2.3e-16 (scripts/fit_math_polys.py)
#define KITE_ASIN_FAST_MAX 0.6
__device__ __forceinline__ double asin_poly(double x) {
    const double u = x * x;
    double p = 7.29997078374671204e-02;
    p = fma(p, u, -1.04137057612019940e-01);
    p = fma(p, u, 9.35985258000532616e-02);
    p = fma(p, u, -3.48182818154024673e-02);
    p = fma(p, u, 2.18831168546623420e-02);
    p = fma(p, u, 6.79360522162927773e-03);
    p = fma(p, u, 1.20052897697550849e-02);
    p = fma(p, u, 1.39170345012287547e-02);
    p = fma(p, u, 1.73561602888368215e-02);
    p = fma(p, u, 2.23720037341655145e-02);
    p = fma(p, u, 3.03819486815959904e-02);
    p = fma(p, u, 4.46428570828007673e-02);
    p = fma(p, u, 7.50000000003340495e-02);
    p = fma(p, u, 1.66666666666666352e-01);
    return fma(x * u, p, x);
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
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
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
