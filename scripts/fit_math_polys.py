#!/usr/bin/env python
"""Derive the polynomial coefficients used by openkite_b200/csrc/kite_math.cuh (asin core, exp core).
Chebyshev interpolation in 60-digit arithmetic (near-minimax), converted to the monomial basis, then the float64
Horner evaluation is checked against mpmath over the interval.  Prints C arrays."""
import mpmath as mp
import numpy as np

mp.mp.dps = 60


def cheb_fit(f, a, b, deg):
    n = deg + 1
    xs = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / n) for k in range(n)]
    ys = [f((b - a) / 2 * x + (a + b) / 2) for x in xs]
    c = [2 / mp.mpf(n) * sum(ys[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / n) for k in range(n)) for j in range(n)]
    c[0] /= 2
    # Chebyshev -> monomial in t, then t = (2u - (a+b))/(b-a)
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        t = [mp.mpf(0)] + [2 * v for v in T[j - 1]]
        for i, v in enumerate(T[j - 2]):
            t[i] -= v
        T.append(t)
    mono_t = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono_t[i] += c[j] * v
    # substitute t = alpha*u + beta
    alpha, beta = 2 / (b - a), -(a + b) / (b - a)
    out = [mp.mpf(0)] * n
    for i, ci in enumerate(mono_t):
        # (alpha u + beta)^i
        for k in range(i + 1):
            out[k] += ci * mp.binomial(i, k) * alpha ** k * beta ** (i - k)
    return out


def asin_core(u):
    # asin(x) = x + x*u*P(u), u = x^2  ->  P(u) = (asin(sqrt(u))/sqrt(u) - 1)/u
    if u == 0:
        return mp.mpf(1) / 6
    s = mp.sqrt(u)
    return (mp.asin(s) / s - 1) / u


def check_asin(coef, xmax):
    c = [float(v) for v in coef]
    worst = 0.0
    for x in np.linspace(-xmax, xmax, 20001):
        u = x * x
        p = c[-1]
        for v in c[-2::-1]:
            p = p * u + v
        y = x + x * u * p
        ref = mp.asin(mp.mpf(float(x)))
        err = abs((mp.mpf(y) - ref)) / max(abs(ref), mp.mpf(1e-300)) if x != 0 else 0
        worst = max(worst, float(err))
    return worst


if __name__ == "__main__":
    for xmax, deg in ((0.5, 10), (0.5, 11), (0.5, 12), (0.6, 12), (0.6, 13), (0.6, 14), (0.65, 14), (0.7072, 16), (0.7072, 18)):
        coef = cheb_fit(asin_core, mp.mpf(0), mp.mpf(xmax) ** 2, deg)
        print("asin |x|<=%.4f deg %d: max rel err %.3e" % (xmax, deg, check_asin(coef, xmax)))


def check_exp(coef, rmax):
    c = [float(v) for v in coef]
    worst = 0.0
    for r in np.linspace(-rmax, rmax, 20001):
        p = c[-1]
        for v in c[-2::-1]:
            p = p * r + v
        ref = mp.exp(mp.mpf(float(r)))
        worst = max(worst, float(abs(mp.mpf(p) - ref) / ref))
    return worst


def emit(name, coef):
    print("// %s" % name)
    print("{" + ", ".join("%.17e" % float(v) for v in coef) + "}")


if __name__ == "__main__":
    print()
    coef = cheb_fit(asin_core, mp.mpf(0), mp.mpf("0.6") ** 2, 13)
    emit("ASIN_P[14]: asin(x) = x + x*u*P(u), u = x^2, |x| <= 0.6, max rel err %.2e" % check_asin(coef, 0.6), coef)
    rmax = float(mp.log(2) / 2) * 1.0001
    for deg in (10, 11, 12):
        coef = cheb_fit(mp.exp, mp.mpf(-rmax), mp.mpf(rmax), deg)
        print("exp deg %d: max rel err %.3e" % (deg, check_exp(coef, rmax)))
    coef = cheb_fit(mp.exp, mp.mpf(-rmax), mp.mpf(rmax), 11)
    emit("EXP_P[12]: exp(r), |r| <= ln2/2", coef)
    ln2 = mp.log(2)
    hi = float(ln2)
    hi = float(np.float64(hi).view(np.uint64) & np.uint64(0xFFFFFFFFF8000000)) if False else hi
    # split ln2 = hi + lo with hi having 32 trailing zero bits so n*hi is exact for |n| < 2^20
    bits = np.float64(float(ln2)).view(np.uint64) & np.uint64(0xFFFFFFFF00000000)
    hi = float(bits.view(np.float64))
    lo = float(ln2 - mp.mpf(hi))
    print("LN2_HI = %.17e, LN2_LO = %.17e, LOG2E = %.17e" % (hi, lo, float(1 / ln2)))
