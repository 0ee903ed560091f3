#!/usr/bin/env python
"""Freeze the FP64 flop counts behind every roofline numerator into openkite_b200/csrc/kite_flops.h.

Two independent counts per unit of work (SURVEY.md 8d asks for an op-counting derivation, frozen in a header):

  ORACLE  what the reference's algorithm needs.  RHS and RK4 step: oracle::Counted on the literal restatement
          (oracle/kite_oracle.hpp).  f + Jx + Ju: joint sympy CSE of the symbolic RHS and its symbolic Jacobians
          (oracle/sympy_oracle.py; the survey's method).  Composites (RK4 + sensitivities, EKF predict, collocation)
          follow from those by the formulas below.  FMA = 2, div / sqrt / transcendental = 1 (internals not counted).
  DEVICE  what the hand-written device algorithm executes: openkite_b200/csrc/kite_model.cuh + kite_math.cuh are compiled
          for the host with `double` replaced by an 8-byte counting scalar (add / mul = 1, fma = 2, MUFU seed = 1,
          compares / selects / abs / copysign = 0), so the count includes the internals of the lean special functions
          (polynomial asin / atan2 / exp, Newton steps) and excludes whatever the analytic-sparse formulation never
          computes.  Tangent products and tableau updates of the kernels are added by formula from the sparsity
          predicates of kite_kernels.cuh.

bench.py divides min(ORACLE, DEVICE) x units by the measured kernel time: a fraction of the FMA peak that neither
credits flops the kernel never executes (DEVICE < ORACLE: the sensitivity kernel) nor the internals of special
functions (DEVICE > ORACLE: the plain rollout).

    python scripts/make_flops.py            # rewrites openkite_b200/csrc/kite_flops.h
    python scripts/make_flops.py --check    # exits 1 if the committed header is stale
"""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "openkite_b200", "csrc")
OUT = os.path.join(CSRC, "kite_flops.h")

NNZ_JX, NNZ_JU, NNZ_ARM = 104, 7, 21       # structural non-zeros; checked against the device predicates below

COUNTING_SCALAR = r"""
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>
struct Tally { long add = 0, mul = 0, fma = 0, div = 0, mufu = 0, other = 0; };
inline Tally& tally() { static Tally t; return t; }
struct real_t {                      // exactly 8 bytes: the bit tricks of kite_math.cuh (memcpy to int64) keep working
    double v;
    real_t() : v(0) {}
    real_t(double a) : v(a) {}
    real_t(float a) : v(a) {}
    real_t(int a) : v(a) {}
    real_t(unsigned long a) : v((double)a) {}
    real_t(long a) : v((double)a) {}
    explicit operator float() const { return (float)v; }
    explicit operator double() const { return v; }
};
inline real_t operator+(real_t a, real_t b) { tally().add++; return real_t(a.v + b.v); }
inline real_t operator-(real_t a, real_t b) { tally().add++; return real_t(a.v - b.v); }
inline real_t operator*(real_t a, real_t b) { tally().mul++; return real_t(a.v * b.v); }
inline real_t operator/(real_t a, real_t b) { tally().div++; return real_t(a.v / b.v); }
inline real_t operator-(real_t a) { return real_t(-a.v); }                      // sign flips fold into the consumer
inline real_t& operator+=(real_t& a, real_t b) { a = a + b; return a; }
inline real_t& operator-=(real_t& a, real_t b) { a = a - b; return a; }
inline real_t& operator*=(real_t& a, real_t b) { a = a * b; return a; }
#define MIXED(op) inline real_t operator op(double a, real_t b) { return real_t(a) op b; } \
                  inline real_t operator op(real_t a, double b) { return a op real_t(b); }
MIXED(+) MIXED(-) MIXED(*) MIXED(/)
inline bool operator>(real_t a, real_t b) { return a.v > b.v; }
inline bool operator<(real_t a, real_t b) { return a.v < b.v; }
inline bool operator>(real_t a, double b) { return a.v > b; }
inline bool operator<(real_t a, double b) { return a.v < b; }
inline bool operator==(real_t a, double b) { return a.v == b; }
inline real_t fma(real_t a, real_t b, real_t c) { tally().fma++; return real_t(std::fma(a.v, b.v, c.v)); }
inline real_t fma(double a, real_t b, real_t c) { return fma(real_t(a), b, c); }
inline real_t fma(real_t a, double b, real_t c) { return fma(a, real_t(b), c); }
inline real_t fma(real_t a, real_t b, double c) { return fma(a, b, real_t(c)); }
inline real_t fma(real_t a, double b, double c) { return fma(a, real_t(b), real_t(c)); }
inline real_t fma(double a, real_t b, double c) { return fma(real_t(a), b, real_t(c)); }
inline real_t fma(double a, double b, real_t c) { return fma(real_t(a), real_t(b), c); }
inline real_t fabs(real_t a) { return real_t(std::fabs(a.v)); }
inline real_t copysign(real_t a, real_t b) { return real_t(std::copysign(a.v, b.v)); }
inline real_t copysign(double a, real_t b) { return real_t(std::copysign(a, b.v)); }
inline real_t fmin(real_t a, real_t b) { tally().other++; return real_t(std::fmin(a.v, b.v)); }
inline real_t fmax(real_t a, real_t b) { tally().other++; return real_t(std::fmax(a.v, b.v)); }
inline real_t fmin(real_t a, double b) { return fmin(a, real_t(b)); }
inline real_t fmax(real_t a, double b) { return fmax(a, real_t(b)); }
inline real_t sqrt(real_t a) { tally().div++; return real_t(std::sqrt(a.v)); }
inline float sqrtf_(float a) { return std::sqrt(a); }
"""

COUNT_MAIN = r"""
#include "counting_scalar.h"
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define KITE_COUNTING_BUILD 1
inline real_t __dadd_rn(real_t a, real_t b) { return a + b; }
inline real_t __dmul_rn(real_t a, real_t b) { return a * b; }
#include "kite_model_counting.cuh"
#include <cstdio>
using namespace kite;
struct NullSink { int nx = 0, nu = 0; void jx(int, int, real_t) { ++nx; } void ju(int, int, real_t) { ++nu; } void aero(real_t, real_t, real_t) {} };
static void report(const char* name, int nx = 0, int nu = 0) {
    Tally& t = tally();
    std::printf("%s add=%ld mul=%ld fma=%ld div=%ld mufu=%ld other=%ld nx=%d nu=%d\n", name, t.add, t.mul, t.fma, t.div, t.mufu, t.other, nx, nu);
    t = Tally();
}
int main() {
    KiteConsts K{};                       // values are irrelevant for the count; avoid divisions by zero only
    real_t* kp = reinterpret_cast<real_t*>(&K);
    for (size_t i = 0; i < sizeof(K) / sizeof(real_t) - 1; ++i) kp[i] = real_t(0.37 + 0.01 * i);
    K.has_arm = 0; K.model_kind = 0;
    real_t x[13], u[3], f[13];
    const double x0[13] = {6.19, -0.028, 0.918, 0.297, -2.2, -0.148, -0.416, -2.26, 1.29, 0.0356, -0.07, 0.8266, 0.557};
    for (int i = 0; i < 13; ++i) x[i] = real_t(x0[i]);
    u[0] = real_t(0.1); u[1] = real_t(0.01); u[2] = real_t(-0.02);
    tally() = Tally();
    { NoSink s; kite_eval<false>(K, K.A, x, u, f, s); report("rhs"); }
    { NullSink s; kite_eval<true>(K, K.A, x, u, f, s); report("rhs_jac", s.nx, s.nu); }
    K.has_arm = 1;
    { NullSink s; kite_eval<true>(K, K.A, x, u, f, s); report("rhs_jac_arm", s.nx, s.nu); }
    K.has_arm = 0;
    { real_t xx[13]; for (int i = 0; i < 13; ++i) xx[i] = x[i]; RkTab rk = make_rk_tab(real_t(1e-3)); tally() = Tally(); rk4_step<false>(K, K.A, xx, u, rk); report("rk4_step"); }
    { NoSink s; rigid_eval<false>(K, x, f, s); report("rigid_rhs"); }
    { NullSink s; rigid_eval<true>(K, x, f, s); report("rigid_rhs_jac", s.nx, s.nu); }
    return 0;
}
"""


def device_counts():
    """Compile the device model with the counting scalar and run it."""
    tmp = tempfile.mkdtemp(prefix="kite_flops_")
    for name in ("kite_model.cuh", "kite_math.cuh"):
        src = open(os.path.join(CSRC, name)).read()
        src = re.sub(r"\bdouble\b", "real_t", src)
        src = src.replace('#include <cuda_runtime.h>', '').replace('#include "kite_math.cuh"', '#include "kite_math_counting.cuh"')
        # MUFU seeds: one special-function-unit instruction each
        src = src.replace("return (real_t)(1.0f / (float)a);", "tally().mufu++; return (real_t)(1.0f / (float)a);")
        src = src.replace("return (real_t)(1.0f / sqrtf((float)a));", "tally().mufu++; return (real_t)(1.0f / sqrtf_((float)a));")
        open(os.path.join(tmp, name.replace(".cuh", "_counting.cuh")), "w").write(src)
    open(os.path.join(tmp, "counting_scalar.h"), "w").write(COUNTING_SCALAR)
    open(os.path.join(tmp, "main.cpp"), "w").write(COUNT_MAIN)
    exe = os.path.join(tmp, "count")
    subprocess.check_call(["g++", "-O0", "-std=c++17", "-I", tmp, "-o", exe, os.path.join(tmp, "main.cpp")])
    out = subprocess.check_output([exe], text=True)
    res = {}
    for ln in out.strip().splitlines():
        name, *kv = ln.split()
        d = {k: int(v) for k, v in (t.split("=") for t in kv)}
        d["flops"] = d["add"] + d["mul"] + 2 * d["fma"] + d["div"] + d["mufu"]
        d["fp64_ops"] = d["add"] + d["mul"] + d["fma"] + d["div"] + d["other"]      # lower bound of FP64-pipe instructions
        res[name] = d
    return res


def oracle_counts():
    from oracle.oracle_py import Oracle, params_from_yaml
    orc = Oracle(params_from_yaml(os.path.join(ROOT, "data", "umx_radian.yaml")))
    return orc.flop_counts()


def sympy_counts():
    """Joint CSE of f, Jx, Ju of the symbolic reference RHS (the survey's method)."""
    import sympy as sp
    import yaml
    from oracle import sympy_oracle as so
    with open(os.path.join(ROOT, "data", "umx_radian.yaml")) as fh:
        cfg = yaml.safe_load(fh)
    cfg.setdefault("tether", {})
    for k in ("rx", "ry", "rz"):
        cfg["tether"].setdefault(k, 0.0)
    x, u, p, f = so.build_rhs(cfg, "kite")
    fm = sp.Matrix(f)
    Jx = fm.jacobian(sp.Matrix(x)); Ju = fm.jacobian(sp.Matrix(u))

    def count(exprs):
        repl, red = sp.cse(exprs, optimizations="basic")
        tot = dict(add=0, mul=0, div=0, special=0)
        for e in [r for _, r in repl] + [t for m in red for t in (list(m) if hasattr(m, "__iter__") else [m])]:
            for node in sp.preorder_traversal(e):
                if isinstance(node, sp.Add):
                    tot["add"] += len(node.args) - 1
                elif isinstance(node, sp.Mul):
                    args = [a for a in node.args if a != -1]
                    tot["mul"] += max(len(args) - 1, 0)
                elif isinstance(node, sp.Pow):
                    e_ = node.exp
                    if e_ == -1:
                        tot["div"] += 1
                    elif e_ == sp.Rational(1, 2):
                        tot["special"] += 1
                    elif e_ == -sp.Rational(1, 2):
                        tot["special"] += 1; tot["div"] += 1
                    elif e_.is_Integer and e_ > 0:
                        tot["mul"] += int(e_) - 1
                    elif e_.is_Integer:
                        tot["mul"] += -int(e_) - 1; tot["div"] += 1
                    elif e_.is_Rational:           # x^(k/2)
                        tot["special"] += 1; tot["mul"] += abs(int(e_ * 2)) // 2; tot["div"] += 1 if e_ < 0 else 0
                elif isinstance(node, sp.Function):
                    tot["special"] += 1
        tot["flops"] = sum(tot.values())
        return tot
    nnz_x = sum(1 for e in Jx if e != 0); nnz_u = sum(1 for e in Ju if e != 0)
    return {"rhs": count([fm]), "rhs_jac": count([fm, Jx, Ju]), "nnz_x": nnz_x, "nnz_u": nnz_u}


def compose(orc, sym, dev):
    """Composite units.  T = 16 tangent columns ([Phi | Gamma]); FMA = 2."""
    T, n = 16, 13
    nnz = NNZ_JX + NNZ_JU
    o, d = {}, {}
    # ---- ORACLE (algorithm of the reference, sparse symbolic Jacobians as SX::jacobian yields them)
    o["RHS"] = orc["rhs"]["flops"]
    o["RK4_STEP"] = orc["rk4_step"]["flops"]
    o["RHS_JAC"] = sym["rhs_jac"]["flops"]
    jac_only = sym["rhs_jac"]["flops"] - sym["rhs"]["flops"]                  # Jacobians given the shared RHS intermediates
    # RK4 + sensitivities: primal step + 4 stage Jacobians + S_i = [Jx_i | Ju_i] D_i for stages 2..4 (stage 1 has the
    # identity seed: a copy, no flops; the survey's 27.8 k counted it as a fourth product) + tableau on the 13 x 16 block
    tableau = 3 * 2 * n * T + 3 * 2 * n * T + 2 * n * T                       # D_i = E + a h S, A += w S, [Phi|Gamma] = E + h/6 A
    # (S_i = [Jx_i | Ju_i] + a h Jx_i S_(i-1): the control Jacobian enters by addition, 7 adds per stage)
    o["RK4_SENS_STEP"] = o["RK4_STEP"] + 4 * jac_only + 3 * (2 * NNZ_JX * T + NNZ_JU) + tableau
    o["RK4_SENS_STEP_SURVEY"] = 27800
    # EKF predict (kiteEKF.cpp:75-98): RK4 step + Jx at the pre-step state + A = I + J dt + two DENSE 13^3 products + W
    o["EKF_PREDICT"] = o["RK4_STEP"] + sym["rhs_jac"]["flops"] + 2 * NNZ_JX + 2 * (2 * n ** 3) + n * n
    # collocation scenario (chebyshev.hpp:241-271, kiteNMPF.cpp:100-107): 11 x (f + J) + scaling + (CompD (x) I) X + tau F
    M = 11
    o["COLLOC_SCENARIO"] = M * (sym["rhs_jac"]["flops"] + 15 + 2 * (NNZ_JX + 1) + 2 * (NNZ_JU + 1) + 19) + 165 * 6 * 2 + 165 * 2
    # ---- DEVICE (what the kernels execute)
    d["RHS"] = dev["rhs"]["flops"]
    d["RK4_STEP"] = dev["rk4_step"]["flops"]
    d["RHS_JAC"] = dev["rhs_jac"]["flops"]
    # k_sens_fused: primal pass (RK4 step) + 4 x (f + J) (the Jacobian pass recomputes f's intermediates) + phase B
    # (phase B: 104 x 16 FMAs per stage for Jx D, and 7 x 8 for the Ju columns -- one FMA per lane of a unit)
    d["RK4_SENS_STEP"] = d["RK4_STEP"] + 4 * d["RHS_JAC"] + 3 * (2 * NNZ_JX * T + 2 * NNZ_JU * 8) + tableau
    # k_ekf_predict_tma: RK4 step whose first stage is the (f + J) evaluation (k1 and the Jacobian share the pre-step point)
    # + two SPARSE products  Q = P + dt (P J^T),  Pn = Q + dt (J Q) + W
    d["EKF_PREDICT"] = d["RK4_STEP"] - d["RHS"] + d["RHS_JAC"] + 2 * (2 * NNZ_JX * n + 2 * n * n) + n * n
    d["COLLOC_SCENARIO"] = M * (d["RHS_JAC"] + 19 + 2 * (NNZ_JX + 1) + 2 * (NNZ_JU + 1) + 15) + 165 * 6 * 2 + 165 * 2
    return o, d


def render(orc, sym, dev):
    o, d = compose(orc, sym, dev)
    L = []
    L.append("// kite_flops.h -- GENERATED by scripts/make_flops.py; do not edit.  Frozen FP64 flop counts per unit of work, the")
    L.append("// numerators of every roofline fraction bench.py reports (DESIGN.md section 5).  FMA = 2 flops.")
    L.append("//   KITE_FLOPS_ORACLE_*: the reference's algorithm (oracle::Counted on the literal restatement for RHS / RK4 step,")
    L.append("//                        joint sympy CSE of f, df/dx, df/du for the Jacobians; special functions count 1).")
    L.append("//   KITE_FLOPS_DEVICE_*: executed by the device source (kite_model.cuh + kite_math.cuh compiled with a counting")
    L.append("//                        scalar: internals of the lean special functions included), products by formula.")
    L.append("#pragma once")
    L.append("")
    L.append("// oracle::Counted, RHS: add %d mul %d div %d special %d" % tuple(orc["rhs"][k] for k in ("add", "mul", "div", "special")))
    L.append("// oracle::Counted, RK4 step: add %d mul %d div %d special %d" % tuple(orc["rk4_step"][k] for k in ("add", "mul", "div", "special")))
    L.append("// sympy CSE, f: add %d mul %d div %d special %d = %d (survey: 435)" % (*[sym["rhs"][k] for k in ("add", "mul", "div", "special")], sym["rhs"]["flops"]))
    L.append("// sympy CSE, f + Jx + Ju: add %d mul %d div %d special %d = %d (survey: 2690); nnz %d + %d" %
             (*[sym["rhs_jac"][k] for k in ("add", "mul", "div", "special")], sym["rhs_jac"]["flops"], sym["nnz_x"], sym["nnz_u"]))
    for k in ("rhs", "rhs_jac", "rhs_jac_arm", "rk4_step", "rigid_rhs", "rigid_rhs_jac"):
        v = dev[k]
        L.append("// device source, %s: add %d mul %d fma %d div %d mufu %d (min/max %d) -> %d flops, >= %d FP64-pipe instructions%s" %
                 (k, v["add"], v["mul"], v["fma"], v["div"], v["mufu"], v["other"], v["flops"], v["fp64_ops"],
                  (", Jacobian entries emitted %d + %d" % (v["nx"], v["nu"])) if v["nx"] else ""))
    L.append("")
    for k in ("RHS", "RK4_STEP", "RHS_JAC", "RK4_SENS_STEP", "RK4_SENS_STEP_SURVEY", "EKF_PREDICT", "COLLOC_SCENARIO"):
        L.append("#define KITE_FLOPS_ORACLE_%s %d" % (k, o[k]))
    for k in ("RHS", "RK4_STEP", "RHS_JAC", "RK4_SENS_STEP", "EKF_PREDICT", "COLLOC_SCENARIO"):
        L.append("#define KITE_FLOPS_DEVICE_%s %d" % (k, d[k]))
    L.append("#define KITE_NNZ_JX %d" % NNZ_JX)
    L.append("#define KITE_NNZ_JU %d" % NNZ_JU)
    L.append("#define KITE_NNZ_JX_ARM_EXTRA %d" % NNZ_ARM)
    L.append("")
    return "\n".join(L)


def read_header(path=OUT):
    """{'ORACLE_RK4_STEP': 1888, ...} from the committed header (used by bench.py and the tests)."""
    out = {}
    for m in re.finditer(r"#define KITE_(FLOPS_\w+|NNZ_\w+) (\d+)", open(path).read()):
        out[m.group(1).replace("FLOPS_", "")] = int(m.group(2))
    return out


def main():
    dev = device_counts()
    assert dev["rhs_jac"]["nx"] == NNZ_JX and dev["rhs_jac"]["nu"] == NNZ_JU, dev["rhs_jac"]
    assert dev["rhs_jac_arm"]["nx"] == NNZ_JX + NNZ_ARM
    orc = oracle_counts()
    sym = sympy_counts()
    assert sym["nnz_x"] == NNZ_JX and sym["nnz_u"] == NNZ_JU, (sym["nnz_x"], sym["nnz_u"])
    txt = render(orc, sym, dev)
    if "--check" in sys.argv:
        ok = os.path.exists(OUT) and open(OUT).read() == txt
        print("kite_flops.h is %s" % ("up to date" if ok else "STALE"))
        sys.exit(0 if ok else 1)
    open(OUT, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
