// Compiles the DEVICE model source (kite_model.cuh) for the host and exposes it to ctypes.
// Test infrastructure only (see cuda_runtime.h stub in this directory).
#include "cuda_runtime.h"
#include "../../openkite_b200/csrc/kite_model.cuh"
#include "../../openkite_b200/csrc/kite_consts.h"
#include <cstring>

using namespace kite;

struct DenseSink {
    double* Jx; double* Ju;
    void jx(int i, int j, double v) { Jx[i * 13 + j] = v; }
    void ju(int i, int j, double v) { Ju[i * 3 + j] = v; }
    void aero(double, double, double) {}
};

extern "C" {
void shim_eval(const kite_params* prm, int kind, const double* x, const double* u, const double* p, double* f,
               double* Jx, double* Ju) {
    KiteConsts K = make_consts(*prm, kind);
    AeroCoef A = K.A;
    if (p) derive_coef(K, p, A);
    double xx[13], uu[3], ff[13];
    std::memcpy(xx, x, sizeof xx); std::memcpy(uu, u, sizeof uu);
    std::memset(Jx, 0, 169 * 8); std::memset(Ju, 0, 39 * 8);
    DenseSink s{Jx, Ju};
    if (kind == 2) model_eval<true, true>(K, A, xx, uu, ff, s);
    else model_eval<false, true>(K, A, xx, uu, ff, s);
    std::memcpy(f, ff, sizeof ff);
}
// the same through the table forms of the special functions (volatile coefficients = the identification-sweep kernels)
void shim_eval_tab(const kite_params* prm, int kind, const double* x, const double* u, const double* p, double* f,
                   double* Jx, double* Ju) {
    KiteConsts K = make_consts(*prm, kind);
    AeroCoef A = K.A;
    if (p) derive_coef(K, p, A);
    const volatile AeroCoef& Av = A;
    double xx[13], uu[3], ff[13];
    std::memcpy(xx, x, sizeof xx); std::memcpy(uu, u, sizeof uu);
    std::memset(Jx, 0, 169 * 8); std::memset(Ju, 0, 39 * 8);
    DenseSink s{Jx, Ju};
    kite_eval<true>(K, Av, xx, uu, ff, s);
    std::memcpy(f, ff, sizeof ff);
}
void shim_rk4(const kite_params* prm, int kind, const double* x, const double* u, const double* p, double h, long n,
              double* xn) {
    KiteConsts K = make_consts(*prm, kind);
    AeroCoef A = K.A;
    if (p) derive_coef(K, p, A);
    double xx[13], uu[3];
    std::memcpy(xx, x, sizeof xx); std::memcpy(uu, u, sizeof uu);
    for (long k = 0; k < n; ++k) { if (kind == 2) rk4_step<true>(K, A, xx, uu, make_rk_tab(h)); else rk4_step<false>(K, A, xx, uu, make_rk_tab(h)); }
    std::memcpy(xn, xx, sizeof xx);
}
}
