/* =====================================================================================
 * kite_casadi.h -- libkite_casadi.so: the engine's single-point functions in CasADi's external-function C convention
 * (SURVEY.md 8f-3).  An unmodified CasADi host loads them by NAME instead of building the SX graphs:
 *
 *     casadi::Function dyn = casadi::external("dynamics",     "libkite_casadi.so");   // replaces kite.cpp:324
 *     casadi::Function jac = casadi::external("dyn_jacobian", "libkite_casadi.so");   // replaces kite.cpp:327-328
 *     casadi::Function aer = casadi::external("Aero",         "libkite_casadi.so");   // replaces kite.cpp:330
 *     casadi::Function rk4 = casadi::external("RK4",          "libkite_casadi.so");   // replaces kite.cpp:332-338
 *     ODESolver solver(dyn, opts);  KiteEKF ekf(rk4, jac);                            // integrator.cpp:7, kiteEKF.cpp:40
 *
 * The model file comes from $KITE_B200_YAML (device from $KITE_B200_DEVICE, default 0) or kite_external_init().
 * kite_casadi_int is `int` for CasADi 3.0 - 3.4 (the reference pins v3.0.0-rc2, README.md:6); build the shim with
 * -DKITE_CASADI_INT64 for CasADi >= 3.5.  Not verifiable against CasADi in this image (not installed): tests/cpp/
 * casadi_external_test.c drives the symbols through dlopen / dlsym exactly as casadi::external does.
 * ===================================================================================== */
#ifndef KITE_CASADI_H
#define KITE_CASADI_H

#ifdef KITE_CASADI_INT64
typedef long long kite_casadi_int;
#else
typedef int kite_casadi_int;
#endif

#ifdef __cplusplus
extern "C" {
#endif

int kite_external_init(const char* yaml_path, int device);   /* optional: bind model file and CUDA device; 0 = ok */
void kite_external_shutdown(void);

/* For NAME in { dynamics, dyn_jacobian, Aero, RK4, dynamics_id, dyn_jacobian_id }:
 *   inputs  dynamics / dyn_jacobian / Aero: (x[13], u[3]);  RK4: (X[13], U[3], dT[1]);  *_id: (x[13], u[3], p[21])
 *   output  dynamics, RK4, dynamics_id: dense 13;  Aero: dense 3;  dyn_jacobian[_id]: sparse 13 x 13, 104 non-zeros in
 *           compressed-column order (125 when the model has a tether arm)
 *   sparsity patterns: {nrow, ncol, colind[ncol + 1], row[nnz]} */
#define KITE_CASADI_DECLARE(NAME)                                                                                      \
    int NAME(const double** arg, double** res, kite_casadi_int* iw, double* w, int mem);                               \
    kite_casadi_int NAME##_n_in(void);                                                                                 \
    kite_casadi_int NAME##_n_out(void);                                                                                \
    const kite_casadi_int* NAME##_sparsity_in(kite_casadi_int i);                                                      \
    const kite_casadi_int* NAME##_sparsity_out(kite_casadi_int i);                                                     \
    int NAME##_work(kite_casadi_int* sz_arg, kite_casadi_int* sz_res, kite_casadi_int* sz_iw, kite_casadi_int* sz_w);  \
    const char* NAME##_name_in(kite_casadi_int i);                                                                     \
    const char* NAME##_name_out(kite_casadi_int i);                                                                    \
    void NAME##_incref(void);                                                                                          \
    void NAME##_decref(void);

KITE_CASADI_DECLARE(dynamics)          /* kite.cpp:324 */
KITE_CASADI_DECLARE(dyn_jacobian)      /* kite.cpp:327-328 */
KITE_CASADI_DECLARE(Aero)              /* kite.cpp:330 */
KITE_CASADI_DECLARE(RK4)               /* kite.cpp:332-338 */
KITE_CASADI_DECLARE(dynamics_id)       /* kite.cpp:575 */
KITE_CASADI_DECLARE(dyn_jacobian_id)   /* kite.cpp:578-579 */

#ifdef __cplusplus
}
#endif
#endif /* KITE_CASADI_H */
