// =====================================================================================
// kite_casadi.cpp -- libkite_casadi.so: the engine behind CasADi's external-function C convention, so that an
// unmodified CasADi host can load the hot path by name,
//       casadi::Function f = casadi::external("dynamics", "libkite_casadi.so");
// in place of the Function objects KiteDynamics builds (reference kite.cpp:324 "dynamics", :328 "dyn_jacobian",
// :330 "Aero", :338 "RK4"; identification variant :575-579 as "dynamics_id" / "dyn_jacobian_id").  SURVEY.md 8f-3.
//
// Per function NAME the library exports what casadi::external resolves:
//   int         NAME(const double** arg, double** res, kite_casadi_int* iw, double* w, int mem);   0 = ok
//   kite_casadi_int NAME_n_in(void), NAME_n_out(void);
//   const kite_casadi_int* NAME_sparsity_in(kite_casadi_int i), NAME_sparsity_out(kite_casadi_int i);
//               compressed-column pattern {nrow, ncol, colind[ncol + 1], row[nnz]}
//   int         NAME_work(kite_casadi_int* sz_arg, kite_casadi_int* sz_res, kite_casadi_int* sz_iw, kite_casadi_int* sz_w);
//   const char* NAME_name_in(kite_casadi_int i), NAME_name_out(kite_casadi_int i);
//   void        NAME_incref(void), NAME_decref(void);
// kite_casadi_int is `int` (CasADi 3.0 - 3.4, which includes the reference's pinned v3.0.0-rc2); build with
// -DKITE_CASADI_INT64 for CasADi >= 3.5 (`long long`).  A NULL arg[i] reads as zeros, a NULL res[i] is skipped, outputs are
// written in the order of the output pattern's non-zeros (column-major), as CasADi expects.
//
// The model comes from the YAML file named by $KITE_B200_YAML (kite_external_init(path, device) sets it explicitly); every
// call evaluates ONE point on the GPU through the B = 1 entry points of libkite_b200.so.  Plain host C++: the header-only
// mirror (include/openkite) does the YAML loading and the staging.
// =====================================================================================
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/openkite/kite.hpp"
#include "kite_sparsity.h"

#ifdef KITE_CASADI_INT64
typedef long long kite_casadi_int;
#else
typedef int kite_casadi_int;
#endif

namespace {

struct Shim {
    std::shared_ptr<openkite::KiteContext> std_ctx, id_ctx;
    bool has_arm = false;
    std::string yaml;
    int device = 0;
};
Shim& shim() { static Shim s; return s; }
std::mutex& shim_mutex() { static std::mutex m; return m; }

// YAML -> parameters (host only: needed for the Jacobian pattern, which depends on the tether arm, before any GPU call)
bool load_params_locked(Shim& S, kite_params& p) {
    try {
        if (S.yaml.empty()) {
            const char* e = std::getenv("KITE_B200_YAML");
            if (!e) return false;
            S.yaml = e;
            if (const char* d = std::getenv("KITE_B200_DEVICE")) S.device = std::atoi(d);
        }
        openkite::KiteProperties props = openkite::kite_utils::LoadProperties(S.yaml);
        p = openkite::kite_utils::to_c_params(props);
        S.has_arm = p.rx != 0.0 || p.ry != 0.0 || p.rz != 0.0;
    } catch (const std::exception&) {
        return false;
    }
    return true;
}
bool model_has_arm() {
    std::lock_guard<std::mutex> lock(shim_mutex());
    kite_params p;
    return load_params_locked(shim(), p) && shim().has_arm;
}
bool ensure(bool id) {
    Shim& S = shim();
    std::lock_guard<std::mutex> lock(shim_mutex());
    std::shared_ptr<openkite::KiteContext>& c = id ? S.id_ctx : S.std_ctx;
    if (c) return true;
    kite_params p;
    if (!load_params_locked(S, p)) return false;
    try {
        c = std::make_shared<openkite::KiteContext>(p, id ? KITE_MODEL_KITE_ID : KITE_MODEL_KITE, S.device);
    } catch (const std::exception&) {
        return false;
    }
    return true;
}

// ---- sparsity patterns, built once ----------------------------------------------------------------
std::vector<kite_casadi_int> dense_pattern(int nrow, int ncol) {
    std::vector<kite_casadi_int> p{nrow, ncol};
    for (int j = 0; j <= ncol; ++j) p.push_back((kite_casadi_int)j * nrow);
    for (int j = 0; j < ncol; ++j) for (int i = 0; i < nrow; ++i) p.push_back(i);
    return p;
}
std::vector<kite_casadi_int> jac_pattern(bool arm) {
    std::vector<kite_casadi_int> p{13, 13}, rows;
    p.push_back(0);
    for (int j = 0; j < 13; ++j) {
        for (int i = 0; i < 13; ++i) if (kite::jx_nz(i, j, arm)) rows.push_back(i);
        p.push_back((kite_casadi_int)rows.size());
    }
    p.insert(p.end(), rows.begin(), rows.end());
    return p;
}
const kite_casadi_int* pat_x() { static const std::vector<kite_casadi_int> p = dense_pattern(13, 1); return p.data(); }
const kite_casadi_int* pat_u() { static const std::vector<kite_casadi_int> p = dense_pattern(3, 1); return p.data(); }
const kite_casadi_int* pat_p() { static const std::vector<kite_casadi_int> p = dense_pattern(21, 1); return p.data(); }
const kite_casadi_int* pat_s() { static const std::vector<kite_casadi_int> p = dense_pattern(1, 1); return p.data(); }
const kite_casadi_int* pat_f3() { static const std::vector<kite_casadi_int> p = dense_pattern(3, 1); return p.data(); }
const kite_casadi_int* pat_jac() {
    // the pattern depends on the tether arm of the loaded model (absent keys = 0: 104 non-zeros); without a loadable
    // YAML the zero-arm pattern of the shipped umx_radian.yaml is reported
    static std::vector<kite_casadi_int> p;
    static bool arm_built = false, built = false;
    const bool arm = model_has_arm();
    if (!built || arm != arm_built) { p = jac_pattern(arm); built = true; arm_built = arm; }
    return p.data();
}

void load(double* dst_d, const double* src, int n, openkite::KiteContext& c) {
    std::vector<double> z;
    if (!src) { z.assign((size_t)n, 0.0); src = z.data(); }       // CasADi: a null input is all zeros
    c.h2d(dst_d, src, (size_t)n);
}

int eval_point(bool id, int what, const double** arg, double** res) {     // what: 0 dynamics, 1 jacobian, 2 aero, 3 RK4
    if (!ensure(id)) return 1;
    openkite::KiteContext& c = *(id ? shim().id_ctx : shim().std_ctx);
    try {
        double* s = c.stage;
        load(s, arg[0], 13, c); load(s + 13, arg[1], 3, c);
        if (id) load(s + 16, arg[2], 21, c);
        const double* p_d = id ? s + 16 : nullptr;
        if (!res[0]) return 0;
        if (what == 0) {
            c.check(kite_rhs_batch(c.ctx, 1, 1, s, s + 13, p_d, s + 64), "kite_rhs_batch");
            c.d2h(res[0], s + 64, 13);
        } else if (what == 2) {
            c.check(kite_aero_batch(c.ctx, 1, 1, s, s + 13, p_d, s + 64), "kite_aero_batch");
            c.d2h(res[0], s + 64, 3);
        } else if (what == 1) {
            c.check(kite_jac_batch(c.ctx, 1, 1, s, s + 13, p_d, s + 64, nullptr), "kite_jac_batch");
            double J[169];
            c.d2h(J, s + 64, 169);                                  // row-major 13 x 13
            const kite_casadi_int* pat = pat_jac();
            const kite_casadi_int* colind = pat + 2; const kite_casadi_int* row = pat + 2 + 14;
            for (int j = 0; j < 13; ++j)
                for (kite_casadi_int t = colind[j]; t < colind[j + 1]; ++t) res[0][t] = J[row[t] * 13 + j];
        } else {
            const double h = arg[2] ? arg[2][0] : 0.0;
            c.check(kite_rk4_rollout(c.ctx, 1, 1, 1, h, s, s + 13, KITE_U_CONST, nullptr, s + 64, nullptr, 0, nullptr, nullptr,
                                     nullptr, 0), "kite_rk4_rollout");
            c.d2h(res[0], s + 64, 13);
        }
    } catch (const std::exception&) {
        return 1;
    }
    return 0;
}

}  // namespace

#define KITE_EXT_COMMON(NAME, NIN)                                                                                        \
    kite_casadi_int NAME##_n_in(void) { return NIN; }                                                                     \
    kite_casadi_int NAME##_n_out(void) { return 1; }                                                                      \
    int NAME##_work(kite_casadi_int* sz_arg, kite_casadi_int* sz_res, kite_casadi_int* sz_iw, kite_casadi_int* sz_w) {    \
        if (sz_arg) *sz_arg = NIN;                                                                                        \
        if (sz_res) *sz_res = 1;                                                                                          \
        if (sz_iw) *sz_iw = 0;                                                                                            \
        if (sz_w) *sz_w = 0;                                                                                              \
        return 0;                                                                                                         \
    }                                                                                                                     \
    void NAME##_incref(void) {}                                                                                           \
    void NAME##_decref(void) {}

extern "C" {

/* Explicit binding (optional): the model file and the CUDA device the external functions evaluate on. */
int kite_external_init(const char* yaml_path, int device) {
    if (!yaml_path) return 1;
    {
        std::lock_guard<std::mutex> lock(shim_mutex());
        Shim& S = shim();
        S.std_ctx.reset(); S.id_ctx.reset();
        S.yaml = yaml_path; S.device = device;
    }
    return ensure(false) ? 0 : 1;
}
void kite_external_shutdown(void) {
    std::lock_guard<std::mutex> lock(shim_mutex());
    shim().std_ctx.reset(); shim().id_ctx.reset();
}

/* ---- "dynamics"(x[13], u[3]) -> xdot[13]                                         kite.cpp:324 ---- */
int dynamics(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(false, 0, arg, res); }
KITE_EXT_COMMON(dynamics, 2)
const kite_casadi_int* dynamics_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : nullptr); }
const kite_casadi_int* dynamics_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_x() : nullptr; }
const char* dynamics_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : nullptr); }
const char* dynamics_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

/* ---- "dyn_jacobian"(x[13], u[3]) -> d f/d x, sparse 13 x 13 (104 non-zeros)      kite.cpp:327-328 ---- */
int dyn_jacobian(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(false, 1, arg, res); }
KITE_EXT_COMMON(dyn_jacobian, 2)
const kite_casadi_int* dyn_jacobian_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : nullptr); }
const kite_casadi_int* dyn_jacobian_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_jac() : nullptr; }
const char* dyn_jacobian_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : nullptr); }
const char* dyn_jacobian_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

/* ---- "Aero"(x[13], u[3]) -> Faero_b[3]                                           kite.cpp:330 ---- */
int Aero(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(false, 2, arg, res); }
KITE_EXT_COMMON(Aero, 2)
const kite_casadi_int* Aero_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : nullptr); }
const kite_casadi_int* Aero_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_f3() : nullptr; }
const char* Aero_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : nullptr); }
const char* Aero_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

/* ---- "RK4"(X[13], U[3], dT[1]) -> X+[13]                                         kite.cpp:332-338 ---- */
int RK4(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(false, 3, arg, res); }
KITE_EXT_COMMON(RK4, 3)
const kite_casadi_int* RK4_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : (i == 2 ? pat_s() : nullptr)); }
const kite_casadi_int* RK4_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_x() : nullptr; }
const char* RK4_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : (i == 2 ? "i2" : nullptr)); }
const char* RK4_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

/* ---- identification variant: "dynamics"(x, u, p[21]) / "dyn_jacobian"(x, u, p)   kite.cpp:575-579 ----
 * (the reference reuses the names "dynamics" / "dyn_jacobian" inside a different KiteDynamics object; one shared library
 * can export a name once, hence the _id suffix) */
int dynamics_id(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(true, 0, arg, res); }
KITE_EXT_COMMON(dynamics_id, 3)
const kite_casadi_int* dynamics_id_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : (i == 2 ? pat_p() : nullptr)); }
const kite_casadi_int* dynamics_id_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_x() : nullptr; }
const char* dynamics_id_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : (i == 2 ? "i2" : nullptr)); }
const char* dynamics_id_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

int dyn_jacobian_id(const double** arg, double** res, kite_casadi_int*, double*, int) { return eval_point(true, 1, arg, res); }
KITE_EXT_COMMON(dyn_jacobian_id, 3)
const kite_casadi_int* dyn_jacobian_id_sparsity_in(kite_casadi_int i) { return i == 0 ? pat_x() : (i == 1 ? pat_u() : (i == 2 ? pat_p() : nullptr)); }
const kite_casadi_int* dyn_jacobian_id_sparsity_out(kite_casadi_int i) { return i == 0 ? pat_jac() : nullptr; }
const char* dyn_jacobian_id_name_in(kite_casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "i1" : (i == 2 ? "i2" : nullptr)); }
const char* dyn_jacobian_id_name_out(kite_casadi_int i) { return i == 0 ? "o0" : nullptr; }

}  // extern "C"
