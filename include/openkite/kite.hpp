// kite.hpp -- C++ host mirror of openKITE's kite_model API (reference: src/kite_model/kite.h:9-173,
// src/kite_model/kite.cpp) on top of the C ABI (include/kite_b200.h).  Same type / method names and argument
// meaning as the reference so that callers (simulator, KiteNMPF, KiteEKF, the tests) read the same; the
// casadi::Function / DM value types are replaced by openkite::Function / DM (dm.hpp).  Every numeric call
// lands in a hand-written CUDA kernel; nothing here computes dynamics on the CPU.
#pragma once
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#include "../kite_b200.h"
#include "dm.hpp"
#include "yaml_lite.hpp"

namespace openkite {

enum IntType { RK4, CVODES, CHEBYCHEV };   // kitemath.h:10

// ---- property structs, field-for-field as in kite.h:9-93 -------------------------------------------
struct PlaneGeometry {
    double WingSpan, MAC, AspectRatio, WingSurfaceArea, TaperRatio, HTailsurface, TailLeverArm, FinSurfaceArea,
        FinLeverArm, AerodynamicCenter;
};
struct PlaneInertia { double Mass, Ixx, Iyy, Izz, Ixz; };
struct PlaneAerodynamics {
    double CL0, CL0_tail, CLa_total, CLa_wing, CLa_tail, e_oswald;
    double CD0_total, CD0_wing, CD0_tail, CYb, CYb_vtail, Cm0, Cma, Cn0, Cnb, Cl0, Clb;
    double CLq, Cmq, CYr, Cnr, Clr, CYp, Clp, Cnp;
    double CLde, CYdr, Cmde, Cndr, Cldr, CDde;
};
struct TetherProperties { double length, Ks, Kd, rx, ry, rz; };
struct KiteProperties {
    std::string Name;
    PlaneGeometry Geometry;
    PlaneInertia Inertia;
    PlaneAerodynamics Aerodynamics;
    TetherProperties Tether;
};
struct AlgorithmProperties {
    IntType Integrator = RK4;
    double sampling_time = 0.02;
};

namespace kite_utils {

/** YAML -> KiteProperties (reference kite.cpp:7-76).  Differences, both documented in DESIGN.md:
 *  tether.rx/ry/rz default to 0.0 when absent (the shipped umx_radian.yaml lacks them);
 *  any other missing key throws std::runtime_error naming the key (yaml-cpp would throw too). */
inline KiteProperties LoadProperties(const std::string& filename) {
    yaml_lite::Document config = yaml_lite::load_file(filename);
    KiteProperties props;
    props.Name = config.str("", "name");
    auto g = [&](const char* k) { return config.num("geometry", k); };
    props.Geometry = {g("b"), g("c"), g("AR"), g("S"), g("lam"), g("St"), g("lt"), g("Sf"), g("lf"), g("Xac")};
    auto in = [&](const char* k) { return config.num("inertia", k); };
    props.Inertia = {in("mass"), in("Ixx"), in("Iyy"), in("Izz"), in("Ixz")};
    auto a = [&](const char* k) { return config.num("aerodynamic", k); };
    PlaneAerodynamics& A = props.Aerodynamics;
    A.CL0 = a("CL0"); A.CL0_tail = a("CL0_tail"); A.CLa_total = a("CLa_total"); A.CLa_wing = a("CLa_wing");
    A.CLa_tail = a("CLa_tail"); A.e_oswald = a("e_oswald");
    A.CD0_total = a("CD0_total"); A.CD0_wing = a("CD0_wing"); A.CD0_tail = a("CD0_tail"); A.CYb = a("CYb");
    A.CYb_vtail = a("CYb_vtail"); A.Cm0 = a("Cm0"); A.Cma = a("Cma"); A.Cn0 = a("Cn0"); A.Cnb = a("Cnb");
    A.Cl0 = a("Cl0"); A.Clb = a("Clb");
    A.CLq = a("CLq"); A.Cmq = a("Cmq"); A.CYr = a("CYr"); A.Cnr = a("Cnr"); A.Clr = a("Clr"); A.CYp = a("CYp");
    A.Clp = a("Clp"); A.Cnp = a("Cnp");
    A.CLde = a("CLde"); A.CYdr = a("CYdr"); A.Cmde = a("Cmde"); A.Cndr = a("Cndr"); A.Cldr = a("Cldr"); A.CDde = a("CDde");
    props.Tether.length = config.num("tether", "length");
    props.Tether.Ks = config.num("tether", "Ks");
    props.Tether.Kd = config.num("tether", "Kd");
    props.Tether.rx = config.num_or("tether", "rx", 0.0);
    props.Tether.ry = config.num_or("tether", "ry", 0.0);
    props.Tether.rz = config.num_or("tether", "rz", 0.0);
    return props;
}

typedef std::chrono::time_point<std::chrono::system_clock> time_point;
inline time_point get_time() { return std::chrono::system_clock::now(); }

/** KiteProperties -> the POD the C ABI takes (only the fields the dynamics consume, kite.cpp:99-175). */
inline kite_params to_c_params(const KiteProperties& p) {
    kite_params c;
    c.b = p.Geometry.WingSpan; c.c = p.Geometry.MAC; c.AR = p.Geometry.AspectRatio; c.S = p.Geometry.WingSurfaceArea;
    c.mass = p.Inertia.Mass; c.Ixx = p.Inertia.Ixx; c.Iyy = p.Inertia.Iyy; c.Izz = p.Inertia.Izz; c.Ixz = p.Inertia.Ixz;
    const PlaneAerodynamics& A = p.Aerodynamics;
    c.CL0 = A.CL0; c.CLa_total = A.CLa_total; c.e_oswald = A.e_oswald; c.CD0_total = A.CD0_total; c.CYb = A.CYb;
    c.Cm0 = A.Cm0; c.Cma = A.Cma; c.Cn0 = A.Cn0; c.Cnb = A.Cnb; c.Cl0 = A.Cl0; c.Clb = A.Clb;
    c.CLq = A.CLq; c.Cmq = A.Cmq; c.CYr = A.CYr; c.Cnr = A.Cnr; c.Clr = A.Clr; c.CYp = A.CYp; c.Clp = A.Clp; c.Cnp = A.Cnp;
    c.CLde = A.CLde; c.CYdr = A.CYdr; c.Cmde = A.Cmde; c.Cndr = A.Cndr; c.Cldr = A.Cldr;
    c.Ks = p.Tether.Ks; c.Kd = p.Tether.Kd; c.tether_length = p.Tether.length;
    c.rx = p.Tether.rx; c.ry = p.Tether.ry; c.rz = p.Tether.rz;
    return c;
}

}  // namespace kite_utils

/** Owns one engine context plus small device staging buffers for the single-point (B = 1) calls. */
class KiteContext {
public:
    KiteContext(const kite_params& p, int model_kind, int device = 0) : params(p), kind(model_kind) {
        int rc = kite_create(&ctx, &p, model_kind, device);
        if (rc != KITE_OK) throw std::runtime_error("kite_create failed (status " + std::to_string(rc) + "): a CUDA device is required, there is no CPU fallback");
        if (kite_ctx_malloc(ctx, (void**)&stage, sizeof(double) * STAGE_DOUBLES) != 0) throw std::runtime_error("device staging allocation failed");
    }
    ~KiteContext() { if (stage) kite_ctx_free(ctx, stage); if (ctx) kite_destroy(ctx); }
    KiteContext(const KiteContext&) = delete;
    KiteContext& operator=(const KiteContext&) = delete;

    void check(int rc, const char* what) const { if (rc != KITE_OK) throw std::runtime_error(std::string(what) + ": " + kite_last_error(ctx)); }
    void h2d(double* dst, const double* src, size_t n) { check(kite_copy_h2d(ctx, dst, src, n * sizeof(double)), "h2d"); }
    void d2h(double* dst, const double* src, size_t n) { check(kite_copy_d2h(ctx, dst, src, n * sizeof(double)), "d2h"); }

    kite_ctx* ctx = nullptr;
    kite_params params;
    int kind;
    static constexpr size_t STAGE_DOUBLES = 4096;
    double* stage = nullptr;     // device scratch for B = 1 calls: [x 13 | u 3 | p 21 | out ...]
};

namespace detail {
inline std::shared_ptr<KiteContext> ctx_of(const Function& f) {
    auto p = std::static_pointer_cast<KiteContext>(f.owner());
    if (!p) throw std::runtime_error("Function '" + f.name() + "' is not bound to a kite engine context");
    return p;
}
}  // namespace detail

/** KiteDynamics (kite.h:105-150): the reference builds CasADi graphs in the constructor; here the constructor
 *  creates a GPU engine context and the getNumeric* methods return Function handles that evaluate on it. */
class KiteDynamics {
public:
    KiteDynamics(const KiteProperties& KiteProps, const AlgorithmProperties& AlgoProps) { init(KiteProps, AlgoProps, false); }
    KiteDynamics(const KiteProperties& KiteProps, const AlgorithmProperties& AlgoProps, const bool& id) { init(KiteProps, AlgoProps, id); }
    virtual ~KiteDynamics() {}

    /** dynamics(x[13], u[3]) -> xdot[13]; identification variant: dynamics(x, u, p[21])   (kite.cpp:324, :575) */
    Function getNumericDynamics() { return NumDynamics; }
    /** RK4(X[13], U[3], dT[1]) -> X+[13]   (kite.cpp:332-338).  Null for the identification variant (kite.cpp:587-615). */
    Function getNumericIntegrator() { return NumIntegrator; }
    /** dyn_jacobian(x, u[, p]) -> d f/d x (13x13)   (kite.cpp:327-328, :578-579) */
    Function getNumericJacobian() { return NumJacobian; }
    /** d f/d u (13x3): not exposed by the reference class, used by the sensitivity / collocation paths */
    Function getNumericControlJacobian() { return NumControlJacobian; }
    /** Aero(x[13], u[3]) -> body-frame aerodynamic force Faero_b[3]   (kite.h:126, kite.cpp:224-234, :330) */
    Function getAeroDynamicForces() { return AeroDynamics; }

    std::shared_ptr<KiteContext> context() const { return Ctx; }
    const KiteProperties& properties() const { return Props; }

private:
    void init(const KiteProperties& KiteProps, const AlgorithmProperties& AlgoProps, bool id) {
        if (AlgoProps.Integrator != RK4)
            std::cerr << "KiteDynamics: only the RK4 integrator is provided by the GPU engine (CVODES is out of scope)\n";
        Props = KiteProps;
        Ctx = std::make_shared<KiteContext>(kite_utils::to_c_params(KiteProps), id ? KITE_MODEL_KITE_ID : KITE_MODEL_KITE);
        std::shared_ptr<KiteContext> c = Ctx;
        std::vector<int> ins = id ? std::vector<int>{13, 3, 21} : std::vector<int>{13, 3};
        NumDynamics = Function("dynamics", ins, {13}, [c, id](const DMVector& a) {
            double* s = c->stage;
            c->h2d(s, a[0].ptr(), 13); c->h2d(s + 13, a[1].ptr(), 3);
            if (id) c->h2d(s + 16, a[2].ptr(), 21);
            c->check(kite_rhs_batch(c->ctx, 1, 1, s, s + 13, id ? s + 16 : nullptr, s + 64), "kite_rhs_batch");
            DM f(13, 1); c->d2h(f.ptr(), s + 64, 13);
            return DMVector{f};
        }, c);
        auto jac = [c, id](const DMVector& a, bool wrt_u) {
            double* s = c->stage;
            c->h2d(s, a[0].ptr(), 13); c->h2d(s + 13, a[1].ptr(), 3);
            if (id) c->h2d(s + 16, a[2].ptr(), 21);
            c->check(kite_jac_batch(c->ctx, 1, 1, s, s + 13, id ? s + 16 : nullptr, s + 64, s + 64 + 169), "kite_jac_batch");
            std::vector<double> buf(169 + 39); c->d2h(buf.data(), s + 64, 169 + 39);
            return wrt_u ? DM::from_row_major(buf.data() + 169, 13, 3) : DM::from_row_major(buf.data(), 13, 13);
        };
        NumJacobian = Function("dyn_jacobian", ins, {169}, [jac](const DMVector& a) { return DMVector{jac(a, false)}; }, c);
        NumControlJacobian = Function("dyn_jacobian_u", ins, {39}, [jac](const DMVector& a) { return DMVector{jac(a, true)}; }, c);
        AeroDynamics = Function("Aero", ins, {3}, [c, id](const DMVector& a) {
            double* s = c->stage;
            c->h2d(s, a[0].ptr(), 13); c->h2d(s + 13, a[1].ptr(), 3);
            if (id) c->h2d(s + 16, a[2].ptr(), 21);
            c->check(kite_aero_batch(c->ctx, 1, 1, s, s + 13, id ? s + 16 : nullptr, s + 64), "kite_aero_batch");
            DM F(3, 1); c->d2h(F.ptr(), s + 64, 3);
            return DMVector{F};
        }, c);
        if (!id) {
            NumIntegrator = Function("RK4", {13, 3, 1}, {13}, [c](const DMVector& a) {
                double* s = c->stage;
                c->h2d(s, a[0].ptr(), 13); c->h2d(s + 13, a[1].ptr(), 3);
                c->check(kite_rk4_rollout(c->ctx, 1, 1, 1, a[2][0], s, s + 13, KITE_U_CONST, nullptr, s + 64, nullptr, 0,
                                          nullptr, nullptr, nullptr, 0), "kite_rk4_rollout");
                DM xn(13, 1); c->d2h(xn.ptr(), s + 64, 13);
                return DMVector{xn};
            }, c);
        }
    }
    KiteProperties Props;
    std::shared_ptr<KiteContext> Ctx;
    Function NumDynamics, NumIntegrator, NumJacobian, NumControlJacobian, AeroDynamics;
};

/** RigidBodyKinematics (kite.h:153-173, kite.cpp:622-661).  The reference's integrator here is CVODES; the engine
 *  provides the fixed-step RK4 map under the name "RK4" (SURVEY.md 8f rank 4). */
class RigidBodyKinematics {
public:
    explicit RigidBodyKinematics(const AlgorithmProperties& AlgoProps) : algo_props(AlgoProps) {
        kite_params p; std::memset(&p, 0, sizeof p);
        p.mass = p.Ixx = p.Iyy = p.Izz = 1.0; p.b = p.c = p.AR = p.S = p.e_oswald = 1.0; p.tether_length = 1.0;
        Ctx = std::make_shared<KiteContext>(p, KITE_MODEL_RIGID_BODY);
        std::shared_ptr<KiteContext> c = Ctx;
        NumDynamics = Function("RB_Dynamics", {13}, {13}, [c](const DMVector& a) {
            double* s = c->stage; c->h2d(s, a[0].ptr(), 13);
            c->check(kite_rhs_batch(c->ctx, 1, 1, s, nullptr, nullptr, s + 64), "kite_rhs_batch");
            DM f(13, 1); c->d2h(f.ptr(), s + 64, 13); return DMVector{f};
        }, c);
        NumJacobian = Function("RB_Jacobian", {13, 3}, {169}, [c](const DMVector& a) {
            double* s = c->stage; c->h2d(s, a[0].ptr(), 13);
            c->check(kite_jac_batch(c->ctx, 1, 1, s, nullptr, nullptr, s + 64, nullptr), "kite_jac_batch");
            std::vector<double> buf(169); c->d2h(buf.data(), s + 64, 169);
            return DMVector{DM::from_row_major(buf.data(), 13, 13)};
        }, c);
        NumIntegartor = Function("RK4", {13, 3, 1}, {13}, [c](const DMVector& a) {
            double* s = c->stage; c->h2d(s, a[0].ptr(), 13);
            c->check(kite_rk4_rollout(c->ctx, 1, 1, 1, a[2][0], s, nullptr, KITE_U_CONST, nullptr, s + 64, nullptr, 0, nullptr,
                                      nullptr, nullptr, 0), "kite_rk4_rollout");
            DM xn(13, 1); c->d2h(xn.ptr(), s + 64, 13); return DMVector{xn};
        }, c);
    }
    virtual ~RigidBodyKinematics() {}
    Function getNumericIntegrator() { return NumIntegartor; }
    Function getNumericJacobian() { return NumJacobian; }
    Function getNumericDynamcis() { return NumDynamics; }   // (sic) spelling of the reference, kite.h:162
    std::shared_ptr<KiteContext> context() const { return Ctx; }

private:
    AlgorithmProperties algo_props;
    std::shared_ptr<KiteContext> Ctx;
    Function NumIntegartor, NumJacobian, NumDynamics;
};

}  // namespace openkite
