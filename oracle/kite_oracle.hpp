// =====================================================================================
//  oracle/kite_oracle.hpp  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
//
//  Scalar CPU restatement of the openKITE hot path, written to follow the reference
//  expression-by-expression so that the CUDA engine (openkite_b200/csrc) can be checked
//  against it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
//  --impl reference legs may build, link or call anything in oracle/.
//
//  PARITY PINNING: the reference's arithmetic lives in CasADi v3.0.0-rc2 (README.md:6), a
//  third-party dependency that is NOT vendored in /root/reference and is not installable
//  here, and the reference's own tests assert no outputs (all BOOST_CHECK(true)).  So this
//  oracle is "parity unpinned" by the reference; it is pinned instead to (i) an independent
//  sympy/mpmath restatement (oracle/sympy_oracle.py, 50-digit evaluation, symbolic
//  Jacobians) and (ii) the survey-time values in SURVEY.md Appendix A.  See DESIGN.md §3.
//
//  Everything is templated on a scalar type T so that the same literal restatement yields
//    T = double          the value path (what CasADi's SX VM computes in IEEE FP64)
//    T = Dual<N>         forward-mode derivatives (what SX::jacobian gives symbolically)
//    T = Counted         algorithmic flop counts for the roofline
//
//  Reference files followed (all under /root/reference/src):
//    kite_model/kite.cpp:90-363      standard KiteDynamics ctor (RHS graph, Jacobian, RK4)
//    kite_model/kite.cpp:365-616     identification variant (21 symbolic aero parameters)
//    kite_model/kite.cpp:622-661     RigidBodyKinematics
//    kite_math/kitemath.cpp:9-51     quat_multiply, quat_inverse, heaviside, rk4_symbolic
//    kite_model/integrator.cpp:86-98 ODESolver::rk4_solve
//    kite_math/pseudospectral/chebyshev.hpp:119-271  collocation operators
//    kite_control/kiteNMPF.cpp:58-111,169-171        augmentation, scaling, AugJacobian
//    kite_estimation/kiteEKF.cpp:6-13,75-126         EKF predict / update
//    kite_control/kite_identification_test.cpp:193-205  fitting cost
// =====================================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace oracle {

// ------------------------------------------------------------------------------------
// Parameters actually consumed by the reference constructors (kite.cpp:99-175).
// ------------------------------------------------------------------------------------
struct Params {
    // geometry (kite.cpp:99-102)
    double b, c, AR, S;
    // inertia (kite.cpp:117-121)
    double Mass, Ixx, Iyy, Izz, Ixz;
    // aerodynamics (kite.cpp:126-161)
    double CL0, CLa_tot, e_o, CD0_tot, CYb, Cm0, Cma, Cn0, Cnb, Cl0, Clb;
    double CLq, Cmq, CYr, Cnr, Clr, CYp, Clp, Cnp;
    double CLde, CYdr, Cmde, Cndr, Cldr;
    // tether (kite.cpp:170-175)
    double Ks, Kd, Lt, rx, ry, rz;
};

enum ModelKind { KITE = 0, KITE_ID = 1, RIGID_BODY = 2 };

// ------------------------------------------------------------------------------------
// Forward-mode dual number with N tangents.
// ------------------------------------------------------------------------------------
template <int N>
struct Dual {
    double v;
    double d[N];
    Dual() : v(0) { for (int i = 0; i < N; ++i) d[i] = 0; }
    Dual(double a) : v(a) { for (int i = 0; i < N; ++i) d[i] = 0; }
};
template <int N> inline Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v + b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v - b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a) { Dual<N> r; r.v = -a.v; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v * b.v; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) { Dual<N> r; r.v = a.v / b.v; for (int i = 0; i < N; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
template <int N> inline Dual<N> operator+(const Dual<N>& a, double b) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator+(double b, const Dual<N>& a) { Dual<N> r = a; r.v += b; return r; }
template <int N> inline Dual<N> operator-(const Dual<N>& a, double b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> inline Dual<N> operator-(double b, const Dual<N>& a) { return Dual<N>(b) - a; }
template <int N> inline Dual<N> operator*(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v * b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] * b; return r; }
template <int N> inline Dual<N> operator*(double b, const Dual<N>& a) { return a * b; }
template <int N> inline Dual<N> operator/(const Dual<N>& a, double b) { Dual<N> r; r.v = a.v / b; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] / b; return r; }
template <int N> inline Dual<N> operator/(double a, const Dual<N>& b) { return Dual<N>(a) / b; }
template <int N> inline Dual<N> chain(const Dual<N>& a, double val, double der) { Dual<N> r; r.v = val; for (int i = 0; i < N; ++i) r.d[i] = der * a.d[i]; return r; }
template <int N> inline Dual<N> sqrt(const Dual<N>& a) { double s = std::sqrt(a.v); return chain(a, s, 0.5 / s); }
template <int N> inline Dual<N> sin(const Dual<N>& a) { return chain(a, std::sin(a.v), std::cos(a.v)); }
template <int N> inline Dual<N> cos(const Dual<N>& a) { return chain(a, std::cos(a.v), -std::sin(a.v)); }
template <int N> inline Dual<N> exp(const Dual<N>& a) { double e = std::exp(a.v); return chain(a, e, e); }
template <int N> inline Dual<N> asin(const Dual<N>& a) { return chain(a, std::asin(a.v), 1.0 / std::sqrt(1.0 - a.v * a.v)); }
template <int N> inline Dual<N> atan2(const Dual<N>& y, const Dual<N>& x) {
    Dual<N> r; r.v = std::atan2(y.v, x.v);
    double den = x.v * x.v + y.v * y.v;
    for (int i = 0; i < N; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / den;
    return r;
}

// ------------------------------------------------------------------------------------
// Op-counting scalar: counts algorithmic adds / muls / divs / special functions of the
// literal restatement (no CSE beyond what the source text itself shares).
// ------------------------------------------------------------------------------------
struct OpTally { long add = 0, mul = 0, div = 0, special = 0; long flops() const { return add + mul + div + special; } };
inline OpTally& tally() { static thread_local OpTally t; return t; }
struct Counted {
    double v;
    bool is_const;   // literal constants fold at graph-build time in CasADi; do not count const*const
    Counted() : v(0), is_const(true) {}
    Counted(double a) : v(a), is_const(true) {}
    Counted(double a, bool c) : v(a), is_const(c) {}
};
inline Counted cnt2(double v, const Counted& a, const Counted& b, long OpTally::*slot) {
    bool c = a.is_const && b.is_const;
    // CasADi SX simplifies x*0, x+0, x*1 at construction; mirror that so zero entries of
    // pure-vector quaternions do not inflate the count.
    if (!c) tally().*slot += 1;
    return Counted(v, c);
}
inline bool is0(const Counted& a) { return a.is_const && a.v == 0.0; }
inline bool is1(const Counted& a) { return a.is_const && a.v == 1.0; }
inline Counted operator+(const Counted& a, const Counted& b) { if (is0(a)) return b; if (is0(b)) return a; return cnt2(a.v + b.v, a, b, &OpTally::add); }
inline Counted operator-(const Counted& a, const Counted& b) { if (is0(b)) return a; if (is0(a)) return Counted(-b.v, b.is_const); return cnt2(a.v - b.v, a, b, &OpTally::add); }
inline Counted operator-(const Counted& a) { return Counted(-a.v, a.is_const); }
inline Counted operator*(const Counted& a, const Counted& b) { if (is0(a) || is0(b)) return Counted(0.0); if (is1(a)) return b; if (is1(b)) return a; return cnt2(a.v * b.v, a, b, &OpTally::mul); }
inline Counted operator/(const Counted& a, const Counted& b) { if (is0(a)) return Counted(0.0); if (is1(b)) return a; return cnt2(a.v / b.v, a, b, &OpTally::div); }
inline Counted cnt1(double v, const Counted& a) { if (!a.is_const) tally().special += 1; return Counted(v, a.is_const); }
inline Counted sqrt(const Counted& a) { return cnt1(std::sqrt(a.v), a); }
inline Counted sin(const Counted& a) { return cnt1(std::sin(a.v), a); }
inline Counted cos(const Counted& a) { return cnt1(std::cos(a.v), a); }
inline Counted exp(const Counted& a) { return cnt1(std::exp(a.v), a); }
inline Counted asin(const Counted& a) { return cnt1(std::asin(a.v), a); }
inline Counted atan2(const Counted& a, const Counted& b) { bool c = a.is_const && b.is_const; if (!c) tally().special += 1; return Counted(std::atan2(a.v, b.v), c); }

using std::sqrt; using std::sin; using std::cos; using std::exp; using std::asin; using std::atan2;

// ------------------------------------------------------------------------------------
// kmath primitives (kitemath.cpp:9-34), literal.
// ------------------------------------------------------------------------------------
template <class T> inline void quat_multiply(const T q1[4], const T q2[4], T out[4]) {
    // s = s1*s2 - dot(v1,v2);  v = cross(v1,v2) + s1*v2 + s2*v1      (kitemath.cpp:9-23)
    const T s1 = q1[0], s2 = q2[0];
    const T* v1 = q1 + 1; const T* v2 = q2 + 1;
    T s = (s1 * s2) - (v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]);
    T c0 = v1[1] * v2[2] - v1[2] * v2[1];
    T c1 = v1[2] * v2[0] - v1[0] * v2[2];
    T c2 = v1[0] * v2[1] - v1[1] * v2[0];
    T o1 = c0 + (s1 * v2[0]) + (s2 * v1[0]);
    T o2 = c1 + (s1 * v2[1]) + (s2 * v1[1]);
    T o3 = c2 + (s1 * v2[2]) + (s2 * v1[2]);
    out[0] = s; out[1] = o1; out[2] = o2; out[3] = o3;
}
template <class T> inline void quat_inverse(const T q[4], T out[4]) {   // kitemath.cpp:25-29
    out[0] = q[0]; out[1] = -q[1]; out[2] = -q[2]; out[3] = -q[3];
}
template <class T> inline T heaviside(const T& x, double K) {           // kitemath.cpp:31-34
    return T(K) / (T(1.0) + exp(T(-4.0) * x));
}
template <class T> inline void cross3(const T a[3], const T b[3], T o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// Order of the 21 identification parameters (kite.cpp:571-572).
enum IdParam { P_CL0 = 0, P_CLa, P_CD0, P_CYb, P_Cm0, P_Cma, P_Cnb, P_Clb, P_CLq, P_Cmq,
               P_CYr, P_Cnr, P_Clr, P_CYp, P_Clp, P_Cnp, P_CLde, P_CYdr, P_Cmde, P_Cndr, P_Cldr, P_COUNT };

inline void nominal_id_params(const Params& P, double p[21]) {
    p[P_CL0] = P.CL0; p[P_CLa] = P.CLa_tot; p[P_CD0] = P.CD0_tot; p[P_CYb] = P.CYb; p[P_Cm0] = P.Cm0;
    p[P_Cma] = P.Cma; p[P_Cnb] = P.Cnb; p[P_Clb] = P.Clb; p[P_CLq] = P.CLq; p[P_Cmq] = P.Cmq;
    p[P_CYr] = P.CYr; p[P_Cnr] = P.Cnr; p[P_Clr] = P.Clr; p[P_CYp] = P.CYp; p[P_Clp] = P.Clp;
    p[P_Cnp] = P.Cnp; p[P_CLde] = P.CLde; p[P_CYdr] = P.CYdr; p[P_Cmde] = P.Cmde; p[P_Cndr] = P.Cndr;
    p[P_Cldr] = P.Cldr;
}

// ------------------------------------------------------------------------------------
// The rigid-wing RHS  xdot = f(x,u[,p])      state x = [v(3) w(3) r(3) q(4)], u = [T dE dR]
//   kind == KITE     : kite.cpp:197-322 (1e-4 regularisers in ss / aoa, fixed coefficients)
//   kind == KITE_ID  : kite.cpp:448-573 (no regularisers, 21 coefficients from `p`)
// ------------------------------------------------------------------------------------
template <class T>
void kite_rhs(const Params& P, ModelKind kind, const T x[13], const T u[3], const T* p /*21 or null*/, T f[13],
              T* faero_out = nullptr /* 3: Faero_b, the output of Function "Aero" (kite.cpp:330) */) {
    const double g = 9.80665;     // kite.cpp:93
    const double ro = 1.2985;     // kite.cpp:94
    const double pi = 3.14159265358979323846;   // casadi::pi
    const double b = P.b, c = P.c, AR = P.AR, S = P.S;
    const double Mass = P.Mass, Ixx = P.Ixx, Iyy = P.Iyy, Izz = P.Izz, Ixz = P.Ixz;
    const double e_o = P.e_o, Cn0 = P.Cn0, Cl0 = P.Cl0;
    const double Ks = P.Ks, Kd = P.Kd, Lt = P.Lt, rx = P.rx, ry = P.ry, rz = P.rz;

    T CL0, CLa_tot, CD0_tot, CYb, Cm0, Cma, Cnb, Clb, CLq, Cmq, CYr, Cnr, Clr, CYp, Clp, Cnp, CLde, CYdr, Cmde, Cndr, Cldr;
    if (kind == KITE_ID) {
        CL0 = p[P_CL0]; CLa_tot = p[P_CLa]; CD0_tot = p[P_CD0]; CYb = p[P_CYb]; Cm0 = p[P_Cm0]; Cma = p[P_Cma];
        Cnb = p[P_Cnb]; Clb = p[P_Clb]; CLq = p[P_CLq]; Cmq = p[P_Cmq]; CYr = p[P_CYr]; Cnr = p[P_Cnr];
        Clr = p[P_Clr]; CYp = p[P_CYp]; Clp = p[P_Clp]; Cnp = p[P_Cnp]; CLde = p[P_CLde]; CYdr = p[P_CYdr];
        Cmde = p[P_Cmde]; Cndr = p[P_Cndr]; Cldr = p[P_Cldr];
    } else {
        CL0 = T(P.CL0); CLa_tot = T(P.CLa_tot); CD0_tot = T(P.CD0_tot); CYb = T(P.CYb); Cm0 = T(P.Cm0); Cma = T(P.Cma);
        Cnb = T(P.Cnb); Clb = T(P.Clb); CLq = T(P.CLq); Cmq = T(P.Cmq); CYr = T(P.CYr); Cnr = T(P.Cnr);
        Clr = T(P.Clr); CYp = T(P.CYp); Clp = T(P.Clp); Cnp = T(P.Cnp); CLde = T(P.CLde); CYdr = T(P.CYdr);
        Cmde = T(P.Cmde); Cndr = T(P.Cndr); Cldr = T(P.Cldr);
    }
    const double eps = (kind == KITE_ID) ? 0.0 : 1e-4;   // kite.cpp:200-201 vs :451-452

    const T v[3] = {x[0], x[1], x[2]};
    const T w[3] = {x[3], x[4], x[5]};
    const T r[3] = {x[6], x[7], x[8]};
    const T q[4] = {x[9], x[10], x[11], x[12]};
    const T Tthrust = u[0], dE = u[1], dR = u[2];

    T V2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];          // :198
    T V = sqrt(V2);                                            // :197

    T ss, aoa;
    if (kind == KITE_ID) { ss = asin(v[1] / V); aoa = atan2(v[2], v[0]); }            // :451-452
    else { ss = asin(v[1] / (V + T(eps))); aoa = atan2(v[2], v[0] + T(eps)); }        // :200-201
    T dyn_press = T(0.5 * ro) * V2;                                                   // :202

    T CLlin = CL0 + CLa_tot * aoa;
    T CD = CD0_tot + (CLlin * CLlin) / T(pi * e_o * AR);                             // :204

    T LIFT = CLlin * dyn_press * T(S) + (T(0.25) * CLq * T(c * S * ro)) * V * w[1];   // :209-210
    T DRAG = CD * dyn_press * T(S);                                                   // :211
    T SF = (CYb * ss + CYdr * dR) * dyn_press * T(S)
         + T(0.25) * (CYr * w[2] + CYp * w[0]) * T(b * ro * S) * V;                   // :212-213

    T q_aoa[4] = {cos(aoa / T(2.0)), T(0.0), sin(aoa / T(2.0)), T(0.0)};              // :217
    T q_ss[4] = {cos(-ss / T(2.0)), T(0.0), T(0.0), sin(-ss / T(2.0))};               // :218
    T qw_b[4], qw_b_inv[4];
    quat_multiply(q_aoa, q_ss, qw_b);                                                 // :220
    quat_inverse(qw_b, qw_b_inv);                                                     // :221

    T Fw[4] = {T(0.0), -DRAG, T(0.0), -LIFT};
    T qF_tmp[4], qF_q[4];
    quat_multiply(qw_b_inv, Fw, qF_tmp);                                              // :224
    quat_multiply(qF_tmp, qw_b, qF_q);                                                // :225
    T Faero_b[3] = {qF_q[1], qF_q[2], qF_q[3]};                                       // :226

    T Zde = (-CLde) * dE * dyn_press * T(S);                                          // :228
    T q_aoa_inv[4]; quat_inverse(q_aoa, q_aoa_inv);
    T Zq[4] = {T(0.0), T(0.0), T(0.0), Zde};
    T FdE_tmp[4], qFdE[4];
    quat_multiply(q_aoa_inv, Zq, FdE_tmp);                                            // :229-230
    quat_multiply(FdE_tmp, q_aoa, qFdE);                                              // :231
    Faero_b[0] = Faero_b[0] + qFdE[1];                                                // :234
    Faero_b[1] = Faero_b[1] + qFdE[2] + SF;
    Faero_b[2] = Faero_b[2] + qFdE[3];
    if (faero_out) { faero_out[0] = Faero_b[0]; faero_out[1] = Faero_b[1]; faero_out[2] = Faero_b[2]; }

    T q_inv[4]; quat_inverse(q, q_inv);
    T gq[4] = {T(0.0), T(0.0), T(0.0), T(g)};
    T qG[4], qG_q[4];
    quat_multiply(q_inv, gq, qG);                                                     // :237-238
    quat_multiply(qG, q, qG_q);                                                       // :239
    T G_b[3] = {qG_q[1], qG_q[2], qG_q[3]};

    T T_b[3] = {Tthrust, T(0.0), T(0.0)};                                             // :243

    T d_ = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);                             // :247
    T Rv = d_ - T(Lt);                                                                // :252
    T Rs[3] = {-Rv * (r[0] / d_), -Rv * (r[1] / d_), -Rv * (r[2] / d_)};              // :253
    T vq[4] = {T(0.0), v[0], v[1], v[2]};
    T qvi[4], qvi_q[4];
    quat_multiply(q, vq, qvi);                                                        // :255
    quat_multiply(qvi, q_inv, qvi_q);                                                 // :256
    T vi[3] = {qvi_q[1], qvi_q[2], qvi_q[3]};
    T rdotvi = r[0] * vi[0] + r[1] * vi[1] + r[2] * vi[2];
    T Rd[3] = {(-r[0] / d_) * rdotvi / d_, (-r[1] / d_) * rdotvi / d_, (-r[2] / d_) * rdotvi / d_};   // :258
    T hv = heaviside(d_ - T(Lt), 1.0);
    T R[3] = {(T(Ks) * Rs[0] + T(Kd) * Rd[0]) * hv, (T(Ks) * Rs[1] + T(Kd) * Rd[1]) * hv,
              (T(Ks) * Rs[2] + T(Kd) * Rd[2]) * hv};                                  // :259

    T Rq[4] = {T(0.0), R[0], R[1], R[2]};
    T qR[4], qR_q[4];
    quat_multiply(q_inv, Rq, qR);                                                     // :262-263
    quat_multiply(qR, q, qR_q);                                                       // :264
    T R_b[3] = {qR_q[1], qR_q[2], qR_q[3]};

    T wxv[3]; cross3(w, v, wxv);
    T v_dot[3];
    for (int i = 0; i < 3; ++i) v_dot[i] = (Faero_b[i] + T_b[i] + R_b[i]) / T(Mass) + G_b[i] - wxv[i];   // :268

    T L = (T(Cl0) + Clb * ss + Cldr * dR) * dyn_press * T(S) * T(b)
        + (Clr * w[2] + Clp * w[0]) * T(0.25 * ro * std::pow(b, 2) * S) * V;          // :274-275
    T M = (Cm0 + Cma * aoa + Cmde * dE) * dyn_press * T(S) * T(c)
        + Cmq * T(0.25 * S * std::pow(c, 2) * ro) * w[1] * V;                         // :278-279
    T N = (T(Cn0) + Cnb * ss + Cndr * dR) * dyn_press * T(S) * T(b)
        + (Cnp * w[0] + Cnr * w[2]) * T(0.25 * S * std::pow(b, 2) * ro) * V;          // :282-283

    T LMN[4] = {T(0.0), L, M, N};
    T T_tmp[4], Trot[4];
    quat_multiply(q_aoa_inv, LMN, T_tmp);                                             // :293-294
    quat_multiply(T_tmp, q_aoa, Trot);                                                // :295
    T Maero[3] = {Trot[1], Trot[2], Trot[3]};

    T arm[3] = {T(rx), T(ry), T(rz)};
    T Mt[3]; cross3(arm, R_b, Mt);                                                    // :299-300

    // J = [[Ixx,0,Ixz],[0,Iyy,0],[Ixz,0,Izz]]  (:286-289);  w_dot = inv(J) (Maero + Mt - w x (J w))  (:302)
    T Jw[3] = {T(Ixx) * w[0] + T(Ixz) * w[2], T(Iyy) * w[1], T(Ixz) * w[0] + T(Izz) * w[2]};
    T wxJw[3]; cross3(w, Jw, wxJw);
    T rhs[3] = {Maero[0] + Mt[0] - wxJw[0], Maero[1] + Mt[1] - wxJw[1], Maero[2] + Mt[2] - wxJw[2]};
    const double det = Ixx * Izz - Ixz * Ixz;
    const double Ji00 = Izz / det, Ji02 = -Ixz / det, Ji11 = 1.0 / Iyy, Ji22 = Ixx / det;
    T w_dot[3] = {T(Ji00) * rhs[0] + T(Ji02) * rhs[2], T(Ji11) * rhs[1], T(Ji02) * rhs[0] + T(Ji22) * rhs[2]};

    // r_dot = vec(q (x) [0,v] (x) q^-1)  (:308-310)  -- same expression as vi
    const double lambda = -5.0;                                                       // :316
    T wq[4] = {T(0.0), w[0], w[1], w[2]};
    T qw[4]; quat_multiply(q, wq, qw);
    T qq1 = (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) - T(1.0);
    T q_dot[4];
    for (int i = 0; i < 4; ++i) q_dot[i] = T(0.5) * qw[i] + T(0.5 * lambda) * q[i] * qq1;   // :317

    f[0] = v_dot[0]; f[1] = v_dot[1]; f[2] = v_dot[2];
    f[3] = w_dot[0]; f[4] = w_dot[1]; f[5] = w_dot[2];
    f[6] = vi[0]; f[7] = vi[1]; f[8] = vi[2];
    f[9] = q_dot[0]; f[10] = q_dot[1]; f[11] = q_dot[2]; f[12] = q_dot[3];
}

// RigidBodyKinematics (kite.cpp:622-661): vdot = wdot = 0, rdot = q v q^-1, lambda = -10.
template <class T>
void rigid_body_rhs(const T x[13], T f[13]) {
    const T v[3] = {x[0], x[1], x[2]};
    const T w[3] = {x[3], x[4], x[5]};
    const T q[4] = {x[9], x[10], x[11], x[12]};
    T q_inv[4]; quat_inverse(q, q_inv);
    T vq[4] = {T(0.0), v[0], v[1], v[2]};
    T qv[4], qv_q[4];
    quat_multiply(q, vq, qv);
    quat_multiply(qv, q_inv, qv_q);
    const double lambda = -10.0;
    T wq[4] = {T(0.0), w[0], w[1], w[2]};
    T qw[4]; quat_multiply(q, wq, qw);
    T qq1 = (q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) - T(1.0);
    for (int i = 0; i < 6; ++i) f[i] = T(0.0);
    f[6] = qv_q[1]; f[7] = qv_q[2]; f[8] = qv_q[3];
    for (int i = 0; i < 4; ++i) f[9 + i] = T(0.5) * qw[i] + T(0.5 * lambda) * q[i] * qq1;
}

template <class T>
inline void model_rhs(const Params& P, ModelKind kind, const T x[13], const T u[3], const T* p, T f[13]) {
    if (kind == RIGID_BODY) rigid_body_rhs(x, f); else kite_rhs(P, kind, x, u, p, f);
}

// ------------------------------------------------------------------------------------
// One classical RK4 step, control held (kitemath.cpp:36-51 == integrator.cpp:86-98).
// ------------------------------------------------------------------------------------
template <class T>
void rk4_step(const Params& P, ModelKind kind, const T x[13], const T u[3], const T* p, const T& h, T xn[13]) {
    T k1[13], k2[13], k3[13], k4[13], xt[13];
    model_rhs(P, kind, x, u, p, k1);
    for (int i = 0; i < 13; ++i) xt[i] = x[i] + T(0.5) * h * k1[i];
    model_rhs(P, kind, xt, u, p, k2);
    for (int i = 0; i < 13; ++i) xt[i] = x[i] + T(0.5) * h * k2[i];
    model_rhs(P, kind, xt, u, p, k3);
    for (int i = 0; i < 13; ++i) xt[i] = x[i] + h * k3[i];
    model_rhs(P, kind, xt, u, p, k4);
    for (int i = 0; i < 13; ++i) xt[i] = k1[i] + T(2.0) * k2[i] + T(2.0) * k3[i] + k4[i];
    for (int i = 0; i < 13; ++i) xn[i] = x[i] + (h / T(6.0)) * xt[i];
}

// State / control Jacobians of the RHS by forward mode (what SX::jacobian yields, kite.cpp:327).
// Jx is 13x13 row-major (Jx[i*13+j] = d f_i / d x_j), Ju is 13x3 row-major.
inline void rhs_jacobian(const Params& P, ModelKind kind, const double x[13], const double u[3], const double* p,
                         double f[13], double Jx[169], double Ju[39]) {
    typedef Dual<16> D;
    D X[13], U[3], F[13], Pd[21];
    for (int i = 0; i < 13; ++i) { X[i] = D(x[i]); X[i].d[i] = 1.0; }
    for (int i = 0; i < 3; ++i) { U[i] = D(u[i]); U[i].d[13 + i] = 1.0; }
    if (p) for (int i = 0; i < 21; ++i) Pd[i] = D(p[i]);
    model_rhs<D>(P, kind, X, U, p ? Pd : nullptr, F);
    for (int i = 0; i < 13; ++i) {
        if (f) f[i] = F[i].v;
        for (int j = 0; j < 13; ++j) Jx[i * 13 + j] = F[i].d[j];
        for (int j = 0; j < 3; ++j) Ju[i * 3 + j] = F[i].d[13 + j];
    }
}

// RK4 step sensitivities Phi = d x+ / d x (13x13 row-major), Gamma = d x+ / d u (13x3 row-major),
// by differentiating the whole RK4 map (cf. MATLAB RK4_JACOBIAN, scripts/matlab/kite_sim.m:300-301).
inline void rk4_step_sens(const Params& P, ModelKind kind, const double x[13], const double u[3], const double* p,
                          double h, double xn[13], double Phi[169], double Gamma[39]) {
    typedef Dual<16> D;
    D X[13], U[3], XN[13], Pd[21];
    for (int i = 0; i < 13; ++i) { X[i] = D(x[i]); X[i].d[i] = 1.0; }
    for (int i = 0; i < 3; ++i) { U[i] = D(u[i]); U[i].d[13 + i] = 1.0; }
    if (p) for (int i = 0; i < 21; ++i) Pd[i] = D(p[i]);
    rk4_step<D>(P, kind, X, U, p ? Pd : nullptr, D(h), XN);
    for (int i = 0; i < 13; ++i) {
        xn[i] = XN[i].v;
        for (int j = 0; j < 13; ++j) Phi[i * 13 + j] = XN[i].d[j];
        for (int j = 0; j < 3; ++j) Gamma[i * 3 + j] = XN[i].d[13 + j];
    }
}

// ------------------------------------------------------------------------------------
// Chebyshev pseudospectral operators (chebyshev.hpp:119-232).
// ------------------------------------------------------------------------------------
inline std::vector<double> cheb_points(int P) {                                   // :119-127
    std::vector<double> x(P + 1);
    for (int k = 0; k <= P; ++k) x[k] = std::cos(double(k) * (M_PI / P));
    return x;
}
inline std::vector<double> cheb_diff_matrix(int P) {                              // :136-153, row-major (P+1)^2
    const int n = P + 1;
    std::vector<double> xs = cheb_points(P), c(n), Dn(n * n), D(n * n);
    for (int k = 0; k < n; ++k) c[k] = std::pow(-1.0, k) * ((k == 0 || k == P) ? 2.0 : 1.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            Dn[i * n + j] = (c[i] * (1.0 / c[j])) / ((xs[i] - xs[j]) + (i == j ? 1.0 : 0.0));
    for (int i = 0; i < n; ++i) {
        double rs = 0;
        for (int j = 0; j < n; ++j) rs += Dn[i * n + j];
        for (int j = 0; j < n; ++j) D[i * n + j] = Dn[i * n + j] - (i == j ? rs : 0.0);
    }
    return D;
}
inline std::vector<double> cheb_quad_weights(int P) {                             // :162-195 (Clenshaw-Curtis)
    const int n = P + 1;
    std::vector<double> w(n, 0.0), v(P - 1, 1.0), theta(n);
    for (int k = 0; k < n; ++k) theta[k] = double(k) * (M_PI / P);
    if (P % 2 == 0) {
        w[0] = 1.0 / (double(P) * P - 1.0); w[P] = w[0];
        for (int k = 1; k <= P / 2 - 1; ++k)
            for (int i = 1; i < P; ++i) v[i - 1] -= 2.0 * std::cos(2.0 * k * theta[i]) / (4.0 * k * k - 1.0);
        for (int i = 1; i < P; ++i) v[i - 1] -= std::cos(P * theta[i]) / (double(P) * P - 1.0);
    } else {
        w[0] = 1.0 / (double(P) * P); w[P] = w[0];
        for (int k = 1; k <= (P - 1) / 2; ++k)
            for (int i = 1; i < P; ++i) v[i - 1] -= 2.0 * std::cos(2.0 * k * theta[i]) / (4.0 * k * k - 1.0);
    }
    for (int i = 1; i < P; ++i) w[i] = 2.0 * v[i - 1] / P;
    return w;
}
// Composite differentiation matrix BEFORE the kron with I_NX: (S*P+1)^2 row-major (:204-232).
inline std::vector<double> cheb_comp_diff_matrix(int P, int S) {
    const int m = S * P + 1, n = P + 1;
    std::vector<double> D = cheb_diff_matrix(P), C(m * m, 0.0);
    if (S < 2) return D;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) C[(m - n + i) * m + (m - n + j)] = D[i * n + j];
    for (int k = 0; k < (S - 1) * P; k += P)
        for (int i = 0; i < P; ++i)
            for (int j = 0; j < n; ++j) C[(k + i) * m + (k + j)] = D[i * n + j];
    return C;
}

// ------------------------------------------------------------------------------------
// NMPC collocation constraint G(z) and its Jacobian (chebyshev.hpp:241-271, kiteNMPF.cpp:58-111,169-171).
//   z = [X (M*15) ; U (M*4)], M = S*P+1 nodes, node 0 = final time.
//   aug dynamics: [f(x13,u3) ; V1 ; Uv]                                        (kiteNMPF.cpp:62-73)
//   scaled:       f_s(xs,us) = Sx * f_aug(Sx^-1 xs, Su^-1 us)  (Sx,Su diagonal) (kiteNMPF.cpp:100-103)
//   G = (CompD (x) I15) X - tau * F,  tau = (tf-t0)/(2S)
// Outputs: G[M*15]; JX[M][15*15] row-major = d f_s(node)/d xs ; JU[M][15*4] = d f_s/d us
//   (the varying blocks of AugJacobian are -tau*JX, -tau*JU; the constant part is CompD (x) I).
// ------------------------------------------------------------------------------------
template <class T>
void aug_scaled_rhs(const Params& P, ModelKind kind, const double sx[15], const double su[4],
                    const T xs[15], const T us[4], T fs[15]) {
    T x[15], u[4], f[13];
    for (int i = 0; i < 15; ++i) x[i] = T(1.0 / sx[i]) * xs[i];      // invSX = solve(Scale_X, I)
    for (int i = 0; i < 4; ++i) u[i] = T(1.0 / su[i]) * us[i];
    model_rhs<T>(P, kind, x, u, nullptr, f);
    for (int i = 0; i < 13; ++i) fs[i] = T(sx[i]) * f[i];
    fs[13] = T(sx[13]) * x[14];
    fs[14] = T(sx[14]) * u[3];
}

inline void colloc_eval(const Params& P, ModelKind kind, int Pord, int S, double t0, double tf,
                        const double sx[15], const double su[4], const double* z,
                        double* G, double* JX, double* JU) {
    const int M = S * Pord + 1;
    const double tau = (tf - t0) / (2.0 * S);
    std::vector<double> C = cheb_comp_diff_matrix(Pord, S);
    const double* X = z; const double* U = z + M * 15;
    typedef Dual<19> D;
    std::vector<double> F(M * 15);
    for (int k = 0; k < M; ++k) {
        D xs[15], us[4], fs[15];
        for (int i = 0; i < 15; ++i) { xs[i] = D(X[k * 15 + i]); xs[i].d[i] = 1.0; }
        for (int i = 0; i < 4; ++i) { us[i] = D(U[k * 4 + i]); us[i].d[15 + i] = 1.0; }
        aug_scaled_rhs<D>(P, kind, sx, su, xs, us, fs);
        for (int i = 0; i < 15; ++i) {
            F[k * 15 + i] = tau * fs[i].v;
            if (JX) for (int j = 0; j < 15; ++j) JX[(k * 15 + i) * 15 + j] = fs[i].d[j];
            if (JU) for (int j = 0; j < 4; ++j) JU[(k * 15 + i) * 4 + j] = fs[i].d[15 + j];
        }
    }
    for (int k = 0; k < M; ++k)
        for (int i = 0; i < 15; ++i) {
            double acc = 0.0;
            for (int l = 0; l < M; ++l) acc += C[k * M + l] * X[l * 15 + i];
            G[k * 15 + i] = acc - F[k * 15 + i];
        }
}

// ------------------------------------------------------------------------------------
// NMPC performance index and its gradient (Chebyshev::CollocateCost chebyshev.hpp:280-333 on the Lagrange / Mayer
// terms of kiteNMPF.cpp:116-143).  Path = circle of given radius / altitude rotated by a quaternion, as both callers
// build it (nmpf_node.cpp:31-39 tilted by pi/8 about y; kite_control_test.cpp:242-247 untilted).
//   residual = Sx[6:9] * path(x_s[13] / Sx[13]) - x_s[6:9]                          (kiteNMPF.cpp:120-122)
//   L(x,u)   = sum Q_c residual_c^2 + W (vref_s - x_s[14])^2 + sum R_m u_s[m]^2      (:123-124)
//   Mayer(x) = sum Q_c residual_c^2 at node 0 (= final time)                         (:141, chebyshev.hpp:291-295)
//   cost     = Mayer(X_0) + sum_seg tau * sum_{m=0..P} w_m L(X_{seg*P+m}, U_{seg*P+m})   (chebyshev.hpp:298-330)
// with Q = 1e2 diag(10,10,100), R = diag(1e-4,1e-1,1e-1,1e-3), W = 1e-3 (kiteNMPF.cpp:32-34) as defaults.
// ------------------------------------------------------------------------------------
struct NmpcCost {
    double Q[3], R[4], W, vref_scaled;          // vref_scaled = Scale_X(14,14) * vel_ref (kiteNMPF.h:34)
    double radius, altitude, q_rot[4];
};
inline NmpcCost nmpc_cost_defaults(const double sx[15], double vel_ref, double radius, double altitude, const double q_rot[4]) {
    NmpcCost c{{1e2 * 1e1, 1e2 * 1e1, 1e2 * 1e2}, {1e-4, 1e-1, 1e-1, 1e-3}, 1e-3, sx[14] * vel_ref, radius, altitude,
               {q_rot[0], q_rot[1], q_rot[2], q_rot[3]}};
    return c;
}
template <class T> inline void path_point(const NmpcCost& c, const T& theta, T out[3]) {
    // Path = [r cos, r sin, alt]; rotated: vec( conj(q) (x) [0,Path] (x) q )          (nmpf_node.cpp:34-39)
    T q[4] = {T(c.q_rot[0]), T(c.q_rot[1]), T(c.q_rot[2]), T(c.q_rot[3])}, qi[4], Pq[4], t1[4], t2[4];
    Pq[0] = T(0.0); Pq[1] = T(c.radius) * cos(theta); Pq[2] = T(c.radius) * sin(theta); Pq[3] = T(c.altitude);
    quat_inverse(q, qi);
    quat_multiply(qi, Pq, t1);
    quat_multiply(t1, q, t2);
    out[0] = t2[1]; out[1] = t2[2]; out[2] = t2[3];
}
template <class T> inline T nmpc_path_cost(const NmpcCost& c, const double sx[15], const T xs[15]) {
    T th = T(1.0 / sx[13]) * xs[13], pp[3], acc = T(0.0);
    path_point<T>(c, th, pp);
    for (int i = 0; i < 3; ++i) { T r = T(sx[6 + i]) * pp[i] - xs[6 + i]; acc = acc + T(c.Q[i]) * (r * r); }
    return acc;
}
template <class T> inline T nmpc_lagrange(const NmpcCost& c, const double sx[15], const T xs[15], const T us[4]) {
    T acc = nmpc_path_cost<T>(c, sx, xs);
    T dv = T(c.vref_scaled) - xs[14];
    acc = acc + T(c.W) * (dv * dv);
    for (int m = 0; m < 4; ++m) acc = acc + T(c.R[m]) * (us[m] * us[m]);
    return acc;
}
// z = [X (M*15) ; U (M*4)]; returns the cost, grad[M*19] (same ordering as z) if non-null.
inline double colloc_cost(const NmpcCost& c, int Pord, int S, double t0, double tf, const double sx[15], const double* z, double* grad) {
    const int M = S * Pord + 1;
    const double tau = (tf - t0) / (2.0 * S);
    std::vector<double> w = cheb_quad_weights(Pord);
    const double* X = z; const double* U = z + M * 15;
    typedef Dual<19> D;
    if (grad) for (int i = 0; i < M * 19; ++i) grad[i] = 0.0;
    double cost = 0.0;
    {   // Mayer term at node 0
        D xs[15];
        for (int i = 0; i < 15; ++i) { xs[i] = D(X[i]); xs[i].d[i] = 1.0; }
        D m = nmpc_path_cost<D>(c, sx, xs);
        cost += m.v;
        if (grad) for (int i = 0; i < 15; ++i) grad[i] += m.d[i];
    }
    for (int k = 0; k < S; ++k) {
        double local = 0.0;
        for (int m = 0; m <= Pord; ++m) {
            const int n = k * Pord + m;
            D xs[15], us[4];
            for (int i = 0; i < 15; ++i) { xs[i] = D(X[n * 15 + i]); xs[i].d[i] = 1.0; }
            for (int i = 0; i < 4; ++i) { us[i] = D(U[n * 4 + i]); us[i].d[15 + i] = 1.0; }
            D L = nmpc_lagrange<D>(c, sx, xs, us);
            local += w[m] * L.v;
            if (grad) {
                for (int i = 0; i < 15; ++i) grad[n * 15 + i] += tau * w[m] * L.d[i];
                for (int i = 0; i < 4; ++i) grad[M * 15 + n * 4 + i] += tau * w[m] * L.d[15 + i];
            }
        }
        cost += tau * local;
    }
    return cost;
}

// ------------------------------------------------------------------------------------
// EKF (kiteEKF.cpp:6-13, 75-126).  P, W are 13x13 row-major.
// ------------------------------------------------------------------------------------
inline void ekf_default_W(double W[169]) {
    const double s[13] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.1, 0.1, 0.01, 0.05, 0.05, 0.05};
    for (int i = 0; i < 169; ++i) W[i] = 0.0;
    for (int i = 0; i < 13; ++i) W[i * 13 + i] = s[i] * s[i];
}
inline void ekf_default_V(double V[49]) {
    const double s[7] = {0.01, 0.01, 0.01, 0.0001, 0.005, 0.005, 0.005};
    for (int i = 0; i < 49; ++i) V[i] = 0.0;
    for (int i = 0; i < 7; ++i) V[i * 7 + i] = s[i] * s[i];
}
inline void ekf_predict(const Params& Pm, ModelKind kind, const double x[13], const double u[3], double dt,
                        const double Pc[169], const double W[169], double xn[13], double Pn[169]) {
    double Jx[169], Ju[39], A[169], AP[169];
    const double h = dt;
    rk4_step<double>(Pm, kind, x, u, nullptr, h, xn);                     // kiteEKF.cpp:78-82
    rhs_jacobian(Pm, kind, x, u, nullptr, nullptr, Jx, Ju);               // Jacobian at the PRE-step state (:93)
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) A[i * 13 + j] = Jx[i * 13 + j] * dt + (i == j ? 1.0 : 0.0);
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) { double a = 0; for (int k = 0; k < 13; ++k) a += A[i * 13 + k] * Pc[k * 13 + j]; AP[i * 13 + j] = a; }
    for (int i = 0; i < 13; ++i)
        for (int j = 0; j < 13; ++j) { double a = 0; for (int k = 0; k < 13; ++k) a += AP[i * 13 + k] * A[j * 13 + k]; Pn[i * 13 + j] = a + W[i * 13 + j]; }   // :94
}
// Update with H = [0_{7x6} I_7] (kiteEKF.cpp:115-125); in-place on x (13) and P (13x13).
// Templated on the scalar so that the tests can measure the update's own conditioning (T = long double against T = double:
// the 7x7 innovation covariance is inverted, and the covariance update P - K H P cancels).
template <class T>
inline void ekf_update_t(const double z[7], const double V[49], double x[13], double Pc[169]) {
    T y[7], Smat[49], Sinv[49], K[13 * 7];
    for (int i = 0; i < 7; ++i) y[i] = T(z[i]) - T(x[6 + i]);
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) Smat[i * 7 + j] = T(Pc[(6 + i) * 13 + (6 + j)]) + T(V[i * 7 + j]);
    // dense inverse by Gauss-Jordan with partial pivoting (DM::solve(S, eye(7)))
    T aug[7][14];
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) { aug[i][j] = Smat[i * 7 + j]; aug[i][7 + j] = T(i == j ? 1.0 : 0.0); }
    for (int c = 0; c < 7; ++c) {
        int piv = c; for (int i = c + 1; i < 7; ++i) if (std::fabs(aug[i][c]) > std::fabs(aug[piv][c])) piv = i;
        if (piv != c) for (int j = 0; j < 14; ++j) { T t = aug[c][j]; aug[c][j] = aug[piv][j]; aug[piv][j] = t; }
        T d = aug[c][c]; for (int j = 0; j < 14; ++j) aug[c][j] /= d;
        for (int i = 0; i < 7; ++i) if (i != c) { T m = aug[i][c]; for (int j = 0; j < 14; ++j) aug[i][j] -= m * aug[c][j]; }
    }
    for (int i = 0; i < 7; ++i) for (int j = 0; j < 7; ++j) Sinv[i * 7 + j] = aug[i][7 + j];
    for (int i = 0; i < 13; ++i) for (int j = 0; j < 7; ++j) { T a = 0; for (int k = 0; k < 7; ++k) a += T(Pc[i * 13 + (6 + k)]) * Sinv[k * 7 + j]; K[i * 7 + j] = a; }
    T xo[13], Po[169];
    for (int i = 0; i < 13; ++i) { T a = 0; for (int k = 0; k < 7; ++k) a += K[i * 7 + k] * y[k]; xo[i] = T(x[i]) + a; }
    for (int i = 0; i < 13; ++i) for (int j = 0; j < 13; ++j) {
        T a = 0; for (int k = 0; k < 7; ++k) a += K[i * 7 + k] * T(Pc[(6 + k) * 13 + j]);   // (K H) P
        Po[i * 13 + j] = T(Pc[i * 13 + j]) - a;
    }
    for (int i = 0; i < 13; ++i) x[i] = (double)xo[i];
    for (int i = 0; i < 169; ++i) Pc[i] = (double)Po[i];
}
inline void ekf_update(const double z[7], const double V[49], double x[13], double Pc[169]) { ekf_update_t<double>(z, V, x, Pc); }

// ------------------------------------------------------------------------------------
// Synthetic-input generator shared by the CPU baseline, the tests and the GPU engine's
// on-device generator.  Counter-based: splitmix64 of a key made of
// (seed, global trajectory index, step, channel) -> 53-bit uniform in [0,1).
// This is workload definition, not reference behaviour.
// ------------------------------------------------------------------------------------
inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
inline double counter_uniform(uint64_t seed, uint64_t traj, uint64_t step, uint64_t channel) {
    uint64_t k = splitmix64(seed ^ splitmix64(traj ^ splitmix64((step << 8) | channel)));
    return double(k >> 11) * (1.0 / 9007199254740992.0);
}
static const uint64_t SYNTH_SEED = 0x6b697465ULL;   // "kite"
static const double SYNTH_X0[13] = {6.1977743e+00, -2.8407148e-02, 9.1815942e-01, 2.9763089e-01, -2.2052198e+00, -1.4827499e-01,
                                    -4.1624807e-01, -2.2601052e+00, 1.2903439e+00, 3.5646195e-02, -6.9986094e-02, 8.2660637e-01,
                                    5.5727089e-01};   // kite_model_test.cpp:58-60
// channel ids: 0..12 = x0 perturbation (step field = 0xFFFFFF), 0..2 = controls at a step
inline void synth_x0(uint64_t traj, double x0[13]) {
    const double amp[13] = {0.5, 0.5, 0.5, 0.2, 0.2, 0.2, 0.2, 0.2, 0.2, 0.05, 0.05, 0.05, 0.05};
    for (int c = 0; c < 13; ++c) x0[c] = SYNTH_X0[c] + amp[c] * (2.0 * counter_uniform(SYNTH_SEED, traj, 0xFFFFFFULL, c) - 1.0);
    double n = std::sqrt(x0[9] * x0[9] + x0[10] * x0[10] + x0[11] * x0[11] + x0[12] * x0[12]);
    for (int c = 9; c < 13; ++c) x0[c] = x0[c] / n;
}
inline void synth_control(uint64_t traj, uint64_t step, double u[3]) {
    const double amax = 8.0 * (M_PI / 180.0);        // +-8 deg; T in [0,0.3]  (kite_control_test.cpp:261-263 ranges)
    u[0] = 0.3 * counter_uniform(SYNTH_SEED, traj, step, 0);
    u[1] = amax * (2.0 * counter_uniform(SYNTH_SEED, traj, step, 1) - 1.0);
    u[2] = amax * (2.0 * counter_uniform(SYNTH_SEED, traj, step, 2) - 1.0);
}

// Identification-sweep parameter sample (config 5): reference coefficients perturbed uniformly inside the bounds of
// kite_identification_test.cpp:127-148 (fractions of |ref|), channel = coefficient index, step field = 0xFFFFFE.
static const double ID_BOUNDS_LO[21] = {-0.1, -0.05, -0.1, -0.5, -0.5, -0.1, -0.5, -0.5, -0.2, -0.3, -0.3, -0.5, -0.5, -0.5, -0.5, -0.3, -0.5, -0.5, -0.5, -0.5, -0.5};
static const double ID_BOUNDS_HI[21] = {0.1, 0.1, 0.25, 0.5, 0.5, 0.30, 0.5, 0.5, 0.2, 0.3, 0.3, 0.5, 0.5, 0.5, 0.5, 1.0, 0.5, 0.5, 0.5, 0.5, 0.5};
inline void synth_id_params(uint64_t traj, const double ref[21], double p[21]) {
    for (int c = 0; c < 21; ++c) {
        const double t = counter_uniform(SYNTH_SEED, traj, 0xFFFFFEULL, (uint64_t)c);
        const double w = ID_BOUNDS_HI[c] - ID_BOUNDS_LO[c];
        const double frac = ID_BOUNDS_LO[c] + w * t;
        p[c] = ref[c] + std::fabs(ref[c]) * frac;
    }
}

// Identification fitting cost (kite_identification_test.cpp:193-205):
//   J = (1/N) sum_j sum_c Q_c (y_cj - x_cj)^2
static const double ID_COST_Q[13] = {1e3, 1e2, 1e2, 1e2, 1e2, 1e2, 1e1, 1e1, 1e2, 1e2, 1e2, 1e2, 1e2};

}  // namespace oracle
