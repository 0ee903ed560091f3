// =====================================================================================
//  oracle/kite_oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see kite_oracle.hpp header).
//  extern "C" wrapper so tests/ and bench.py's cpu_baseline / --impl reference legs can
//  call the scalar CPU restatement through ctypes.  All arrays are HOST memory, row-major
//  "array of structs" (one trajectory's 13 states contiguous) -- deliberately NOT the SoA
//  device layout of the engine, so layout bugs in the engine cannot cancel out.
// =====================================================================================
#include "kite_oracle.hpp"

#include <chrono>
#include <cstring>
#include <thread>

using namespace oracle;

static inline const Params& as_params(const double* p39) { return *reinterpret_cast<const Params*>(p39); }
static_assert(sizeof(Params) == 39 * sizeof(double), "Params must be 39 packed doubles");

template <class F>
static void parallel_for(long n, int nthreads, F&& body) {
    if (nthreads <= 1 || n < 2) { body(0, n); return; }
    std::vector<std::thread> th;
    long chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        long lo = t * chunk, hi = lo + chunk > n ? n : lo + chunk;
        if (lo >= hi) break;
        th.emplace_back([=, &body]() { body(lo, hi); });
    }
    for (auto& t : th) t.join();
}

template <class T>
static void id_cost_rollout_t(const double* prm, long n, long nsteps, double h, const double* x0, const double* u,
                              const double* y, const double* p, double* cost, double* xf, int nthreads) {
    const Params& P = as_params(prm);
    parallel_for(n, nthreads, [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i) {
            T x[13], xn[13], uk[3], pk[21];
            for (int c = 0; c < 13; ++c) x[c] = T(x0[c]);
            for (int c = 0; c < 21; ++c) pk[c] = T(p[21 * i + c]);
            T acc = T(0.0);
            for (long k = 0; k < nsteps; ++k) {
                for (int c = 0; c < 3; ++c) uk[c] = T(u[3 * k + c]);
                rk4_step<T>(P, KITE_ID, x, uk, pk, T(h), xn);
                for (int c = 0; c < 13; ++c) x[c] = xn[c];
                T e = T(0.0);
                for (int c = 0; c < 13; ++c) { T d = T(y[k * 13 + c]) - x[c]; e += T(ID_COST_Q[c]) * (d * d); }
                acc += e;
            }
            cost[i] = (double)(acc * (T(1.0) / T((double)nsteps)));
            if (xf) for (int c = 0; c < 13; ++c) xf[13 * i + c] = (double)x[c];
        }
    });
}
extern "C" {

int orc_num_params() { return 39; }

void orc_rhs(const double* prm, int kind, long n, const double* x, const double* u, const double* p, double* f) {
    const Params& P = as_params(prm);
    for (long i = 0; i < n; ++i)
        model_rhs<double>(P, (ModelKind)kind, x + 13 * i, u + 3 * i, p ? p + 21 * i : nullptr, f + 13 * i);
}

// Function "Aero"(x,u) -> Faero_b (kite.cpp:330); kind KITE or KITE_ID
void orc_aero(const double* prm, int kind, long n, const double* x, const double* u, const double* p, double* F) {
    const Params& P = as_params(prm);
    double f[13];
    for (long i = 0; i < n; ++i) kite_rhs<double>(P, (ModelKind)kind, x + 13 * i, u + 3 * i, p ? p + 21 * i : nullptr, f, F + 3 * i);
}

void orc_jac(const double* prm, int kind, long n, const double* x, const double* u, const double* p, double* Jx, double* Ju) {
    const Params& P = as_params(prm);
    for (long i = 0; i < n; ++i)
        rhs_jacobian(P, (ModelKind)kind, x + 13 * i, u + 3 * i, p ? p + 21 * i : nullptr, nullptr, Jx + 169 * i, Ju + 39 * i);
}

// u_mode: 0 = u[n][3] held for all steps; 1 = u[n][nsteps][3]; 2 = shared u[nsteps][3];
//         3 = synthetic counter-based controls AND x0 (x0/u ignored), global index = traj0 + i
// traj (optional): [n][nsteps+1][13] including the initial state.
void orc_rk4_rollout(const double* prm, int kind, long n, long nsteps, double h, const double* x0, const double* u,
                     int u_mode, const double* p, long traj0, double* xf, double* traj, int nthreads) {
    const Params& P = as_params(prm);
    parallel_for(n, nthreads, [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i) {
            double x[13], xn[13], uk[3];
            if (u_mode == 3) synth_x0((uint64_t)(traj0 + i), x); else std::memcpy(x, x0 + 13 * i, sizeof x);
            if (traj) std::memcpy(traj + (i * (nsteps + 1)) * 13, x, sizeof x);
            for (long k = 0; k < nsteps; ++k) {
                switch (u_mode) {
                    case 0: std::memcpy(uk, u + 3 * i, sizeof uk); break;
                    case 1: std::memcpy(uk, u + (i * nsteps + k) * 3, sizeof uk); break;
                    case 2: std::memcpy(uk, u + k * 3, sizeof uk); break;
                    default: synth_control((uint64_t)(traj0 + i), (uint64_t)k, uk); break;
                }
                rk4_step<double>(P, (ModelKind)kind, x, uk, p ? p + 21 * i : nullptr, h, xn);
                std::memcpy(x, xn, sizeof x);
                if (traj) std::memcpy(traj + (i * (nsteps + 1) + k + 1) * 13, x, sizeof x);
            }
            std::memcpy(xf + 13 * i, x, sizeof x);
        }
    });
}

void orc_rk4_sens(const double* prm, int kind, long n, double h, const double* x, const double* u, const double* p,
                  double* xn, double* Phi, double* Gamma) {
    const Params& P = as_params(prm);
    for (long i = 0; i < n; ++i)
        rk4_step_sens(P, (ModelKind)kind, x + 13 * i, u + 3 * i, p ? p + 21 * i : nullptr, h, xn + 13 * i, Phi + 169 * i,
                      Gamma + 39 * i);
}

// Rollout with per-step sensitivities: x0[n][13], u[n][nsteps][3] -> xs[n][nsteps][13] (states AFTER each step),
// Phi[n][nsteps][169], Gamma[n][nsteps][39].
void orc_rk4_sens_rollout(const double* prm, int kind, long n, long nsteps, double h, const double* x0, const double* u,
                          double* xs, double* Phi, double* Gamma, int nthreads) {
    const Params& P = as_params(prm);
    parallel_for(n, nthreads, [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i) {
            double x[13], xn[13];
            std::memcpy(x, x0 + 13 * i, sizeof x);
            for (long k = 0; k < nsteps; ++k) {
                long o = i * nsteps + k;
                rk4_step_sens(P, (ModelKind)kind, x, u + o * 3, nullptr, h, xn, Phi + o * 169, Gamma + o * 39);
                std::memcpy(xs + o * 13, xn, sizeof xn);
                std::memcpy(x, xn, sizeof x);
            }
        }
    });
}

int orc_colloc_nodes(int P, int S) { return S * P + 1; }

// z[n][M*15 + M*4]; G[n][M*15]; JX[n][M][15][15]; JU[n][M][15][4]; prm_batch != null => per-scenario params [n][39]
void orc_colloc_eval(const double* prm, const double* prm_batch, int kind, int Pord, int S, double t0, double tf,
                     const double* sx, const double* su, long n, const double* z, double* G, double* JX, double* JU,
                     int nthreads) {
    const int M = S * Pord + 1;
    parallel_for(n, nthreads, [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i) {
            const Params& P = as_params(prm_batch ? prm_batch + 39 * i : prm);
            colloc_eval(P, (ModelKind)kind, Pord, S, t0, tf, sx, su, z + i * (M * 19), G + i * (M * 15),
                        JX ? JX + i * (M * 225) : nullptr, JU ? JU + i * (M * 60) : nullptr);
        }
    });
}

// NMPC performance index + gradient: z[n][M*19]; cost[n]; grad[n][M*19] (may be null).  cc = {Q[3],R[4],W,vref_scaled,
// radius,altitude,q_rot[4]} (15 doubles).
void orc_colloc_cost(const double* cc, int Pord, int S, double t0, double tf, const double* sx, long n, const double* z,
                     double* cost, double* grad, int nthreads) {
    NmpcCost c;
    std::memcpy(&c, cc, sizeof c);
    const int M = S * Pord + 1;
    parallel_for(n, nthreads, [&](long lo, long hi) {
        for (long i = lo; i < hi; ++i)
            cost[i] = colloc_cost(c, Pord, S, t0, tf, sx, z + i * (M * 19), grad ? grad + i * (M * 19) : nullptr);
    });
}

void orc_ekf_predict(const double* prm, int kind, long n, const double* x, const double* u, double dt, const double* Pc,
                     const double* W, double* xn, double* Pn) {
    const Params& P = as_params(prm);
    for (long i = 0; i < n; ++i)
        ekf_predict(P, (ModelKind)kind, x + 13 * i, u + 3 * i, dt, Pc + 169 * i, W, xn + 13 * i, Pn + 169 * i);
}

void orc_ekf_update(long n, const double* z, const double* V, double* x, double* Pc) {
    for (long i = 0; i < n; ++i) ekf_update(z + 7 * i, V, x + 13 * i, Pc + 169 * i);
}

void orc_ekf_defaults(double* W, double* V) { ekf_default_W(W); ekf_default_V(V); }

// Identification Monte-Carlo cost: shared x0[13], shared control log u[nsteps][3], measurement y[nsteps][13]
// (y[j] = measured state after step j+1), per-sample parameters p[n][21]; id-variant RHS.
void orc_id_cost_rollout(const double* prm, long n, long nsteps, double h, const double* x0, const double* u,
                         const double* y, const double* p, double* cost, double* xf, int nthreads) {
    id_cost_rollout_t<double>(prm, n, nsteps, h, x0, u, y, p, cost, xf, nthreads);
}
// The same rollout in 80-bit extended precision: the yardstick that tells the oracle's own round-off growth over a long
// horizon apart from an engine error (tests/test_gpu_fullsize.py).
void orc_id_cost_rollout_ld(const double* prm, long n, long nsteps, double h, const double* x0, const double* u,
                            const double* y, const double* p, double* cost, double* xf, int nthreads) {
    id_cost_rollout_t<long double>(prm, n, nsteps, h, x0, u, y, p, cost, xf, nthreads);
}
void orc_ekf_update_ld(long n, const double* z, const double* V, double* x, double* P) {
    for (long i = 0; i < n; ++i) ekf_update_t<long double>(z + 7 * i, V, x + 13 * i, P + 169 * i);
}

void orc_cheb_points(int P, double* out) { auto v = cheb_points(P); std::memcpy(out, v.data(), v.size() * 8); }
void orc_cheb_diff(int P, double* out) { auto v = cheb_diff_matrix(P); std::memcpy(out, v.data(), v.size() * 8); }
void orc_cheb_weights(int P, double* out) { auto v = cheb_quad_weights(P); std::memcpy(out, v.data(), v.size() * 8); }
void orc_cheb_compdiff(int P, int S, double* out) { auto v = cheb_comp_diff_matrix(P, S); std::memcpy(out, v.data(), v.size() * 8); }

void orc_synth_x0(long traj0, long n, double* x0) { for (long i = 0; i < n; ++i) synth_x0((uint64_t)(traj0 + i), x0 + 13 * i); }
// u[n][nsteps][3]
void orc_synth_id_params(long traj0, long n, const double* ref21, double* p) {
    for (long i = 0; i < n; ++i) synth_id_params((uint64_t)(traj0 + i), ref21, p + 21 * i);
}
void orc_synth_controls(long traj0, long n, long nsteps, double* u) {
    for (long i = 0; i < n; ++i)
        for (long k = 0; k < nsteps; ++k) synth_control((uint64_t)(traj0 + i), (uint64_t)k, u + (i * nsteps + k) * 3);
}

// Algorithmic flop counts of the literal restatement (Counted scalar).  out[4*k..]: add, mul, div, special for
// k = 0: RHS f(x,u); 1: one RK4 state-step.  Evaluated at the reference test state (kite_model_test.cpp:58-60).
void orc_flop_counts(const double* prm, long* out) {
    const Params& P = as_params(prm);
    Counted x[13], u[3], f[13];
    for (int i = 0; i < 13; ++i) x[i] = Counted(SYNTH_X0[i], false);
    u[0] = Counted(0.1, false); u[1] = Counted(0.01, false); u[2] = Counted(-0.01, false);
    tally() = OpTally();
    kite_rhs<Counted>(P, KITE, x, u, nullptr, f);
    OpTally t = tally();
    out[0] = t.add; out[1] = t.mul; out[2] = t.div; out[3] = t.special;
    tally() = OpTally();
    Counted xn[13];
    rk4_step<Counted>(P, KITE, x, u, nullptr, Counted(1e-3, false), xn);
    t = tally();
    out[4] = t.add; out[5] = t.mul; out[6] = t.div; out[7] = t.special;
}

// CPU baseline: synthetic config-2 workload slice [traj0, traj0+n) x nsteps on `nthreads` host threads.
// Returns elapsed seconds (wall clock around the parallel region only).
double orc_bench_rollout(const double* prm, long traj0, long n, long nsteps, double h, int nthreads, double* xf) {
    auto t0 = std::chrono::steady_clock::now();
    orc_rk4_rollout(prm, KITE, n, nsteps, h, nullptr, nullptr, 3, nullptr, traj0, xf, nullptr, nthreads);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
