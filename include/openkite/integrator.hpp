// integrator.hpp -- host mirror of openKITE's ODESolver (reference: src/kite_model/integrator.h:10-60,
// integrator.cpp:7-109, :245-273).  Same constructor / solve / updateParams / getParams / dim_x / dim_u surface.
// Only the RK4 branch is on the hot path: solve() is ONE classical RK4 step of size dt with the control held
// (integrator.cpp:86-98; the time loop lives in the caller, simulator.cpp:43-51) and runs as a CUDA kernel.
// The CVODES (adaptive BDF) and CHEBYCHEV (dense Newton) branches are out of scope and throw.
#pragma once
#include <iostream>

#include "kite.hpp"

namespace openkite {

class ODESolver {
public:
    ODESolver(const Function& rhs, const Dict& params = Dict()) : RHS(rhs) {
        // defaults of integrator.cpp:12-25
        Parameters["method"] = CVODES;
        Parameters["tf"] = 1;
        Parameters["restart"] = 0;
        Parameters["max_iter"] = 300;
        Parameters["tol"] = 1e-8;
        Parameters["poly_order"] = 10;
        if (!params.empty()) updateParams(params);
        nx = RHS.nnz_out();                 // integrator.cpp:40-41
        nu = RHS.nnz_in() - nx;
        Ctx = detail::ctx_of(RHS);
        if (nx != 13) throw std::invalid_argument("ODESolver: the GPU engine integrates the 13-state kite / rigid-body RHS");
        has_params = (RHS.n_in() == 3);     // identification variant dynamics(x,u,p): nu counts u and p
        if (has_params) nu = RHS.size_in(1);
    }
    virtual ~ODESolver() {}

    /** One integration step of length dt from x0 with control u held (integrator.cpp:245-273). */
    DM solve(const DM& x0, const DM& u, const double& dt) {
        const int method = (int)Parameters["method"];
        switch (method) {
            case RK4: return rk4_solve(x0, u, dt);
            case CVODES: throw std::runtime_error("ODESolver: CVODES is outside the GPU hot path (SURVEY.md section 2 row 3)");
            case CHEBYCHEV: throw std::runtime_error("ODESolver: the Chebyshev-Newton solver is outside the GPU hot path");
            default: throw std::runtime_error("ODESolver: unknown method");
        }
    }

    /** Batched-rollout entry point (new; BASELINE.json north_star): B trajectories x N steps on the device.
     *  Pointers are DEVICE pointers in the SoA layout of include/kite_b200.h. */
    void rollout_device(long B, long N, double h, const double* x0_d, const double* u_d, kite_u_mode u_mode, double* xf_d,
                        const double* p_d = nullptr, int32_t* status_d = nullptr) {
        Ctx->check(kite_rk4_rollout(Ctx->ctx, B, B, N, h, x0_d, u_d, (int)u_mode, p_d, xf_d, nullptr, 0, nullptr, nullptr,
                                    status_d, 0), "kite_rk4_rollout");
    }
    /** Same with HOST buffers (SoA [13][B], controls per u_mode); copies are pipelined inside the engine. */
    void rollout_host(long B, long N, double h, const double* x0_h, const double* u_h, kite_u_mode u_mode, double* xf_h,
                      const double* p_h = nullptr, int32_t* status_h = nullptr) {
        Ctx->check(kite_rk4_rollout_host(Ctx->ctx, B, N, h, x0_h, u_h, (int)u_mode, p_h, xf_h, nullptr, nullptr, status_h),
                   "kite_rk4_rollout_host");
    }

    void updateParams(const Dict& params) {
        for (Dict::const_iterator it = params.begin(); it != params.end(); ++it) {
            if (Parameters.count(it->first) > 0) Parameters[it->first] = it->second;
            else std::cout << "Unknown parameter: " << it->first << "\n";        // integrator.cpp:107
        }
    }
    Dict getParams() { return Parameters; }
    int dim_x() { return nx; }
    int dim_u() { return nu; }

private:
    DM rk4_solve(const DM& x0, const DM& u, const double& dt) {
        if (x0.numel() != 13) throw std::invalid_argument("ODESolver::solve: x0 must have 13 elements");
        double* s = Ctx->stage;
        Ctx->h2d(s, x0.ptr(), 13);
        const double* u_d = nullptr; const double* p_d = nullptr;
        if (Ctx->kind != KITE_MODEL_RIGID_BODY) {
            if (has_params) {                 // u carries [controls ; parameters] like DMVector{x0, u} would for a 3-input RHS
                if (u.numel() != 3 + 21) throw std::invalid_argument("ODESolver::solve: expected u = [T dE dR ; p(21)] for the identification model");
                Ctx->h2d(s + 13, u.ptr(), 24); u_d = s + 13; p_d = s + 16;
            } else {
                if (u.numel() != 3) throw std::invalid_argument("ODESolver::solve: u must have 3 elements");
                Ctx->h2d(s + 13, u.ptr(), 3); u_d = s + 13;
            }
        }
        Ctx->check(kite_rk4_rollout(Ctx->ctx, 1, 1, 1, dt, s, u_d, KITE_U_CONST, p_d, s + 64, nullptr, 0, nullptr, nullptr,
                                    nullptr, 0), "kite_rk4_rollout");
        DM xn(13, 1);
        Ctx->d2h(xn.ptr(), s + 64, 13);
        return xn;
    }

    Function RHS;
    Dict Parameters;
    int nx, nu;
    bool has_params = false;
    std::shared_ptr<KiteContext> Ctx;
};

}  // namespace openkite
