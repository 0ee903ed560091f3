// kiteEKF.hpp -- host mirror of openKITE's KiteEKF (reference: src/kite_estimation/kiteEKF.h:12-62,
// kiteEKF.cpp:6-126).  Same constructor / setter / getter / propagate / _estimate surface.  The predict step
// (kiteEKF.cpp:75-98) and the update step (:108-126) run as CUDA kernels (kite_ekf_predict_batch / _update_batch);
// a batched device entry point is added for ensembles of filters.
#pragma once
#include <iostream>

#include "kite.hpp"

namespace openkite {

class KiteEKF {
public:
    /** experimentally defined values (kiteEKF.cpp:6-13) */
    static DM default_process_covariance() {
        const double s[13] = {0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.1, 0.1, 0.01, 0.05, 0.05, 0.05};   // [SIGMA_V, SIGMA_W, SIGMA_R, SIGMA_Q]
        DM W(13, 13); for (int i = 0; i < 13; ++i) W(i, i) = s[i] * s[i]; return W;
    }
    static DM default_measurement_covariance() {
        const double s[7] = {0.01, 0.01, 0.01, 0.0001, 0.005, 0.005, 0.005};
        DM V(7, 7); for (int i = 0; i < 7; ++i) V(i, i) = s[i] * s[i]; return V;
    }
    static DM default_measurement_matrix() { DM H(7, 13); for (int i = 0; i < 7; ++i) H(i, 6 + i) = 1.0; return H; }

    KiteEKF(const KiteProperties& KiteProps, const AlgorithmProperties& AlgoProps) {
        Kite = std::make_shared<KiteDynamics>(KiteProps, AlgoProps);
        init(Kite->getNumericIntegrator(), Kite->getNumericJacobian());
    }
    explicit KiteEKF(std::shared_ptr<KiteDynamics> obj_Kite) : Kite(obj_Kite) { init(Kite->getNumericIntegrator(), Kite->getNumericJacobian()); }
    KiteEKF(const Function& _Dynamics, const Function& _Jacobian) { init(_Dynamics, _Jacobian); }
    virtual ~KiteEKF() {}

    void setProcessCovariance(const DM& _W) { W = _W; }
    void setMeasurementCovariance(const DM& _V) { V = _V; }
    void setEstimationCovariance(const DM& _P) { Cov_Est = _P; }
    void setEstimation(const DM& _estimation) { State_Est = _estimation; }
    void setControl(const DM& _control) { Control = _control; }
    void setTime(const double& _current_time) { tstamp = _current_time; }
    DM getEstimation() { return State_Est; }
    DM getEstimationCovariance() { return Cov_Est; }
    double getTimeStamp() { return tstamp; }

    void estimate(const DM& measurement, const double& _tstamp) {
        double dt = _tstamp - tstamp;
        _estimate(measurement, dt);
    }
    void _estimate(const DM& measurement, const double& _dt) {
        propagate(_dt);
        // update step with H = [0 I7] (kiteEKF.cpp:115-125) on the device
        double* s = Ctx->stage;
        std::vector<double> P = Cov_Est.row_major(), Vr = V.row_major();
        Ctx->h2d(s, State_Est.ptr(), 13);
        Ctx->h2d(s + 16, measurement.ptr(), 7);
        Ctx->h2d(s + 32, P.data(), 169);
        Ctx->check(kite_ekf_update_batch(Ctx->ctx, 1, 1, s + 16, Vr.data(), s, s + 32), "kite_ekf_update_batch");
        DM x(13, 1); Ctx->d2h(x.ptr(), s, 13);
        Ctx->d2h(P.data(), s + 32, 169);
        State_Est = x; Cov_Est = DM::from_row_major(P.data(), 13, 13);
    }

    /** x+ = RK4(x,u,dt);  A = I + Jx(x,u) dt (at the pre-step state);  P+ = A P A^T + W   (kiteEKF.cpp:75-98) */
    void propagate(const double& _dt) {
        const bool rk4 = m_Integrator.name().find("RK4") != std::string::npos;
        if (!rk4 && m_Integrator.name().find("CVODES") != std::string::npos)
            // kiteEKF.cpp:83-88: the CVODES branch integrates with SUNDIALS, which is outside the GPU path;
            // failing loudly beats silently skipping the covariance propagation
            throw std::runtime_error("KiteEKF::propagate: a CVODES integrator is not supported by the GPU engine (use RK4)");
        if (!rk4) std::cout << "WARNING: Unknown intergrator! \n";          // kiteEKF.cpp:89-90, then falls through like the reference
        if (State_Est.numel() != 13) throw std::invalid_argument("KiteEKF::propagate: set the estimation first");
        DM u = Control.is_empty() ? DM::zeros(3) : Control;
        double* s = Ctx->stage;
        std::vector<double> P = Cov_Est.row_major(), Wr = W.row_major();
        Ctx->h2d(s, State_Est.ptr(), 13);
        Ctx->h2d(s + 13, u.ptr(), 3);
        Ctx->h2d(s + 32, P.data(), 169);
        // layout in the staging buffer: x[0:13] u[13:16] P[32:201] xn[208:221] Pn[224:393] work[400:532]
        Ctx->check(kite_ekf_predict_batch(Ctx->ctx, 1, 1, _dt, s, Ctx->kind == KITE_MODEL_RIGID_BODY ? nullptr : s + 13, s + 32,
                                          Wr.data(), s + 208, s + 224, s + 400), "kite_ekf_predict_batch");
        DM x(13, 1); Ctx->d2h(x.ptr(), s + 208, 13);
        Ctx->d2h(P.data(), s + 224, 169);
        // the covariance is propagated whichever integrator branch was taken (kiteEKF.cpp:92-97); with an unknown
        // integrator the reference assigns a default-constructed (empty) state
        State_Est = rk4 ? x : DM();
        Cov_Est = DM::from_row_major(P.data(), 13, 13);
    }

    /** Batched predict for an ensemble of B filters (device SoA pointers, include/kite_b200.h). */
    void propagate_device(long B, double dt, const double* x_d, const double* u_d, const double* P_d, double* xn_d,
                          double* Pn_d, void* work_d) {
        std::vector<double> Wr = W.row_major();
        Ctx->check(kite_ekf_predict_batch(Ctx->ctx, B, B, dt, x_d, u_d, P_d, Wr.data(), xn_d, Pn_d, work_d), "kite_ekf_predict_batch");
    }

private:
    void init(const Function& integ, const Function& jac) {
        m_Integrator = integ; m_Jacobian = jac;
        Ctx = detail::ctx_of(jac);
        W = default_process_covariance();
        V = default_measurement_covariance();
        H = default_measurement_matrix();
        Cov_Est = 10.0 * W;                                   // kiteEKF.cpp:26
        auto us = std::chrono::duration_cast<std::chrono::microseconds>(kite_utils::get_time().time_since_epoch()).count();
        tstamp = static_cast<double>(us) * 1e-6;
    }
    std::shared_ptr<KiteDynamics> Kite;
    std::shared_ptr<KiteContext> Ctx;
    DM State_Est, Cov_Est, Control;
    Function m_Integrator, m_Jacobian;
    DM W, V, H;
    double tstamp = 0.0;
};

}  // namespace openkite
