// chebyshev.hpp -- host mirror of openKITE's Chebyshev<BaseClass,PolyOrder,NumSegments,NX,NU,NP> collocation
// class (reference: src/kite_math/pseudospectral/chebyshev.hpp:6-271) for the numeric (DM) use the NMPC makes of it.
// The one-time operators (points, D, Clenshaw-Curtis weights, composite D) are built on the host exactly as the
// reference builds them; the per-iteration work -- G(X,U) = (CompD (x) I) X - tau F(X,U) and its Jacobian
// (kiteNMPF.cpp:94-111,169-171) -- is evaluated by the CUDA kernel k_colloc_eval through kite_colloc_eval.
#pragma once
#include <cmath>

#include "kite.hpp"

namespace openkite {

template <int PolyOrder, int NumSegments, int NX, int NU, int NP>
class Chebyshev {
public:
    static constexpr int NODES = NumSegments * PolyOrder + 1;

    Chebyshev() {
        _Points = CollocPoints();
        _D = DiffMatrix();
        _QuadWeights = QuadWeights();
        _CompDiff = CompDiffBlock();
    }
    virtual ~Chebyshev() {}

    DM D() { return _D; }
    /** composite differentiation matrix kron(CompDiff, I_NX) as the reference returns it (chebyshev.hpp:231) */
    DM CompD() {
        DM K(NODES * NX, NODES * NX);
        for (int i = 0; i < NODES; ++i) for (int j = 0; j < NODES; ++j) { double v = _CompDiff(i, j); if (v != 0.0) for (int k = 0; k < NX; ++k) K(i * NX + k, j * NX + k) = v; }
        return K;
    }
    /** the (S*P+1)^2 block matrix before the Kronecker product (what the engine consumes) */
    DM CompDBlock() { return _CompDiff; }
    DM CPoints() { return _Points; }
    DM QWeights() { return _QuadWeights; }

    /** Collocated dynamics of the scaled augmented kite model (kiteNMPF.cpp:58-111): the returned object evaluates
     *  G(z) and d G / d z for z = [X (NODES*15) ; U (NODES*4)].  Requires NX = 15, NU = 4. */
    class Collocated {
    public:
        Collocated(std::shared_ptr<KiteContext> ctx, const DM& compd, double tau, const DM& Sx, const DM& Su)
            : Ctx(ctx), compD_rm(compd.row_major()), tau_(tau), compd_(compd) {
            for (int i = 0; i < 15; ++i) sx[i] = Sx.size2() > 1 ? Sx(i, i) : Sx[i];
            for (int i = 0; i < 4; ++i) su[i] = Su.size2() > 1 ? Su(i, i) : Su[i];
            const size_t nd = (size_t)NODES * (19 + 15 + 225 + 60) + 1;
            if (kite_ctx_malloc(Ctx->ctx, (void**)&buf, sizeof(double) * nd) != 0) throw std::runtime_error("Collocated: device allocation failed");
            nnz_ = kite_colloc_nnz_per_node(Ctx->ctx);
            prow_.resize((size_t)nnz_); pcol_.resize((size_t)nnz_);
            if (kite_colloc_sparsity(Ctx->ctx, prow_.data(), pcol_.data()) != nnz_) throw std::runtime_error("Collocated: sparsity query failed");
        }
        ~Collocated() { if (buf) kite_ctx_free(Ctx->ctx, buf); }
        Collocated(const Collocated&) = delete;

        /** constraint residual G (NODES*15), as DynamicConstraints / nlp_g (kiteNMPF.cpp:151) */
        DM G(const DM& z) { DM g, j; eval(z, g, j, false); return g; }
        /** dense d G / d z (NODES*15 x NODES*19), as AugJacobian (kiteNMPF.cpp:169-171) */
        DM Jacobian(const DM& z) { DM g, j; eval(z, g, j, true); return j; }
        void eval(const DM& z, DM& Gout, DM& Jout, bool want_jac = true) {
            const int M = NODES;
            std::vector<double> jv;
            eval_values(z, Gout, jv, want_jac);
            if (!want_jac) return;
            // assemble the dense matrix from the structural non-zeros: [kron(CompD, I15) - tau blkdiag(JX_k) | -tau blkdiag(JU_k)]
            Jout = DM(M * 15, M * 19);
            for (int k = 0; k < M; ++k) {
                for (int l = 0; l < M; ++l) { double v = compd_(k, l); if (v != 0.0) for (int i = 0; i < 15; ++i) Jout(k * 15 + i, l * 15 + i) = v; }
                for (int s = 0; s < nnz_; ++s) {
                    const int i = prow_[s], j = pcol_[s];
                    const double v = -tau_ * jv[(size_t)k * nnz_ + s];
                    if (j < 15) Jout(k * 15 + i, k * 15 + j) += v;
                    else Jout(k * 15 + i, M * 15 + k * 4 + (j - 15)) = v;
                }
            }
        }
        /** AugJacobian in compressed-column storage, the form the reference holds it in (a sparse casadi::SX Jacobian,
         *  kiteNMPF.cpp:169-171): colptr [NODES*19 + 1], rowind [nnz], values [nnz].  The pattern is the union of
         *  kron(CompD, I15) and the per-node structural non-zeros; it does not depend on z. */
        void JacobianCCS(const DM& z, std::vector<int>& colptr, std::vector<int>& rowind, std::vector<double>& values) {
            const int M = NODES;
            DM g; std::vector<double> jv;
            eval_values(z, g, jv, true);
            colptr.assign((size_t)M * 19 + 1, 0); rowind.clear(); values.clear();
            for (int col = 0; col < M * 19; ++col) {
                const bool isx = col < M * 15;
                const int k = isx ? col / 15 : (col - M * 15) / 4;          // node owning this column's varying block
                const int jl = isx ? col % 15 : 15 + (col - M * 15) % 4;    // column inside the 15 x 19 node block
                for (int row = 0; row < M * 15; ++row) {
                    const int kr = row / 15, il = row % 15;
                    double v = 0.0; bool nz = false;
                    if (isx && il == jl && compd_(kr, k) != 0.0) { v = compd_(kr, k); nz = true; }
                    if (kr == k) {
                        for (int s = 0; s < nnz_; ++s)
                            if (prow_[s] == il && pcol_[s] == jl) { v -= tau_ * jv[(size_t)k * nnz_ + s]; nz = true; break; }
                    }
                    if (nz) { rowind.push_back(row); values.push_back(v); }
                }
                colptr[(size_t)col + 1] = (int)rowind.size();
            }
        }
        /** structural non-zeros per node block and their (row, column) inside the 15 x 19 block, CCS order */
        int nnz_per_node() const { return nnz_; }
        const std::vector<int>& pattern_rows() const { return prow_; }
        const std::vector<int>& pattern_cols() const { return pcol_; }
        /** batched device entry point (config 4): see kite_colloc_eval in include/kite_b200.h */
        void eval_device(long B, const double* z_d, const double* p_d, double* G_d, double* JX_d, double* JU_d, double* gnorm_d) {
            Ctx->check(kite_colloc_eval(Ctx->ctx, B, B, NODES, compD_rm.data(), tau_, sx, su, z_d, p_d, G_d, JX_d, JU_d, gnorm_d), "kite_colloc_eval");
        }
        /** the same with sparse node blocks: JV_d [NODES * nnz_per_node()][B] */
        void eval_device_sparse(long B, const double* z_d, const double* p_d, double* G_d, double* JV_d, double* gnorm_d) {
            Ctx->check(kite_colloc_eval_sparse(Ctx->ctx, B, B, NODES, compD_rm.data(), tau_, sx, su, z_d, p_d, G_d, JV_d, gnorm_d), "kite_colloc_eval_sparse");
        }
        double tau() const { return tau_; }

    private:
        /** G and the structural non-zeros of the node blocks ([NODES][nnz]) through kite_colloc_eval_sparse (B = 1) */
        void eval_values(const DM& z, DM& Gout, std::vector<double>& jv, bool want_jac) {
            const int M = NODES;
            if (z.numel() != M * 19) throw std::invalid_argument("Collocated::eval: z must have NODES*(15+4) elements");
            double* z_d = buf; double* G_d = z_d + M * 19; double* JV_d = G_d + M * 15; double* gn_d = JV_d + M * 285;
            Ctx->h2d(z_d, z.ptr(), (size_t)M * 19);
            if (want_jac) Ctx->check(kite_colloc_eval_sparse(Ctx->ctx, 1, 1, M, compD_rm.data(), tau_, sx, su, z_d, nullptr, G_d, JV_d, gn_d), "kite_colloc_eval_sparse");
            else Ctx->check(kite_colloc_eval(Ctx->ctx, 1, 1, M, compD_rm.data(), tau_, sx, su, z_d, nullptr, G_d, nullptr, nullptr, gn_d), "kite_colloc_eval");
            Gout = DM(M * 15, 1);
            Ctx->d2h(Gout.ptr(), G_d, (size_t)M * 15);
            if (!want_jac) return;
            jv.resize((size_t)M * nnz_);
            Ctx->d2h(jv.data(), JV_d, jv.size());
        }
        std::shared_ptr<KiteContext> Ctx;
        std::vector<double> compD_rm;
        double tau_;
        DM compd_;
        double sx[15], su[4];
        double* buf = nullptr;
        int nnz_ = 0;
        std::vector<int> prow_, pcol_;
    };

    /** chebyshev.hpp:241-271 with the scaled augmented kite ODE of kiteNMPF.cpp:100-107 as `dynamics`. */
    std::shared_ptr<Collocated> CollocateDynamics(KiteDynamics& kite, const DM& ScaleX, const DM& ScaleU, const double& t0, const double& tf) {
        static_assert(NX == 15 && NU == 4 && NP == 0, "the GPU collocation evaluator implements the NMPC's augmented kite model (15 states, 4 controls)");
        const double t_scale = (tf - t0) / (2 * NumSegments);
        return std::make_shared<Collocated>(kite.context(), _CompDiff, t_scale, ScaleX, ScaleU);
    }

    /** Collocated performance index of the path-following NMPC (chebyshev.hpp:280-333 applied to the Lagrange and Mayer
     *  terms of kiteNMPF.cpp:116-143).  The reference takes the two terms as casadi::Function objects; here they are
     *  described by their parameters (weights, scaled reference velocity, circular path) and evaluated on the GPU. */
    class CollocatedCost {
    public:
        CollocatedCost(std::shared_ptr<KiteContext> ctx, const DM& qw, double tau, const DM& Sx, const kite_nmpc_cost& c)
            : Ctx(ctx), tau_(tau), cost_(c) {
            for (int i = 0; i <= PolyOrder; ++i) qw_[i] = qw.size1() > 1 ? qw[i] : qw(0, i);
            for (int i = 0; i < 15; ++i) sx[i] = Sx.size2() > 1 ? Sx(i, i) : Sx[i];
            if (kite_ctx_malloc(Ctx->ctx, (void**)&buf, sizeof(double) * ((size_t)NODES * 38 + 1)) != 0) throw std::runtime_error("CollocatedCost: device allocation failed");
        }
        ~CollocatedCost() { if (buf) kite_ctx_free(Ctx->ctx, buf); }
        CollocatedCost(const CollocatedCost&) = delete;
        /** performance_idx(z), as PerformanceIndex / nlp_f (kiteNMPF.cpp:143,152) */
        double operator()(const DM& z) { DM g; return eval(z, g, false); }
        /** value and gradient with respect to z = [X ; U] */
        double eval(const DM& z, DM& grad, bool want_grad = true) {
            const int M = NODES;
            if (z.numel() != M * 19) throw std::invalid_argument("CollocatedCost::eval: z must have NODES*(15+4) elements");
            double* z_d = buf; double* g_d = z_d + M * 19; double* c_d = g_d + M * 19;
            Ctx->h2d(z_d, z.ptr(), (size_t)M * 19);
            Ctx->check(kite_colloc_cost(Ctx->ctx, 1, 1, PolyOrder, NumSegments, qw_, tau_, sx, &cost_, z_d, c_d, want_grad ? g_d : nullptr), "kite_colloc_cost");
            double c = 0.0;
            Ctx->d2h(&c, c_d, 1);
            if (want_grad) { grad = DM(M * 19, 1); Ctx->d2h(grad.ptr(), g_d, (size_t)M * 19); }
            return c;
        }
        /** batched device entry point: see kite_colloc_cost in include/kite_b200.h */
        void eval_device(long B, const double* z_d, double* cost_d, double* grad_d) {
            Ctx->check(kite_colloc_cost(Ctx->ctx, B, B, PolyOrder, NumSegments, qw_, tau_, sx, &cost_, z_d, cost_d, grad_d), "kite_colloc_cost");
        }

    private:
        std::shared_ptr<KiteContext> Ctx;
        double tau_;
        kite_nmpc_cost cost_;
        double qw_[PolyOrder + 1], sx[15];
        double* buf = nullptr;
    };
    std::shared_ptr<CollocatedCost> CollocateCost(KiteDynamics& kite, const kite_nmpc_cost& terms, const DM& ScaleX, const double& t0, const double& tf) {
        static_assert(NX == 15 && NU == 4 && NP == 0, "the GPU cost evaluator implements the NMPC's augmented kite model (15 states, 4 controls)");
        return std::make_shared<CollocatedCost>(kite.context(), _QuadWeights, (tf - t0) / (2 * NumSegments), ScaleX, terms);
    }
    /** Lagrange / Mayer weights and reference velocity with the defaults of KiteNMPF (kiteNMPF.cpp:32-34,42-43; kiteNMPF.h:34). */
    static kite_nmpc_cost DefaultCost(const DM& ScaleX, double vel_ref, double radius, double altitude, const DM& q_rot) {
        kite_nmpc_cost c;
        c.Q[0] = 1e2 * 1e1; c.Q[1] = 1e2 * 1e1; c.Q[2] = 1e2 * 1e2;
        c.R[0] = 1e-4; c.R[1] = 1e-1; c.R[2] = 1e-1; c.R[3] = 1e-3;
        c.W = 1e-3;
        c.vref_scaled = (ScaleX.size2() > 1 ? ScaleX(14, 14) : ScaleX[14]) * vel_ref;
        c.path_radius = radius; c.path_altitude = altitude;
        for (int i = 0; i < 4; ++i) c.path_q[i] = q_rot[i];
        return c;
    }

private:
    /** Chebyshev-Gauss-Lobatto points x_k = cos(k pi / P) (chebyshev.hpp:119-127); index 0 is the FINAL time. */
    static DM CollocPoints() { DM X(PolyOrder + 1, 1); for (int k = 0; k <= PolyOrder; ++k) X[k] = std::cos(k * (M_PI / PolyOrder)); return X; }
    /** Trefethen's differentiation matrix with the negative-sum diagonal (chebyshev.hpp:136-153). */
    static DM DiffMatrix() {
        const int n = PolyOrder + 1;
        DM x = CollocPoints(), c(n, 1), Dn(n, n), Dm(n, n);
        for (int k = 0; k < n; ++k) c[k] = ((k % 2) ? -1.0 : 1.0) * ((k == 0 || k == PolyOrder) ? 2.0 : 1.0);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Dn(i, j) = (c[i] * (1.0 / c[j])) / ((x[i] - x[j]) + (i == j ? 1.0 : 0.0));
        for (int i = 0; i < n; ++i) { double rs = 0; for (int j = 0; j < n; ++j) rs += Dn(i, j); for (int j = 0; j < n; ++j) Dm(i, j) = Dn(i, j) - (i == j ? rs : 0.0); }
        return Dm;
    }
    /** Clenshaw-Curtis weights (chebyshev.hpp:162-195), returned as a 1 x (P+1) row like the reference. */
    static DM QuadWeights() {
        const int P = PolyOrder;
        DM w(1, P + 1);
        std::vector<double> v(P - 1, 1.0), theta(P + 1);
        for (int k = 0; k <= P; ++k) theta[k] = k * (M_PI / P);
        if (P % 2 == 0) {
            w(0, 0) = 1.0 / (double(P) * P - 1); w(0, P) = w(0, 0);
            for (int k = 1; k <= P / 2 - 1; ++k) for (int i = 1; i < P; ++i) v[i - 1] -= 2 * std::cos(2 * k * theta[i]) / (4.0 * k * k - 1);
            for (int i = 1; i < P; ++i) v[i - 1] -= std::cos(P * theta[i]) / (double(P) * P - 1);
        } else {
            w(0, 0) = 1.0 / (double(P) * P); w(0, P) = w(0, 0);
            for (int k = 1; k <= (P - 1) / 2; ++k) for (int i = 1; i < P; ++i) v[i - 1] -= 2 * std::cos(2 * k * theta[i]) / (4.0 * k * k - 1);
        }
        for (int i = 1; i < P; ++i) w(0, i) = 2 * v[i - 1] / P;
        return w;
    }
    /** composite block matrix (chebyshev.hpp:204-229): last segment gets the full D, earlier ones its first P rows */
    static DM CompDiffBlock() {
        const int m = NODES, n = PolyOrder + 1;
        DM Dm = DiffMatrix();
        if (NumSegments < 2) return Dm;
        DM C(m, m);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) C(m - n + i, m - n + j) = Dm(i, j);
        for (int k = 0; k < (NumSegments - 1) * PolyOrder; k += PolyOrder)
            for (int i = 0; i < PolyOrder; ++i) for (int j = 0; j < n; ++j) C(k + i, k + j) = Dm(i, j);
        return C;
    }
    DM _D, _CompDiff, _Points, _QuadWeights;
};

}  // namespace openkite
