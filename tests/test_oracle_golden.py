"""CPU-only: pins the C++ dual-number oracle (oracle/kite_oracle.hpp) to the independent sympy/mpmath
goldens (tests/golden/golden.json, made by scripts/make_golden.py) and to SURVEY.md Appendix A."""
import numpy as np
import pytest

from conftest import assert_close
from oracle.oracle_py import KITE, KITE_ID, RIGID_BODY, Oracle, params_from_yaml

TOL = 2e-12   # two independent float64-vs-50-digit restatements


def test_rhs_and_jacobians(oracle, golden):
    for name, c in golden["rhs"].items():
        f = oracle.rhs(c["x"], c["u"])[0]
        Jx, Ju = oracle.jac(c["x"], c["u"])
        assert_close(f, c["f"], TOL, what=f"f[{name}]")
        assert_close(Jx[0], c["Jx"], TOL, what=f"Jx[{name}]")
        assert_close(Ju[0], c["Ju"], TOL, what=f"Ju[{name}]")


def test_aero_function(oracle, golden):
    """Function "Aero"(x, u) -> Faero_b (kite.cpp:224-234, :330; KiteDynamics::getAeroDynamicForces, kite.h:126)."""
    for name, c in golden["rhs"].items():
        assert_close(oracle.aero(c["x"], c["u"])[0], c["aero"], TOL, what=f"aero[{name}]")
    for name, c in golden["rhs_id"].items():
        assert_close(oracle.aero(c["x"], c["u"], c["p"], kind=KITE_ID)[0], c["aero"], TOL, what=f"aero id[{name}]")
    # no thrust, tether or gravity in it: the force is invariant under T, position and attitude
    c = golden["rhs"]["model_test"]
    x = np.array(c["x"]); x2 = x.copy(); x2[6:9] += 1.0; x2[9:13] = [1.0, 0.0, 0.0, 0.0]
    assert np.array_equal(oracle.aero(x, [0.1, 0.02, 0.01]), oracle.aero(x2, [0.3, 0.02, 0.01]))


def test_jacobian_sparsity_matches_survey(oracle, golden):
    # 104 + 7 structural non-zeros when the tether arm is zero
    c = golden["rhs"]["model_test_u"]
    Jx, Ju = oracle.jac(c["x"], c["u"])
    assert int((Jx[0] != 0).sum()) == 104
    assert int((Ju[0] != 0).sum()) == 7


def test_rk4_step_and_sensitivities(oracle, golden):
    for name, c in golden["rk4_step"].items():
        xn, Phi, Gam = oracle.rk4_sens(c["x"], c["u"], c["h"])
        assert_close(xn[0], c["xn"], TOL, what=f"xn[{name}]")
        assert_close(Phi[0], c["Phi"], 1e-11, what=f"Phi[{name}]")
        assert_close(Gam[0], c["Gamma"], 1e-11, what=f"Gamma[{name}]")


def test_config1_rollout(oracle, golden):
    """BASELINE.json configs[0]: 10 s at 1 ms, open loop."""
    c = golden["rollout_config1"]
    xf, traj = oracle.rollout(c["x0"], c["u"], 10000, c["h"], want_traj=True)
    for k, ref in c["states_after"].items():
        assert_close(traj[0, int(k)], ref, 1e-11, what=f"state after {k} steps")
    # survey-time independent value
    survey = [2.7030845834920783, -0.13244357494175155, 0.90667392034009757, 0.0039031400464208724,
              -0.76289756833608439, -0.35559584048370014, 2.2435320190344083, -2.6743848407577450,
              1.3030254843563795, -0.34692557553911291, -0.36780711330237016, 0.39073452318968758,
              0.76921200252509402]
    assert_close(xf[0], survey, 1e-11, what="Appendix A rollout")


def test_survey_appendix_rhs(oracle):
    x = [1.5, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0]
    f = oracle.rhs(x, [0.1, 0, 0])[0]
    ref = [2.230498149273e+00, 1.946105464361e-02, 9.593304974432e+00, 0, 8.894917655786e-02, 0, 1.5, 0, 0, 0, 0, 0, 0]
    assert_close(f, ref, 1e-12, what="Appendix A rhs")


def test_ekf_predict(oracle, golden):
    c = golden["ekf_predict"]
    xn, Pn = oracle.ekf_predict(c["x"], c["u"], c["dt"], np.array(c["P"])[None], c["W"])
    assert_close(xn[0], c["xn"], TOL, what="ekf xn")
    assert_close(Pn[0], c["Pn"], 1e-11, what="ekf Pn")
    assert abs(np.trace(Pn[0]) - 15.45098842411765) < 1e-11
    W, V = oracle.ekf_defaults()
    assert_close(W, c["W"], 1e-15, what="default W")


def test_ekf_update_is_kalman(oracle, golden):
    """Update step (kiteEKF.cpp:108-126) against a numpy restatement."""
    c = golden["ekf_predict"]
    W, V = oracle.ekf_defaults()
    x = np.array(c["xn"]); P = np.array(c["Pn"])
    z = np.array([1.4522, -3.1274, -1.7034, -0.5455, -0.2382, -0.2922, -0.7485])   # kite_control_test.cpp:51
    H = np.hstack([np.zeros((7, 6)), np.eye(7)])
    K = P @ H.T @ np.linalg.inv(H @ P @ H.T + V)
    x_ref = x + K @ (z - H @ x)
    P_ref = (np.eye(13) - K @ H) @ P
    xo, Po = oracle.ekf_update(z, V, x, P[None])
    assert_close(xo[0], x_ref, 1e-10, what="ekf update x")
    assert_close(Po[0], P_ref, 1e-9, scale=1e-3, what="ekf update P")


@pytest.mark.parametrize("case", ["colloc_generics_P10_S1", "colloc_nmpc_P5_S2_scaled"])
def test_collocation(oracle, golden, case):
    c = golden[case]
    G, JX, JU = oracle.colloc_eval(c["z"], c["P"], c["S"], c["t0"], c["tf"], c["sx"], c["su"])
    assert_close(G[0], c["G"], 1e-11, what="G")
    assert_close(JX[0], c["JX"], 1e-11, what="JX")
    assert_close(JU[0], c["JU"], 1e-11, what="JU")


def test_chebyshev_operators(oracle, golden):
    for name, c in golden["cheb"].items():
        P, S = c["P"], c["S"]
        assert_close(oracle.cheb_points(P), c["points"], 1e-14, what="points")
        assert_close(oracle.cheb_diff(P), c["D"], 1e-13, what="D")
        assert_close(oracle.cheb_weights(P), c["weights"], 1e-14, what="weights")
        assert_close(oracle.cheb_compdiff(P, S), c["compD"], 1e-13, what="compD")
        # closed forms: D00 = (2P^2+1)/6, Clenshaw-Curtis weights sum to 2
        assert abs(oracle.cheb_diff(P)[0, 0] - (2 * P * P + 1) / 6.0) < 1e-12
        assert abs(oracle.cheb_weights(P).sum() - 2.0) < 1e-14


def test_compdiff_differentiates_quadratic(oracle):
    P, S, tf = 5, 2, 1.0
    C = oracle.cheb_compdiff(P, S)
    tau = tf / (2 * S)
    xs = oracle.cheb_points(P)
    # node times, node 0 = final time
    t = np.concatenate([(xs[:-1] + 1) * tau + tau * 2 * (S - 1 - s) for s in range(S)] + [[0.0]])
    assert np.abs(C @ t ** 2 - tau * 2 * t).max() < 5e-15


def test_tether_arm(yaml_path, golden):
    c = golden["tether_arm"]
    prm = params_from_yaml(yaml_path)
    prm[36:39] = c["tether_arm"]
    o = Oracle(prm)
    assert_close(o.rhs(c["x"], c["u"])[0], c["f"], TOL, what="f arm")
    Jx, Ju = o.jac(c["x"], c["u"])
    assert_close(Jx[0], c["Jx"], TOL, what="Jx arm")
    assert int((Jx[0] != 0).sum()) == 104 + 21
    xn, Phi, Gam = o.rk4_sens(c["x"], c["u"], c["h"])
    assert_close(Phi[0], c["Phi"], 1e-11, what="Phi arm")


def test_identification_variant(oracle, golden):
    for name, c in golden["rhs_id"].items():
        f = oracle.rhs(c["x"], c["u"], p=c["p"], kind=KITE_ID)[0]
        Jx, Ju = oracle.jac(c["x"], c["u"], p=c["p"], kind=KITE_ID)
        assert_close(f, c["f"], TOL, what=f"id f[{name}]")
        assert_close(Jx[0], c["Jx"], TOL, what=f"id Jx[{name}]")
        xf = oracle.rollout(c["x"], c["u"], 1, c["h"], p=c["p"], kind=KITE_ID)
        assert_close(xf[0], c["xn"], TOL, what=f"id xn[{name}]")


def test_rigid_body(oracle, golden):
    c = golden["rigid_body"]
    f = oracle.rhs(c["x"], c["u"], kind=RIGID_BODY)[0]
    assert_close(f, c["f"], TOL, what="rb f")
    Jx, _ = oracle.jac(c["x"], c["u"], kind=RIGID_BODY)
    assert_close(Jx[0], c["Jx"], TOL, what="rb Jx")
    xn, Phi, _ = oracle.rk4_sens(c["x"], c["u"], c["h"], kind=RIGID_BODY)
    assert_close(xn[0], c["xn"], TOL, what="rb xn")
    assert_close(Phi[0], c["Phi"], 1e-11, what="rb Phi")


def test_invariants(oracle):
    """Closed-form checks that need no oracle."""
    rng = np.random.default_rng(0)
    x = oracle.synth_x0(0, 64)
    u = rng.uniform(-0.1, 0.1, (64, 3))
    f = oracle.rhs(x, u)
    q = x[:, 9:13]
    assert np.abs((f[:, 9:13] * q).sum(1)).max() < 1e-14            # qdot . q = 0 on the unit sphere
    # lateral symmetry: v1 = w0 = w2 = dR = 0 and q a pure pitch rotation keep lateral components zero
    xs = np.array([5.0, 0, 0.4, 0, 0.3, 0, 1.0, 0, -2.5, np.cos(0.2), 0, np.sin(0.2), 0])
    fs = oracle.rhs(xs, [0.1, 0.02, 0.0])[0]
    assert abs(fs[1]) < 1e-14 and abs(fs[3]) < 1e-13 and abs(fs[5]) < 1e-13 and abs(fs[7]) < 1e-14


def test_flop_counts_near_survey(oracle):
    fc = oracle.flop_counts()
    assert abs(fc["rhs"]["flops"] - 435) / 435 < 0.10          # SURVEY.md 8d figure
    assert abs(fc["rk4_step"]["flops"] - 1909) / 1909 < 0.10
    assert fc["rhs"]["special"] == 9


def test_counter_rng_is_sharding_invariant(oracle):
    a = oracle.synth_controls(0, 8, 5)
    b = oracle.synth_controls(4, 4, 5)
    assert np.array_equal(a[4:], b)
    assert a[..., 0].min() >= 0 and a[..., 0].max() <= 0.3
    x0 = oracle.synth_x0(0, 16)
    assert np.abs(np.linalg.norm(x0[:, 9:13], axis=1) - 1).max() < 1e-15


def test_python_collocation_operators_match_oracle(oracle):
    """openkite_b200/collocation.py (product-side host helper) against the oracle's operators and closed forms."""
    from openkite_b200.collocation import colloc_points, comp_diff_matrix, diff_matrix, quad_weights
    for P, S in ((5, 2), (2, 3), (10, 1), (4, 3)):
        assert_close(colloc_points(P), oracle.cheb_points(P), 1e-15, what="points")
        assert_close(diff_matrix(P), oracle.cheb_diff(P), 1e-13, what="D")
        assert_close(quad_weights(P), np.ravel(oracle.cheb_weights(P)), 1e-13, what="weights")
        assert_close(comp_diff_matrix(P, S), oracle.cheb_compdiff(P, S), 1e-13, what="compD")
    assert abs(diff_matrix(5)[0, 0] - 8.5) < 1e-13 and abs(quad_weights(5).sum() - 2.0) < 1e-14


def _nmpc_case(golden):
    c = golden["colloc_nmpc_P5_S2_scaled"]
    q = (np.cos(np.pi / 8), 0.0, np.sin(np.pi / 8), 0.0)          # nmpf_node.cpp:35
    return np.array(c["z"]), np.array(c["sx"]), q


def nmpc_cost_closed_form(z, sx, q, P=5, S=2, tf=1.0, vel_ref=0.05, radius=2.65, alt=0.0):
    """Independent numpy restatement of the NMPC performance index (rotation-matrix form of the tilted circle)."""
    from openkite_b200.collocation import quad_weights
    M = S * P + 1
    X, U = z[:M * 15].reshape(M, 15), z[M * 15:].reshape(M, 4)
    s0, v = q[0], np.array(q[1:])
    Rm = (s0 * s0 - v @ v) * np.eye(3) + 2 * np.outer(v, v) + 2 * s0 * np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    Q = np.array([1e3, 1e3, 1e4]); R = np.array([1e-4, 1e-1, 1e-1, 1e-3]); W = 1e-3
    def pathcost(x):
        th = x[13] / sx[13]
        pp = Rm.T @ np.array([radius * np.cos(th), radius * np.sin(th), alt])
        r = sx[6:9] * pp - x[6:9]
        return float(Q @ (r * r))
    w, tau = quad_weights(P), tf / (2 * S)
    cost = pathcost(X[0])
    for k in range(S):
        for m in range(P + 1):
            n = k * P + m
            cost += tau * w[m] * (pathcost(X[n]) + W * (sx[14] * vel_ref - X[n, 14]) ** 2 + float(R @ (U[n] ** 2)))
    return cost


def test_nmpc_cost_and_gradient(oracle, golden):
    """chebyshev.hpp:280-333 + kiteNMPF.cpp:116-143: oracle cost vs an independent closed form, gradient vs central differences."""
    z, sx, q = _nmpc_case(golden)
    cc = oracle.nmpc_cost_params(sx, q_rot=q)
    cost, g = oracle.colloc_cost(z, 5, 2, 0.0, 1.0, sx, cc)
    assert abs(cost[0] - nmpc_cost_closed_form(z, sx, q)) <= 1e-12 * abs(cost[0])
    rng = np.random.default_rng(5)
    for i in rng.choice(209, 40, replace=False):
        e = np.zeros(209); e[i] = 1e-6
        fd = (nmpc_cost_closed_form(z + e, sx, q) - nmpc_cost_closed_form(z - e, sx, q)) / 2e-6
        assert abs(fd - g[0, i]) <= 1e-6 * max(1.0, abs(g[0, i])), i
    # structure: only r (6..8), theta (13), theta_dot (14) and the controls carry gradient
    gx = g[0, :165].reshape(11, 15)
    assert np.all(gx[:, [0, 1, 2, 3, 4, 5, 9, 10, 11, 12]] == 0.0)


def test_flops_header_matches_sources():
    """openkite_b200/csrc/kite_flops.h (the roofline numerators bench.py reads) is generated by scripts/make_flops.py:
    the op-counted figures must follow from the CURRENT oracle and device sources (the sympy CSE count is regenerated by
    the script itself; here only its frozen value is sanity-checked against the survey's figure)."""
    import importlib.util
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_flops", os.path.join(ROOT, "scripts", "make_flops.py"))
    mf = importlib.util.module_from_spec(spec); spec.loader.exec_module(mf)
    hdr = mf.read_header()
    dev = mf.device_counts()
    orc = mf.oracle_counts()
    assert hdr["ORACLE_RHS"] == orc["rhs"]["flops"] and hdr["ORACLE_RK4_STEP"] == orc["rk4_step"]["flops"]
    assert hdr["DEVICE_RHS"] == dev["rhs"]["flops"] and hdr["DEVICE_RK4_STEP"] == dev["rk4_step"]["flops"]
    assert hdr["DEVICE_RHS_JAC"] == dev["rhs_jac"]["flops"]
    assert dev["rhs_jac"]["nx"] == hdr["NNZ_JX"] == 104 and dev["rhs_jac"]["nu"] == hdr["NNZ_JU"] == 7
    assert abs(hdr["ORACLE_RHS_JAC"] - 2690) / 2690 < 0.15            # SURVEY.md 8d figure, own CSE count within 15 %
    assert abs(hdr["ORACLE_RK4_STEP"] - 1909) / 1909 < 0.02
    assert hdr["DEVICE_RK4_SENS_STEP"] < hdr["ORACLE_RK4_SENS_STEP"] < hdr["ORACLE_RK4_SENS_STEP_SURVEY"]
