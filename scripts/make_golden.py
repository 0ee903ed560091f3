#!/usr/bin/env python
"""Generate tests/golden/golden.json from the sympy/mpmath oracle (oracle/sympy_oracle.py).

The INPUTS are the fixed points the reference's own tests / launch files use (cited per case); the
OUTPUTS are 50-digit mpmath evaluations of the sympy restatement rounded to float64.  The reference's
tests assert no outputs (all BOOST_CHECK(true)), CasADi is not installable here, so these vectors are the
repo's own pin ("parity unpinned" by the reference -- see DESIGN.md section 3).

Run:  python scripts/make_golden.py        (takes a few minutes; writes tests/golden/golden.json)
"""
import json
import os
import sys
import time

import mpmath as mp
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sympy_oracle as so  # noqa: E402


def fl(v):
    if isinstance(v, (list, tuple)):
        return [fl(t) for t in v]
    if isinstance(v, mp.matrix):
        return [[float(v[i, j]) for j in range(v.cols)] for i in range(v.rows)]
    return float(v)


# ---- inputs pinned by the reference ---------------------------------------------------------
X_MODEL_TEST = [6.1977743e+00, -2.8407148e-02, 9.1815942e-01, 2.9763089e-01, -2.2052198e+00, -1.4827499e-01,
                -4.1624807e-01, -2.2601052e+00, 1.2903439e+00, 3.5646195e-02, -6.9986094e-02, 8.2660637e-01,
                5.5727089e-01]                                   # kite_model_test.cpp:58-59
U_MODEL_TEST = [0.1, 0.0, 0.0]                                   # kite_model_test.cpp:60
X_CONTROL_TEST = [1.5, 0, 0, 0, 0, 0, 0, 1.0, 0, 1, 0, 0.0, 0.0]  # kite_control_test.cpp:252
X_LAUNCH_SIM = [4.4, 0.44, 1.73, 0.81, -1.73, -1.53, -0.46, -2.68, 0.64, -0.0289, 0.1587, 0.4304, 0.8881]  # launch/simulator.launch:3
X_LAUNCH_HIL = [1.5, 0, 0, 0, 0, 0, -3, 0, -2, 0.7071, 0, 0, 0.7071]   # launch/hw_in_the_loop.launch:3
X_EKF_TEST = [6.0026, -0.3965, 0.1705, 0.4414, -0.2068, 0.9293, 1.4634, -3.1765, -1.7037, -0.5486, -0.2354, -0.2922,
              -0.7471]                                           # kite_control_test.cpp:52-55
X_RIGID_TEST = [4.318732, 0.182552, 0.254833, 1.85435, -0.142882, -0.168359, -0.229383, -0.0500282, -0.746832,
                0.189409, -0.836349, -0.48178, 0.180367]         # kite_control_test.cpp:28-29
X_MODEL_TEST_ALT = [0.318732, 0.182552, 0.254833, 1.85435, -0.142882, -0.168359, -0.229383, -0.0500282, -0.746832,
                    0.189409, -0.836349, -0.48178, 0.180367]     # kite_model_test.cpp:54-55 (commented alt state)
U_MODEL_TEST_ALT = [0.3, 5 * 3.141592653589793 / 180, -2 * 3.141592653589793 / 180]   # kite_model_test.cpp:56
EKF_DT = 0.0084                                                  # kite_control_test.cpp:49
# 209-vector at which full_generics_test evaluates AugJacobian (kite_control_test.cpp:582-598)
Z209 = [0.322159, -1.7086, 0.0187737, 0.571611, -0.463085, -3.98942, -1.00108, 2.67555, -1.67471, 0.685529, 0.179051, 0.601294,
        -0.234078, -8.93944, -6.15368, 0.320558, -1.81794, -0.048227, 0.697989, -0.487516, -3.89813, -0.978354, 2.70897, -1.6673, 0.693579,
        0.203702, 0.597764, -0.194959, -8.78899, -6.13793, 0.254425, -1.85413, 0.141461, 0.852417, -0.100141, -3.66057, -0.922693, 2.80918,
        -1.6289, 0.712096, 0.26544, 0.576425, -0.0820664, -8.3567, -6.02738, 7.80537e-09, -2.46967, 2.21734e-09, 1.44569, -0.464535,
        -2.70741, -0.862032, 2.98398, -1.51414, 0.725987, 0.318053, 0.529904, 0.0847142, -7.7014, -5.81885, 0.143081, -2.19291, 0.185359,
        1.49807, -1.06445, -1.31733, -0.826037, 3.2155, -1.30218, 0.705787, 0.300448, 0.509612, 0.259055, -6.9092, -5.54832, 0.455407,
        -1.48802, 0.925243, 0.939966, -0.431242, -0.183387, -0.971463, 3.33088, -1.10238, 0.683697, 0.242655, 0.507753, 0.363209, -6.07502,
        -5.24718, 1.30346, -1.11843, 0.985835, 0.52808, 0.239268, 0.253541, -1.13998, 3.37537, -0.955375, 0.696112, 0.198938, 0.49205,
        0.386145, -5.28823, -4.94104, 3.23749, -0.562678, 1.1487, 0.185406, 0.798954, 0.427541, -1.37814, 3.14841, -0.761349, 0.725335,
        0.182191, 0.46157, 0.378231, -4.61785, -4.66725, 4.98206, 0.581094, 1.0498, -0.361074, 1.36802, -0.142782, -1.43793, 2.9391,
        -0.593935, 0.758427, 0.200775, 0.420269, 0.347893, -4.11561, -4.44789, 5.96975, 0.969053, 0.86408, -0.503073, 1.45039, -0.880998,
        -1.73889, 2.39822, -0.451459, 0.761644, 0.241508, 0.383887, 0.382548, -3.79856, -4.30918, 6.17882, 1.08634, 0.852976, -0.562302,
        1.54206, -1.17198, -0.95457, 2.91599, -0.201047, 0.867553, 0.374954, 0.230215, 0.231888, -3.79856, -4.30918, 0.121739, 0.136834,
        0.136834, -0.165329, 0.101, 0.116373, -0.136834, -1.03451, 0.199, -0.0223408, 0.0984078, -1.81909, 0.129195, -0.136834, 0.136834,
        -1.91874, 0.101, -0.135728, 0.0148018, -1.94518, 0.101, -0.106827, 0.0835699, -1.96381, 0.122227, -0.136834, -0.115507, -1.9685,
        0.101, -0.136834, -0.136834, -1.96848, 0.199, -0.135985, -0.128318, -1.9641, 0.101, -0.136834, 0.00392875, -1.93056, 0.199,
        -0.136834, 0.0482633, -1.83182]
# NMPC scaling (nmpf_node.cpp:50-51)
SCALE_X = [0.1, 1 / 3.0, 1 / 3.0, 1 / 2.0, 1 / 5.0, 1 / 2.0, 1 / 3.0, 1 / 3.0, 1 / 3.0, 1.0, 1.0, 1.0, 1.0, 1 / 6.28, 1 / 6.28]
SCALE_U = [1 / 0.15, 1 / 0.2618, 1 / 0.2618, 1 / 5.0]
EKF_SIG = [0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.1, 0.1, 0.01, 0.05, 0.05, 0.05]   # kiteEKF.cpp:6-11


def add_aero_only():
    """Adds the output of Function "Aero" (kite.cpp:330) to the existing rhs cases without regenerating the rest."""
    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    out = json.load(open(path))
    cfg = yaml.safe_load(open(os.path.join(ROOT, "data", "umx_radian.yaml")))
    cfg.setdefault("tether", {})
    kite = so.SymModel(cfg, "kite")
    for name, c in out["cases"]["rhs"].items():
        c["aero"] = fl(kite.aero(c["x"], c["u"]))
    kid = so.SymModel(cfg, "kite_id")
    for name, c in out["cases"]["rhs_id"].items():
        c["aero"] = fl(kid.aero(c["x"], c["u"], c["p"]))
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("aero added to %d + %d cases" % (len(out["cases"]["rhs"]), len(out["cases"]["rhs_id"])))


def main():
    if "--aero-only" in sys.argv:
        return add_aero_only()
    t0 = time.time()
    cfg = yaml.safe_load(open(os.path.join(ROOT, "data", "umx_radian.yaml")))
    cfg.setdefault("tether", {})
    out = {"generator": "scripts/make_golden.py (sympy %s, mpmath dps=50)" % so.sp.__version__, "cases": {}}
    C = out["cases"]

    kite = so.SymModel(cfg, "kite")
    print("built kite model %.1fs" % (time.time() - t0)); sys.stdout.flush()

    # ---- RHS + Jacobians at reference-pinned points ----------------------------------------
    pts = {
        "model_test": (X_MODEL_TEST, U_MODEL_TEST),
        "control_test": (X_CONTROL_TEST, U_MODEL_TEST),
        "launch_sim": (X_LAUNCH_SIM, [0.0, 0.0, 0.0]),
        "launch_hil": (X_LAUNCH_HIL, [0.1, 0.05, -0.03]),
        "ekf_test": (X_EKF_TEST, [0.0, 0.0, 0.0]),
        "model_test_alt": (X_MODEL_TEST_ALT, U_MODEL_TEST_ALT),
        "model_test_u": (X_MODEL_TEST, [0.25, -0.11, 0.09]),
    }
    C["rhs"] = {}
    for name, (x, u) in pts.items():
        f = kite.f(x, u)
        Jx, Ju = kite.jac(x, u)
        C["rhs"][name] = dict(kind="kite", x=x, u=u, f=fl(f), Jx=fl(Jx), Ju=fl(Ju), aero=fl(kite.aero(x, u)))

    # ---- RK4 single steps with sensitivities -------------------------------------------------
    C["rk4_step"] = {}
    for name, (x, u, h) in {
        "model_test_h1ms": (X_MODEL_TEST, U_MODEL_TEST, 1e-3),
        "model_test_h20ms": (X_MODEL_TEST, [0.25, -0.11, 0.09], 0.02),
        "launch_sim_h100ms": (X_LAUNCH_SIM, [0.15, 0.05, -0.04], 0.1),
        "ekf_test_dt": (X_EKF_TEST, [0.0, 0.0, 0.0], EKF_DT),
    }.items():
        xn, Phi, Gam = kite.rk4_step_sens(x, u, h)
        C["rk4_step"][name] = dict(kind="kite", x=x, u=u, h=h, xn=fl(xn), Phi=fl(Phi), Gamma=fl(Gam))
    print("rhs/rk4 done %.1fs" % (time.time() - t0)); sys.stdout.flush()

    # ---- config 1: 10 s at 1 ms open loop (BASELINE.json configs[0]) ---------------------------
    x = [mp.mpf(t) for t in X_MODEL_TEST]
    marks = {}
    for k in range(1, 10001):
        x = kite.rk4_step(x, U_MODEL_TEST, 1e-3)
        if k in (1, 10, 100, 1000, 5000, 10000):
            marks[str(k)] = fl(x)
    C["rollout_config1"] = dict(kind="kite", x0=X_MODEL_TEST, u=U_MODEL_TEST, h=1e-3, states_after=marks)
    print("rollout done %.1fs" % (time.time() - t0)); sys.stdout.flush()

    # ---- EKF predict (kite_control_test.cpp:49-55 inputs, P = 10 W) ------------------------------
    W = [[(EKF_SIG[i] ** 2 if i == j else 0.0) for j in range(13)] for i in range(13)]
    P0 = [[10.0 * W[i][j] for j in range(13)] for i in range(13)]
    xn, Pn = so.ekf_predict(kite, X_EKF_TEST, [0.0, 0.0, 0.0], EKF_DT, P0, W)
    C["ekf_predict"] = dict(kind="kite", x=X_EKF_TEST, u=[0.0, 0.0, 0.0], dt=EKF_DT, P=P0, W=W, xn=fl(xn), Pn=fl(Pn))

    # ---- collocation: full_generics_test point (P=10, S=1, unscaled, tf=1) -----------------------
    one15, one4 = [1.0] * 15, [1.0] * 4
    G, JX, JU = so.colloc_eval(kite, Z209, 10, 1, 0.0, 1.0, one15, one4)
    C["colloc_generics_P10_S1"] = dict(kind="kite", P=10, S=1, t0=0.0, tf=1.0, sx=one15, su=one4, z=Z209, G=fl(G),
                                       JX=fl(JX), JU=fl(JU))
    # ---- collocation: NMPC config (P=5, S=2, nmpf_node scaling), same nodes scaled -----------------
    zs = []
    for k in range(11):
        zs += [SCALE_X[i] * Z209[k * 15 + i] for i in range(15)]
    for k in range(11):
        zs += [SCALE_U[i] * Z209[165 + k * 4 + i] for i in range(4)]
    G, JX, JU = so.colloc_eval(kite, zs, 5, 2, 0.0, 1.0, SCALE_X, SCALE_U)
    C["colloc_nmpc_P5_S2_scaled"] = dict(kind="kite", P=5, S=2, t0=0.0, tf=1.0, sx=SCALE_X, su=SCALE_U, z=zs, G=fl(G),
                                         JX=fl(JX), JU=fl(JU))
    print("colloc done %.1fs" % (time.time() - t0)); sys.stdout.flush()

    # ---- Chebyshev operators ----------------------------------------------------------------------
    C["cheb"] = {}
    for P, S in ((5, 2), (10, 1), (2, 3)):       # NMPC; full_generics_test; pseudo_test (kite_control_test.cpp:163)
        C["cheb"]["P%d_S%d" % (P, S)] = dict(P=P, S=S, points=fl(so.cheb_points(P)), D=fl(so.cheb_diff(P)),
                                             weights=fl(so.cheb_weights(P)), compD=fl(so.cheb_compdiff(P, S)))

    # ---- tether arm != 0 (exercises Mt = arm x R_b, kite.cpp:299-300) ------------------------------
    cfg2 = json.loads(json.dumps(cfg)); cfg2["tether"].update(rx=0.012, ry=-0.004, rz=0.021)
    kite_arm = so.SymModel(cfg2, "kite")
    f = kite_arm.f(X_MODEL_TEST, [0.25, -0.11, 0.09]); Jx, Ju = kite_arm.jac(X_MODEL_TEST, [0.25, -0.11, 0.09])
    xn, Phi, Gam = kite_arm.rk4_step_sens(X_MODEL_TEST, [0.25, -0.11, 0.09], 0.02)
    C["tether_arm"] = dict(kind="kite", tether_arm=[0.012, -0.004, 0.021], x=X_MODEL_TEST, u=[0.25, -0.11, 0.09], f=fl(f),
                           Jx=fl(Jx), Ju=fl(Ju), h=0.02, xn=fl(xn), Phi=fl(Phi), Gamma=fl(Gam))
    print("arm done %.1fs" % (time.time() - t0)); sys.stdout.flush()

    # ---- identification variant (no regularisers, 21 parameters) ----------------------------------
    kid = so.SymModel(cfg, "kite_id")
    aer = cfg["aerodynamic"]
    pnom = [float(aer[k]) for k in so.ID_PARAM_NAMES]
    ppert = [v * (1 + 0.07 * ((-1) ** i) * (1 + (i % 5) / 5.0)) for i, v in enumerate(pnom)]
    C["rhs_id"] = {}
    for name, (x, u, p) in {"nominal": (X_MODEL_TEST, [0.25, -0.11, 0.09], pnom),
                            "perturbed": (X_LAUNCH_SIM, [0.15, 0.05, -0.04], ppert)}.items():
        f = kid.f(x, u, p); Jx, Ju = kid.jac(x, u, p)
        xn = kid.rk4_step(x, u, 1e-3, p)
        C["rhs_id"][name] = dict(kind="kite_id", x=x, u=u, p=p, f=fl(f), Jx=fl(Jx), Ju=fl(Ju), h=1e-3, xn=fl(xn))

    # ---- rigid body (kite_control_test.cpp:28-29) -------------------------------------------------
    rb = so.SymModel(cfg, "rigid_body")
    f = rb.f(X_RIGID_TEST, [0, 0, 0]); Jx, Ju = rb.jac(X_RIGID_TEST, [0, 0, 0])
    xn, Phi, Gam = rb.rk4_step_sens(X_RIGID_TEST, [0, 0, 0], 0.02)
    C["rigid_body"] = dict(kind="rigid_body", x=X_RIGID_TEST, u=[0.0, 0.0, 0.0], f=fl(f), Jx=fl(Jx), Ju=fl(Ju), h=0.02,
                           xn=fl(xn), Phi=fl(Phi), Gamma=fl(Gam))

    path = os.path.join(ROOT, "tests", "golden", "golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path, "%.1fs" % (time.time() - t0))


if __name__ == "__main__":
    main()
