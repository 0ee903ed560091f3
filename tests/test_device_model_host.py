"""CPU-only: compiles the DEVICE model source (openkite_b200/csrc/kite_model.cuh) for the host through a stub
cuda_runtime.h (tests/cpu_shim) and checks its hand-derived arithmetic -- attitude-matrix RHS, analytic Jacobian
blocks, register RK4 -- against the goldens and the oracle.  Catches model bugs without a GPU round trip."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import assert_close

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "cpu_shim")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def shim():
    so = os.path.join(SHIM, "libmodel_host_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", SHIM, "-o", so,
                           os.path.join(SHIM, "model_host_check.cpp")])
    return C.CDLL(so)


def P(a):
    return a.ctypes.data_as(dp)


def shim_eval(shim, prm, kind, x, u, p=None):
    x = np.array(x, float); u = np.array(u, float)
    f = np.zeros(13); Jx = np.zeros((13, 13)); Ju = np.zeros((13, 3))
    pp = None if p is None else P(np.array(p, float))
    shim.shim_eval(P(prm), kind, P(x), P(u), pp, P(f), P(Jx), P(Ju))
    return f, Jx, Ju


@pytest.fixture(scope="module")
def prm(yaml_path):
    from oracle.oracle_py import params_from_yaml
    return params_from_yaml(yaml_path)      # same 39-double order as struct kite_params


def test_table_forms_of_special_functions(shim, prm, golden):
    """asin_red / table exponential (the kernels with per-trajectory coefficients) against the goldens."""
    for n, c in list(golden["rhs"].items()) + list(golden["rhs_id"].items()):
        x = np.array(c["x"], float); u = np.array(c["u"], float)
        f = np.zeros(13); Jx = np.zeros((13, 13)); Ju = np.zeros((13, 3))
        kind = 1 if "p" in c else 0
        pp = P(np.array(c["p"], float)) if "p" in c else None
        shim.shim_eval_tab(P(prm), kind, P(x), P(u), pp, P(f), P(Jx), P(Ju))
        assert_close(f, c["f"], 1e-12, what=f"f tab[{n}]")
        assert_close(Jx, c["Jx"], 1e-12, what=f"Jx tab[{n}]")


def test_device_model_vs_golden(shim, prm, golden):
    for n, c in golden["rhs"].items():
        f, Jx, Ju = shim_eval(shim, prm, 0, c["x"], c["u"])
        assert_close(f, c["f"], 1e-12, what=f"f[{n}]")
        assert_close(Jx, c["Jx"], 1e-12, what=f"Jx[{n}]")
        assert_close(Ju, c["Ju"], 1e-12, what=f"Ju[{n}]")
    c = golden["tether_arm"]
    p2 = prm.copy(); p2[36:39] = c["tether_arm"]
    f, Jx, Ju = shim_eval(shim, p2, 0, c["x"], c["u"])
    assert_close(f, c["f"], 1e-12, what="f arm"); assert_close(Jx, c["Jx"], 1e-12, what="Jx arm")
    assert int((Jx != 0).sum()) == 125
    for n, c in golden["rhs_id"].items():
        f, Jx, Ju = shim_eval(shim, prm, 1, c["x"], c["u"], c["p"])
        assert_close(f, c["f"], 1e-12, what="id f"); assert_close(Jx, c["Jx"], 1e-12, what="id Jx")
    c = golden["rigid_body"]
    f, Jx, Ju = shim_eval(shim, prm, 2, c["x"], c["u"])
    assert_close(f, c["f"], 1e-12, what="rb f"); assert_close(Jx, c["Jx"], 1e-12, what="rb Jx")


def test_device_model_vs_oracle_random(shim, prm, oracle):
    x = oracle.synth_x0(0, 200); u = oracle.synth_controls(0, 200, 1)[:, 0, :]
    rf = oracle.rhs(x, u); rJx, rJu = oracle.jac(x, u)
    for i in range(200):
        f, Jx, Ju = shim_eval(shim, prm, 0, x[i], u[i])
        assert_close(f, rf[i], 1e-12, what="f"); assert_close(Jx, rJx[i], 1e-11, what="Jx"); assert_close(Ju, rJu[i], 1e-12, what="Ju")


def test_device_model_all_quadrants(shim, prm, oracle):
    """Branch-free angle evaluation (kite_math.cuh asin_sc / atan2_sc): large sideslip, post-stall angles of attack,
    backwards flight in both rear quadrants, pure sideways / vertical flight, and rest."""
    x = oracle.synth_x0(0, 12)
    vs = [[1.0, 0.2, 4.0], [-3.0, 0.1, 0.5], [-3.0, 0.1, -0.5], [2.0, 3.0, 0.1], [2.0, -3.0, 0.1], [0.0, 0.0, 0.0],
          [0.3, 5.0, 0.2], [1e-3, 0.0, 4.0], [-2.0, 0.0, 0.0], [4.0, 2.9, 2.9], [-1.0, -2.0, -3.0], [5.0, 0.0, -4.9]]
    for i, v in enumerate(vs):
        x[i, 0:3] = v
    u = oracle.synth_controls(0, 12, 1)[:, 0, :]
    rf = oracle.rhs(x, u); rJx, rJu = oracle.jac(x, u)
    for i in range(12):
        f, Jx, Ju = shim_eval(shim, prm, 0, x[i], u[i])
        assert_close(f, rf[i], 1e-12, what="f[%d]" % i)
        if i != 5:     # at rest the reference Jacobian itself is 0/0 in places
            assert_close(Jx, rJx[i], 1e-10, what="Jx[%d]" % i); assert_close(Ju, rJu[i], 1e-12, what="Ju[%d]" % i)


def test_device_rk4_config1(shim, prm, golden):
    c = golden["rollout_config1"]
    xn = np.zeros(13)
    shim.shim_rk4(P(prm), 0, P(np.array(c["x0"])), P(np.array(c["u"])), None, C.c_double(c["h"]), C.c_long(10000), P(xn))
    assert_close(xn, c["states_after"]["10000"], 1e-11, what="10 s rollout")
