"""The C++ host mirror of the reference API (include/openkite/*.hpp): KiteProperties/LoadProperties, KiteDynamics,
ODESolver, KiteEKF, Chebyshev -- built with g++ against libkite_b200.so and run as tests/cpp/host_api_test.
CPU part: YAML loader, option dict, collocation operators.  GPU part: the reference's own test scenarios
(ode_solver_test, ekf_test, full_generics_test, ...) with expected values from tests/golden and the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_api_test")


@pytest.fixture(scope="module")
def host_bin():
    from openkite_b200 import build
    build.build()
    subprocess.check_call(["bash", os.path.join(ROOT, "tests", "cpp", "build_host_tests.sh")])
    return BIN


def test_host_cpu_only(host_bin, yaml_path):
    r = subprocess.run([host_bin, "--cpu-only", "--yaml", yaml_path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout


@pytest.mark.gpu
def test_host_api_on_gpu(host_bin, yaml_path, golden, oracle, tmp_path):
    lines = []

    def put(tag, v):
        v = np.asarray(v, dtype=np.float64).ravel()
        lines.append("%s %d %s" % (tag, v.size, " ".join(repr(float(t)) for t in v)))

    c = golden["rhs"]["model_test"]
    put("rhs_model_test", c["f"]); put("jx_model_test", c["Jx"]); put("aero", c["aero"])
    r = golden["rollout_config1"]["states_after"]
    put("config1_after_1", r["1"]); put("config1_after_1000", r["1000"])
    e = golden["ekf_predict"]
    put("ekf_xn", e["xn"]); put("ekf_Pn", e["Pn"])
    W, V = oracle.ekf_defaults()
    z = np.array([1.4522, -3.1274, -1.7034, -0.5455, -0.2382, -0.2922, -0.7485])
    xu, Pu = oracle.ekf_update(z, V, np.array(e["xn"]), np.array(e["Pn"])[None])
    put("ekf_est", xu[0])
    put("config1_after_1000b", r["1000"])
    rb = golden["rigid_body"]
    put("rb_f", rb["f"]); put("rb_xn", rb["xn"])
    cc = golden["colloc_nmpc_P5_S2_scaled"]
    M, tau = 11, 0.25
    compD = np.array(golden["cheb"]["P5_S2"]["compD"])
    J = np.zeros((165, 209))
    J[:, :165] = np.kron(compD, np.eye(15))
    JX, JU = np.array(cc["JX"]), np.array(cc["JU"])
    for k in range(M):
        J[k * 15:(k + 1) * 15, k * 15:(k + 1) * 15] -= tau * JX[k]
        J[k * 15:(k + 1) * 15, 165 + k * 4:165 + (k + 1) * 4] = -tau * JU[k]
    put("colloc_sx", cc["sx"]); put("colloc_su", cc["su"]); put("colloc_z", cc["z"]); put("colloc_G", cc["G"]); put("colloc_J", J)
    ncc = oracle.nmpc_cost_params(cc["sx"], q_rot=(np.cos(np.pi / 8), 0.0, np.sin(np.pi / 8), 0.0))
    ncost, ngrad = oracle.colloc_cost(np.array(cc["z"]), 5, 2, 0.0, 1.0, cc["sx"], ncc)
    put("nmpc_cost", ncost); put("nmpc_grad", ngrad[0])
    idc = golden["rhs_id"]["nominal"]
    put("id_p", idc["p"]); put("id_f", idc["f"])
    gpath = tmp_path / "host_golden.txt"
    gpath.write_text("\n".join(lines) + "\n")
    r = subprocess.run([host_bin, "--yaml", yaml_path, "--golden", str(gpath)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host_api_test: 0 failures" in r.stdout
