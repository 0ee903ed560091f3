"""the CasADi external-function seam (libkite_casadi.so, include/kite_casadi.h).  CasADi itself is not
installed in this image, so tests/cpp/casadi_external_test.c plays its part: dlopen + dlsym of NAME, NAME_n_in, NAME_n_out,
NAME_sparsity_in / _out, NAME_work and a call with the (arg, res, iw, w, mem) convention, for the Function names the
reference builds (kite.cpp:324 "dynamics", :328 "dyn_jacobian", :330 "Aero", :338 "RK4")."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "cpp")


@pytest.fixture(scope="module")
def driver():
    from openkite_b200 import build
    build.build()
    exe = os.path.join(HERE, "casadi_external_test")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-o", exe, os.path.join(HERE, "casadi_external_test.c"), "-ldl", "-lm"])
    return exe, build.LIB_CASADI


def test_shim_exports_every_declared_symbol(driver):
    _, lib = driver
    L = C.CDLL(lib)
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "kite_casadi.h")).read(), flags=re.S)
    names = [n for n in re.findall(r"KITE_CASADI_DECLARE\((\w+)\)", txt) if n != "NAME"]
    assert names == ["dynamics", "dyn_jacobian", "Aero", "RK4", "dynamics_id", "dyn_jacobian_id"]
    for n in names:
        for suffix in ("", "_n_in", "_n_out", "_sparsity_in", "_sparsity_out", "_work", "_name_in", "_name_out", "_incref", "_decref"):
            assert hasattr(L, n + suffix), n + suffix
    assert hasattr(L, "kite_external_init") and hasattr(L, "kite_external_shutdown")


def test_patterns_without_gpu(driver, yaml_path):
    """Arity, work sizes and sparsity patterns are answered without touching the GPU (what casadi::external asks at load)."""
    exe, lib = driver
    r = subprocess.run([exe, lib, "--patterns"], capture_output=True, text=True, env=dict(os.environ, KITE_B200_YAML=yaml_path))
    assert r.returncode == 0, r.stdout + r.stderr
    # the pattern matches the engine's own sparsity query and the oracle's numerical Jacobian
    from openkite_b200 import load_library
    L = load_library()
    L.kite_jac_sparsity.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rows, cols = (C.c_int * 169)(), (C.c_int * 169)()
    assert L.kite_jac_sparsity(0, 0, 0, rows, cols) == 104
    assert L.kite_jac_sparsity(0, 1, 0, None, None) == 125 and L.kite_jac_sparsity(0, 0, 1, None, None) == 7
    assert L.kite_jac_sparsity(2, 0, 0, None, None) == 49 and L.kite_jac_sparsity(2, 0, 1, None, None) == 0
    from oracle.oracle_py import Oracle, params_from_yaml
    orc = Oracle(params_from_yaml(yaml_path))
    Jx, Ju = orc.jac(orc.synth_x0(5, 1), orc.synth_controls(5, 1, 1)[:, 0, :])
    nz = {(int(i), int(j)) for i, j in zip(*np.nonzero(Jx[0]))}
    assert nz == {(rows[k], cols[k]) for k in range(104)}


@pytest.mark.gpu
def test_external_functions_on_gpu(driver, yaml_path, golden, oracle, tmp_path):
    exe, lib = driver
    c = golden["rhs"]["model_test"]
    s = golden["rk4_step"]["model_test_h1ms"]
    assert s["x"] == c["x"] and s["u"] == c["u"]
    pid = golden["rhs_id"]["perturbed"]["p"]
    lines = []

    def put(tag, v):
        v = np.asarray(v, dtype=np.float64).ravel()
        lines.append("%s %d %s" % (tag, v.size, " ".join(repr(float(t)) for t in v)))

    put("x", c["x"]); put("u", c["u"]); put("f", c["f"]); put("Jx", c["Jx"]); put("aero", c["aero"])
    put("h", [s["h"]]); put("xn", s["xn"]); put("p", pid)
    put("f_id", oracle.rhs(c["x"], c["u"], pid, kind=1)[0])
    gp = tmp_path / "casadi_golden.txt"
    gp.write_text("\n".join(lines) + "\n")
    r = subprocess.run([exe, lib, "--golden", str(gp)], capture_output=True, text=True, env=dict(os.environ, KITE_B200_YAML=yaml_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
