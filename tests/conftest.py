"""pytest configuration: registers the `gpu` marker and shared fixtures.

`-m "not gpu"` tests: oracle vs committed golden vectors, host logic, C-ABI symbol export (no GPU needed).
`-m gpu` tests: parity of the CUDA engine (through the C-ABI) against the oracle and the goldens.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def yaml_path():
    return os.path.join(ROOT, "data", "umx_radian.yaml")


@pytest.fixture(scope="session")
def oracle(yaml_path):
    from oracle.oracle_py import Oracle, params_from_yaml

    return Oracle(params_from_yaml(yaml_path))


# Parity metric: |a - b| <= rtol * max(|b|, scale)
def assert_close(a, b, rtol=1e-9, scale=1.0, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    fin_a, fin_b = np.isfinite(a), np.isfinite(b)
    assert np.array_equal(fin_a, fin_b), f"{what}: non-finite sets differ"
    err = np.abs(a - b)[fin_b] / np.maximum(np.abs(b[fin_b]), scale)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= rtol, f"{what}: max scaled error {worst:.3e} > {rtol:.1e}"
    return worst
