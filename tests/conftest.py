import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unhex(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "oracle_cases.json")) as fh:
        d = json.load(fh)
    return d["cases"]


@pytest.fixture(scope="session")
def golden_configs():
    """Members of BASELINE configs 4 and 5 and per-attempt logs, integrated once by the unmodified reference
    (tools/make_golden_configs.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "config_cases.json")) as fh:
        d = json.load(fh)
    return d["cases"]


@pytest.fixture(scope="session")
def golden_defaults():
    """What the reference's 4-argument irk::odeint(func, t0, t1, y0) returns (tools/make_golden_defaults.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "odeint_defaults.json")) as fh:
        return json.load(fh)["cases"]


def parse_reference_log(text):
    """Rows (step, t, dt, err, iters) of the reference's per-attempt log, irk.hpp:575-577, 805-810."""
    rows = []
    seen_header = False
    for line in text.splitlines():
        f = line.split()
        if f[:2] == ["Rehuel:", "step"]:
            seen_header = True
        elif seen_header and len(f) == 6 and f[0] == "Rehuel:":
            rows.append([float(x) for x in f[1:]])
    return np.array(rows)


def cases_of(golden, kind):
    return [c for c in golden if c["kind"] == kind]


@pytest.fixture(scope="session")
def port():
    """The plain-C restatement; built on demand (gcc only)."""
    from oracle import refsolver
    if not os.path.exists(refsolver.LIB_PATHS["port"]):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return refsolver.RefSolver("port")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference over the Armadillo stand-in (prebuilt oracle/_ref; skip if absent)."""
    from oracle import refsolver
    if not os.path.exists(refsolver.LIB_PATHS["reference"]):
        if os.path.exists("/root/reference/irk.hpp"):
            import subprocess
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    return refsolver.RefSolver("reference")


def scaled_err(y, yref, rtol, atol):
    """SURVEY.md 8(d) parity metric: max_i |y_i - yref_i| / (atol + rtol*|yref_i|), per member."""
    y = np.asarray(y, dtype=np.float64)
    yref = np.asarray(yref, dtype=np.float64)
    return np.max(np.abs(y - yref) / (atol + rtol * np.abs(yref)), axis=0)
