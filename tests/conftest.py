"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # Nothing is skipped silently on a GPU box: a gpu-marked test without a
    # device fails in the product's loader ("CUDA extension/device missing").
    pass


@pytest.fixture(scope="session")
def golden():
    class G(object):
        sed = np.load(os.path.join(GOLDEN, "golden_sed.npz"))
        response = np.load(os.path.join(GOLDEN, "golden_response.npz"))
        like = np.load(os.path.join(GOLDEN, "golden_like.npz"))
        results = np.load(os.path.join(GOLDEN, "golden_results.npz"))
    return G


@pytest.fixture(scope="session")
def oracle():
    import mbb_oracle
    return mbb_oracle


class _Observed(object):
    """Observed maxima of the parity tests (max relative error etc.), written to
    gpurun_out/observed_errors.json at the end of the session when that directory exists -- the
    numbers DESIGN.md section 6 quotes are copied from there into profiles/."""

    def __init__(self):
        self.rows = {}

    def record(self, what, value):
        v = float(value)
        self.rows[what] = max(self.rows.get(what, v), v) if "fraction" not in what else min(self.rows.get(what, v), v)


_OBSERVED = _Observed()


@pytest.fixture(scope="session")
def observed():
    return _OBSERVED


def pytest_sessionfinish(session, exitstatus):
    out = os.path.join(ROOT, "gpurun_out")
    if _OBSERVED.rows and os.path.isdir(out):
        import json
        path = os.path.join(out, "observed_errors.json")
        old = {}
        if os.path.exists(path):
            try:
                old = json.load(open(path))
            except Exception:
                old = {}
        old.update(_OBSERVED.rows)
        json.dump(old, open(path, "w"), indent=1, sort_keys=True)


VARIANTS = [("thin_noalpha", True, True), ("thin_alpha", True, False),
            ("thick_noalpha", False, True), ("thick_alpha", False, False)]


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    r = np.where((a == b) | (np.isnan(a) & np.isnan(b)), 0.0, r)
    # every comparison of a GPU test leaves its maximum behind (per test; see _Observed)
    cur = os.environ.get("PYTEST_CURRENT_TEST", "")
    if "_gpu.py::" in cur and r.size:
        _OBSERVED.record("max relerr in " + cur.split("::", 1)[1].split(" ")[0], np.max(r))
    return r
