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


VARIANTS = [("thin_noalpha", True, True), ("thin_alpha", True, False),
            ("thick_noalpha", False, True), ("thick_alpha", False, False)]


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    r = np.where((a == b) | (np.isnan(a) & np.isnan(b)), 0.0, r)
    return r
