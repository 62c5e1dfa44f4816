"""CPU: the C-ABI shared library builds, loads, and exports every symbol that
include/mbb_b200.h declares; the product fails loudly without a CUDA device."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mbb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mbb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from mbb_emcee_b200 import _native
    lib = ctypes.CDLL(_native.library_path())
    syms = _header_symbols()
    assert len(syms) >= 19
    for name in syms:
        assert hasattr(lib, name), "libmbb_b200.so does not export %s" % name
    assert sorted(_native.EXPORTED_SYMBOLS) == syms
    assert _native.load_library().mbb_version() >= 100


def test_no_cpu_fallback():
    """Without a device the product raises; it never computes on the host."""
    from mbb_emcee_b200 import _native, likelihood
    if _native.load_library().mbb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    like = likelihood()
    like.set_phot([250.0, 350.0, 500.0], [30.0, 40.0, 30.0], [3.0, 4.0, 3.0])
    with pytest.raises(_native.MBBNativeError):
        like([10.0, 2.0, 100.0, 3.0, 30.0])


def test_product_does_not_import_oracle():
    """No module of the product package may reference the oracle."""
    pkg = os.path.join(ROOT, "mbb_emcee_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mbb_oracle" not in text and "ref_harness" not in text, f
