"""TEST INFRASTRUCTURE -- not product code.

Harness that makes the *unmodified* reference (aconley/mbb_emcee, mounted
read-only at /root/reference) importable in this container, so that golden
vectors can be generated from it (tests/golden/make_golden.py) and so the
oracle restatement (oracle/mbb_oracle.py) can be validated against it.

Nothing here is imported by the product package ``mbb_emcee_b200``.

What it does (recipe verified in SURVEY.md Appendix B):

1. Compiles the reference's only native file, ``mbb_emcee/fnu.pyx``, from where
   it lies under the reference tree into ``oracle/_ref/`` (git-ignored) as the
   top-level module ``fnu`` (the reference does ``import fnu``,
   mbb_emcee/modified_blackbody.py:7).  No reference source is copied into
   the repository; only build outputs land in ``oracle/_ref``.
2. Installs two compatibility shims the reference needs on numpy>=1.24 /
   python>=3.10: ``numpy.float`` (modified_blackbody.py:90,92,116,457,459;
   results.py:552) and ``collections.Iterable`` (utility.py:17).
3. Registers stub modules for the third-party imports that are absent here
   (astropy, h5py, emcee); only ``astropy.io.ascii.read`` needs behaviour
   (response.py:198,699; likelihood.py:254).
4. Optionally applies the two one-line fixes documented in SURVEY.md 2e that
   are required to *run* parts of the path on a modern numpy
   (``wband[0]`` in response._setup_alma, response.py:477-487; ``_normwave``
   for GHz deltas, response.py:352-361).  They are applied by wrapping
   methods at import time; the reference files are never modified.
"""
from __future__ import annotations

import glob
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
DEFAULT_REF = os.environ.get("MBB_REFERENCE_DIR", "/root/reference")


def reference_available(refdir: str = DEFAULT_REF) -> bool:
    return os.path.isfile(os.path.join(refdir, "mbb_emcee", "fnu.pyx"))


def ref_fnu_built() -> bool:
    return len(glob.glob(os.path.join(REF_OUT, "fnu*.so"))) > 0


def build_ref_fnu(refdir: str = DEFAULT_REF, force: bool = False) -> str:
    """Cythonize + compile <ref>/mbb_emcee/fnu.pyx into oracle/_ref/fnu*.so."""
    os.makedirs(REF_OUT, exist_ok=True)
    have = glob.glob(os.path.join(REF_OUT, "fnu*.so"))
    if have and not force:
        return have[0]
    if not reference_available(refdir):
        raise RuntimeError("reference tree not present at %s" % refdir)
    import subprocess
    import sysconfig

    import numpy

    pyx = os.path.join(refdir, "mbb_emcee", "fnu.pyx")
    c_out = os.path.join(REF_OUT, "fnu.c")
    # ``-X cpow=True``: the reference predates Cython 3, whose new default
    # (cpow=False) routes ``cx**(-alpha)`` on typed doubles (fnu.pyx:49,103)
    # through C99 complex ``cpow`` ("soft complex") instead of libm ``pow`` --
    # a toolchain artefact worth ~4 ulp on the power-law side.  cpow=True
    # restores the Cython 0.x code generation the reference was written for,
    # and agrees with its own numpy twin (modified_blackbody.py:474,486).
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "-X", "cpow=True",
                           pyx, "-o", c_out])
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    so = os.path.join(REF_OUT, "fnu" + ext)
    # Same flags the reference's setup.py implies: -lm, numpy include
    # (setup.py:11-13); plain -O2, no -ffast-math.
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fno-strict-aliasing",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           c_out, "-o", so, "-lm"]
    subprocess.check_call(cmd)
    return so


def import_ref_fnu():
    """Import the compiled reference ``fnu`` module (from oracle/_ref)."""
    if not ref_fnu_built():
        build_ref_fnu()
    if REF_OUT not in sys.path:
        sys.path.insert(0, REF_OUT)
    return importlib.import_module("fnu")


def _ascii_read(path, comment=None, **kw):
    """Stand-in for astropy.io.ascii.read: rows of float-or-str tokens."""
    rows = []
    with open(path, "r") as fh:
        for line in fh:
            s = line.strip()
            if not s or s.startswith("#"):
                continue
            toks = []
            for t in s.split():
                try:
                    toks.append(float(t))
                except ValueError:
                    toks.append(t)
            rows.append(tuple(toks))
    return rows


class _Unit(object):
    def __init__(self, name, cm):
        self.name = name
        self.cm = cm


class _Quantity(object):
    """Just enough of astropy.units.Quantity for results.py:95-98,665,778."""

    def __init__(self, value, unit):
        self.value = float(value)
        self.unit = unit

    def to(self, unit):
        return _Quantity(self.value * self.unit.cm / unit.cm, unit)


def _install_stubs():
    import collections
    import collections.abc

    import numpy

    if not hasattr(numpy, "float"):
        numpy.float = float
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable

    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    if "astropy" not in sys.modules:
        astropy = mod("astropy")
        io = mod("astropy.io")
        ascii_ = mod("astropy.io.ascii")
        fits = mod("astropy.io.fits")
        units = mod("astropy.units")
        cosmo = mod("astropy.cosmology")
        astropy.io = io
        io.ascii = ascii_
        io.fits = fits
        astropy.units = units
        astropy.cosmology = cosmo
        ascii_.read = _ascii_read
        units.Quantity = _Quantity
        units.Mpc = _Unit("Mpc", 3.0856775814913673e24)
        units.cm = _Unit("cm", 1.0)
    if "h5py" not in sys.modules:
        mod("h5py")
    if "emcee" not in sys.modules:
        emcee = mod("emcee")

        class EnsembleSampler(object):  # placeholder (mbb_fit.py:80)
            def __init__(self, nwalkers, dim, lnpostfn, threads=1, **kw):
                self.k = nwalkers
                self.dim = dim
                self.lnprobfn = lnpostfn

        emcee.EnsembleSampler = EnsembleSampler


def import_reference(refdir: str = DEFAULT_REF, apply_fixes: bool = True):
    """Return the imported reference package ``mbb_emcee``."""
    if not reference_available(refdir):
        raise RuntimeError("reference tree not present at %s" % refdir)
    import warnings

    import_ref_fnu()
    _install_stubs()
    if refdir not in sys.path:
        sys.path.insert(0, refdir)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("mbb_emcee")
    if apply_fixes and not getattr(ref, "_oracle_fixes_applied", False):
        _apply_fixes()
        ref._oracle_fixes_applied = True
    return ref


INSTALLED_REF = os.path.normpath(os.path.join(HERE, "..", "baseline", "_ref"))


def installed_reference_available(path: str = INSTALLED_REF) -> bool:
    """The reference as `pip install --target baseline/_ref` left it (package + compiled fnu)."""
    return os.path.isfile(os.path.join(path, "mbb_emcee", "likelihood.py")) and \
        len(glob.glob(os.path.join(path, "fnu*.so"))) > 0


def import_installed_reference(path: str = INSTALLED_REF, apply_fixes: bool = True):
    """Import the UNMODIFIED reference from its pip-installed copy (baseline/_ref: git-ignored,
    travels to the GPU box, where /root/reference does not exist).  Same shims as
    import_reference; the native module is the one the reference's own setup.py built."""
    if not installed_reference_available(path):
        raise RuntimeError("no installed reference at %s" % path)
    import warnings

    _install_stubs()
    if path not in sys.path:
        sys.path.insert(0, path)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = importlib.import_module("mbb_emcee")
    if apply_fixes and not getattr(ref, "_oracle_fixes_applied", False):
        _apply_fixes()
        ref._oracle_fixes_applied = True
    return ref


def import_any_reference():
    """(module, where) -- the reference tree when it is mounted (build container), else the
    installed copy; (None, reason) when neither exists."""
    try:
        if reference_available():
            return import_reference(), "/root/reference (fnu.pyx compiled into oracle/_ref)"
        if installed_reference_available():
            return import_installed_reference(), "baseline/_ref (pip install of the reference)"
    except Exception as exc:      # pragma: no cover - reported by the caller
        return None, "reference import failed: %r" % (exc,)
    return None, "neither /root/reference nor baseline/_ref is present"


def _apply_fixes():
    """Wrap (never edit) two reference methods so they run on modern numpy.

    * response._setup_alma (response.py:443-491): ``wband`` is an index array,
      so numpy>=1.16 linspace returns (13,1) arrays.  Flatten the outputs --
      numerically identical to indexing ``wband[0]``.
    * response._setup_delta (response.py:336-374): ``_normwave`` is assigned
      only in the 'thz' frequency branch; pre-set it from the value so GHz/MHz
      /Hz deltas work, exactly as the thz branch computes it.
    """
    import numpy

    rmod = sys.modules["mbb_emcee.response"]
    cls = rmod.response
    orig_alma = cls._setup_alma
    orig_delta = cls._setup_delta

    def _setup_alma(self, cent, xtype='freq', xunit='gHz'):
        xv, rs = _call_alma(orig_alma, self, cent, xtype, xunit)
        return xv, rs

    def _call_alma(fn, self, cent, xtype, xunit):
        # run the original with numpy.linspace temporarily squeezing (1,)-shaped
        # endpoints, which is all the wband[0] fix amounts to
        real_linspace = numpy.linspace

        def lin(a, b, n):
            return real_linspace(float(numpy.ravel(a)[0]), float(numpy.ravel(b)[0]), n)

        rmod.numpy.linspace = lin
        try:
            return fn(self, cent, xtype, xunit)
        finally:
            rmod.numpy.linspace = real_linspace

    def _setup_delta(self, val, xtyp, xun):
        if xtyp == 'freq':
            scale = {'hz': 1e-9, 'mhz': 1e-3, 'ghz': None, 'thz': 1e3}.get(xun)
            nf = val if scale is None else scale * val
            self._normwave = 299792458e-3 / nf
        return orig_delta(self, val, xtyp, xun)

    cls._setup_alma = _setup_alma
    cls._setup_delta = _setup_delta


if __name__ == "__main__":
    print("built:", build_ref_fnu(force="--force" in sys.argv))
    ref = import_reference()
    mbb = ref.modified_blackbody(10.0, 2.0, 800.0, 2.0, 45.0)
    print(mbb([250.0, 350.0, 500.0, 850.0]))
