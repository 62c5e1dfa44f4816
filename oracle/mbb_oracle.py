"""TEST INFRASTRUCTURE -- the CPU oracle.  Not product code.

A functional CPU restatement of the hot path of aconley/mbb_emcee, used ONLY as
the checker in tests/, in ``__graft_entry__.smoke()`` and as the timed
``cpu_baseline`` / ``--impl reference`` leg of bench.py.  The product package
``mbb_emcee_b200`` never imports it and has no CPU fallback.

Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
against (a) the reference's own known-answer vectors
(mbb_emcee/tests/test_modified_blackbody.py:18-20,34-36,57-69;
test_response.py:24-29,38-41) and (b) golden vectors produced by executing the
unmodified reference in the build container (tests/golden/make_golden.py ->
tests/golden/*.npz), to 0 ulp where the arithmetic is the same libm calls in
the same order and to <=1e-15 otherwise.

Third-party arithmetic the reference leans on and this file therefore also
uses (same installed versions): glibc libm ``pow/expm1/exp``
(via ``math`` and numpy), ``scipy.special.lambertw``,
``scipy.optimize.brentq`` (defaults xtol=2e-12, rtol=4*eps, maxiter=100),
``scipy.integrate.quad`` (QUADPACK dqagse, epsabs=epsrel=1.49e-8, limit=50),
``numpy.linalg.inv``.  scipy 1.18.1 / numpy 2.3.5 here; the reference pins
only ``scipy>0.8``, ``numpy>1.7`` (setup.py:35-36).

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

import ctypes
import glob
import math
import os
import sys
from collections import namedtuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# constants, modified_blackbody.py:15-18
C_UM = 299792458e6        # um / s
H = 6.6260693e-34
K = 1.3806505e-23
UM_TO_GHZ = 299792458e-3

# ----------------------------------------------------------------------------
# native node loops: the reference's compiled fnu.pyx if built, else the C port
# ----------------------------------------------------------------------------
_fnu_ref = None
_fnu_port = None


def _load_ref_fnu():
    global _fnu_ref
    if _fnu_ref is None:
        cands = glob.glob(os.path.join(HERE, "_ref", "fnu*.so"))
        if not cands:
            return None
        refdir = os.path.join(HERE, "_ref")
        if refdir not in sys.path:
            sys.path.insert(0, refdir)
        try:
            import fnu  # noqa: the reference's Cython module
            _fnu_ref = fnu
        except Exception:
            _fnu_ref = False
    return _fnu_ref or None


def _load_port():
    global _fnu_port
    if _fnu_port is None:
        so = os.path.join(HERE, "_build", "libfnu_port.so")
        if not os.path.exists(so):
            import subprocess
            subprocess.check_call(["make", "-C", HERE, "-s"])
        lib = ctypes.CDLL(so)
        dp = ctypes.POINTER(ctypes.c_double)
        d = ctypes.c_double
        lib.oracle_fnu_thin_noalpha.argtypes = [dp, ctypes.c_size_t, d, d, d, dp]
        lib.oracle_fnu_thin_walpha.argtypes = [dp, ctypes.c_size_t, d, d, d, d, d, d, dp]
        lib.oracle_fnu_thick_noalpha.argtypes = [dp, ctypes.c_size_t, d, d, d, d, dp]
        lib.oracle_fnu_thick_walpha.argtypes = [dp, ctypes.c_size_t, d, d, d, d, d, d, d, dp]
        _fnu_port = lib
    return _fnu_port


def native_kind():
    """'reference' if the reference's own compiled fnu is in use, else 'port'."""
    return "reference" if _load_ref_fnu() is not None else "port"


def _ptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------
# SED model: per-walker constants  (modified_blackbody.py:168-337)
# ----------------------------------------------------------------------------
SED = namedtuple("SED", "T beta lambda0 alpha fnorm wavenorm noalpha opthin "
                        "hcokt xnorm x0 normfac xmerge kappa")


def merge_residual(x, alpha, beta, x0):
    """modified_blackbody.py:122-151 (alpha_merge_eqn)."""
    try:
        t = (x / x0)**beta
        bterm = t / math.expm1(t)
    except OverflowError:
        bterm = 0.0
    return x - (1.0 - math.exp(-x)) * (3.0 + alpha + beta * bterm)


def make_sed(T, beta, lambda0, alpha, fnorm, wavenorm=500.0, noalpha=False,
             opthin=False):
    """Per-walker constants, modified_blackbody.py:200-337."""
    from scipy.optimize import brentq
    from scipy.special import lambertw

    T = float(T)
    beta = float(beta)
    fnorm = float(fnorm)
    wavenorm = float(wavenorm)
    noalpha = bool(noalpha)
    opthin = bool(opthin)
    alpha_f = None if noalpha else float(alpha)
    lam0_f = None if opthin else float(lambda0)
    if (not noalpha) and alpha <= 0.0:                       # :219-221
        raise ValueError("alpha must be positive.  You gave: {:.5g}".format(alpha_f))
    if beta < 0.0:                                           # :222-224
        raise ValueError("beta must be non-negative.  You gave: {:.5g}".format(beta))

    hcokt = H * C_UM / (K * T)                               # :228
    x0 = None if opthin else hcokt / lambda0                 # :232 (raw lambda0)
    xnorm = hcokt / wavenorm                                 # :233
    xmerge = kappa = None
    if opthin:
        if noalpha:                                          # :240-241
            normfac = fnorm * math.expm1(xnorm) / xnorm**(3.0 + beta)
        else:                                                # :253-269
            a = 3.0 + alpha_f + beta
            xmerge = a + lambertw(-a * math.exp(-a)).real
            kappa = xmerge**(3.0 + alpha_f + beta) / math.expm1(xmerge)
            if xnorm > xmerge:
                normfac = fnorm * xnorm**alpha_f / kappa
            else:
                normfac = fnorm * math.expm1(xnorm) / xnorm**(3.0 + beta)
    else:
        if noalpha:                                          # :274-276
            normfac = -fnorm * math.expm1(xnorm) / \
                (math.expm1(-(xnorm / x0)**beta) * xnorm**3)
        else:                                                # :286-337
            a = 0.1
            aval = merge_residual(a, alpha_f, beta, x0)
            it = 0
            while aval >= 0.0:
                a /= 2.0
                aval = merge_residual(a, alpha_f, beta, x0)
                if it > 100:
                    raise ValueError("Couldn't bracket low alpha merge point")
                it += 1
            b = 15.0
            bval = merge_residual(b, alpha_f, beta, x0)
            it = 0
            while bval <= 0.0:
                b *= 2.0
                bval = merge_residual(b, alpha_f, beta, x0)
                if it > 100:
                    raise ValueError("Couldn't bracket high alpha merge point")
                it += 1
            xmerge = brentq(merge_residual, a, b, args=(alpha_f, beta, x0), disp=True)
            kappa = -xmerge**(3 + alpha_f) * \
                math.expm1(-(xmerge / x0)**beta) / math.expm1(xmerge)
            if xnorm > xmerge:
                normfac = fnorm * xnorm**alpha_f / kappa
            else:
                ef = math.expm1(-(xnorm / x0)**beta)
                normfac = -fnorm * math.expm1(xnorm) / (xnorm**3 * ef)
    return SED(T, beta, lam0_f, alpha_f, fnorm, wavenorm, noalpha, opthin,
               hcokt, xnorm, x0, normfac, xmerge, kappa)


def wavemerge(s):
    """modified_blackbody.py:383-389."""
    return None if s.noalpha else s.hcokt / s.xmerge


# ----------------------------------------------------------------------------
# f_nu evaluation
# ----------------------------------------------------------------------------
def fnu_native(s, freq, impl=None):
    """Array path: modified_blackbody._f_nu_c (:493-533) -> fnu.pyx loops."""
    freq = np.ascontiguousarray(freq, dtype=np.float64)
    ref = _load_ref_fnu() if impl in (None, "reference") else None
    if impl == "reference" and ref is None:
        raise RuntimeError("oracle/_ref/fnu*.so not built")
    if ref is not None:
        if s.opthin:
            if s.noalpha:
                return ref.fnueval_thin_noalpha(freq, s.T, s.beta, s.normfac)
            return ref.fnueval_thin_walpha(freq, s.T, s.beta, s.alpha, s.normfac,
                                           s.xmerge, s.kappa)
        if s.noalpha:
            return ref.fnueval_thick_noalpha(freq, s.T, s.beta, s.x0, s.normfac)
        return ref.fnueval_thick_walpha(freq, s.T, s.beta, s.x0, s.alpha,
                                        s.normfac, s.xmerge, s.kappa)
    lib = _load_port()
    out = np.empty_like(freq)
    n = freq.size
    if s.opthin:
        if s.noalpha:
            lib.oracle_fnu_thin_noalpha(_ptr(freq), n, s.T, s.beta, s.normfac, _ptr(out))
        else:
            lib.oracle_fnu_thin_walpha(_ptr(freq), n, s.T, s.beta, s.alpha, s.normfac,
                                       s.xmerge, s.kappa, _ptr(out))
    else:
        if s.noalpha:
            lib.oracle_fnu_thick_noalpha(_ptr(freq), n, s.T, s.beta, s.x0, s.normfac,
                                         _ptr(out))
        else:
            lib.oracle_fnu_thick_walpha(_ptr(freq), n, s.T, s.beta, s.x0, s.alpha,
                                        s.normfac, s.xmerge, s.kappa, _ptr(out))
    return out


def fnu_numpy(s, freq):
    """Scalar/numpy path: modified_blackbody.f_nu (:441-491).

    Note x = (h/(kT)) * 1e9 * freq -- a different rounding of the same number
    as the native path's (1e9*h/(kT)) * freq."""
    frequency = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    hokt = H / (K * s.T)
    x = hokt * 1e9 * frequency
    if s.opthin:
        if s.noalpha:
            return s.normfac * x**(3.0 + s.beta) / np.expm1(x)
        out = np.zeros_like(frequency)
        pw = x > s.xmerge
        out[pw] = s.kappa * x[pw]**(-s.alpha)
        out[~pw] = x[~pw]**(3.0 + s.beta) / np.expm1(x[~pw])
        out *= s.normfac
        return out
    if s.noalpha:
        return -s.normfac * np.expm1(-(x / s.x0)**s.beta) * x**3 / np.expm1(x)
    out = np.zeros_like(frequency)
    pw = x > s.xmerge
    out[pw] = s.kappa * x[pw]**(-s.alpha)
    out[~pw] = -np.expm1(-(x[~pw] / s.x0)**s.beta) * x[~pw]**3 / np.expm1(x[~pw])
    out *= s.normfac
    return out


def sed_call(s, wave, impl=None):
    """modified_blackbody.__call__ (:535-554): array -> native, scalar -> numpy."""
    if isinstance(wave, np.ndarray) and wave.ndim == 0 or np.isscalar(wave):
        return fnu_numpy(s, UM_TO_GHZ / float(wave))
    return fnu_native(s, UM_TO_GHZ / np.asanyarray(wave, dtype=np.float64), impl)


# ----------------------------------------------------------------------------
# peak wavelength (modified_blackbody.py:556-637)
# ----------------------------------------------------------------------------
def snu_deriv(s, x):
    """modified_blackbody._snudev (:556-579)."""
    if s.opthin:
        ef = math.expm1(x)
        return x**(2.0 + s.beta) * (3.0 + s.beta) / ef - \
            math.exp(x) * x**(3.0 + s.beta) / ef**2
    ef = math.expm1(x)
    xx0 = x / s.x0
    try:
        xx0b = xx0**s.beta
        eb = -math.expm1(-xx0b)
        return 3 * x**2 * eb / ef - math.exp(x) * x**3 * eb / ef**2 + \
            s.beta * x**3 * math.exp(-xx0b) * xx0b / (x * ef)
    except OverflowError:
        return 3 * x**2 / ef - math.exp(x) * x**3 / ef**2


def max_wave(s):
    """modified_blackbody.max_wave (:581-637)."""
    from scipy.optimize import brentq
    xbb = 2.82144
    if s.opthin and s.beta == 0:
        return C_UM / (xbb * K * s.T / H)
    a = xbb / 2.0
    av = snu_deriv(s, a)
    it = 0
    while av <= 0.0:
        if it > 20:
            raise Exception("Couldn't bracket maximum from low frequency side")
        a /= 2.0
        av = snu_deriv(s, a)
        it += 1
    b = xbb * 2.0
    bv = snu_deriv(s, b)
    it = 0
    while bv >= 0.0:
        if it > 20:
            raise Exception("Couldn't bracket maximum from high frequency side")
        b *= 2.0
        bv = snu_deriv(s, b)
        it += 1
    xmax = brentq(lambda x: snu_deriv(s, x), a, b, disp=True)
    return C_UM / (xmax * K * s.T / H)


def freq_integrate(s, minwave, maxwave):
    """modified_blackbody.freq_integrate (:639-674); erg/s/cm^2."""
    from scipy.integrate import quad
    minwave, maxwave = float(minwave), float(maxwave)
    if minwave <= 0.0:
        raise ValueError("Minimum wavelength must be > 0.0")
    if minwave > maxwave:
        minwave, maxwave = maxwave, minwave
    val = quad(lambda f: fnu_numpy(s, f), UM_TO_GHZ / maxwave, UM_TO_GHZ / minwave)[0]
    return 1e-17 * val


# ----------------------------------------------------------------------------
# passband flux (response.py:544-576); tables come from the product's host
# table builder, itself pinned bit-for-bit to the reference's attributes by
# tests/test_host_api_cpu.py::test_response_tables_match_reference_bit_for_bit
# ----------------------------------------------------------------------------
Band = namedtuple("Band", "isdelta normwave wave sedmult normfac")


def band_from_response(r):
    """Snapshot the fields response.__call__ touches (works for the reference's
    response objects and for the product's)."""
    if r._isdelta:
        return Band(True, float(r._normwave), None, None, None)
    return Band(False, None, np.array(r._wave), np.array(r._sedmult), float(r._normfac))


def band_flux(s, band, impl=None):
    if band.isdelta:
        return sed_call(s, band.normwave)                     # scalar -> numpy path
    return (sed_call(s, band.wave, impl) * band.sedmult).sum() * band.normfac


# ----------------------------------------------------------------------------
# likelihood (likelihood.py:643-834)
# ----------------------------------------------------------------------------
class LikeSpec(object):
    """Data + limits + priors, with the reference's defaults
    (likelihood.py:73, 83-92)."""

    def __init__(self, wavenorm=500.0, noalpha=False, opthin=False):
        self.wavenorm = float(wavenorm)
        self.noalpha = bool(noalpha)
        self.opthin = bool(opthin)
        self.lowlim = np.array([1, 0.1, 1, 0.1, 1e-3])
        inf = float("inf")
        self.has_uplim = [False, True, False, True, False, False]
        self.uplim = np.array([inf, 20.0, inf, 20.0, inf, inf])
        self.has_gprior = [False] * 6
        self.gprior_mean = np.zeros(6)
        self.gprior_ivar = np.ones(6)
        self.bands = None        # list[Band] when response-integrating
        self.wave = None         # float64[nb] data wavelengths otherwise
        self.flux = None
        self.ivar = None
        self.invcov = None

    def set_phot(self, first, flux, flux_unc):
        """likelihood.set_phot (:158-232).  ``first``: wavelengths [um], or a
        list of Band (response mode; data wavelengths must then be given by
        ``eff_wave``)."""
        if len(first) and isinstance(first[0], Band):
            self.bands, self.wave = list(first), None
        else:
            self.bands, self.wave = None, np.asarray(first, dtype=np.float64)
        self.flux = np.asarray(flux)
        self.ivar = 1.0 / np.asarray(flux_unc)**2
        self.invcov = None

    def auto_lambda0_uplim(self, maxwave):
        """:227-229 -- first set_phot latches lambda0 <= 3*max(wave)."""
        if not self.has_uplim[2]:
            self.has_uplim[2] = True
            self.uplim[2] = 3.0 * maxwave

    def set_cov(self, cov):
        self.invcov = np.linalg.inv(np.asarray(cov))          # :356

    def set_uplim(self, idx, val):
        self.has_uplim[idx] = True
        self.uplim[idx] = val

    def set_gprior(self, idx, mean, sigma):
        self.has_gprior[idx] = True
        self.gprior_mean[idx] = float(mean)
        self.gprior_ivar[idx] = 1.0 / (float(sigma)**2)       # :570


def loglike(spec, pars, impl=None):
    """likelihood.__call__ (:790-834) for one parameter vector."""
    if len(pars) != 5:
        raise ValueError("pars is not of expected length 5")
    for i in range(5):                                        # :666-668
        if pars[i] < spec.lowlim[i]:
            return float("-inf")
    s = make_sed(pars[0], pars[1], pars[2], pars[3], pars[4],
                 wavenorm=spec.wavenorm, noalpha=spec.noalpha, opthin=spec.opthin)
    if spec.bands is not None:                                # :813-815 (+ravel)
        model = np.array([np.ravel(band_flux(s, b, impl))[0] for b in spec.bands])
    else:                                                     # :817
        model = sed_call(s, spec.wave, impl)
    diff = spec.flux - model
    if spec.invcov is not None:                               # :823
        ll = -0.5 * np.dot(diff, np.dot(spec.invcov, diff))
    else:                                                     # :825
        ll = -0.5 * np.sum(diff**2 * spec.ivar)
    # soft upper limits (:701-717)
    pen = 0.0
    for i in range(5):
        if spec.has_uplim[i]:
            lim = spec.uplim[i]
            if pars[i] > lim:
                limvar = (0.02 * (lim - spec.lowlim[i]))**2
                pen -= 0.5 * (pars[i] - lim)**2 / limvar
    if spec.has_uplim[5]:
        val = max_wave(s)
        lim = spec.uplim[5]
        if val > lim:
            limvar = (0.02 * lim)**2
            pen -= 0.5 * (val - lim)**2 / limvar
    ll += pen
    # Gaussian priors (:737-752)
    if any(spec.has_gprior):
        pr = 0.0
        for i in range(5):
            if spec.has_gprior[i]:
                d = pars[i] - spec.gprior_mean[i]
                pr -= 0.5 * spec.gprior_ivar[i] * d**2
        if spec.has_gprior[5]:
            d = max_wave(s) - spec.gprior_mean[5]
            pr -= 0.5 * spec.gprior_ivar[5] * d**2
        ll += pr
    return float(ll)


def loglike_batch(spec, P, impl=None):
    P = np.asarray(P, dtype=np.float64)
    return np.array([loglike(spec, P[i], impl) for i in range(P.shape[0])])


# ----------------------------------------------------------------------------
# chain post-processing (results.py:534-801, 1267-1326)
# ----------------------------------------------------------------------------
def map_chain(chain, func):
    """results._map_chain (:553-566): per-walker sequential allclose-dedupe."""
    nw, ns = chain.shape[0:2]
    out = np.empty((nw, ns), dtype=np.float64)
    for w in range(nw):
        prev = chain[w, 0, :]
        out[w, 0] = func(prev)
        for t in range(1, ns):
            cur = chain[w, t, :]
            if np.allclose(prev, cur):
                out[w, t] = out[w, t - 1]
            else:
                out[w, t] = func(cur)
                prev = cur
    return out


def peaklambda_step(step):
    """results.compute_peaklambda inner (:574-580).

    Reference defect, replicated: ``_map_chain(peaklambda_inner)`` is called
    WITHOUT the opthin/noalpha keywords (:580), so the SED is always built
    optically thick with alpha (the closure's defaults, :574) and with the
    default wavenorm=500, whatever the fit used."""
    return max_wave(make_sed(step[0], step[1], step[2], step[3], step[4],
                             opthin=False, noalpha=False))


def lir_step(step, z, lammin, lammax, opthin, noalpha):
    """results.mbb_freqint.__call__ (:1314-1326); wavenorm default 500."""
    opz = 1.0 + z
    s = make_sed(step[0], step[1], step[2], step[3], step[4], opthin=opthin,
                 noalpha=noalpha)
    return freq_integrate(s, lammin * opz, lammax * opz)


LIR_PREFAC = 3.11749657e4           # results.py:666  (times dl[Mpc]^2)
MPC_CM = 3.0856775814913673e24


def dustmass_consts(z, wavenorm, kappa_wave, lumdist_mpc):
    """results.compute_dustmass precomputation (:778-791)."""
    dl = lumdist_mpc * MPC_CM
    dl2 = dl**2
    opz = 1.0 + z
    wavenorm_rest = wavenorm / opz
    nunorm_rest = 299792458e6 / wavenorm_rest
    temp_fac = 6.6260693e-27 * nunorm_rest / 1.38065e-16
    bnu_fac = 2 * 6.6260693e-27 * nunorm_rest**3 / 299792458e2**2
    knu_fac = wavenorm_rest / kappa_wave
    return opz, bnu_fac, temp_fac, knu_fac, dl2


def dustmass_step(step, kappa, wavenorm, opthin, opz, bnu_fac, temp_fac, knu_fac, dl2):
    """results._dmass_calc (:726-744)."""
    msolar8 = 1.97792e41
    T = step[0] * opz
    beta = step[1]
    S_nu = step[4] * 1e-26
    B_nu = bnu_fac / math.expm1(temp_fac / T)
    K_nu = 10.0 * kappa * knu_fac**(-beta)
    dm = dl2 * S_nu / (opz * K_nu * B_nu * msolar8)
    if not opthin:
        tau = (step[2] / wavenorm)**beta
        dm *= -tau / math.expm1(-tau)
    return dm


# ----------------------------------------------------------------------------
# emcee 2.2-semantics stretch move -- the sampler the
# reference drives (mbb_fit.py:80-81, 533-542).  emcee is a third-party
# dependency absent from /root/reference and not installed: parity at the
# sampler boundary is anchored on this published algorithm (Goodman & Weare
# 2010; emcee 2.2.1 ensemble.py ``_propose_stretch``), NOT on executed emcee.
# ----------------------------------------------------------------------------
def stretch_chain(lnprob_rows, p0, nsteps, rstate, a=2.0):
    """Run nsteps of the two-half stretch move.

    lnprob_rows: callable (m,5) ndarray -> (m,) ndarray.
    Returns chain[k,nsteps,dim], lnprob[k,nsteps], naccepted[k]."""
    p = np.array(p0, dtype=np.float64)
    k, dim = p.shape
    half = k // 2
    lnprob = np.asarray(lnprob_rows(p), dtype=np.float64)
    chain = np.empty((k, nsteps, dim))
    lnp = np.empty((k, nsteps))
    nacc = np.zeros(k)
    first, second = slice(half), slice(half, k)
    for i in range(nsteps):
        for S0, S1 in ((first, second), (second, first)):
            s = p[S0]
            c = p[S1]
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.0) * rstate.rand(Ns) + 1) ** 2.0 / a
            rint = rstate.randint(Nc, size=(Ns,))
            q = c[rint] - zz[:, np.newaxis] * (c[rint] - s)
            newlnp = np.asarray(lnprob_rows(q), dtype=np.float64)
            lnpdiff = (dim - 1.0) * np.log(zz) + newlnp - lnprob[S0]
            accept = lnpdiff > np.log(rstate.rand(len(lnpdiff)))
            if np.any(accept):
                lnprob[S0][accept] = newlnp[accept]
                p[S0][accept] = q[accept]
                nacc[S0][accept] += 1
        chain[:, i, :] = p
        lnp[:, i] = lnprob
    return chain, lnp, nacc
