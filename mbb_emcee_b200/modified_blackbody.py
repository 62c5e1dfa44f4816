"""Modified blackbody and blackbody SEDs, evaluated on the GPU.

Host-side mirror of the reference's ``modified_blackbody`` / ``blackbody``
classes (reference mbb_emcee/modified_blackbody.py:21-119, 154-674): same
constructor, properties and call conventions.  All arithmetic -- the
per-object constants (``normfac``, merge point, ``kappa``), ``f_nu``, the peak
wavelength and the frequency integral -- is done by the CUDA kernels behind
the C ABI (include/mbb_b200.h); nothing is computed on the CPU.

These objects are the single-SED convenience layer.  Ensembles of walkers go
through ``likelihood.__call__`` where setup, passband integration and the
chi-square are one fused kernel.
"""
import numpy

from . import _native
from .utility import isiterable

__all__ = ["modified_blackbody", "blackbody"]

# constants as hardwired by the reference (modified_blackbody.py:15-18)
c = 299792458e6       # um / s
h = 6.6260693e-34     # J s
k = 1.3806505e-23     # J / K
um_to_GHz = 299792458e-3


class modified_blackbody(object):
    """A modified grey body, f_nu ~ (1 - exp(-(nu/nu0)^beta)) B_nu(T), with an
    optional nu^-alpha power law joined on the Wien side.  Instances are static.
    """

    def __init__(self, T, beta, lambda0, alpha, fnorm, wavenorm=500.0,
                 noalpha=False, opthin=False, context=None):
        """Same parameters as reference modified_blackbody.py:168-198.

        ``context`` (extension): the ``_native.Context`` to evaluate on;
        defaults to the process-wide one for the current device.
        """
        self._T = float(T)
        self._beta = float(beta)
        self._hasalpha = not bool(noalpha)
        self._alpha = float(alpha) if self._hasalpha else None
        self._fnorm = float(fnorm)
        self._wavenorm = float(wavenorm)
        self._opthin = bool(opthin)
        self._lambda0 = None if self._opthin else float(lambda0)

        if self._hasalpha and self._alpha <= 0.0:
            raise ValueError("alpha must be positive.  You gave: "
                             "{:.5g}".format(self._alpha))
        if self._beta < 0.0:
            raise ValueError("beta must be non-negative.  You gave: "
                             "{:.5g}".format(self._beta))

        self._hcokt = h * c / (k * self._T)
        self._xnorm = self._hcokt / self._wavenorm
        if not self._opthin:
            self._x0 = self._hcokt / self._lambda0

        self._ctx = context if context is not None else _native.default_context()
        # unused parameters still travel to the device; give them inert values
        self._pars = numpy.array([self._T, self._beta,
                                  1.0 if self._opthin else self._lambda0,
                                  1.0 if not self._hasalpha else self._alpha,
                                  self._fnorm], dtype=numpy.float64)
        consts, status = self._device(lambda ctx: ctx.sed_consts(self._pars))
        _native.raise_for_status(status, self._pars)
        self._normfac = float(consts[0, 0])
        if self._hasalpha:
            self._xmerge = float(consts[0, 1])
            self._kappa = float(consts[0, 2])

    def _device(self, fn):
        ctx = self._ctx
        ctx.set_model(self._wavenorm, self._opthin, not self._hasalpha)
        return fn(ctx)

    # -------------------------------------------------------------- properties
    @property
    def T(self):
        """Temperature / (1+z) in K"""
        return self._T

    @property
    def beta(self):
        return self._beta

    @property
    def lambda0(self):
        """lambda_0 (1+z) in microns, None if optically thin"""
        return None if self._opthin else self._lambda0

    @property
    def alpha(self):
        return self._alpha if self._hasalpha else None

    @property
    def fnorm(self):
        """Normalization flux at wavenorm in mJy"""
        return self._fnorm

    @property
    def wavenorm(self):
        return self._wavenorm

    @property
    def has_alpha(self):
        return self._hasalpha

    @property
    def optically_thin(self):
        return self._opthin

    @property
    def wavemerge(self):
        """Merge wavelength in microns (None without the power law)"""
        if not self._hasalpha:
            return None
        return self._hcokt / self._xmerge

    def __repr__(self):
        lam0 = "None" if self._opthin else "{:.2g}".format(self._lambda0)
        alp = "None" if not self._hasalpha else "{:.2g}".format(self._alpha)
        out = "modified_blackbody({:.2g}, {:.2g}, {:s}, {:s}, {:.2g}".format(
            self._T, self._beta, lam0, alp, self._fnorm)
        if not self._hasalpha:
            out += ", noalpha=True"
        if self._opthin:
            out += ", opthin=True"
        return out + ", wavenorm={:.2g})".format(self._wavenorm)

    def __str__(self):
        out = "modified_blackbody(T: {:.2g} beta: {:.2g}".format(self._T, self._beta)
        if not self._opthin:
            out += " lambda0: {:.2g}".format(self._lambda0)
        if self._hasalpha:
            out += " alpha: {:.2g}".format(self._alpha)
        return out + " fnorm: {:.2g} wavenorm: {:.2g})".format(self._fnorm,
                                                                self._wavenorm)

    # -------------------------------------------------------------- evaluation
    def _eval(self, freq_ghz, scalar_path):
        out, status = self._device(
            lambda ctx: ctx.fnu(self._pars, freq_ghz, scalar_path=scalar_path))
        _native.raise_for_status(status, self._pars)
        return out[0]

    def f_nu(self, freq):
        """f_nu [mJy] at frequencies in GHz (reference :441-491; its numpy
        formulation, x = (h/kT)*1e9*nu)."""
        if not isiterable(freq):
            frequency = numpy.asarray([freq], dtype=numpy.float64)
        else:
            frequency = numpy.asanyarray(freq, dtype=numpy.float64)
        return self._eval(frequency.ravel(), True).reshape(frequency.shape)

    def _f_nu_c(self, freq):
        """Array path of the reference (:493-533 -> fnu.pyx), x = (1e9*h/kT)*nu."""
        if not isiterable(freq):
            frequency = numpy.asarray([freq], dtype=numpy.float64)
        else:
            frequency = numpy.asanyarray(freq, dtype=numpy.float64)
        return self._eval(frequency.ravel(), False).reshape(frequency.shape)

    def __call__(self, wave):
        """f_nu [mJy] at wavelengths in microns (reference :535-554): arrays
        take the array path, scalars the numpy path (1-element array back)."""
        if isiterable(wave):
            return self._f_nu_c(um_to_GHz /
                                numpy.asanyarray(wave, dtype=numpy.float64))
        return self.f_nu(um_to_GHz / float(wave))

    def max_wave(self):
        """Wavelength of maximum f_nu in microns (reference :581-637)."""
        consts, status = self._device(
            lambda ctx: ctx.sed_consts(self._pars, want_peak=True))
        _native.raise_for_status(status, self._pars)
        return float(consts[0, 5])

    def freq_integrate(self, minwave, maxwave, method="quadpack"):
        """Integral of f_nu over [minwave, maxwave] microns, erg/s/cm^2
        (reference :639-674).  ``method``: "quadpack" replays the reference's
        scipy.integrate.quad call (its value to ~1e-15), "gauss" is the
        fixed-rule quadrature (the true integral to ~1e-15)."""
        minwave = float(minwave)
        maxwave = float(maxwave)
        if minwave <= 0.0:
            raise ValueError("Minimum wavelength must be > 0.0")
        if minwave > maxwave:
            minwave, maxwave = maxwave, minwave
        # unit prefactor so the chain kernel returns 1e-17 * integral
        chain = self._pars.reshape(1, 1, 5)
        self._ctx.set_lir_method(method)
        _, lir, _, status = self._device(
            lambda ctx: ctx.chain_post(chain, 2, z=0.0, dl_mpc=-1.0, lir_min=minwave,
                                       lir_max=maxwave))
        _native.raise_for_status(status, self._pars)
        return float(lir[0, 0])


class blackbody(modified_blackbody):
    """A plain blackbody normalised to ``fnorm`` at ``wavenorm``.

    Reference mbb_emcee/modified_blackbody.py:21-119.  It is exactly the
    optically-thin, no-alpha grey body with beta = 0 (same normalisation
    ``fnorm*expm1(x_n)/x_n**3`` and same ``normfac * x**3 / expm1(x)``), so it
    shares the device path.  One deliberate difference: for array input the
    reference converts wavelength to frequency with ``c`` in um/s instead of
    um*GHz (:117), which yields frequencies 1e9 too large; here arrays and
    scalars agree.
    """

    def __init__(self, T, fnorm, wavenorm=500.0, context=None):
        modified_blackbody.__init__(self, T, 0.0, None, None, fnorm,
                                    wavenorm=wavenorm, noalpha=True,
                                    opthin=True, context=context)

    def __repr__(self):
        return "blackbody({:.2g}, {:.2g}, wavenorm={:.2g})".format(
            self._T, self._fnorm, self._wavenorm)

    def __str__(self):
        return "blackbody(T: {:.2g} fnorm: {:.2g} wavenorm: {:.2g})".format(
            self._T, self._fnorm, self._wavenorm)

    def __call__(self, wave):
        if isiterable(wave):
            return self.f_nu(um_to_GHz / numpy.asanyarray(wave, dtype=numpy.float64))
        return self.f_nu(um_to_GHz / float(wave))
