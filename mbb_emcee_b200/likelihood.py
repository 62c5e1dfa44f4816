"""Data, limits, priors -- and the batched GPU log-likelihood.

Host-side mirror of the reference's ``likelihood`` class (reference
mbb_emcee/likelihood.py:17-834): same constructor, setters, getters and
defaults.  ``__call__`` accepts what emcee hands a log-probability function --
one parameter vector ``(5,)`` -> float -- and also a whole block ``(n, 5)`` ->
``ndarray[n]`` (emcee >= 3 ``vectorize=True``, or the ``pool`` object returned
by :meth:`likelihood.as_pool` for emcee 2).  Either way the work is one launch
of the fused CUDA kernel (per-walker SED setup, passband integration,
chi-square, soft upper limits, Gaussian priors incl. the lambda_peak terms)
through ``mbb_loglike`` of the C ABI.  There is no CPU evaluation path.
"""
import copy

import numpy as np

from . import _native
from .modified_blackbody import modified_blackbody
from .response import response, response_set, special_types
from .utility import read_text_table

__all__ = ["likelihood"]


class likelihood(object):
    """Class holding data, defining likelihood"""

    _param_order = {'t': 0, 't/(1+z)': 0, 'beta': 1, 'lambda0': 2,
                    'lambda0*(1+z)': 2, 'lambda_0': 2, 'lambda_0*(1+z)': 2,
                    'alpha': 3, 'fnorm': 4, 'f500': 4}

    def __init__(self, photfile=None, covfile=None, covextn=0,
                 wavenorm=500.0, noalpha=False, opthin=False,
                 response=False, responsefile=None, responsedir=None,
                 device=None):
        """Same parameters as reference likelihood.py:24-59, plus ``device``
        (CUDA ordinal; default: env MBB_B200_DEVICE / LOCAL_RANK / 0)."""
        self._wavenorm = float(wavenorm)
        self._noalpha = bool(noalpha)
        self._opthin = bool(opthin)

        # lower limits on every parameter (reference :73)
        self._lowlim = np.array([1, 0.1, 1, 0.1, 1e-3])

        self._limprior_order = copy.copy(self._param_order)
        self._limprior_order.update({'lambda_peak': 5, 'peaklam': 5,
                                     'lambdapeak': 5, 'peak_lambda': 5})

        # beta and alpha have soft upper limits by default (reference :83-85)
        inf = float("inf")
        self._has_uplim = [False, True, False, True, False, False]
        self._uplim = np.array([inf, 20.0, inf, 20.0, inf, inf])

        self._any_gprior = False
        self._has_gprior = [False] * 6
        self._gprior_mean = np.zeros(6)
        self._gprior_sigma = np.zeros(6)
        self._gprior_ivar = np.ones(6)

        self._device = device
        self._ctx = None
        self._dirty = True
        self._math_mode = _native.MATH_FAST
        self._cov_solver = "inverse"

        self._response_integrate = False
        if response:
            self.read_responses(responsefile, responsedir=responsedir)

        self._data_read = False
        self._has_covmatrix = False
        if photfile is not None:
            self.read_phot(photfile)
            if covfile is not None:
                if not isinstance(covfile, str):
                    raise TypeError("covfile must be string-like")
                self.read_cov(covfile, extn=covextn)
            # reference :110-111 -- only on the file path
            self._lowlim[4] = 1e-3 * self._flux.min()
        elif covfile is not None:
            raise Exception("Can't pass in covfile if no photfile")

        self._badval = float("-inf")

    # a live CUDA context cannot cross a fork/pickle boundary (emcee threads>1)
    def __getstate__(self):
        raise TypeError("likelihood objects hold a CUDA context and cannot be "
                        "pickled; the GPU evaluates whole ensembles per call, "
                        "use nthreads=1")

    # ------------------------------------------------------------ properties
    @property
    def wavenorm(self):
        return self._wavenorm

    @property
    def noalpha(self):
        return self._noalpha

    @property
    def opthin(self):
        return self._opthin

    @property
    def response_integrate(self):
        return self._response_integrate

    @property
    def math_mode(self):
        """0 = reference evaluation order (pow per node), 1 = restructured
        exp-only node arithmetic (default; see csrc/mbb_model.cuh), 2 = 1 plus
        the 32-point Gauss rules of the tabulated passbands wherever a
        per-walker bound shows they reproduce the full node sum to rounding
        (csrc/mbb_gaussrule.h; 4-9x fewer nodes per evaluation)."""
        return self._math_mode

    @math_mode.setter
    def math_mode(self, mode):
        self._math_mode = int(mode)
        self._dirty = True

    # ------------------------------------------------------------------ data
    def read_responses(self, responsefile=None, responsedir=None):
        """Load a response set; turns on passband integration (reference
        :139-156)."""
        self._responsewheel = response_set(responsefile, dir=responsedir)
        self._response_integrate = True
        self._dirty = True

    def set_phot(self, firstarg, flux, flux_unc):
        """Set photometry (reference :158-232).  ``firstarg``: response names
        when integrating passbands, else wavelengths in microns.  Wipes any
        covariance matrix."""
        if self._response_integrate:
            if not isinstance(firstarg[0], str):
                raise ValueError("Expecting response string name")
            self._responses = []
            for name in firstarg:
                if name not in self._responsewheel:
                    parts = name.split('_')
                    if len(parts) > 1 and parts[1].lower() in special_types:
                        self._responsewheel.add_special(name)
                    else:
                        raise ValueError("Unknown filter response "
                                         "{:s}".format(name))
                self._responses.append(self._responsewheel[name])
            self._response_names = [r.name for r in self._responses]
            self._wave = np.array([r.effective_wavelength
                                   for r in self._responses])
        else:
            self._wave = np.asarray(firstarg, dtype=np.float64)

        self._ndata = len(self._wave)
        if self._ndata == 0:
            raise ValueError("No elements in wavelength vector")
        self._flux = np.asarray(flux)
        self._flux_unc = np.asarray(flux_unc)
        if self._ndata != len(self._flux):
            raise ValueError("wave not same length as flux")
        if self._ndata != len(self._flux_unc):
            raise ValueError("wave not same length as flux_unc")
        self._ivar = 1.0 / self._flux_unc**2

        # latch: lambda0 can't be constrained beyond 3x the reddest point
        if not self._has_uplim[2]:
            self._has_uplim[2] = True
            self._uplim[2] = 3.0 * self._wave.max()

        self._data_read = True
        self._has_covmatrix = False
        self._dirty = True

    def read_phot(self, filename):
        """Three-column text photometry: wavelength [um] or response name,
        flux density [mJy], uncertainty [mJy] (reference :234-262)."""
        if not isinstance(filename, str):
            raise TypeError("filename must be string-like")
        rows = read_text_table(filename)
        if len(rows) == 0:
            raise IOError("No data read from %s" % filename)
        self.set_phot([r[0] for r in rows], [r[1] for r in rows],
                      [r[2] for r in rows])

    @property
    def data_read(self):
        return self._data_read

    @property
    def ndata(self):
        return self._ndata if self._data_read else 0

    @property
    def data_wave(self):
        return self._wave if self._data_read else None

    @property
    def response_names(self):
        return getattr(self, '_response_names', None)

    def has_response(self, name):
        return hasattr(self, '_responsewheel') and name in self._responsewheel

    def get_response(self, name):
        if not hasattr(self, '_responsewheel'):
            return None
        return self._responsewheel[name]

    @property
    def data_flux(self):
        return self._flux if self._data_read else None

    @property
    def data_flux_unc(self):
        """Uncertainties in mJy (sqrt of the covariance diagonal if one is
        set; they are then not used by the fit)."""
        if not self._data_read:
            return None
        if self._has_covmatrix:
            return np.sqrt(np.diag(self._covmatrix))
        return self._flux_unc

    def set_cov(self, covmatrix):
        """Flux covariance matrix in mJy^2 (reference :330-357).  The device
        uses its explicit inverse, like the reference (np.linalg.inv, :356)."""
        if not self._data_read:
            raise Exception("Can't set covariance matrix without photometry")
        covmatrix = np.asarray(covmatrix)
        if len(covmatrix.shape) != 2:
            raise ValueError("Covariance matrix is not 2 dimensional")
        if covmatrix.shape[0] != covmatrix.shape[1]:
            raise ValueError("Covariance matrix from is not square: "
                             "%d by %d" % covmatrix.shape)
        if covmatrix.shape[0] != self._ndata:
            raise ValueError("Covariance matrix doesn't have same number of "
                             "datapoints as photometry; {0:d} vs. "
                             "{1:d}".format(covmatrix.shape[0], self._ndata))
        self._covmatrix = covmatrix
        self._invcovmatrix = np.linalg.inv(self._covmatrix)
        self._has_covmatrix = True
        self._dirty = True

    def read_cov(self, filename, extn=0):
        """Covariance matrix from a FITS file (needs astropy), a ``.npy`` file
        or a whitespace-separated text matrix (reference :359-376)."""
        if not self._data_read:
            raise Exception("Can't read in covaraince matrix without phot")
        if filename.endswith(".npy"):
            self.set_cov(np.load(filename))
            return
        if filename.endswith((".txt", ".dat")):
            self.set_cov(np.loadtxt(filename))
            return
        try:
            import astropy.io.fits
        except ImportError:
            raise ImportError("reading a FITS covariance matrix needs astropy; "
                              "pass a .npy/.txt file or call set_cov()")
        hdu = astropy.io.fits.open(filename)
        self.set_cov(hdu[extn].data)

    @property
    def cov_solver(self):
        """How the chi-square of a full covariance is formed on the device (extension):
        ``"inverse"`` (default) -- diff . inv(C) . diff with the explicit inverse, the
        reference's arithmetic (likelihood.py:356, 823); ``"cholesky"`` -- the covariance is
        factored once on the host, C = L L', and the device forms |L^-1 diff|^2 by forward
        substitution (no inverse is ever formed).  The two agree to ~cond(C) * 1e-16."""
        return self._cov_solver

    @cov_solver.setter
    def cov_solver(self, value):
        if value not in ("inverse", "cholesky"):
            raise ValueError("cov_solver must be 'inverse' or 'cholesky'")
        if value != self._cov_solver:
            self._cov_solver = value
            self._dirty = True

    @property
    def has_data_covmatrix(self):
        return self._has_covmatrix

    @property
    def data_covmatrix(self):
        return self._covmatrix if self._has_covmatrix else None

    @property
    def data_invcovmatrix(self):
        return self._invcovmatrix if self._has_covmatrix else None

    # -------------------------------------------------------- limits / priors
    def get_paramindex(self, paramname):
        return self._param_order[paramname]

    def _pidx(self, param, table):
        return table[param.lower()] if isinstance(param, str) else int(param)

    def set_lowlim(self, param, val):
        self._lowlim[self._pidx(param, self._param_order)] = val
        self._dirty = True

    def lowlim(self, param):
        return self._lowlim[self._pidx(param, self._param_order)]

    @property
    def lowlims(self):
        return self._lowlim

    def set_uplim(self, param, val):
        i = self._pidx(param, self._limprior_order)
        self._has_uplim[i] = True
        self._uplim[i] = val
        self._dirty = True

    def has_uplim(self, param):
        return self._has_uplim[self._pidx(param, self._limprior_order)]

    def uplim(self, param):
        i = self._pidx(param, self._limprior_order)
        return self._uplim[i] if self._has_uplim[i] else None

    @property
    def has_uplims(self):
        return self._has_uplim

    @property
    def uplims(self):
        return self._uplim

    def set_gaussian_prior(self, param, mean, sigma):
        i = self._pidx(param, self._limprior_order)
        self._any_gprior = True
        self._has_gprior[i] = True
        self._gprior_mean[i] = float(mean)
        self._gprior_sigma[i] = float(sigma)
        self._gprior_ivar[i] = 1.0 / (float(sigma)**2)
        self._dirty = True

    @property
    def has_gpriors(self):
        return self._has_gprior

    @property
    def gprior_means(self):
        return self._gprior_mean

    @property
    def gprior_sigmas(self):
        return self._gprior_sigma

    @property
    def gprior_ivars(self):
        return self._gprior_ivar

    def has_gaussian_prior(self, param):
        return self._has_gprior[self._pidx(param, self._limprior_order)]

    def get_gaussian_prior(self, param):
        if not self._any_gprior:
            return None
        i = self._pidx(param, self._limprior_order)
        if not self._has_gprior[i]:
            return None
        return (self._gprior_mean[i], self._gprior_sigma[i])

    # ----------------------------------------------------------- device side
    @property
    def context(self):
        """The ``_native.Context`` this likelihood evaluates on (created on
        first use; raises if the CUDA library/device is missing)."""
        if self._ctx is None:
            dev = _native.default_device() if self._device is None else self._device
            self._ctx = _native.Context(dev)
        return self._ctx

    def band_tables(self):
        """(band_off, node_wave, node_weight, scalar_path) for the kernels.

        Passband mode: one band per response, nodes = its wavelength table,
        weights = sedmult*normfac (response.py:575-576 folded into one
        factor); a delta response is a single node evaluated through the
        reference's scalar path (response.py:565-569).  Plain-wavelength mode:
        one single-node band per datum, array path (likelihood.py:817).
        """
        if not self._data_read:
            raise Exception("Data not read")
        off, waves, weights, scalar = [0], [], [], []
        if self._response_integrate:
            for r in self._responses:
                wv, wt, isdelta = r.node_table()
                waves.append(np.asarray(wv, dtype=np.float64))
                weights.append(np.asarray(wt, dtype=np.float64))
                scalar.append(1 if isdelta else 0)
                off.append(off[-1] + len(wv))
        else:
            for w in self._wave:
                waves.append(np.array([w]))
                weights.append(np.array([1.0]))
                scalar.append(0)
                off.append(off[-1] + 1)
        return (np.array(off, dtype=np.int32), np.concatenate(waves),
                np.concatenate(weights), np.array(scalar, dtype=np.uint8))

    def _stage(self):
        """Push model flags, node tables, data, limits and priors to the device."""
        ctx = self.context
        ctx.set_model(self._wavenorm, self._opthin, self._noalpha)
        ctx.set_math_mode(self._math_mode)
        ctx.set_bands(*self.band_tables())
        if self._has_covmatrix and self._cov_solver == "cholesky":
            ctx.set_data(self._flux, chol=np.linalg.cholesky(np.asarray(self._covmatrix, dtype=np.float64)))
        elif self._has_covmatrix:
            ctx.set_data(self._flux, cinv=self._invcovmatrix)
        else:
            ctx.set_data(self._flux, ivar=self._ivar)
        ctx.set_priors(self._lowlim, self._has_uplim, self._uplim,
                       self._has_gprior, self._gprior_mean, self._gprior_ivar)
        self._dirty = False

    def _set_sed(self, pars):
        if len(pars) != 5:
            raise ValueError("pars is not of expected length 5")
        self._sed = modified_blackbody(pars[0], pars[1], pars[2], pars[3],
                                       pars[4], wavenorm=self._wavenorm,
                                       noalpha=self._noalpha,
                                       opthin=self._opthin)

    def get_sed(self, pars, wave):
        """Model SED [mJy] at ``wave`` [um] for one parameter vector
        (reference :770-788)."""
        self._set_sed(pars)
        return self._sed(wave)

    def evaluate(self, pars):
        """(lnlike[n], status[n]) for a block of parameter vectors, without
        raising on per-walker failures."""
        if not self._data_read:
            raise Exception("Data not read")
        if self._dirty:
            self._stage()
        P = np.ascontiguousarray(pars, dtype=np.float64)
        if P.ndim == 1:
            P = P.reshape(1, -1)
        if P.shape[1] != 5:
            raise ValueError("pars is not of expected length 5")
        return self._ctx.loglike(P)

    def __call__(self, pars):
        """log P(pars | data) incl. limits and priors (reference :790-834).

        ``pars`` of shape (5,) returns a float; shape (n, 5) returns
        ndarray[n].  ``-inf`` below a lower limit.  Failures that make the
        reference raise (bracketing, non-convergence, overflow) raise the
        same exception types here.
        """
        P = np.asarray(pars, dtype=np.float64)
        lnl, status = self.evaluate(P)
        _native.raise_for_status(status, P)
        if P.ndim == 1:
            return float(lnl[0])
        return lnl

    def as_pool(self):
        """A ``pool``-like object for emcee 2.x: ``pool.map(fn, rows)`` stacks
        the rows and evaluates them in ONE device call, whatever ``fn`` is."""
        return _BatchPool(self)


class _BatchPool(object):
    def __init__(self, like):
        self._like = like

    def map(self, fn, rows):
        rows = list(rows)
        if len(rows) == 0:
            return []
        return list(self._like(np.vstack([np.asarray(r, dtype=np.float64)
                                          for r in rows])))

    def close(self):
        pass
