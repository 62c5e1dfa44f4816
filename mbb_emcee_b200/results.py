"""Fit results and chain post-processing on the GPU.

Host-side mirror of the reference's ``mbb_results`` (reference
mbb_emcee/results.py:29-1262).  The three chain-wide computations --
``compute_peaklambda``, ``compute_lir``, ``compute_dustmass`` (reference
:570-581, :627-674, :746-801), serial Python double loops in the reference --
run as CUDA kernels over the whole chain (``mbb_chain_post`` of the C ABI),
including the sequential allclose-dedupe of ``_map_chain`` (:553-566).

Reference behaviours kept on purpose:
  * L_IR and the peak wavelength are computed with the *default* wavenorm=500,
    whatever the fit used (:575-577, :1323-1325); dust mass uses the fit's.
  * ``compute_peaklambda`` always treats the SED as optically thick with alpha
    (:574-580 never forwards opthin/noalpha).
Not carried over: HDF5 (de)serialisation (h5py is not a dependency; use
``save``/``load`` = ``.npz`` with the same key names) and the astropy cosmology
lookup (pass ``lumdist`` in Mpc, or have astropy installed).
"""
from __future__ import print_function, division

import copy

import numpy as np

from . import _native
from .mbb_fit import mbb_fitter
from .modified_blackbody import modified_blackbody

__all__ = ["mbb_results"]

_MPC_CM = 3.0856775814913673e24


class mbb_results(object):
    """Holds results of fit"""

    _param_order = {'t': 0, 't/(1+z)': 0, 'beta': 1, 'lambda0': 2,
                    'lambda0*(1+z)': 2, 'lambda_0': 2, 'lambda_0*(1+z)': 2,
                    'alpha': 3, 'fnorm': 4, 'f500': 4}

    def __init__(self, fit=None, h5file=None, redshift=None, lumdist=None,
                 cosmo_type='WMAP9', device=None, devices=None):
        if h5file is not None:
            raise NotImplementedError("HDF5 results files are not supported "
                                      "(h5py is not a dependency); use "
                                      "mbb_results.load(<npz>)")
        self._fitset = False
        self._has_lir = False
        self._lir_min = None
        self._lir_max = None
        self._has_dustmass = False
        self._kappa = None
        self._kappa_wave = None
        self._has_peaklambda = False
        self._device = device
        self._ctx = None
        # devices=[0, 1, ...]: the chain's walker rows are cut into one contiguous shard per GPU
        # (the per-walker repeat scan of _map_chain stays local), each driven from its own host
        # thread through its own context, results landing in disjoint rows of one page-locked
        # array -- no collective
        self._devices = None if devices is None else [int(d) for d in devices]
        self._shard_ctx = {}

        self._z = None if redshift is None else float(redshift)
        if lumdist is None:
            self._has_lumdist = False
        else:
            self._has_lumdist = True
            self._lumdist = float(getattr(lumdist, "value", lumdist))   # Mpc
        if cosmo_type is None:
            raise ValueError("Cosmology type must not be none -- "
                             "maybe you should just use the default")
        self._cosmo_type = cosmo_type
        if fit is not None:
            self.process_fit(fit)

    # ------------------------------------------------------------------ setup
    def process_fit(self, fit):
        """Take over chain, data and settings from a finished ``mbb_fitter``
        (reference results.py:107-180)."""
        if not isinstance(fit, mbb_fitter):
            raise ValueError("Input is not of type mbb_fit")
        self._fitset = True
        self._noalpha = fit.noalpha
        self._opthin = fit.opthin
        self._wavenorm = fit.wavenorm
        self._nwalkers = fit.nwalkers
        self._lowlim = fit.like.lowlims
        self._has_uplim = fit.like.has_uplims
        self._uplim = fit.like.uplims
        self._has_gprior = fit.like.has_gpriors
        self._gprior_mean = fit.like.gprior_means
        self._gprior_sigma = fit.like.gprior_sigmas
        self._gprior_ivar = fit.like.gprior_ivars
        self._fixed = fit._fixed
        self._ndata = fit.like.ndata
        self._response_integrate = fit.like.response_integrate
        if self._response_integrate:
            self._responsewheel = fit.like._responsewheel
        elif hasattr(self, '_responsewheel'):
            del self._responsewheel
        self._data_wave = fit.like.data_wave
        self._data_flux = fit.like.data_flux
        self._data_flux_unc = fit.like.data_flux_unc
        self._has_covmatrix = bool(fit.like.has_data_covmatrix)
        if self._has_covmatrix:
            self._covmatrix = fit.like.data_covmatrix
            self._invcovmatrix = fit.like.data_invcovmatrix
        if self._device is None:
            self._device = fit.like._device
        self.set_chain(fit.sampler.chain, fit.sampler.lnprobability)

    def set_chain(self, chain, lnprobability):
        """Install a chain [nwalkers, nsteps, 5] and its log-probabilities."""
        self.chain = np.asarray(chain, dtype=np.float64)
        self.lnprobability = np.asarray(lnprobability, dtype=np.float64)
        self.par_central_values = np.array([self.par_cen(i) for i in range(5)])
        flat = self.lnprobability.argmax()
        idx = np.unravel_index(flat, self.lnprobability.shape)
        self._best_fit = (self.chain[idx[0], idx[1], :],
                          self.lnprobability[idx[0], idx[1]], idx)
        self._has_lir = self._has_dustmass = self._has_peaklambda = False
        self._lir_min = self._lir_max = self._kappa = self._kappa_wave = None
        for name in ('lir', 'dustmass', 'peaklambda'):
            if hasattr(self, name):
                delattr(self, name)

    @classmethod
    def from_chain(cls, chain, lnprobability=None, wavenorm=500.0, noalpha=False,
                   opthin=False, redshift=None, lumdist=None, device=None, devices=None):
        """Post-process a bare chain (extension; used for sharded chains)."""
        self = cls(redshift=redshift, lumdist=lumdist, device=device, devices=devices)
        self._fitset = True
        self._noalpha, self._opthin = bool(noalpha), bool(opthin)
        self._wavenorm = float(wavenorm)
        chain = np.asarray(chain, dtype=np.float64)
        self._nwalkers = chain.shape[0]
        self._lowlim = np.array([1, 0.1, 1, 0.1, 1e-3])
        self._has_uplim = [False] * 6
        self._uplim = np.full(6, np.inf)
        self._has_gprior = [False] * 6
        self._gprior_mean = np.zeros(6)
        self._gprior_sigma = np.zeros(6)
        self._gprior_ivar = np.ones(6)
        self._fixed = [False] * 5
        self._ndata = 0
        self._response_integrate = False
        self._has_covmatrix = False
        self._data_wave = self._data_flux = self._data_flux_unc = None
        if lnprobability is None:
            lnprobability = np.zeros(chain.shape[:2])
        self.set_chain(chain, lnprobability)
        return self

    @property
    def context(self):
        if self._ctx is None:
            dev = _native.default_device() if self._device is None else self._device
            self._ctx = _native.Context(dev)
        return self._ctx

    # ------------------------------------------------------------- properties
    @property
    def redshift(self):
        return self._z

    @property
    def opthin(self):
        return self._opthin if self._fitset else None

    @property
    def noalpha(self):
        return self._noalpha if self._fitset else None

    @property
    def wavenorm(self):
        return self._wavenorm if self._fitset else None

    @property
    def response_integrate(self):
        return self._response_integrate if self._fitset else None

    @property
    def cosmo_type(self):
        return self._cosmo_type

    @property
    def lumdist(self):
        """Luminosity distance in Mpc (float)."""
        if self._has_lumdist:
            return self._lumdist
        if self._z is None:
            raise Exception("Need redshift to be set to compute lumdist")
        if self._z <= -1:
            raise ValueError("Redshift is less than -1: {:f}".format(self._z))
        try:
            import astropy.cosmology
        except ImportError:
            raise ImportError("no lumdist given and astropy is not available "
                              "for the cosmology lookup; pass lumdist= [Mpc]")
        cosmo = getattr(astropy.cosmology, self._cosmo_type)
        dl = cosmo.luminosity_distance(self._z)
        return float(getattr(dl, "value", dl))

    @property
    def best_fit(self):
        return self._best_fit if self._fitset else None

    @property
    def best_fit_chisq(self):
        return -2.0 * self._best_fit[1] if self._fitset else None

    def best_fit_sed(self, wave):
        if not self._fitset:
            return None
        p = self._best_fit[0]
        sed = modified_blackbody(p[0], p[1], p[2], p[3], p[4],
                                 wavenorm=self._wavenorm, noalpha=self._noalpha,
                                 opthin=self._opthin)
        return sed(wave)

    @property
    def data(self):
        if not self._fitset:
            return None
        return (self._data_wave, self._data_flux, self._data_flux_unc)

    @property
    def covmatrix(self):
        if not self._fitset or not self._has_covmatrix:
            return None
        return self._covmatrix

    # ------------------------------------------------------- chain statistics
    def _parcen_internal(self, array, percentile, lowlim=None, uplim=None):
        """Mean and +/- percentile half-widths (reference results.py:314-369)."""
        if not self._fitset:
            raise Exception("Fit not available")
        pcnt = float(percentile)
        if pcnt < 0 or pcnt > 100:
            raise ValueError("Invalid percentile {:f}".format(pcnt))
        pval = 0.5 * (100 - pcnt)
        arr = array
        if lowlim is not None or uplim is not None:
            keep = np.ones(arr.shape, dtype=bool)
            if lowlim is not None:
                keep &= arr >= float(lowlim)
            if uplim is not None:
                keep &= arr <= float(uplim)
            if not keep.any():
                raise Exception("No elements survive lower/upper limit clipping")
            arr = arr[keep]
        mn = arr.mean()
        perc = np.percentile(arr, [pval, 100 - pval])
        return np.array([mn, perc[1] - mn, mn - perc[0]])

    def _paridx(self, param):
        if isinstance(param, str):
            return self._param_order[param.lower()]
        idx = int(param)
        if idx < 0 or idx > 5:
            raise ValueError("invalid parameter index {:d}".format(idx))
        return idx

    def parameter_chain(self, param):
        if not self._fitset:
            return None
        return self.chain[:, :, self._paridx(param)].flatten()

    def par_cen(self, param, percentile=68.3, lowlim=None, uplim=None):
        if not self._fitset:
            return None
        if percentile <= 0 or percentile >= 100.0:
            raise ValueError("percentile needs to be between 0 and 100")
        return self._parcen_internal(self.parameter_chain(param), percentile,
                                     lowlim=lowlim, uplim=uplim)

    def par_lowlim(self, param, percentile=68.3):
        if not self._fitset:
            return None
        if percentile <= 0 or percentile >= 100.0:
            raise ValueError("percentile needs to be between 0 and 100")
        return np.percentile(self.parameter_chain(self._paridx(param)),
                             100 - percentile)

    def par_uplim(self, param, percentile=68.3):
        if not self._fitset:
            return None
        if percentile <= 0 or percentile >= 100.0:
            raise ValueError("percentile needs to be between 0 and 100")
        return np.percentile(self.parameter_chain(self._paridx(param)),
                             percentile)

    # ------------------------------------------------------------ peak lambda
    @property
    def has_peaklambda(self):
        return self._has_peaklambda

    @property
    def peaklambda_chain(self):
        return self.peaklambda.flatten() if self._has_peaklambda else None

    def peaklambda_cen(self, percentile=68.3, lowlim=None, uplim=None):
        if not self._has_peaklambda:
            return None
        return self._parcen_internal(self.peaklambda.flatten(), percentile,
                                     lowlim=lowlim, uplim=uplim)

    def _post(self, which, wavenorm, opthin, noalpha, lir_method=None, **kw):
        if not self._devices:
            ctx = self.context
            ctx.set_model(wavenorm, opthin, noalpha)
            if lir_method is not None:
                ctx.set_lir_method(lir_method)
            pk, lir, dm, status = ctx.chain_post(self.chain, which, **kw)
            _native.raise_for_status(status, self.chain)
            return pk, lir, dm
        import threading
        from .sharding import shard_range
        chain = np.ascontiguousarray(self.chain, dtype=np.float64)
        nw, ns = chain.shape[0], chain.shape[1]
        world = len(self._devices)
        out = [(_native.pinned_empty((nw, ns)) if which & bit else None) for bit in (1, 2, 4)]
        status = _native.pinned_empty((nw, ns), np.int32)
        errors = [None] * world

        def shard(r):
            try:
                lo, hi = shard_range(nw, r, world)
                if hi <= lo:
                    return
                ctx = self._shard_ctx.get(r)
                if ctx is None:
                    ctx = self._shard_ctx[r] = _native.Context(self._devices[r])
                ctx.set_model(wavenorm, opthin, noalpha)
                if lir_method is not None:
                    ctx.set_lir_method(lir_method)
                ctx.chain_post_into(chain[lo:hi], which, *[None if o is None else o[lo:hi] for o in out],
                                    status=status[lo:hi], **kw)
            except BaseException as exc:
                errors[r] = exc

        threads = [threading.Thread(target=shard, args=(r,)) for r in range(world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for exc in errors:
            if exc is not None:
                raise exc
        _native.raise_for_status(status, chain)
        return tuple(out)

    def compute_peaklambda(self):
        """Observer-frame wavelength of peak f_nu [um] for every chain sample."""
        if not self._fitset:
            raise Exception("Fit results not loaded")
        self.peaklambda = self._post(1, 500.0, False, False)[0]
        self._has_peaklambda = True

    # ------------------------------------------------------------------- L_IR
    @property
    def has_lir(self):
        return self._has_lir

    @property
    def lir_wavelength(self):
        return (self._lir_min, self._lir_max)

    @property
    def lir_chain(self):
        return self.lir.flatten() if self._has_lir else None

    def lir_cen(self, percentile=68.3, lowlim=None, uplim=None):
        if not self._has_lir:
            return None
        return self._parcen_internal(self.lir.flatten(), percentile,
                                     lowlim=lowlim, uplim=uplim)

    def compute_lir(self, wavemin=8.0, wavemax=1000.0, maxidx=None, method="quadpack"):
        """L_IR in 10^12 L_sun for every chain sample (reference :627-674).

        ``method`` (extension): "quadpack" (default) replays the reference's
        scipy.integrate.quad call and returns the reference's numbers to
        ~1e-15; "gauss" is a fixed-rule quadrature accurate to ~1e-15 of the
        true integral (the reference itself is only good to ~1e-8 there)."""
        if not self._fitset:
            raise Exception("Fit results not loaded")
        if self._z is None:
            raise Exception("Redshift must be set to compute L_IR")
        if maxidx is not None:
            raise NotImplementedError("maxidx is dead code in the reference "
                                      "(results.py:547-550 assigns into a tuple)")
        self._lir_min = float(wavemin)
        self._lir_max = float(wavemax)
        if self._lir_min <= 0:
            raise ValueError("Invalid wavemin: {:f}".format(self._lir_min))
        if self._lir_max <= 0:
            raise ValueError("Invalid wavemax: {:f}".format(self._lir_max))
        if self._lir_min > self._lir_max:
            self._lir_min, self._lir_max = self._lir_max, self._lir_min
        self.lir = self._post(2, 500.0, self._opthin, self._noalpha, lir_method=method, z=self._z,
                              dl_mpc=self.lumdist, lir_min=self._lir_min,
                              lir_max=self._lir_max)[1]
        self._has_lir = True

    # -------------------------------------------------------------- dust mass
    @property
    def has_dustmass(self):
        return self._has_dustmass

    @property
    def dust_kappa(self):
        return self._kappa

    @property
    def dust_kappa_wavelength(self):
        return self._kappa_wave

    @property
    def dustmass_chain(self):
        return self.dustmass.flatten() if self._has_dustmass else None

    def dustmass_cen(self, percentile=68.3, lowlim=None, uplim=None):
        if not self._has_dustmass:
            return None
        return self._parcen_internal(self.dustmass.flatten(), percentile,
                                     lowlim=lowlim, uplim=uplim)

    def compute_dustmass(self, kappa=2.64, kappa_wave=125.0, maxidx=None):
        """Dust mass in 10^8 M_sun for every chain sample (reference :746-801)."""
        if not self._fitset:
            raise Exception("Fit not processed")
        if self._z is None:
            raise Exception("Redshift must be set to compute dust mass")
        if maxidx is not None:
            raise NotImplementedError("maxidx is dead code in the reference")
        self._kappa = float(kappa)
        self._kappa_wave = float(kappa_wave)
        if self._kappa <= 0:
            raise ValueError("Invalid (non-positive) kappa "
                             "{:f}".format(self._kappa))
        if self._kappa_wave <= 0:
            raise ValueError("Invalid (non-positive) kappa wavelength "
                             "{:f}".format(self._kappa_wave))
        self.dustmass = self._post(4, self._wavenorm, self._opthin, self._noalpha,
                                   z=self._z, dl_mpc=self.lumdist,
                                   kappa=self._kappa,
                                   kappa_wave=self._kappa_wave)[2]
        self._has_dustmass = True

    # ---------------------------------------------------------- predicted flux
    def _predict_flux(self, spec, maxidx=None):
        """Predicted flux density [mJy] for every chain sample at a wavelength
        (float, microns) or through a named response of the fit (reference
        results.py:895-944; like the reference, the SED is built with the
        default wavenorm=500)."""
        if maxidx is not None:
            raise NotImplementedError("maxidx is dead code in the reference")
        if isinstance(spec, str):
            if not self._response_integrate:
                raise Exception("Asked for response integration, but no response "
                                "functions available from original fit")
            if spec not in self._responsewheel:
                raise ValueError("Do not have response function matching "
                                 "{:s}".format(spec))
            wv, wt, isdelta = self._responsewheel[spec].node_table()
        else:
            w = float(spec)
            if w <= 0:
                raise ValueError("Invalid wavelength {:f}".format(w))
            wv, wt, isdelta = np.array([w]), np.array([1.0]), True
        ctx = self.context
        ctx.set_model(500.0, self._opthin, self._noalpha)
        ctx.set_bands(np.array([0, len(wv)], dtype=np.int32), wv, wt,
                      np.array([1 if isdelta else 0], dtype=np.uint8))
        flux, status = ctx.chain_flux(self.chain, 0)
        _native.raise_for_status(status, self.chain)
        return flux

    def predflux_cen(self, spec, percentile=68.3, maxidx=None, lowlim=None, uplim=None):
        """Central confidence interval of the predicted flux (reference :946-985)."""
        if not self._fitset:
            return None
        return self._parcen_internal(self._predict_flux(spec, maxidx).flatten(), percentile,
                                     lowlim=lowlim, uplim=uplim)

    # ---------------------------------------------------------------- choices
    def choice(self, nsamples=1, getpeaklambda=False, getlir=False,
               getdustmass=False):
        """Random draws from the chain (reference results.py:803-893)."""
        if not self._fitset or nsamples == 0:
            return None
        if getpeaklambda and not self._has_peaklambda:
            raise Exception("Peak lambda not computed")
        if getlir and not self._has_lir:
            raise Exception("LIR not computed")
        if getdustmass and not self._has_dustmass:
            raise Exception("Dustmass not computed")
        extras = [a for flag, a in ((getpeaklambda, 'peaklambda'), (getlir, 'lir'),
                                    (getdustmass, 'dustmass')) if flag]
        if nsamples == 1:
            iw = np.random.randint(0, self.chain.shape[0])
            it = np.random.randint(0, self.chain.shape[1])
            return (self.chain[iw, it, :],) + tuple(getattr(self, a)[iw, it]
                                                    for a in extras)
        pars = np.empty((nsamples, 5), dtype=np.float64)
        outs = [np.empty(nsamples, dtype=np.float64) for _ in extras]
        for i in range(nsamples):
            iw = np.random.randint(0, self.chain.shape[0])
            it = np.random.randint(0, self.chain.shape[1])
            pars[i, :] = self.chain[iw, it, :]
            for o, a in zip(outs, extras):
                o[i] = getattr(self, a)[iw, it]
        return (pars,) + tuple(outs)

    # ------------------------------------------------------------------- I/O
    def save(self, filename):
        """Write the results to ``.npz`` using the reference's HDF5 key names
        (results.py:987-1158) flattened with '/'."""
        if not self._fitset:
            raise Exception("Fit not processed")
        d = {"z": np.nan if self._z is None else self._z,
             "Noalpha": self._noalpha, "Opthin": self._opthin,
             "Nwalkers": self._nwalkers, "Wavenorm": self._wavenorm,
             "Lowlim": self._lowlim, "HasUplim": np.array(self._has_uplim),
             "Uplim": self._uplim, "HasGaussianPrior": np.array(self._has_gprior),
             "GaussianPriorMean": self._gprior_mean,
             "GaussianPriorSigma": self._gprior_sigma,
             "GaussianPriorIVar": self._gprior_ivar,
             "ResponseIntegrate": self._response_integrate,
             "Fixed": np.array(self._fixed), "Ndata": self._ndata,
             "Chain/Chain": self.chain, "Chain/LogLike": self.lnprobability,
             "Chain/ParamCentralValues": self.par_central_values,
             "Chain/BestFitParams": self._best_fit[0],
             "Chain/BestFitLogLike": self._best_fit[1],
             "Chain/BestFitIndex": np.array(self._best_fit[2])}
        if self._has_lumdist:
            d["LumDist"] = self._lumdist
        if self._data_wave is not None:
            d["Data/Wave"] = np.asarray(self._data_wave)
            d["Data/FluxDensity"] = np.asarray(self._data_flux)
            d["Data/FluxDensityUnc"] = np.asarray(self._data_flux_unc)
        if self._has_covmatrix:
            d["Data/Covmatrix"] = self._covmatrix
            d["Data/InvCovmatrix"] = self._invcovmatrix
        if self._has_lir:
            d["Ancillary/Lir"] = self.lir
            d["Ancillary/LirMin"], d["Ancillary/LirMax"] = self._lir_min, self._lir_max
        if self._has_dustmass:
            d["Ancillary/Dustmass"] = self.dustmass
            d["Ancillary/DustKappa"] = self._kappa
            d["Ancillary/DustKappaWave"] = self._kappa_wave
        if self._has_peaklambda:
            d["Ancillary/PeakLambda"] = self.peaklambda
        np.savez_compressed(filename, **d)

    @classmethod
    def load(cls, filename, device=None):
        with np.load(filename) as f:
            z = float(f["z"])
            self = cls.from_chain(f["Chain/Chain"], f["Chain/LogLike"],
                                  wavenorm=float(f["Wavenorm"]),
                                  noalpha=bool(f["Noalpha"]), opthin=bool(f["Opthin"]),
                                  redshift=None if np.isnan(z) else z,
                                  lumdist=float(f["LumDist"]) if "LumDist" in f else None,
                                  device=device)
            self._lowlim = f["Lowlim"]
            self._has_uplim = list(f["HasUplim"])
            self._uplim = f["Uplim"]
            self._has_gprior = list(f["HasGaussianPrior"])
            self._gprior_mean = f["GaussianPriorMean"]
            self._gprior_sigma = f["GaussianPriorSigma"]
            self._gprior_ivar = f["GaussianPriorIVar"]
            self._fixed = list(f["Fixed"])
            self._ndata = int(f["Ndata"])
            if "Data/Wave" in f:
                self._data_wave = f["Data/Wave"]
                self._data_flux = f["Data/FluxDensity"]
                self._data_flux_unc = f["Data/FluxDensityUnc"]
            if "Ancillary/Lir" in f:
                self.lir, self._has_lir = f["Ancillary/Lir"], True
                self._lir_min = float(f["Ancillary/LirMin"])
                self._lir_max = float(f["Ancillary/LirMax"])
            if "Ancillary/Dustmass" in f:
                self.dustmass, self._has_dustmass = f["Ancillary/Dustmass"], True
                self._kappa = float(f["Ancillary/DustKappa"])
                self._kappa_wave = float(f["Ancillary/DustKappaWave"])
            if "Ancillary/PeakLambda" in f:
                self.peaklambda, self._has_peaklambda = f["Ancillary/PeakLambda"], True
        return self

    def __str__(self):
        if not self._fitset:
            return "<Uninitialized mbb_results object>"
        lines = []

        def par_line(i, tag, unit):
            if self._fixed[i]:
                return "{:s}: {:0.2f} (fixed){:s}".format(
                    tag, self.chain[:, :, i].mean(), " " + unit if unit else "")
            c = self.par_central_values[i]
            s = "{:s}: {:0.2f} +{:0.2f} -{:0.2f} (low lim: {:0.2f}".format(
                tag, c[0], c[1], c[2], self._lowlim[i])
            if self._has_uplim[i]:
                s += " upper lim: {:0.2f}".format(self._uplim[i])
            if self._has_gprior[i]:
                s += " prior: {:0.2f} {:0.2f}".format(self._gprior_mean[i],
                                                      self._gprior_sigma[i])
            return s + ")" + (" " + unit if unit else "")

        lines.append(par_line(0, "T/(1+z)", "[K]"))
        lines.append(par_line(1, "beta", ""))
        lines.append(par_line(4, "fnorm", "[mJy]"))
        lines.append("Optically thin case assumed" if self._opthin
                     else par_line(2, "lambda0 (1+z)", "[um]"))
        lines.append("Alpha not used" if self._noalpha else par_line(3, "alpha", ""))
        if self._has_uplim[5] or self._has_gprior[5]:
            s = "Lambda_peak prior"
            if self._has_uplim[5]:
                s += " upper lim: {:0.2f}".format(self._uplim[5])
            if self._has_gprior[5]:
                s += " prior: {:0.2f} {:0.2f}".format(self._gprior_mean[5],
                                                      self._gprior_sigma[5])
            lines.append(s)
        if self.has_peaklambda:
            lines.append("Lambda peak: {:0.1f} +{:0.1f} -{:0.1f} "
                         "[um]".format(*self.peaklambda_cen()))
        if self.has_lir:
            lines.append("L_IR({:0.1f} to {:0.1f}um): {:0.2f} +{:0.2f} -{:0.2f} "
                         "[10^12 L_sun]".format(*(self.lir_wavelength +
                                                  tuple(self.lir_cen()))))
        if self.has_dustmass:
            lines.append("M_d(kappa={0:0.2f}, lam={1:0.1f}um): {2:0.2f} "
                         "+{3:0.2f} -{4:0.2f} [10^8 M_sun]".format(
                             self.dust_kappa, self.dust_kappa_wavelength,
                             *self.dustmass_cen()))
        lines.append("Number of data points: {:d}".format(self._ndata))
        lines.append("ChiSquare of best fit point: "
                     "{:0.2f}".format(self.best_fit_chisq))
        return "\n".join(lines)
