"""Many sources, GPU-resident ensemble samplers, one fit per source.

``batch_fitter`` is the many-source counterpart of ``mbb_fitter`` (SURVEY.md 8f
row 1; BASELINE configs[4]: 1e5 sources x 512 walkers).  It has no class of its
own in the reference -- there one would loop ``mbb_fitter`` + ``mbb_results``
over sources -- but it performs exactly what each of those fits does (reference
mbb_fit.py:524-543): burn-in, reset, main run of an ensemble of ``nwalkers``
walkers in the 5 parameters, all sources sharing the model flags, band set,
limits and priors, each with its own photometry; and it returns what
``mbb_results`` would derive from each chain (reference results.py:314-431):
per-parameter mean and spread, the best-fitting sample, the acceptance
fraction, and -- on request -- the thinned chain itself.  Proposals,
log-probability, accept/reject and the posterior accumulation all run on the
device (``mbb_ensemble_fit``); the photometry and the starting ensembles go
in, the final ensembles and the summaries come out.

With ``devices=[0, 1, ...]`` the sources are cut into contiguous shards, one per
GPU, each driven from its own host thread through its own context (SURVEY.md
8e): no collective, every GPU copies its results straight into its slice of
the shared page-locked output arrays.  The random numbers of a source depend on
its global index only, so the fit is the same however it is sharded.
"""
import threading

import numpy as np

from . import _native
from .likelihood import likelihood
from .sharding import shard_range

__all__ = ["batch_fitter", "batch_fit_result"]

_PARNAMES = ("T", "beta", "lambda0", "alpha", "fnorm")


class batch_fit_result(dict):
    """What ``batch_fitter.run`` returns.  A dict (``pos``, ``lnprob``,
    ``acceptance_fraction``, ``status``, ``stats`` and optionally ``chain``,
    ``chain_lnprob``) with the per-source posterior summaries as properties;
    arrays are indexed [source][parameter] in the order T, beta, lambda0,
    alpha, fnorm."""

    def _s(self, lo, n=5):
        return self["stats"][:, lo:lo + n]

    @property
    def nsamples(self):
        """Samples behind the summaries (walkers x recorded iterations), per source."""
        return self["stats"][:, _native.FS_N]

    @property
    def mean(self):
        return self._s(_native.FS_MEAN)

    @property
    def std(self):
        """Posterior standard deviation (ddof=1)."""
        n = self.nsamples[:, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.sqrt(self._s(_native.FS_M2) / (n - 1.0))

    @property
    def min(self):
        return self._s(_native.FS_MIN)

    @property
    def max(self):
        return self._s(_native.FS_MAX)

    @property
    def best_fit(self):
        """The recorded sample of largest log-probability (reference
        results.py best_fit: the chain's argmax of lnprobability)."""
        return self._s(_native.FS_BEST)

    @property
    def best_lnprob(self):
        return self["stats"][:, _native.FS_BESTLNP]

    @property
    def mean_acceptance(self):
        """emcee's acceptance_fraction averaged over walkers, main run only."""
        return self["stats"][:, _native.FS_ACC]

    def par_cen(self, param, percentile=68.3):
        """[nsrc][3]: mean, upper and lower uncertainty of one parameter, the
        triple ``mbb_results.par_cen`` returns (reference results.py:399-431).
        From the chain's percentiles when a chain was recorded, else from the
        moments (Gaussian equivalent of the central ``percentile`` interval)."""
        i = _PARNAMES.index(param) if isinstance(param, str) and param in _PARNAMES else \
            (likelihood._param_order[param.lower()] if isinstance(param, str) else int(param))
        mean = self.mean[:, i]
        if "chain" in self:
            ch = self["chain"][:, :, :, i]                       # [nrec][nsrc][nw]
            ch = np.moveaxis(ch, 1, 0).reshape(ch.shape[1], -1)
            pval = 0.5 * (100.0 - float(percentile))
            lo, hi = np.percentile(ch, [pval, 100.0 - pval], axis=1)
            return np.stack([mean, hi - mean, mean - lo], axis=1)
        from scipy.special import erfinv
        k = np.sqrt(2.0) * erfinv(float(percentile) / 100.0)
        sd = self.std[:, i] * k
        return np.stack([mean, sd, sd], axis=1)


class batch_fitter(object):
    def __init__(self, nwalkers=512, wavenorm=500.0, noalpha=False, opthin=False,
                 response=False, responsefile=None, responsedir=None, device=None, devices=None):
        if nwalkers % 2 or nwalkers <= 10:
            raise ValueError("nwalkers must be even and > 10")
        self._nwalkers = int(nwalkers)
        # a likelihood object carries the shared settings (bands, limits, priors)
        self.like = likelihood(wavenorm=wavenorm, noalpha=noalpha, opthin=opthin,
                               response=response, responsefile=responsefile,
                               responsedir=responsedir, device=device)
        if devices is None:
            devices = [_native.default_device() if device is None else int(device)]
        self._devices = [int(d) for d in devices]
        if not self._devices:
            raise ValueError("devices must be a non-empty list of CUDA ordinals")
        # one context per shard (an ordinal may appear more than once: its shards then share the GPU);
        # the batch's own contexts -- self.like keeps its own for single calls
        self._ctxs = {}
        self._fixed = [False] * 5
        self._steps_done = 0
        self._flux = None
        self._ivar = self._cinv = self._chol = None

    @property
    def nwalkers(self):
        return self._nwalkers

    @property
    def nsources(self):
        return self._flux.shape[0]

    @property
    def devices(self):
        return list(self._devices)

    # limits / priors: same calls as mbb_fitter, applied to every source
    def set_lowlim(self, param, val):
        self.like.set_lowlim(param, val)

    def set_uplim(self, param, val):
        self.like.set_uplim(param, val)

    def set_gaussian_prior(self, param, mean, sigma):
        self.like.set_gaussian_prior(param, mean, sigma)

    def fix_param(self, param):
        idx = self.like._param_order[param.lower()] if isinstance(param, str) else int(param)
        self._fixed[idx] = True

    def set_data(self, bands, flux, flux_unc=None, covmatrix=None, cholesky=False):
        """bands: wavelengths [um] or response names (shared by all sources);
        flux[nsrc][nb] in mJy; exactly one of flux_unc[nsrc][nb] and
        covmatrix[nsrc][nb][nb].  ``cholesky``: factor the covariances once on the host and let
        the device form |L^-1 diff|^2 (mbb_set_data_chol) instead of using explicit inverses."""
        if (flux_unc is None) == (covmatrix is None):
            raise ValueError("give exactly one of flux_unc and covmatrix")
        flux = np.atleast_2d(np.asarray(flux, dtype=np.float64))
        nsrc, nb = flux.shape
        if covmatrix is not None:
            cov = np.asarray(covmatrix, dtype=np.float64).reshape(nsrc, nb, nb)
            unc0 = np.sqrt(np.diagonal(cov[0]))
            self._chol = np.linalg.cholesky(cov) if cholesky else None
            self._cinv = None if cholesky else np.linalg.inv(cov)
            self._ivar = None
        else:
            unc = np.atleast_2d(np.asarray(flux_unc, dtype=np.float64))
            if unc.shape != flux.shape:
                raise ValueError("flux_unc must have the shape of flux")
            unc0 = unc[0]
            self._ivar = 1.0 / unc**2
            self._cinv = self._chol = None
        # the template likelihood sees source 0 (sets bands, lambda0 auto-limit)
        self.like.set_phot(bands, flux[0], unc0)
        self._flux = flux

    def _context(self, shard):
        ctx = self._ctxs.get(shard)
        if ctx is None:
            ctx = self._ctxs[shard] = _native.Context(self._devices[shard])
        return ctx

    def _stage(self, shard=0, lo=0, hi=None):
        """Everything the kernels need for sources [lo, hi) onto the shard's device.  Done at
        every run: the tables are small, and nothing can be stale."""
        if self._flux is None:
            raise Exception("Data not set, needed to do fit")
        hi = self.nsources if hi is None else hi
        like = self.like
        ctx = self._context(shard)
        ctx.set_model(like.wavenorm, like.opthin, like.noalpha)
        ctx.set_math_mode(like.math_mode)
        ctx.set_bands(*like.band_tables())
        if self._chol is not None:
            ctx.set_data(self._flux[lo:hi], chol=self._chol[lo:hi])
        elif self._cinv is not None:
            ctx.set_data(self._flux[lo:hi], cinv=self._cinv[lo:hi])
        else:
            ctx.set_data(self._flux[lo:hi], ivar=self._ivar[lo:hi])
        ctx.set_priors(like.lowlims, like.has_uplims, like.uplims, like.has_gpriors,
                       like.gprior_means, like.gprior_ivars)
        return ctx

    def generate_initial_values(self, initvals, initsigma, seed=None):
        """[nsrc][nwalkers][5] Gaussian balls obeying the limits, as
        mbb_fitter.generate_initial_values does per source (reference
        mbb_fit.py:362-479); ``initvals`` is [5] or [nsrc][5]."""
        rng = np.random.RandomState(seed)
        nsrc, nw = self.nsources, self._nwalkers
        init = np.broadcast_to(np.asarray(initvals, dtype=np.float64), (nsrc, 5)).copy()
        sig = np.asarray(initsigma, dtype=np.float64)
        low = np.asarray(self.like.lowlims, dtype=np.float64)
        up = np.where(self.like.has_uplims[:5], self.like.uplims[:5], np.inf)
        # centres outside the limits are moved 2 sigma inside (or to the middle)
        for i in range(5):
            span = up[i] - low[i]
            below, above = init[:, i] < low[i], init[:, i] > up[i]
            if self._fixed[i] and (below.any() or above.any()):
                raise ValueError("Some fixed parameters outside limits")
            if np.isfinite(span) and 2.0 * sig[i] >= span:
                init[below | above, i] = low[i] + 0.5 * span
            else:
                init[below, i] = low[i] + 2 * sig[i]
                init[above, i] = up[i] - 2 * sig[i]
        p0 = _native.pinned_empty((nsrc, nw, 5)) if self._can_pin() else np.empty((nsrc, nw, 5))
        for i in range(5):
            if self._fixed[i]:
                p0[:, :, i] = init[:, None, i]
                continue
            v = init[:, None, i] + sig[i] * rng.standard_normal((nsrc, nw))
            for _ in range(100):
                bad = (v < low[i]) | (v > up[i])
                nbad = int(bad.sum())
                if nbad == 0:
                    break
                centres = np.broadcast_to(init[:, None, i], v.shape)[bad]
                v[bad] = centres + sig[i] * rng.standard_normal(nbad)
            else:
                raise Exception("Too many iterations initializing param {:d}".format(i))
            p0[:, :, i] = v
        return p0

    @staticmethod
    def _can_pin():
        try:
            return _native.load_library().mbb_device_count() > 0
        except _native.MBBNativeError:
            return False

    def run(self, nburn, nsteps, p0, seed=0, a=2.0, thin=1, chain=False):
        """Burn in, reset, main run (reference mbb_fit.py:524-543, per source), and the
        posterior summaries of every source (reference results.py:314-431).

        p0[nsrc][nwalkers][5]: starting ensembles (``generate_initial_values``).
        thin: every thin-th iteration of the main run is recorded (summaries and chain).
        chain: also return chain[nsteps//thin][nsrc][nwalkers][5] and chain_lnprob.
        Returns a ``batch_fit_result``."""
        if nburn < 0 or nsteps <= 0:
            raise ValueError("nburn must be >= 0 and nsteps positive")
        nsrc, nw = self.nsources, self._nwalkers
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (nsrc, nw, 5):
            raise ValueError("p0 must have shape (nsources, nwalkers, 5)")
        thin = max(int(thin), 1)
        nrec = int(nsteps) // thin
        alloc = _native.pinned_empty
        pos = alloc((nsrc, nw, 5))
        pos[...] = p0
        out = batch_fit_result(pos=pos, lnprob=alloc((nsrc, nw)),
                               naccept=alloc((nsrc, nw), np.int32), status=alloc((nsrc, nw), np.int32),
                               stats=alloc((nsrc, _native.FIT_NSTATS)))
        if chain:
            out["chain"] = alloc((nrec, nsrc, nw, 5))
            out["chain_lnprob"] = alloc((nrec, nsrc, nw))
        world = len(self._devices)
        errors = [None] * world

        def shard(r):
            try:
                lo, hi = shard_range(nsrc, r, world)
                if hi <= lo:
                    return
                ctx = self._stage(r, lo, hi)
                ctx.ensemble_fit_into(out["pos"][lo:hi], out["lnprob"][lo:hi], nburn, nsteps,
                                      naccept=out["naccept"][lo:hi], status=out["status"][lo:hi],
                                      stats=out["stats"][lo:hi],
                                      chain=out["chain"][:, lo:] if chain else None,
                                      chain_lnprob=out["chain_lnprob"][:, lo:] if chain else None,
                                      chain_nsrc=nsrc, seed=seed, src0=lo, a=a, thin=thin)
            except BaseException as exc:        # re-raised on the calling thread
                errors[r] = exc

        if world == 1:
            shard(0)
        else:
            threads = [threading.Thread(target=shard, args=(r,)) for r in range(world)]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
        for exc in errors:
            if exc is not None:
                raise exc
        _native.raise_for_status(out["status"].ravel(), out["pos"].reshape(-1, 5))
        self._steps_done = nburn + nsteps
        out["acceptance_fraction"] = out["naccept"] / float(nsteps)
        return out
