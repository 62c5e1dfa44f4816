"""Many sources, one GPU-resident ensemble sampler.

``batch_fitter`` is the many-source counterpart of ``mbb_fitter`` (SURVEY.md 8f
row 1; BASELINE configs[4]: 1e5 sources x 512 walkers).  It has no class of its
own in the reference -- there one would loop ``mbb_fitter`` over sources -- but
it performs exactly what each of those fits asks emcee for (reference
mbb_fit.py:525-542): stretch-move iterations of an ensemble of ``nwalkers``
walkers in the 5 parameters, all sources sharing the model flags, band set,
limits and priors, each with its own photometry.  Proposals, log-probability
and accept/reject all run on the device (``mbb_ensemble_run``); only the
photometry goes in and the final ensemble / summary comes out.
"""
import numpy as np

from . import _native
from .likelihood import likelihood

__all__ = ["batch_fitter"]


class batch_fitter(object):
    def __init__(self, nwalkers=512, wavenorm=500.0, noalpha=False, opthin=False,
                 response=False, responsefile=None, responsedir=None, device=None):
        if nwalkers % 2 or nwalkers <= 10:
            raise ValueError("nwalkers must be even and > 10")
        self._nwalkers = int(nwalkers)
        # a likelihood object carries the shared settings (bands, limits, priors)
        self.like = likelihood(wavenorm=wavenorm, noalpha=noalpha, opthin=opthin,
                               response=response, responsefile=responsefile,
                               responsedir=responsedir, device=device)
        self._fixed = [False] * 5
        self._steps_done = 0
        self._staged = False

    @property
    def nwalkers(self):
        return self._nwalkers

    @property
    def nsources(self):
        return self._flux.shape[0]

    # limits / priors: same calls as mbb_fitter, applied to every source
    def set_lowlim(self, param, val):
        self.like.set_lowlim(param, val)
        self._staged = False

    def set_uplim(self, param, val):
        self.like.set_uplim(param, val)
        self._staged = False

    def set_gaussian_prior(self, param, mean, sigma):
        self.like.set_gaussian_prior(param, mean, sigma)
        self._staged = False

    def fix_param(self, param):
        idx = self.like._param_order[param.lower()] if isinstance(param, str) else int(param)
        self._fixed[idx] = True

    def set_data(self, bands, flux, flux_unc=None, covmatrix=None):
        """bands: wavelengths [um] or response names (shared by all sources);
        flux[nsrc][nb] in mJy; flux_unc[nsrc][nb] or covmatrix[nsrc][nb][nb]."""
        flux = np.atleast_2d(np.asarray(flux, dtype=np.float64))
        nb = flux.shape[1]
        # the template likelihood sees source 0 (sets bands, lambda0 auto-limit)
        unc0 = np.ones(nb) if flux_unc is None else np.atleast_2d(flux_unc)[0]
        self.like.set_phot(bands, flux[0], unc0)
        self._flux = flux
        if covmatrix is not None:
            cov = np.asarray(covmatrix, dtype=np.float64).reshape(flux.shape[0], nb, nb)
            self._cinv = np.linalg.inv(cov)
            self._ivar = None
        else:
            self._ivar = 1.0 / np.atleast_2d(np.asarray(flux_unc, dtype=np.float64))**2
            self._cinv = None
        self._staged = False

    def _stage(self):
        like = self.like
        ctx = like.context
        ctx.set_model(like.wavenorm, like.opthin, like.noalpha)
        ctx.set_math_mode(like.math_mode)
        ctx.set_bands(*like.band_tables())
        if self._cinv is not None:
            ctx.set_data(self._flux, cinv=self._cinv)
        else:
            ctx.set_data(self._flux, ivar=self._ivar)
        ctx.set_priors(like.lowlims, like.has_uplims, like.uplims, like.has_gpriors,
                       like.gprior_means, like.gprior_ivars)
        self._staged = True

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
        p0 = np.empty((nsrc, nw, 5))
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

    def run(self, nburn, nsteps, p0, seed=0, a=2.0):
        """Burn in, then the main run (reference mbb_fit.py:524-543 per source).
        Returns a dict: pos[nsrc][nw][5], lnprob[nsrc][nw],
        acceptance_fraction[nsrc][nw] (main run), status[nsrc][nw]."""
        if not self._staged:
            self._stage()
        p0 = np.asarray(p0, dtype=np.float64)
        if p0.shape != (self.nsources, self._nwalkers, 5):
            raise ValueError("p0 must have shape (nsources, nwalkers, 5)")
        ctx = self.like.context
        pos, lnp, nacc, st = p0, None, None, None
        done = 0
        if nburn > 0:
            pos, lnp, nacc, st = ctx.ensemble_run(pos, nburn, seed=seed, step0=0, a=a)
            _native.raise_for_status(st.ravel(), pos.reshape(-1, 5))
            done = nburn
        pos, lnp, nacc, st = ctx.ensemble_run(pos, nsteps, seed=seed, step0=done, a=a, lnprob=lnp)
        _native.raise_for_status(st.ravel(), pos.reshape(-1, 5))
        self._steps_done = done + nsteps
        return {"pos": pos, "lnprob": lnp,
                "acceptance_fraction": nacc / float(max(nsteps, 1)), "status": st}
