"""Fit driver: owns the GPU likelihood and the ensemble sampler.

Drop-in for the reference's ``mbb_fitter`` (mbb_emcee/mbb_fit.py:13-563): same
constructor arguments, the same limit / prior / fixing calls,
``generate_initial_values`` and ``run``.  Everything about limits and priors
lives in the likelihood object; the calls of that name here are installed as
delegates (the loop at the end of this module) rather than written out one by one.  The one
structural change against the reference is that the sampler receives a *block*
log-probability: every half-ensemble proposal is one CUDA launch instead of
``nwalkers/2`` Python calls.  ``nthreads`` is accepted for API compatibility,
but a CUDA context cannot be forked into a multiprocessing pool, so values
other than 1 are refused.
"""
from __future__ import print_function

import os

import numpy as np

from .ensemble import EnsembleSampler
from .likelihood import likelihood

__all__ = ["mbb_fitter"]

_NPAR = 5
_MAX_REDRAWS = 100


def _make_sampler(nwalkers, like, where="host", seed=0):
    """The package's emcee-2.2 restatement by default; MBB_B200_USE_EMCEE=1 asks for an
    installed emcee instead (block log-probability either way); ``where="device"``: the
    whole run on the GPU (device_sampler.DeviceEnsembleSampler, counter-based random numbers)."""
    if where == "device":
        from .device_sampler import DeviceEnsembleSampler
        return DeviceEnsembleSampler(nwalkers, like, seed=seed)
    if where != "host":
        raise ValueError("sampler must be 'host' or 'device'")
    if os.environ.get("MBB_B200_USE_EMCEE", "0") == "1":
        try:
            import emcee
        except ImportError:
            emcee = None
        if emcee is not None:
            major = int(str(getattr(emcee, "__version__", "2")).split(".")[0])
            if major >= 3:
                return emcee.EnsembleSampler(nwalkers, _NPAR, like, vectorize=True)
            return emcee.EnsembleSampler(nwalkers, _NPAR, like, pool=like.as_pool())
    return EnsembleSampler(nwalkers, _NPAR, like, vectorize=True)


def _delegate(name):
    def call(self, *args):
        return getattr(self.like, name)(*args)
    call.__name__ = name
    call.__doc__ = "likelihood.%s (reference mbb_fit.py:176-360 forwards the same way)." % name
    return call


class mbb_fitter(object):
    """Fits a modified blackbody to photometry with an affine-invariant ensemble sampler."""

    _param_order = {'t': 0, 't/(1+z)': 0, 'beta': 1, 'lambda0': 2,
                    'lambda0*(1+z)': 2, 'lambda_0': 2, 'lambda_0*(1+z)': 2,
                    'alpha': 3, 'fnorm': 4, 'f500': 4, 'lambda_peak': 5,
                    'peaklam': 5}
    _parnames = np.array(['T/(1+z)', 'Beta', 'Lambda0*(1+z)', 'Alpha', 'Fnorm'])

    def __init__(self, nwalkers=250, photfile=None, covfile=None,
                 covextn=0, response=False, responsefile=None,
                 responsedir=None, wavenorm=500.0, noalpha=False,
                 opthin=False, nthreads=1, device=None, sampler="host", seed=0):
        """Arguments of reference mbb_fit.py:26-67, plus ``device`` (CUDA ordinal) and
        ``sampler``: "host" -- emcee's stretch move and random stream on the host, one device
        call per half-ensemble (chains equal emcee's for the same state); "device" -- the whole
        burn-in and main run inside the library (statistically equivalent chains from Philox
        draws keyed by ``seed``; BASELINE configs[0-2] in milliseconds)."""
        if int(nthreads) != 1:
            raise ValueError("nthreads != 1 is not supported: the whole "
                             "ensemble is evaluated by one GPU launch and a "
                             "CUDA context cannot be shared with forked workers")
        self._settings = dict(noalpha=noalpha, opthin=opthin, wavenorm=float(wavenorm),
                              nwalkers=int(nwalkers), nthreads=1)
        self.like = likelihood(photfile=photfile, covfile=covfile, covextn=covextn,
                               wavenorm=wavenorm, noalpha=noalpha, opthin=opthin,
                               response=response, responsefile=responsefile,
                               responsedir=responsedir, device=device)
        self.sampler = _make_sampler(self.nwalkers, self.like, sampler, seed)
        self._sampled = False
        self._fixed = [False] * _NPAR          # T, beta, lambda0, alpha, fnorm

    # settings are read-only, as in the reference
    noalpha = property(lambda self: self._settings["noalpha"])
    opthin = property(lambda self: self._settings["opthin"])
    wavenorm = property(lambda self: self._settings["wavenorm"])
    nwalkers = property(lambda self: self._settings["nwalkers"])
    nthreads = property(lambda self: self._settings["nthreads"])
    sampled = property(lambda self: self._sampled)
    fixed = property(lambda self: self._fixed)
    response_integrate = property(lambda self: self.like.response_integrate)
    # (kept for code that reached into the reference's private names)
    _noalpha, _opthin = noalpha, opthin
    _wavenorm, _nwalkers, _nthreads = wavenorm, nwalkers, nthreads

    # ------------------------------------------------------------------ data
    def read_data(self, photfile, covfile=None, covextn=0,
                  responsefile=None, responsedir=None):
        """Photometry (and covariance) from files; a responsefile turns on
        passband integration (reference mbb_fit.py:121-154)."""
        if responsefile is not None:
            self.like.read_responses(responsefile, responsedir=responsedir)
        self.like.read_phot(photfile)
        if covfile is not None:
            self.like.read_cov(covfile, extn=covextn)

    def set_data(self, wave, flux, flux_unc, covmatrix=None):
        self.like.set_phot(wave, flux, flux_unc)
        if covmatrix is not None:
            self.like.set_cov(covmatrix)

    # ------------------------------------------------- fixing, limits, priors
    def _pidx(self, param):
        return self._param_order[param.lower()] if isinstance(param, str) else int(param)

    def fix_param(self, param):
        self._fixed[self._pidx(param)] = True

    def unfix_param(self, param):
        self._fixed[self._pidx(param)] = False

    # ---------------------------------------------------------- initial ball
    def _ball_centres(self, initvals, initsigma):
        """Centre of each parameter's starting ball: the requested value when it obeys the
        limits, else 2 sigma inside the violated limit (the middle of a range narrower than
        4 sigma); reference mbb_fit.py:398-443."""
        lo = np.array([self.lowlim(i) for i in range(_NPAR)], dtype=np.float64)
        bounded = np.array([bool(self.has_uplim(i)) for i in range(_NPAR)])
        hi = np.array([self.uplim(i) if bounded[i] else np.inf for i in range(_NPAR)], dtype=np.float64)
        want = np.asarray(initvals, dtype=np.float64)
        sig = np.asarray(initsigma, dtype=np.float64)
        below, above = want < lo, bounded & (want > hi)
        stuck = np.asarray(self._fixed) & (below | above)
        if stuck.any():
            raise ValueError("Some fixed parameters outside limits: "
                             "{:s}".format(', '.join(self._parnames[stuck.nonzero()[0]])))
        centre = want.copy()
        for i in np.nonzero(below | above)[0]:
            if not bounded[i]:
                centre[i] = lo[i] + 2 * sig[i]
                continue
            span = hi[i] - lo[i]
            if span <= 0:
                raise ValueError("Limits on parameter {:d} cross".format(int(i)))
            if 2.0 * sig[i] >= span:
                centre[i] = lo[i] + 0.5 * span
            else:
                centre[i] = lo[i] + 2 * sig[i] if below[i] else hi[i] - 2 * sig[i]
        return centre, lo, hi

    def generate_initial_values(self, initvals, initsigma):
        """nwalkers x 5 starting positions obeying the limits (reference
        mbb_fit.py:362-479): per free parameter a Gaussian ball around its centre
        (``_ball_centres``) whose out-of-range members are redrawn until none is left; a fixed
        parameter gets its one value.  Draws come from the global ``np.random`` stream in the
        reference's order -- ``randn(nwalkers)`` per free parameter, then ``randn(#rejected)``
        per redraw round (:449, :461) -- so a seeded run starts from the reference's ensemble."""
        if len(initvals) != _NPAR:
            raise ValueError("Initial values not expected length")
        if len(initsigma) != _NPAR:
            raise ValueError("Initial sigma values not expected length")
        centre, lo, hi = self._ball_centres(initvals, initsigma)
        nw = self.nwalkers
        p0 = np.empty((nw, _NPAR))
        for i in range(_NPAR):
            if self._fixed[i]:
                p0[:, i] = centre[i]
                continue
            col = initsigma[i] * np.random.randn(nw) + centre[i]
            for _ in range(_MAX_REDRAWS + 1):
                reject = np.nonzero((col > hi[i]) | (col < lo[i]))[0]
                if reject.size == 0:
                    break
                col[reject] = initsigma[i] * np.random.randn(reject.size) + centre[i]
            else:
                if np.any((col > hi[i]) | (col < lo[i])):
                    raise Exception("Too many iterations initializing param {:d}".format(i))
            p0[:, i] = col
        return p0

    # -------------------------------------------------------------------- run
    def _used_params(self):
        """Indices whose start values are checked: lambda0 / alpha only when the model uses
        them (reference mbb_fit.py:497-513)."""
        skip = ({2} if self.opthin else set()) | ({3} if self.noalpha else set())
        return [i for i in range(_NPAR) if i not in skip]

    def _check_start(self, p0):
        for i in self._used_params():
            if self.has_uplim(i) and p0[:, i].max() > self.uplim(i):
                raise ValueError("Upper limit initial value violation for "
                                 "{:s}".format(self._parnames[i]))
            if p0[:, i].min() < self.lowlim(i):
                raise ValueError("Lower limit initial value violation for "
                                 "{:s}".format(self._parnames[i]))

    def _report(self, nburn):
        print("  Fit complete")
        print("   Mean acceptance fraction:", np.mean(self.sampler.acceptance_fraction))
        try:
            acor = self.sampler.acor
        except (ImportError, RuntimeError):
            return
        print("   Autocorrelation time: ")
        print("    Number of burn in steps ({:d}) should be larger than these".format(nburn))
        for i in self._used_params():
            print("\t{:<9s} {:f}".format(("T", "beta", "lambda0", "alpha", "fnorm")[i] + ":", acor[i]))

    def run(self, nburn, nsteps, p0, verbose=False):
        """Burn in from ``p0``, forget it, then the main chain from where the burn-in ended,
        continuing its random state -- the sequence of reference mbb_fit.py:481-563
        (reset / run_mcmc(p0, nburn) / reset / run_mcmc(pos, nsteps, rstate0))."""
        if not self.like.data_read:
            raise Exception("Data not read, needed to do fit")
        for what, n in (("burn in", nburn), ("main chain", nsteps)):
            if n <= 0:
                raise ValueError("Invalid (non-positive) number of {:s} steps: {:d}".format(what, n))
        self._check_start(np.asarray(p0))
        say = print if verbose else (lambda *a: None)
        say("Starting fit" + ("\n  Using response integration" if self.response_integrate else ""))
        self._sampled = False
        self.sampler.reset()
        say("  Doing burn in with {:d} steps".format(nburn))
        burn = self.sampler.run_mcmc(p0, nburn)
        self.sampler.reset()
        say("  Doing main chain with {:d} steps".format(nsteps))
        self.sampler.run_mcmc(burn[0], nsteps, rstate0=burn[2])
        self._sampled = True
        if verbose:
            self._report(nburn)


for _name in ("set_lowlim", "lowlim", "set_uplim", "has_uplim", "uplim", "set_gaussian_prior",
              "has_gaussian_prior", "get_gaussian_prior"):
    setattr(mbb_fitter, _name, _delegate(_name))
del _name
