"""Fit driver: owns the GPU likelihood and the ensemble sampler.

Host-side mirror of the reference's ``mbb_fitter`` (reference
mbb_emcee/mbb_fit.py:13-563): same constructor, pass-through setters,
``generate_initial_values`` and ``run``.  The one structural change is that the
sampler receives a *block* log-probability: every half-ensemble proposal is
one CUDA launch instead of ``nwalkers/2`` Python calls.  ``nthreads`` is
accepted for API compatibility but a CUDA context cannot be forked into a
multiprocessing pool, so values other than 1 are refused.
"""
from __future__ import print_function

import os

import numpy as np

from .ensemble import EnsembleSampler
from .likelihood import likelihood

__all__ = ["mbb_fitter"]


def _make_sampler(nwalkers, like):
    if os.environ.get("MBB_B200_USE_EMCEE", "0") == "1":
        try:
            import emcee
        except ImportError:
            emcee = None
        if emcee is not None:
            major = int(str(getattr(emcee, "__version__", "2")).split(".")[0])
            if major >= 3:
                return emcee.EnsembleSampler(nwalkers, 5, like, vectorize=True)
            return emcee.EnsembleSampler(nwalkers, 5, like, pool=like.as_pool())
    return EnsembleSampler(nwalkers, 5, like, vectorize=True)


class mbb_fitter(object):
    """ Does fit"""

    _param_order = {'t': 0, 't/(1+z)': 0, 'beta': 1, 'lambda0': 2,
                    'lambda0*(1+z)': 2, 'lambda_0': 2, 'lambda_0*(1+z)': 2,
                    'alpha': 3, 'fnorm': 4, 'f500': 4, 'lambda_peak': 5,
                    'peaklam': 5}

    _parnames = np.array(['T/(1+z)', 'Beta', 'Lambda0*(1+z)',
                          'Alpha', 'Fnorm'])

    def __init__(self, nwalkers=250, photfile=None, covfile=None,
                 covextn=0, response=False, responsefile=None,
                 responsedir=None, wavenorm=500.0, noalpha=False,
                 opthin=False, nthreads=1, device=None):
        """Same parameters as reference mbb_fit.py:26-67 (+ ``device``)."""
        self._noalpha = noalpha
        self._opthin = opthin
        self._wavenorm = float(wavenorm)
        self._nwalkers = int(nwalkers)
        self._nthreads = int(nthreads)
        if self._nthreads != 1:
            raise ValueError("nthreads != 1 is not supported: the whole "
                             "ensemble is evaluated by one GPU launch and a "
                             "CUDA context cannot be shared with forked workers")
        self.like = likelihood(photfile=photfile, covfile=covfile,
                               covextn=covextn, wavenorm=wavenorm,
                               noalpha=noalpha, opthin=opthin,
                               response=response, responsefile=responsefile,
                               responsedir=responsedir, device=device)
        self.sampler = _make_sampler(self._nwalkers, self.like)
        self._sampled = False
        # order: T, beta, lambda0, alpha, fnorm
        self._fixed = [False, False, False, False, False]

    @property
    def noalpha(self):
        return self._noalpha

    @property
    def opthin(self):
        return self._opthin

    @property
    def wavenorm(self):
        return self._wavenorm

    @property
    def nwalkers(self):
        return self._nwalkers

    @property
    def nthreads(self):
        return self._nthreads

    @property
    def sampled(self):
        return self._sampled

    @property
    def fixed(self):
        return self._fixed

    @property
    def response_integrate(self):
        return self.like.response_integrate

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
        return self._param_order[param.lower()] if isinstance(param, str) \
            else int(param)

    def fix_param(self, param):
        self._fixed[self._pidx(param)] = True

    def unfix_param(self, param):
        self._fixed[self._pidx(param)] = False

    def set_lowlim(self, param, val):
        self.like.set_lowlim(param, val)

    def lowlim(self, param):
        return self.like.lowlim(param)

    def set_uplim(self, param, val):
        self.like.set_uplim(param, val)

    def has_uplim(self, param):
        return self.like.has_uplim(param)

    def uplim(self, param):
        return self.like.uplim(param)

    def set_gaussian_prior(self, param, mean, sigma):
        self.like.set_gaussian_prior(param, mean, sigma)

    def has_gaussian_prior(self, param):
        return self.like.has_gaussian_prior(param)

    def get_gaussian_prior(self, param):
        return self.like.get_gaussian_prior(param)

    # ---------------------------------------------------------- initial ball
    def generate_initial_values(self, initvals, initsigma):
        """nwalkers x 5 starting positions obeying the limits (reference
        mbb_fit.py:362-479): Gaussian ball, out-of-range draws redrawn;
        centres outside the limits are moved 2 sigma inside (or to the middle
        of a narrow range); fixed parameters get one identical value.  Uses
        the global ``np.random`` stream like the reference (:449, :461)."""
        if len(initvals) != 5:
            raise ValueError("Initial values not expected length")
        if len(initsigma) != 5:
            raise ValueError("Initial sigma values not expected length")

        outside = [False] * 5
        for i, val in enumerate(initvals):
            if val < self.lowlim(i):
                outside[i] = True
            elif self.has_uplim(i) and val > self.uplim(i):
                outside[i] = True

        fixed_and_outside = np.logical_and(self._fixed, outside)
        if fixed_and_outside.any():
            bad = ', '.join(self._parnames[fixed_and_outside.nonzero()[0]])
            raise ValueError("Some fixed parameters outside limits: "
                             "{:s}".format(bad))

        centre = np.zeros(5)
        for i in range(5):
            if not outside[i]:
                centre[i] = initvals[i]
            elif self.has_uplim(i):
                span = self.uplim(i) - self.lowlim(i)
                if span <= 0:
                    raise ValueError("Limits on parameter {:d} cross".format(i))
                if 2.0 * initsigma[i] >= span:
                    centre[i] = self.lowlim(i) + 0.5 * span
                elif initvals[i] < self.lowlim(i):
                    centre[i] = self.lowlim(i) + 2 * initsigma[i]
                else:
                    centre[i] = self.uplim(i) - 2 * initsigma[i]
            else:
                centre[i] = self.lowlim(i) + 2 * initsigma[i]

        p0 = np.zeros((self._nwalkers, 5))
        maxiters = 100
        for i in range(5):
            if self._fixed[i]:
                p0[:, i] = centre[i] * np.ones(self._nwalkers)
                continue
            lo = self.lowlim(i)
            has_hi = self.has_uplim(i)
            hi = self.uplim(i)

            def outliers(v):
                if has_hi:
                    return np.logical_or(v > hi, v < lo).nonzero()[0]
                return np.nonzero(v < lo)[0]

            pvec = initsigma[i] * np.random.randn(self._nwalkers) + centre[i]
            bad = outliers(pvec)
            iters = 0
            while len(bad) > 0:
                pvec[bad] = initsigma[i] * np.random.randn(len(bad)) + centre[i]
                iters += 1
                bad = outliers(pvec)
                if iters > maxiters:
                    raise Exception("Too many iterations initializing param "
                                    "{:d}".format(i))
            p0[:, i] = pvec
        return p0

    # -------------------------------------------------------------------- run
    def run(self, nburn, nsteps, p0, verbose=False):
        """Burn in, reset, main chain (reference mbb_fit.py:481-563)."""
        if not self.like.data_read:
            raise Exception("Data not read, needed to do fit")
        if verbose:
            print("Starting fit")
            if self.response_integrate:
                print("  Using response integration")

        for i in range(5):
            if i == 2 and self._opthin:
                continue
            if i == 3 and self._noalpha:
                continue
            if self.has_uplim(i) and p0[:, i].max() > self.uplim(i):
                raise ValueError("Upper limit initial value violation for "
                                 "{:s}".format(self._parnames[i]))
            if p0[:, i].min() < self.lowlim(i):
                raise ValueError("Lower limit initial value violation for "
                                 "{:s}".format(self._parnames[i]))

        self.sampler.reset()
        self._sampled = False
        if nburn <= 0:
            raise ValueError("Invalid (non-positive) number of burn in steps: "
                             "{:d}".format(nburn))
        if verbose:
            print("  Doing burn in with {:d} steps".format(nburn))
        burn = self.sampler.run_mcmc(p0, nburn)
        pos, rstate = burn[0], burn[2]

        self.sampler.reset()
        if nsteps <= 0:
            raise ValueError("Invalid (non-positive) number of main chain "
                             "steps: {:d}".format(nsteps))
        if verbose:
            print("  Doing main chain with {:d} steps".format(nsteps))
        self.sampler.run_mcmc(pos, nsteps, rstate0=rstate)
        self._sampled = True

        if verbose:
            print("  Fit complete")
            print("   Mean acceptance fraction:",
                  np.mean(self.sampler.acceptance_fraction))
            try:
                acor = self.sampler.acor
                print("   Autocorrelation time: ")
                print("    Number of burn in steps ({:d}) should be larger "
                      "than these".format(nburn))
                print("\tT:        {:f}".format(acor[0]))
                print("\tbeta:     {:f}".format(acor[1]))
                if not self._opthin:
                    print("\tlambda0:  {:f}".format(acor[2]))
                if not self._noalpha:
                    print("\talpha:    {:f}".format(acor[3]))
                print("\tfnorm:    {:f}".format(acor[4]))
            except (ImportError, RuntimeError):
                pass
