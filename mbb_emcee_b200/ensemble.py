"""Affine-invariant ensemble sampler with emcee 2.x semantics, batched.

The reference drives ``emcee.EnsembleSampler`` through the emcee 1.x/2.x API
(reference mbb_emcee/mbb_fit.py:80-81, 525-550; results.py:154-155):
``EnsembleSampler(nwalkers, dim, lnpostfn, threads=)``, ``reset()``,
``run_mcmc(p0, N, rstate0=) -> (pos, lnprob, rstate)``, ``chain``,
``lnprobability``, ``acceptance_fraction``, ``acor``.  emcee is a third-party
dependency that is neither under /root/reference nor installed, so this module
restates its published algorithm (Goodman & Weare 2010 stretch move, as
scheduled by emcee 2.2.1: two half-ensembles per iteration, RNG consumption
per half-step ``rand(Ns)``, ``randint(Nc, Ns)``, ``rand(Ns)`` from a private
``numpy.random.RandomState``) with ONE difference that matters on a GPU: the
log-probability of a whole half-ensemble is requested in a single call
``lnpostfn(q[Ns, dim]) -> [Ns]`` instead of a Python ``map`` over rows.  The
chain produced for a given RNG state is the same either way.

``mbb_fitter`` uses the real ``emcee`` when ``MBB_B200_USE_EMCEE=1`` and it is
importable, otherwise this class.
"""
import numpy as np

__all__ = ["EnsembleSampler", "integrated_time"]


def _autocorr_function(x):
    """Normalised autocorrelation of x along axis 0 (FFT)."""
    x = np.atleast_1d(x)
    n = x.shape[0]
    nfft = 1
    while nfft < 2 * n:
        nfft *= 2
    f = np.fft.fft(x - np.mean(x, axis=0), n=nfft, axis=0)
    acf = np.fft.ifft(f * np.conjugate(f), axis=0)[:n].real
    return acf / acf[0]


def integrated_time(x, low=10, high=None, step=1, c=10):
    """Integrated autocorrelation time with Sokal's self-consistent window
    (the emcee 2.x estimator): smallest window M with M > c * tau(M)."""
    x = np.atleast_1d(x)
    size = 0.5 * x.shape[0]
    if c * low >= size:
        raise RuntimeError("The chain is too short")
    f = _autocorr_function(x)
    if high is None:
        high = int(size / c)
    tau = None
    for M in np.arange(low, high, step).astype(int):
        tau = 1 + 2 * np.sum(f[1:M], axis=0)
        if np.all(tau > 1.0) and M > c * tau.max():
            return tau
    raise RuntimeError("The chain is too short to reliably estimate the "
                       "autocorrelation time")


class EnsembleSampler(object):
    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, args=None, kwargs=None,
                 threads=1, pool=None, vectorize=None):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers <= 2 * dim:
            raise ValueError("The number of walkers needs to be more than "
                             "twice the dimension of your parameter space.")
        self.k = int(nwalkers)
        self.dim = int(dim)
        self.lnprobfn = lnpostfn
        self.a = float(a)
        self.args = [] if args is None else args
        self.kwargs = {} if kwargs is None else kwargs
        self.threads = int(threads)
        self.pool = pool
        # None = try a block call first and remember whether it worked
        self._vectorize = vectorize
        self._random = np.random.mtrand.RandomState()
        self.reset()

    # -- RNG state, as emcee exposes it
    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        try:
            self._random.set_state(state)
        except Exception:
            pass

    def reset(self):
        self.naccepted = np.zeros(self.k)
        self._chain = np.empty((self.k, 0, self.dim))
        self._lnprob = np.empty((self.k, 0))
        self.iterations = 0
        self._last_run_mcmc_result = None

    @property
    def chain(self):
        """Walker positions, shape (nwalkers, nsteps, dim)."""
        return self._chain

    @property
    def flatchain(self):
        s = self._chain.shape
        return self._chain.reshape(s[0] * s[1], s[2])

    @property
    def lnprobability(self):
        """Log-probability of every stored position, (nwalkers, nsteps)."""
        return self._lnprob

    @property
    def acceptance_fraction(self):
        return self.naccepted / self.iterations

    @property
    def acor(self):
        return integrated_time(np.mean(self._chain, axis=0))

    def get_autocorr_time(self, **kw):
        return integrated_time(np.mean(self._chain, axis=0), **kw)

    # -- log-probability of a block of positions
    def _get_lnprob(self, p):
        p = np.asarray(p, dtype=np.float64)
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite.")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN.")
        out = None
        if self.pool is not None:
            out = np.array([float(v) for v in self.pool.map(self.lnprobfn, list(p))])
        elif self._vectorize is not False:
            try:
                res = np.asarray(self.lnprobfn(p, *self.args, **self.kwargs), dtype=np.float64)
                if res.shape == (p.shape[0],):
                    out = res
                    self._vectorize = True
                elif self._vectorize:
                    raise ValueError("vectorized log-probability returned shape %s" % (res.shape,))
                else:
                    self._vectorize = False
            except (TypeError, ValueError, IndexError):
                if self._vectorize:
                    raise
                self._vectorize = False
        if out is None:
            out = np.array([float(self.lnprobfn(row, *self.args, **self.kwargs)) for row in p])
        if np.any(np.isnan(out)):
            raise ValueError("lnprob returned NaN.")
        return out

    # -- the move
    def sample(self, p0, lnprob0=None, rstate0=None, iterations=1, storechain=True):
        if rstate0 is not None:
            self.random_state = rstate0
        p = np.array(p0, dtype=np.float64)
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        half = self.k // 2
        first, second = slice(half), slice(half, self.k)
        lnprob = self._get_lnprob(p) if lnprob0 is None else np.array(lnprob0, dtype=np.float64)
        if np.any(np.isnan(lnprob)):
            raise ValueError("The initial lnprob was NaN.")
        if storechain:
            n0 = self._chain.shape[1]
            self._chain = np.concatenate(
                (self._chain, np.zeros((self.k, iterations, self.dim))), axis=1)
            self._lnprob = np.concatenate((self._lnprob, np.zeros((self.k, iterations))), axis=1)
        for i in range(int(iterations)):
            self.iterations += 1
            for S0, S1 in ((first, second), (second, first)):
                s = p[S0]
                c = p[S1]
                Ns, Nc = len(s), len(c)
                zz = ((self.a - 1.0) * self._random.rand(Ns) + 1) ** 2.0 / self.a
                rint = self._random.randint(Nc, size=(Ns,))
                q = c[rint] - zz[:, np.newaxis] * (c[rint] - s)
                newlnprob = self._get_lnprob(q)
                lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnprob - lnprob[S0]
                accept = lnpdiff > np.log(self._random.rand(len(lnpdiff)))
                if np.any(accept):
                    lnprob[S0][accept] = newlnprob[accept]
                    p[S0][accept] = q[accept]
                    self.naccepted[S0][accept] += 1
            if storechain:
                self._chain[:, n0 + i, :] = p
                self._lnprob[:, n0 + i] = lnprob
            yield p, lnprob, self.random_state

    def run_mcmc(self, pos0, N, rstate0=None, lnprob0=None, **kwargs):
        """N iterations from ``pos0``; returns (pos, lnprob, rstate)."""
        if pos0 is None:
            if self._last_run_mcmc_result is None:
                raise ValueError("Cannot have pos0=None if run_mcmc has never been called.")
            pos0 = self._last_run_mcmc_result[0]
            if lnprob0 is None:
                lnprob0 = self._last_run_mcmc_result[1]
            if rstate0 is None:
                rstate0 = self._last_run_mcmc_result[2]
        results = None
        for results in self.sample(pos0, lnprob0, rstate0, iterations=N, **kwargs):
            pass
        self._last_run_mcmc_result = results[:3]
        return results
