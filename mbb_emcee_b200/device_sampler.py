"""The stretch move of a single-source fit run ON the device.

``mbb_fitter`` drives an ensemble sampler through the emcee 2.x surface
(reference mbb_fit.py:80-81, 525-550; results.py:154-155): ``reset()``,
``run_mcmc(p0, N, rstate0=) -> (pos, lnprob, rstate)``, ``chain``,
``lnprobability``, ``acceptance_fraction``, ``acor``.  The host sampler
(``ensemble.EnsembleSampler``) keeps emcee's random stream and pays one
likelihood call (a 125-row host buffer, ~40-80 us) per half-step; this class
keeps that surface but hands the WHOLE run to ``mbb_ensemble_fit``: proposals,
log-probability, accept / reject and the chain record all stay on the GPU
(csrc/mbb_ensemble.cuh; delta-band FAST configurations run every iteration of
the call inside one kernel launch with the ensemble in shared memory) and the
chain comes back once.  BASELINE configs[0] (250 walkers, 50 + 1000 steps):
milliseconds instead of 0.24 s.

The random numbers are counter-based Philox draws, not emcee's Mersenne
Twister stream: the chain is statistically equivalent to, not bit-identical
with, the host sampler's (tests replay it on the host with the same draws and
compare posterior moments of the two samplers).  ``random_state`` is therefore
the pair (seed, iterations done), which is what continuing a run needs.
"""
import numpy as np

from . import _native
from .ensemble import integrated_time

__all__ = ["DeviceEnsembleSampler"]


class DeviceEnsembleSampler(object):
    def __init__(self, nwalkers, like, a=2.0, seed=0):
        if nwalkers % 2 != 0:
            raise ValueError("The number of walkers must be even.")
        if nwalkers <= 10:
            raise ValueError("The number of walkers needs to be more than "
                             "twice the dimension of your parameter space.")
        self.k = int(nwalkers)
        self.dim = 5
        self.a = float(a)
        self._like = like
        self._seed = int(seed)
        self._step = 0           # iterations drawn from this seed so far (the Philox counter)
        self._pinned = {}        # page-locked staging, kept between calls (allocation costs ms)
        self.reset()

    def _staging(self, name, shape, dtype=np.float64):
        """A page-locked array of ``shape``, carved from a buffer that is reused while it is
        large enough."""
        n = int(np.prod(shape))
        buf = self._pinned.get(name)
        if buf is None or buf.size < n or buf.dtype != np.dtype(dtype):
            buf = self._pinned[name] = _native.pinned_empty((max(n, 1),), dtype)
        return buf[:n].reshape(shape)

    # emcee exposes the generator state; here: what reproduces / continues the stream
    @property
    def random_state(self):
        return ("philox4x32-10", self._seed, self._step)

    @random_state.setter
    def random_state(self, state):
        if isinstance(state, tuple) and len(state) == 3 and state[0] == "philox4x32-10":
            self._seed, self._step = int(state[1]), int(state[2])
        elif isinstance(state, (int, np.integer)):
            self._seed, self._step = int(state), 0
        # (a Mersenne-Twister state of the host sampler has no meaning here: ignored,
        # as emcee ignores states it cannot set)

    def reset(self):
        self.naccepted = np.zeros(self.k)
        self._chain = np.empty((self.k, 0, self.dim))
        self._lnprob = np.empty((self.k, 0))
        self.iterations = 0
        self._last = None

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
        return self._lnprob

    @property
    def acceptance_fraction(self):
        return self.naccepted / self.iterations

    @property
    def acor(self):
        return integrated_time(np.mean(self._chain, axis=0))

    def get_autocorr_time(self, **kw):
        return integrated_time(np.mean(self._chain, axis=0), **kw)

    def run_mcmc(self, pos0, N, rstate0=None, lnprob0=None, storechain=True, **kwargs):
        """N iterations from ``pos0`` in one library call; returns (pos, lnprob, rstate)."""
        if rstate0 is not None:
            self.random_state = rstate0
        if pos0 is None:
            if self._last is None:
                raise ValueError("Cannot have pos0=None if run_mcmc has never been called.")
            pos0, lnprob0 = self._last[0], (self._last[1] if lnprob0 is None else lnprob0)
        p = np.array(pos0, dtype=np.float64)
        if p.shape != (self.k, self.dim):
            raise ValueError("p0 must have shape (nwalkers, dim)")
        if np.any(np.isinf(p)):
            raise ValueError("At least one parameter value was infinite.")
        if np.any(np.isnan(p)):
            raise ValueError("At least one parameter value was NaN.")
        N = int(N)
        like = self._like
        if not like.data_read:
            raise Exception("Data not read")
        if like._dirty:
            like._stage()
        ctx = like.context
        alloc = self._staging
        pos = alloc("pos", (1, self.k, 5))
        pos[0] = p
        lnp = alloc("lnp", (1, self.k))
        have = lnprob0 is not None
        if have:
            lnp[0] = np.asarray(lnprob0, dtype=np.float64)
        nacc = alloc("nacc", (1, self.k), np.int32)
        status = alloc("status", (1, self.k), np.int32)
        chain = alloc("chain", (N, 1, self.k, 5)) if storechain and N > 0 else None
        clnp = alloc("clnp", (N, 1, self.k)) if storechain and N > 0 else None
        # (nburn = 0: the whole call is "main run", so every iteration is counted and recorded)
        ctx.ensemble_fit_into(pos, lnp, 0, N, naccept=nacc, status=status, chain=chain, chain_lnprob=clnp,
                              have_lnprob=have, seed=self._seed, step0=self._step, a=self.a, thin=1)
        _native.raise_for_status(status.ravel(), pos.reshape(-1, 5))
        if np.any(np.isnan(lnp)):
            raise ValueError("lnprob returned NaN.")
        self._step += N
        self.iterations += N
        self.naccepted = self.naccepted + nacc[0]
        if chain is not None:
            self._chain = np.concatenate((self._chain, np.moveaxis(chain[:, 0], 0, 1)), axis=1)
            self._lnprob = np.concatenate((self._lnprob, clnp[:, 0].T), axis=1)
        self._last = (np.array(pos[0]), np.array(lnp[0]), self.random_state)
        return self._last
