"""Synthetic photometry configurations (BASELINE.json ``configs``).

Pure numpy, no model evaluation of its own: callers hand in a function that
returns the noise-free model fluxes (the CUDA likelihood machinery in the
product and in bench.py; the reference itself in tests/golden/make_golden.py),
so the same seeded noise realisation is used everywhere.

All randomness is ``numpy.random.RandomState(seed)`` (legacy MT19937, stable
across numpy versions).
"""
import numpy as np

__all__ = ["CONFIGS", "noisy_photometry", "walker_cloud", "cfg3_covariance",
           "random_walk_chain"]

P0_SIGMA = np.array([2.0, 0.2, 100.0, 0.3, 5.0])   # reference run_mbb_emcee.py:285-291

CONFIGS = {
    # 6 delta-function bands, optically thin, no alpha (BASELINE configs[0])
    "cfg1": dict(seed=101, response=False,
                 bands=[70.0, 100.0, 160.0, 250.0, 350.0, 500.0],
                 truth=(12.0, 1.8, 1300.0, 4.0, 30.0), opthin=True, noalpha=True,
                 wavenorm=500.0),
    # full passband integration, optically thick + alpha join (configs[1])
    "cfg2": dict(seed=102, response=True,
                 bands=["PACS_100um", "PACS_160um", "SPIRE_250um", "SPIRE_350um",
                        "SPIRE_500um", "SCUBA2_850um"],
                 truth=(14.0, 1.8, 400.0, 3.0, 30.0), opthin=False, noalpha=False,
                 wavenorm=500.0),
    # interferometric specials + full covariance + limits/priors (configs[2])
    "cfg3": dict(seed=103, response=True,
                 bands=["SPIRE_250um", "SPIRE_350um", "SPIRE_500um", "SCUBA2_850um",
                        "ALMA_alma_345", "ALMA_alma_230", "SMA_dsb_230_8_2",
                        "PdBI_box_135_3.6"],
                 truth=(14.0, 1.8, 400.0, 3.0, 30.0), opthin=False, noalpha=False,
                 wavenorm=500.0,
                 uplims=[("T", 60.0), ("lambda_peak", 400.0)],
                 gpriors=[("beta", 1.8, 0.3), ("lambda_peak", 300.0, 60.0)]),
    # chain post-processing (configs[3])
    "cfg4": dict(seed=104, truth=(14.0, 1.8, 400.0, 3.0, 30.0), opthin=False,
                 noalpha=False, wavenorm=500.0, z=2.0, lumdist=1.6e4,
                 kappa=2.64, kappa_wave=125.0, lir=(8.0, 1000.0)),
    # many-source batch (configs[4]); primary band set = cfg1's
    "cfg5": dict(seed=105, response=False,
                 bands=[70.0, 100.0, 160.0, 250.0, 350.0, 500.0],
                 opthin=True, noalpha=True, wavenorm=500.0,
                 nsources=100000, nwalkers=512),
    # the same batch on cfg2's tabulated band set (cfg5 "secondary")
    "cfg5p": dict(seed=106, response=True,
                  bands=["PACS_100um", "PACS_160um", "SPIRE_250um", "SPIRE_350um",
                         "SPIRE_500um", "SCUBA2_850um"],
                  truth=(12.0, 1.8, 1300.0, 4.0, 30.0), opthin=True, noalpha=True,
                  wavenorm=500.0, nsources=8192, nwalkers=512),
}


def noisy_photometry(model_flux, rng):
    """sigma_i = max(0.1*model_i, 1 mJy); data = model + N(0, sigma)."""
    model_flux = np.asarray(model_flux, dtype=np.float64)
    unc = np.maximum(0.1 * model_flux, 1.0)
    flux = model_flux + unc * rng.standard_normal(model_flux.shape)
    return flux, unc


def cfg3_covariance(flux, unc, spire_idx=(0, 1, 2), rho=0.3, calfrac=0.05):
    """C = diag(sigma^2) + calfrac^2 f f^T + rho-coupled SPIRE block (SPD)."""
    flux = np.asarray(flux, dtype=np.float64)
    unc = np.asarray(unc, dtype=np.float64)
    cov = np.diag(unc**2) + calfrac**2 * np.outer(flux, flux)
    for i in spire_idx:
        for j in spire_idx:
            if i != j:
                cov[i, j] += rho * unc[i] * unc[j]
    return cov


def walker_cloud(truth, n, rng, lowlim, uplim=None, sigma=P0_SIGMA):
    """n parameter vectors ~ N(truth, sigma), redrawn until inside limits --
    the region an emcee ensemble actually explores
    (reference mbb_fit.py:449-477 draws its initial ball the same way)."""
    truth = np.asarray(truth, dtype=np.float64)
    out = np.empty((n, 5))
    for i in range(5):
        v = sigma[i] * rng.standard_normal(n) + truth[i]
        hi = np.inf if uplim is None else uplim[i]
        bad = np.nonzero((v < lowlim[i]) | (v > hi))[0]
        it = 0
        while bad.size:
            v[bad] = sigma[i] * rng.standard_normal(bad.size) + truth[i]
            bad = np.nonzero((v < lowlim[i]) | (v > hi))[0]
            it += 1
            if it > 1000:
                raise RuntimeError("could not draw inside limits for param %d" % i)
        out[:, i] = v
    return out


def random_walk_chain(truth, nwalkers, nsteps, rng, accept=0.35,
                      lowlim=(1, 0.1, 1, 0.1, 1e-3), sigma=P0_SIGMA, step=0.05):
    """A chain-shaped array [nwalkers, nsteps, 5] with exact repeats where a
    proposal is 'rejected' (emcee stores the unchanged position), for
    exercising the chain post-processing path (configs[3])."""
    truth = np.asarray(truth, dtype=np.float64)
    lowlim = np.asarray(lowlim, dtype=np.float64)
    chain = np.empty((nwalkers, nsteps, 5))
    pos = walker_cloud(truth, nwalkers, rng, lowlim, sigma=0.3 * np.asarray(sigma))
    for t in range(nsteps):
        move = rng.random_sample(nwalkers) < accept
        prop = pos + step * np.asarray(sigma) * rng.standard_normal((nwalkers, 5))
        ok = move & np.all(prop >= lowlim * 1.5, axis=1)
        pos = np.where(ok[:, None], prop, pos)
        chain[:, t, :] = pos
    return chain
