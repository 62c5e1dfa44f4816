"""Synthetic photometry configurations (BASELINE.json ``configs``).

Pure numpy, no model evaluation of its own: callers hand in a function that
returns the noise-free model fluxes (the CUDA likelihood machinery in the
product and in bench.py; the reference itself in tests/golden/make_golden.py),
so the same seeded noise realisation is used everywhere.

All randomness is ``numpy.random.RandomState(seed)`` (legacy MT19937, stable
across numpy versions).
"""
import numpy as np

__all__ = ["CONFIGS", "SAMPLE_PHOT", "sample_problem", "noisy_photometry", "walker_cloud",
           "cfg3_covariance", "random_walk_chain"]

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


# One representative source per configuration (truth + one seeded noise draw, rounded): the
# photometry of the bounded single-source sample that bench.py's CPU arms time and that its GPU arm
# re-evaluates as a cross-check.  Literal numbers, so that both arms hold the same data without
# either evaluating a model.
SAMPLE_PHOT = {
    "cfg1": ([2.844, 4.846, 43.249, 73.149, 59.524, 29.042], [1.0, 1.0, 3.965, 6.964, 5.589, 3.0]),
    "cfg2": ([6.959, 23.43, 54.617, 44.904, 35.288, 8.339], [1.0, 2.144, 4.939, 4.946, 3.123, 1.0]),
    "cfg3": ([43.221, 48.168, 32.427, 7.622, 6.647, 4.551, 2.646, 0.795],
             [4.939, 4.946, 3.123, 1.0, 1.0, 1.0, 1.0, 1.0]),
    "cfg5": ([-0.108, 3.605, 36.768, 70.053, 61.505, 29.616], [1.0, 1.0, 3.965, 6.964, 5.589, 3.0]),
    "cfg5p": ([7.184, 34.353, 72.031, 60.65, 27.961, 6.771], [1.0, 3.813, 6.892, 5.687, 3.168, 1.0]),
    # the reference's default model (optically thick + alpha) on cfg5's six delta bands
    "default_model": ([2.1, 9.7, 38.4, 55.2, 44.9, 29.3], [1.0, 1.0, 3.84, 5.52, 4.49, 2.93]),
}


def sample_problem(name, n, seed=None):
    """(cfg, flux, unc, cov-or-None, P[n][5]) of the bounded single-source sample of a workload:
    literal photometry (SAMPLE_PHOT) and n walker positions ~ N(truth, P0_SIGMA) inside the default
    lower limits -- a function of (name, n, seed) only, identical in every process that calls it."""
    cfg = dict(CONFIGS["cfg5" if name == "default_model" else name])
    if name == "default_model":
        cfg.update(opthin=False, noalpha=False, truth=(14.0, 1.8, 400.0, 3.0, 30.0))
    truth = np.asarray(cfg.get("truth", (12.0, 1.8, 1300.0, 4.0, 30.0)), dtype=np.float64)
    flux, unc = (np.array(x, dtype=np.float64) for x in SAMPLE_PHOT[name])
    cov = cfg3_covariance(flux, unc) if name == "cfg3" else None
    rng = np.random.RandomState(cfg["seed"] + 77 if seed is None else seed)
    P = walker_cloud(truth, n, rng, np.array([1, 0.1, 1, 0.1, 1e-3]) * 1.01)
    return cfg, flux, unc, cov, P


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
