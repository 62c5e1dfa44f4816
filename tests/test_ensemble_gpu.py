"""GPU: the device-resident ensemble sampler (mbb_ensemble_run / batch_fitter).

Replay: the same Philox draws regenerated in numpy (tests/philox_np.py) drive
the emcee-2.2 stretch move on the host with the CPU oracle as log-probability;
the device chain must end at bit-identical positions with the same acceptance
counts, log-probabilities within 1e-12."""
import numpy as np
import pytest

import philox_np
from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _problem(oracle, response, nsrc, nw, seed):
    from mbb_emcee_b200 import batch_fitter
    rng = np.random.RandomState(seed)
    if response:
        bands = ["PACS_160um", "SPIRE_250um", "SPIRE_500um", "PdBI_box_135_3.6"]
        bf = batch_fitter(nwalkers=nw, response=True, device=0)
        truth = (14.0, 1.8, 400.0, 3.0, 30.0)
    else:
        bands = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
        bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, device=0)
        bf.fix_param('alpha')
        truth = (12.0, 1.8, 1300.0, 4.0, 30.0)
    nb = len(bands)
    flux = rng.uniform(10, 80, (nsrc, nb))
    unc = np.maximum(0.1 * flux, 1.0)
    bf.set_data(bands, flux, unc)
    p0 = bf.generate_initial_values(truth, [2, 0.2, 100, 0.3, 5.0], seed=seed)
    specs = []
    for s in range(nsrc):
        sp = oracle.LikeSpec(500.0, bf.like.noalpha, bf.like.opthin)
        if response:
            sp.set_phot([oracle.band_from_response(r) for r in bf.like._responses], flux[s], unc[s])
        else:
            sp.set_phot(bands, flux[s], unc[s])
        sp.has_uplim = list(bf.like.has_uplims)
        sp.uplim = np.array(bf.like.uplims)
        specs.append(sp)
    return bf, p0, specs


# (False, ...): delta bands -> fused one-kernel half-step, for walker counts where a CTA covers
# several sources, part of one and a ragged tail; (True, ...): tabulated bands -> propose /
# evaluate / accept kernels
@pytest.mark.parametrize("response,nsrc,nw,nsteps", [(False, 5, 16, 25), (False, 11, 64, 12), (False, 5, 250, 6),
                                                     (True, 2, 12, 8)])
def test_device_sampler_replays_on_host(oracle, response, nsrc, nw, nsteps):
    bf, p0, specs = _problem(oracle, response, nsrc, nw, 11)
    bf._stage()
    ctx = bf.like.context
    pos, lnp, nacc, st = ctx.ensemble_run(p0, nsteps, seed=0x1234ABCD5678, a=2.0)
    assert (st <= 1).all()
    rpos, rlnp, rnacc = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, nsteps,
                                         0x1234ABCD5678)
    assert np.array_equal(pos, rpos)
    assert np.array_equal(nacc, rnacc)
    assert relerr(lnp, rlnp).max() < TOL
    assert 0 < nacc.sum() < nsrc * nw * nsteps
    # continuing a run (burn-in then main chain) == one long run
    p1, l1, n1, _ = ctx.ensemble_run(p0, nsteps - 3, seed=0x1234ABCD5678)
    p2, l2, n2, _ = ctx.ensemble_run(p1, 3, seed=0x1234ABCD5678, step0=nsteps - 3, lnprob=l1)
    assert np.array_equal(p2, pos) and np.array_equal(l2, lnp) and np.array_equal(n1 + n2, nacc)
    if not response:
        assert (pos[:, :, 3] == 4.0).all()            # a fixed parameter never moves


def test_batch_fitter_recovers_truth():
    """Statistical sanity on synthetic sources with known parameters."""
    from mbb_emcee_b200 import batch_fitter, modified_blackbody
    rng = np.random.RandomState(3)
    nsrc, nw = 64, 64
    waves = np.array([70.0, 100.0, 160.0, 250.0, 350.0, 500.0])
    T = rng.uniform(10, 20, nsrc)
    flux = np.empty((nsrc, 6))
    for s in range(nsrc):
        flux[s] = modified_blackbody(T[s], 1.8, None, None, 40.0, opthin=True, noalpha=True)(waves)
    unc = np.maximum(0.05 * flux, 0.5)
    flux = flux + unc * rng.standard_normal(flux.shape)
    bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, device=0)
    bf.fix_param('alpha')
    bf.set_data(waves, flux, unc)
    init = np.column_stack([T + 1.0, np.full(nsrc, 1.7), np.full(nsrc, 1300.0), np.full(nsrc, 4.0),
                            np.full(nsrc, 35.0)])
    p0 = bf.generate_initial_values(init, [2, 0.2, 100, 0.3, 5.0], seed=1)
    out = bf.run(150, 150, p0, seed=99)
    acc = out["acceptance_fraction"].mean()
    assert 0.15 < acc < 0.75
    Tfit = out["pos"][:, :, 0].mean(axis=1)
    assert np.median(np.abs(Tfit - T) / T) < 0.05
    assert np.isfinite(out["lnprob"]).all()


@pytest.mark.parametrize("opthin,noalpha", [(True, True), (False, False)])
@pytest.mark.parametrize("nsrc,nw", [(1500, 512), (1001, 250), (700, 64), (37, 1024)])
def test_sampler_kernels_agree_bitwise(monkeypatch, opthin, noalpha, nsrc, nw):
    """The two device forms of a half-step -- the fused kernel and propose / evaluate /
    accept -- draw the same numbers and do the same arithmetic: identical positions,
    log-probabilities and acceptance counts over many sources."""
    from mbb_emcee_b200 import batch_fitter
    rng = np.random.RandomState(nsrc)
    bands = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    bf = batch_fitter(nwalkers=nw, opthin=opthin, noalpha=noalpha, device=0)
    if noalpha:
        bf.fix_param('alpha')
    flux = rng.uniform(10, 80, (nsrc, 6))
    bf.set_data(bands, flux, np.maximum(0.1 * flux, 1.0))
    truth = (12.0, 1.8, 1300.0, 4.0, 30.0) if opthin else (14.0, 1.8, 400.0, 3.0, 30.0)
    p0 = bf.generate_initial_values(truth, [2, 0.2, 100, 0.3, 5.0], seed=5)
    bf._stage()
    ctx = bf.like.context
    runs = []
    for env in ({}, {"MBB_B200_NO_FUSED_SAMPLER": "1"}):
        monkeypatch.delenv("MBB_B200_NO_FUSED_SAMPLER", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        n0 = ctx.launch_count()
        runs.append(ctx.ensemble_run(p0, 5, seed=77) + (ctx.launch_count() - n0,))
    a, b = runs
    assert a[4] == 11 and b[4] > 11                     # 1 + 2 launches per iteration when fused
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert 0 < a[2].sum() < nsrc * nw * 5
