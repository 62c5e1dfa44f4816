"""GPU: the device-resident ensemble sampler (mbb_ensemble_fit / mbb_ensemble_run / batch_fitter).

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
    ctx = bf._stage()
    pos, lnp, nacc, st = ctx.ensemble_run(p0, nsteps, seed=0x1234ABCD5678, a=2.0)
    assert (st <= 1).all()
    rpos, rlnp, rnacc = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, nsteps,
                                         0x1234ABCD5678)
    assert np.array_equal(pos, rpos)
    assert np.array_equal(nacc, rnacc)
    assert relerr(lnp, rlnp).max() < TOL
    assert 0 < nacc.sum() < nsrc * nw * nsteps
    # emcee 2.2's ORIGINAL acceptance test, lnpdiff > log(u) (the device uses the log-free
    # form u < z^4 exp(dlnp)): the same chain
    lpos, llnp, lnacc = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, nsteps,
                                         0x1234ABCD5678, log_form=True)
    assert np.array_equal(pos, lpos) and np.array_equal(nacc, lnacc)
    # continuing a run (burn-in then main chain) == one long run
    p1, l1, n1, _ = ctx.ensemble_run(p0, nsteps - 3, seed=0x1234ABCD5678)
    p2, l2, n2, _ = ctx.ensemble_run(p1, 3, seed=0x1234ABCD5678, step0=nsteps - 3, lnprob=l1)
    assert np.array_equal(p2, pos) and np.array_equal(l2, lnp) and np.array_equal(n1 + n2, nacc)
    if not response:
        assert (pos[:, :, 3] == 4.0).all()            # a fixed parameter never moves
    # a shard of a larger source list: global source offset in the RNG counter, recorded chain
    out = ctx.ensemble_fit(p0, 0, nsteps, seed=99, src0=1000, stats=False, chain=True)
    r = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, nsteps, 99, src0=1000, chain=True)
    assert np.array_equal(out["pos"], r[0]) and np.array_equal(out["naccept"], r[2])
    assert np.array_equal(out["chain"], r[3])
    assert relerr(out["chain_lnprob"], r[4]).max() < TOL
    assert not np.array_equal(out["pos"], pos)


def _numpy_stats(ch, chl):
    """Reference summaries of a recorded chain ch[nrec][nsrc][nw][5], chl[nrec][nsrc][nw]."""
    nrec, nsrc, nw, _ = ch.shape
    flat = np.moveaxis(ch, 1, 0).reshape(nsrc, nrec * nw, 5)
    fl = np.moveaxis(chl, 1, 0).reshape(nsrc, nrec * nw)
    mean = flat.mean(axis=1)
    m2 = ((flat - mean[:, None, :])**2).sum(axis=1)
    ibest = fl.argmax(axis=1)
    return dict(n=np.full(nsrc, nrec * nw), mean=mean, m2=m2, min=flat.min(axis=1), max=flat.max(axis=1),
                best_lnp=fl.max(axis=1), best=flat[np.arange(nsrc), ibest])


# resident kernel: several sources per CTA (nw 16, 64), one per CTA (512), threads looping over
# walkers (1024); tabulated bands: propose / evaluate / accept + the per-record summary kernel
@pytest.mark.parametrize("response,nsrc,nw,nburn,nsteps,thin", [
    (False, 37, 16, 5, 24, 1), (False, 21, 64, 0, 18, 3), (False, 9, 512, 4, 12, 2), (False, 3, 1024, 2, 9, 1),
    (True, 3, 12, 3, 10, 2)])
def test_fit_summaries_equal_numpy_over_the_chain(oracle, response, nsrc, nw, nburn, nsteps, thin):
    from mbb_emcee_b200 import _native as nat
    bf, p0, _ = _problem(oracle, response, nsrc, nw, 21)
    ctx = bf._stage()
    out = ctx.ensemble_fit(p0, nburn, nsteps, seed=31, chain=True, thin=thin)
    assert (out["status"] <= 1).all()
    ch, chl, S = out["chain"], out["chain_lnprob"], out["stats"]
    assert ch.shape == (nsteps // thin, nsrc, nw, 5)
    # the chain is what burn-in + main run through the plain sampler produce (continuation by step0)
    pb, lb, _, _ = ctx.ensemble_run(p0, nburn, seed=31) if nburn else (p0, None, None, None)
    full = ctx.ensemble_fit(pb, 0, nsteps, seed=31, step0=nburn, lnprob=lb, stats=False, chain=True)
    assert np.array_equal(full["chain"][thin - 1::thin][:nsteps // thin], ch)
    assert np.array_equal(full["pos"], out["pos"]) and np.array_equal(full["lnprob"], out["lnprob"])
    assert np.array_equal(full["naccept"], out["naccept"])          # burn-in moves are not counted
    assert np.array_equal(ch[-1], out["pos"]) if nsteps % thin == 0 else True
    want = _numpy_stats(ch, chl)
    assert np.array_equal(S[:, nat.FS_N], want["n"])
    assert np.array_equal(S[:, nat.FS_MIN:nat.FS_MIN + 5], want["min"])
    assert np.array_equal(S[:, nat.FS_MAX:nat.FS_MAX + 5], want["max"])
    assert np.array_equal(S[:, nat.FS_BESTLNP], want["best_lnp"])
    assert np.array_equal(S[:, nat.FS_BEST:nat.FS_BEST + 5], want["best"])
    assert relerr(S[:, nat.FS_MEAN:nat.FS_MEAN + 5], want["mean"]).max() < 1e-13
    m2 = S[:, nat.FS_M2:nat.FS_M2 + 5]
    free = want["m2"] > 0
    assert relerr(m2[free], want["m2"][free]).max() < 1e-10 and (m2[~free] == 0).all()
    assert np.allclose(S[:, nat.FS_ACC], out["naccept"].sum(axis=1) / float(nw * nsteps), rtol=1e-14)
    assert 0 < out["naccept"].sum() < nsrc * nw * nsteps


def test_fit_chain_streams_to_host_in_segments(oracle, monkeypatch):
    """Host chain output is cut into segments (two device buffers, D2H behind the sampler);
    the summaries are merged across segments.  Same numbers as one device-resident pass."""
    import torch
    from mbb_emcee_b200 import _native as nat
    nsrc, nw, nburn, nsteps, thin = 300, 512, 3, 14, 2
    bf, p0, _ = _problem(oracle, False, nsrc, nw, 5)
    ctx = bf._stage()
    nrec = nsteps // thin
    dev = torch.device("cuda:0")
    P = torch.as_tensor(p0, device=dev).contiguous()
    L = torch.empty((nsrc, nw), dtype=torch.float64, device=dev)
    A = torch.zeros((nsrc, nw), dtype=torch.int32, device=dev)
    S = torch.zeros((nsrc, nat.FIT_NSTATS), dtype=torch.float64, device=dev)
    C = torch.empty((nrec, nsrc, nw, 5), dtype=torch.float64, device=dev)
    CL = torch.empty((nrec, nsrc, nw), dtype=torch.float64, device=dev)
    ctx.ensemble_fit_device(nsrc, nw, nburn, nsteps, P.data_ptr(), L.data_ptr(), seed=8, naccept_ptr=A.data_ptr(),
                            stats_ptr=S.data_ptr(), chain_ptr=C.data_ptr(), chain_lnprob_ptr=CL.data_ptr(),
                            thin=thin)
    ctx.sync()
    # one record of this problem is 7.4 MB: the 256 MB segment budget holds all 7; force small segments
    monkeypatch.setenv("MBB_B200_CHAIN_SEGMENT_MB", "16")
    out = ctx.ensemble_fit(p0, nburn, nsteps, seed=8, chain=True, thin=thin)
    assert np.array_equal(out["chain"], C.cpu().numpy()) and np.array_equal(out["chain_lnprob"], CL.cpu().numpy())
    assert np.array_equal(out["pos"], P.cpu().numpy()) and np.array_equal(out["naccept"], A.cpu().numpy())
    s1, s2 = out["stats"], S.cpu().numpy()
    exact = [nat.FS_N] + list(range(nat.FS_MIN, nat.FS_BEST + 5)) + [nat.FS_ACC]
    assert np.array_equal(s1[:, exact], s2[:, exact])
    assert relerr(s1[:, nat.FS_MEAN:nat.FS_MEAN + 5], s2[:, nat.FS_MEAN:nat.FS_MEAN + 5]).max() < 1e-13
    free = s2[:, nat.FS_M2:nat.FS_M2 + 5] > 0
    assert relerr(s1[:, nat.FS_M2:nat.FS_M2 + 5][free], s2[:, nat.FS_M2:nat.FS_M2 + 5][free]).max() < 1e-10


def test_fit_host_pipeline_equals_device(oracle):
    """Without a chain the host call cuts the sources into chunks that flow through three slots
    (upload / sampler / download of neighbouring chunks overlap); the numbers are those of one
    device-resident call, page-locked or pageable buffers alike."""
    import torch
    from mbb_emcee_b200 import _native as nat
    nsrc, nw, nburn, nsteps = 5000, 16, 3, 8
    bf, p0, _ = _problem(oracle, False, nsrc, nw, 6)
    ctx = bf._stage()
    dev = torch.device("cuda:0")
    P = torch.as_tensor(np.ascontiguousarray(p0), device=dev).contiguous()
    L = torch.empty((nsrc, nw), dtype=torch.float64, device=dev)
    A = torch.zeros((nsrc, nw), dtype=torch.int32, device=dev)
    S = torch.zeros((nsrc, nat.FIT_NSTATS), dtype=torch.float64, device=dev)
    ctx.ensemble_fit_device(nsrc, nw, nburn, nsteps, P.data_ptr(), L.data_ptr(), seed=12, src0=7,
                            naccept_ptr=A.data_ptr(), stats_ptr=S.data_ptr(), thin=2)
    ctx.sync()
    n0 = ctx.launch_count()
    out = ctx.ensemble_fit(p0, nburn, nsteps, seed=12, src0=7, thin=2)                  # pageable arrays
    assert ctx.launch_count() - n0 == 4                                              # 2 chunks x (lnprob + sampler)
    pin = {k: nat.pinned_empty(v.shape, v.dtype) for k, v in out.items()}
    pin["pos"][...] = p0
    ctx.ensemble_fit_into(pin["pos"], pin["lnprob"], nburn, nsteps, naccept=pin["naccept"], status=pin["status"],
                          stats=pin["stats"], seed=12, src0=7, thin=2)
    for res in (out, pin):
        assert np.array_equal(res["pos"], P.cpu().numpy()) and np.array_equal(res["lnprob"], L.cpu().numpy())
        assert np.array_equal(res["naccept"], A.cpu().numpy()) and np.array_equal(res["stats"], S.cpu().numpy())
        assert (res["status"] <= 1).all()


def test_batch_fitter_shards_are_the_same_fit():
    """devices=[0, 0, 0]: three shards (three contexts, three host threads) fill disjoint slices of
    the shared page-locked outputs; bit-identical to the unsharded fit (global source index in the
    RNG counter), chain included."""
    from mbb_emcee_b200 import batch_fitter
    rng = np.random.RandomState(4)
    nsrc, nw = 101, 64
    bands = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(10, 80, (nsrc, 6))
    unc = np.maximum(0.1 * flux, 1.0)
    res = []
    for devices in ([0], [0, 0, 0]):
        bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, devices=devices)
        bf.fix_param('alpha')
        bf.set_data(bands, flux, unc)
        p0 = bf.generate_initial_values((12.0, 1.8, 1300.0, 4.0, 30.0), [2, 0.2, 100, 0.3, 5.0], seed=2)
        res.append(bf.run(6, 10, p0, seed=5, thin=2, chain=True))
    a, b = res
    for key in ("pos", "lnprob", "naccept", "status", "stats", "chain", "chain_lnprob"):
        assert np.array_equal(a[key], b[key]), key
    assert a["chain"].shape == (5, nsrc, nw, 5)
    cen = a.par_cen("T")
    assert cen.shape == (nsrc, 3) and (cen[:, 1:] > 0).all()


def test_batch_fitter_posteriors_match_host_sampler(oracle):
    """The fit a user gets: per-source posterior mean / sigma from the device summaries against
    (a) the known truths and (b) an independent host run of the emcee-2.2 stretch move
    (oracle.stretch_chain: Mersenne-Twister draws, original log-form acceptance, CPU oracle as
    log-probability) on the same sources, within Monte-Carlo error."""
    from mbb_emcee_b200 import batch_fitter, modified_blackbody
    rng = np.random.RandomState(3)
    nsrc, nw, nburn, nsteps = 64, 64, 200, 400
    waves = np.array([70.0, 100.0, 160.0, 250.0, 350.0, 500.0])
    T = rng.uniform(10, 20, nsrc)
    flux = np.empty((nsrc, 6))
    for s in range(nsrc):
        flux[s] = modified_blackbody(T[s], 1.8, None, None, 40.0, opthin=True, noalpha=True)(waves)
    unc = np.maximum(0.05 * flux, 0.5)
    flux = flux + unc * rng.standard_normal(flux.shape)
    bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, device=0)
    bf.fix_param('alpha')
    bf.fix_param('lambda0')
    bf.set_data(waves, flux, unc)
    init = np.column_stack([T + 1.0, np.full(nsrc, 1.7), np.full(nsrc, 1300.0), np.full(nsrc, 4.0),
                            np.full(nsrc, 35.0)])
    p0 = bf.generate_initial_values(init, [2, 0.2, 100, 0.3, 5.0], seed=1)
    out = bf.run(nburn, nsteps, p0, seed=99)
    acc = out.mean_acceptance
    assert (0.15 < acc).all() and (acc < 0.8).all()
    assert np.allclose(acc, out["acceptance_fraction"].mean(axis=1), rtol=1e-12)
    assert np.isfinite(out["lnprob"]).all() and (out.nsamples == nw * nsteps).all()
    # (a) truths: pulls of T are standard-normal-ish
    pull = (out.mean[:, 0] - T) / out.std[:, 0]
    assert np.abs(pull).max() < 5.0 and np.abs(pull.mean()) < 0.5 and 0.5 < pull.std() < 1.6
    assert (out.best_lnprob >= out["lnprob"].max(axis=1)).all()
    assert (out.std[:, 3] == 0).all() and (out.mean[:, 3] == 4.0).all()           # fixed parameters
    # (b) host sampler on 8 of the sources
    free = [0, 1, 4]
    for s in range(0, nsrc, 8):
        sp = oracle.LikeSpec(500.0, True, True)
        sp.set_phot(list(waves), flux[s], unc[s])
        sp.has_uplim = list(bf.like.has_uplims)
        sp.uplim = np.array(bf.like.uplims)
        ch, _, nacc = oracle.stretch_chain(lambda Q: oracle.loglike_batch(sp, Q), p0[s], nburn + nsteps,
                                           np.random.RandomState(100 + s))
        main = ch[:, nburn:, :].reshape(-1, 5)
        hm, hs = main.mean(axis=0), main.std(axis=0, ddof=1)
        # autocorrelation ~30 steps -> ~850 independent samples per run: error of the mean ~ sigma/29
        assert (np.abs(out.mean[s, free] - hm[free]) < 0.25 * hs[free]).all(), (s, out.mean[s], hm, hs)
        assert (np.abs(out.std[s, free] / hs[free] - 1.0) < 0.25).all(), (s, out.std[s], hs)
        assert abs(acc[s] - nacc.mean() / (nburn + nsteps)) < 0.08


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
    ctx = bf._stage()
    runs = []
    for env in ({}, {"MBB_B200_NO_FUSED_SAMPLER": "1"}):
        monkeypatch.delenv("MBB_B200_NO_FUSED_SAMPLER", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        n0 = ctx.launch_count()
        runs.append(ctx.ensemble_run(p0, 5, seed=77) + (ctx.launch_count() - n0,))
    a, b = runs
    assert a[4] == 2 and b[4] > 11                      # initial log-probability + ONE resident launch
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert 0 < a[2].sum() < nsrc * nw * 5


def test_with_installed_emcee(oracle):
    """When emcee itself is importable (it is not in the build image): its EnsembleSampler,
    handed the product's likelihood and the same RandomState, walks the chain of the product's
    own emcee-2.2 restatement (mbb_emcee_b200/ensemble.py) -- bit for bit under emcee 2.x, and
    statistically (posterior mean within 5 MC sigma) under emcee 3.x, whose move scheduling
    draws differently."""
    emcee = pytest.importorskip("emcee")
    from mbb_emcee_b200 import likelihood, synthetic
    from mbb_emcee_b200.ensemble import EnsembleSampler
    cfg = synthetic.CONFIGS["cfg1"]
    like = likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"], device=0)
    _, flux, unc, _, P = synthetic.sample_problem("cfg1", 64)
    like.set_phot(cfg["bands"], flux, unc)
    ours = EnsembleSampler(64, 5, like)
    ours.random_state = np.random.RandomState(7).get_state()
    ours.run_mcmc(P, 60)
    major = int(emcee.__version__.split(".")[0])
    if major < 3:
        theirs = emcee.EnsembleSampler(64, 5, like)
        theirs.random_state = np.random.RandomState(7).get_state()
        theirs.run_mcmc(P, 60)
        assert np.array_equal(theirs.chain, ours.chain)
    else:
        theirs = emcee.EnsembleSampler(64, 5, like, vectorize=True)
        theirs.run_mcmc(P, 400, progress=False)
        a, b = theirs.get_chain()[200:, :, 0].ravel(), ours.chain[:, 30:, 0].ravel()
        assert abs(a.mean() - b.mean()) < 5 * np.hypot(a.std() / np.sqrt(64), b.std() / np.sqrt(64))


@pytest.mark.parametrize("response", [False, True])
def test_mbb_fitter_device_sampler(oracle, response):
    """mbb_fitter(sampler="device"): the reference's run sequence (mbb_fit.py:524-543 -- burn in,
    reset, main chain from where the burn-in ended) with the whole stretch move inside the
    library.  The recorded chain is what a host replay with the same Philox draws and the CPU
    oracle gives (positions bit for bit), and its posterior agrees with the host sampler's
    (emcee's random stream) within Monte-Carlo error."""
    from mbb_emcee_b200 import mbb_fitter, mbb_results
    rng = np.random.RandomState(5)
    if response:
        bands = ["PACS_160um", "SPIRE_250um", "SPIRE_350um", "SPIRE_500um"]
        kw = dict(response=True)
        nw, nburn, nsteps = 40, 4, 12
    else:
        bands = [100.0, 160.0, 250.0, 350.0, 500.0, 850.0]
        kw = dict()
        nw, nburn, nsteps = 250, 5, 20
    truth = np.array([14.0, 1.8, 400.0, 3.0, 30.0])
    fits = {}
    for where in ("device", "host"):
        fit = mbb_fitter(nwalkers=nw, device=0, sampler=where, seed=77, **kw)
        if not fits:
            f0 = fit.like.get_sed(truth, np.array([100.0, 160.0, 250.0, 350.0, 500.0, 850.0]))
            flux = (f0 if not response else f0[1:5]) * (1.0 + 0.03 * rng.standard_normal(len(bands)))
            unc = 0.08 * flux
        fit.set_data(bands, flux, unc)
        fits[where] = fit
    fit = fits["device"]
    np.random.seed(3)
    p0 = fit.generate_initial_values(truth, [2, 0.2, 100, 0.3, 5.0])
    fit.run(nburn, nsteps, p0)
    ch, lnp = fit.sampler.chain, fit.sampler.lnprobability
    assert ch.shape == (nw, nsteps, 5) and lnp.shape == (nw, nsteps)
    assert fit.sampler.iterations == nsteps and fit.sampler.random_state[2] == nburn + nsteps
    # host replay: burn-in and main run are one Philox stream (iterations 0 .. nburn + nsteps)
    sp = oracle.LikeSpec(500.0, False, False)
    if response:
        sp.set_phot([oracle.band_from_response(r) for r in fit.like._responses], flux, unc)
    else:
        sp.set_phot(bands, flux, unc)
    sp.has_uplim = list(fit.like.has_uplims)
    sp.uplim = np.array(fit.like.uplims)
    r = philox_np.replay(lambda s, Q: oracle.loglike_batch(sp, Q), p0[None], nburn + nsteps, 77, chain=True)
    assert np.array_equal(ch, np.moveaxis(r[3][nburn:, 0], 0, 1))
    assert relerr(lnp, r[4][nburn:, 0].T).max() < TOL
    acc = fit.sampler.acceptance_fraction
    assert acc.shape == (nw,) and 0.05 < acc.mean() < 0.95
    res = mbb_results(fit=fit)
    assert np.isfinite(res.par_cen("T")).all()
    if response:
        return
    # posterior against the host sampler on a longer run: means within 5 standard errors
    # (each chain has >= 2e4 samples; integrated autocorrelation ~ 40 iterations)
    stats = {}
    for where in ("device", "host"):
        f = fits[where]
        np.random.seed(11)
        q0 = f.generate_initial_values(truth, [2, 0.2, 100, 0.3, 5.0])
        f.run(150, 400, q0)
        flat = f.sampler.chain[:, ::40].reshape(-1, 5)
        stats[where] = (flat.mean(axis=0), flat.std(axis=0), flat.shape[0])
    (m1, s1, n1), (m2, s2, n2) = stats["device"], stats["host"]
    for i in (0, 1, 4):
        assert abs(m1[i] - m2[i]) < 5.0 * np.hypot(s1[i], s2[i]) / np.sqrt(n1 / 2.0)
        assert 0.6 < s1[i] / s2[i] < 1.6


@pytest.mark.parametrize("opthin,noalpha,cov", [(False, False, False), (True, True, False), (False, False, True)])
def test_fused_half_step_over_passbands_is_bit_identical(monkeypatch, opthin, noalpha, cov):
    """Small ensembles over tabulated passbands take a half-step in two launches (proposal + setup,
    node sums + accept); MBB_B200_NO_FUSED_HALFSTEP restores propose / setup / nodes / accept.  Same
    draws, same arithmetic: the chains are bit-identical, with diagonal errors and a covariance."""
    from mbb_emcee_b200 import batch_fitter
    rng = np.random.RandomState(3)
    bands = ["PACS_160um", "SPIRE_250um", "SPIRE_350um", "SPIRE_500um", "SCUBA2_850um"]
    nsrc, nw = 3, 50
    bf = batch_fitter(nwalkers=nw, opthin=opthin, noalpha=noalpha, response=True, device=0)
    flux = rng.uniform(10, 80, (nsrc, 5))
    unc = np.maximum(0.1 * flux, 1.0)
    if cov:
        C = np.stack([np.diag(unc[k]**2) + 0.04**2 * np.outer(flux[k], flux[k]) for k in range(nsrc)])
        bf.set_data(bands, flux, covmatrix=C)
    else:
        bf.set_data(bands, flux, unc)
    p0 = bf.generate_initial_values((14.0, 1.8, 400.0, 3.0, 30.0), [2, 0.2, 100, 0.3, 5.0], seed=2)
    p0[0, 3, 0] = 0.2                                  # a walker below the T limit: never accepted into
    ctx = bf._stage()
    a = ctx.ensemble_fit(p0, 3, 9, seed=5, chain=True, thin=2)
    l0 = ctx.launch_count()
    ctx.ensemble_fit(p0, 3, 9, seed=5, chain=True, thin=2)
    fused_launches = ctx.launch_count() - l0
    monkeypatch.setenv("MBB_B200_NO_FUSED_HALFSTEP", "1")
    l0 = ctx.launch_count()
    b = ctx.ensemble_fit(p0, 3, 9, seed=5, chain=True, thin=2)
    plain_launches = ctx.launch_count() - l0
    for k in ("pos", "lnprob", "naccept", "status", "chain", "chain_lnprob", "stats"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k
    assert 0 < a["naccept"].sum() < nsrc * nw * 9
    assert fused_launches < plain_launches
