"""GPU: parity of the CUDA path with the reference, through the C ABI.

Checks (all calls go product API -> ctypes -> libmbb_b200.so -> sm_100a kernels):
  * golden vectors recorded from the executed reference (tests/golden);
  * the CPU oracle on seeded random inputs;
  * at BASELINE's full size (1e5 sources x 512 walkers) size-independent
    properties plus an oracle-checked random sample;
  * chain bit-identity under the emcee 2.2 stretch move for a fixed RNG seed.

Tolerance (north star): 1e-12 relative on f_nu and log-likelihood.
"""
import numpy as np
import pytest

from conftest import VARIANTS, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    from mbb_emcee_b200 import _native
    return _native.Context(0)


# --------------------------------------------------------------------------- SED
@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
@pytest.mark.parametrize("wavenorm", [500.0, 250.0])
def test_fnu_and_constants_golden(ctx, golden, name, opthin, noalpha, wavenorm):
    g = golden.sed
    tag = "%s_wn%d" % (name, int(wavenorm))
    P = g[tag + "_P"]
    ctx.set_model(wavenorm, opthin, noalpha)
    freq = 299792458e-3 / g["waves"]
    arr, st = ctx.fnu(P, freq, scalar_path=False)
    assert (st == 0).all()
    assert relerr(arr, g[tag + "_fnu_array"]).max() < TOL
    sc, st = ctx.fnu(P, freq, scalar_path=True)
    assert relerr(sc, g[tag + "_fnu_scalar"]).max() < TOL
    # the FAST arithmetic of the likelihood kernels (per-walker setup incl. the merge-point
    # solve, lean exp family, node formulas) on the same grid
    from mbb_emcee_b200 import _native
    fast, st = ctx.fnu(P, freq, math_mode=_native.MATH_FAST)
    assert (st == 0).all()
    assert relerr(fast, g[tag + "_fnu_array"]).max() < TOL
    assert not np.array_equal(fast, arr)
    c, st = ctx.sed_consts(P, want_peak=True)
    assert (st == 0).all()
    assert relerr(c[:, 0], g[tag + "_normfac"]).max() < TOL
    if not noalpha:
        assert relerr(c[:, 1], g[tag + "_xmerge"]).max() < TOL
        assert relerr(c[:, 2], g[tag + "_kappa"]).max() < TOL
    assert relerr(c[:, 5], g[tag + "_maxwave"]).max() < TOL


def test_reference_known_answers():
    """The reference's own tests (mbb_emcee/tests/test_modified_blackbody.py),
    run against the product classes."""
    from numpy.testing import assert_allclose
    from mbb_emcee_b200 import modified_blackbody
    mbb = modified_blackbody(10.0, 2.0, 800.0, 2.0, 45.0)
    assert mbb.has_alpha and not mbb.optically_thin
    assert_allclose(mbb(500), 45.0, atol=1e-4)
    wave = np.array([250.0, 350.0, 500.0, 850.0])
    assert_allclose(mbb(wave), [21.96268738, 39.53249977, 45.0, 22.06274444], rtol=1e-4)
    mbb = modified_blackbody(15.0, 1.8, 200.0, 3.0, 50.0, opthin=True)
    assert mbb.has_alpha and mbb.optically_thin and mbb.lambda0 is None
    assert_allclose(mbb(wave), [178.34976, 111.03026, 50.0, 10.880588], rtol=1e-4)
    thin = modified_blackbody(15.0, 1.8, 5.0, 3.0, 50.0, opthin=True)
    thick = modified_blackbody(15.0, 1.8, 5.0, 3.0, 50.0, opthin=False)
    w2 = np.array([500.0, 850.0, 1100.0, 2500.0])
    assert_allclose(thin(w2), thick(w2), rtol=1e-3)
    assert modified_blackbody(20.0, 1.9, None, 3.5, 50.0, noalpha=True, opthin=True).wavemerge is None
    assert_allclose(modified_blackbody(20.0, 1.9, None, 3.5, 50.0, opthin=True).wavemerge, 85.66065, rtol=1e-3)
    assert_allclose(modified_blackbody(35.0, 2.2, None, 2.8, 50.0, opthin=True).wavemerge, 51.40211, rtol=1e-3)
    assert modified_blackbody(20.0, 1.9, 250.0, 3.5, 50.0, noalpha=True).wavemerge is None
    assert_allclose(modified_blackbody(20.0, 1.9, 250.0, 3.5, 50.0).wavemerge, 109.5506829, rtol=1e-3)
    assert_allclose(modified_blackbody(40.0, 1.5, 600.0, 3.0, 50.0).wavemerge, 60.10021595, rtol=1e-3)
    with pytest.raises(ValueError):
        modified_blackbody(10.0, 2.0, 800.0, -1.0, 45.0)
    with pytest.raises(ValueError):
        modified_blackbody(10.0, -2.0, 800.0, 2.0, 45.0)


# -------------------------------------------------------------------- likelihood
def _make_like(golden, cfgname, mode):
    from mbb_emcee_b200 import likelihood, synthetic
    cfg = synthetic.CONFIGS[cfgname]
    g = golden.like
    like = likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                      response=cfg["response"], device=0)
    like.math_mode = mode
    like.set_phot(cfg["bands"], g[cfgname + "_flux"], g[cfgname + "_unc"])
    if cfgname + "_cov" in g:
        like.set_cov(g[cfgname + "_cov"])
    for nm, v in cfg.get("uplims", []):
        like.set_uplim(nm, v)
    for nm, m, s in cfg.get("gpriors", []):
        like.set_gaussian_prior(nm, m, s)
    return cfg, like


def _oracle_spec(oracle, like):
    spec = oracle.LikeSpec(like.wavenorm, like.noalpha, like.opthin)
    if like.response_integrate:
        spec.set_phot([oracle.band_from_response(r) for r in like._responses],
                      like.data_flux, like._flux_unc)
    else:
        spec.set_phot(like.data_wave, like.data_flux, like._flux_unc)
    spec.lowlim = np.array(like.lowlims, dtype=np.float64)
    spec.has_uplim = list(like.has_uplims)
    spec.uplim = np.array(like.uplims, dtype=np.float64)
    spec.has_gprior = list(like.has_gpriors)
    spec.gprior_mean = np.array(like.gprior_means)
    spec.gprior_ivar = np.array(like.gprior_ivars)
    if like.has_data_covmatrix:
        spec.set_cov(like.data_covmatrix)
    return spec


@pytest.mark.parametrize("cfgname", ["cfg1", "cfg2", "cfg3"])
@pytest.mark.parametrize("mode", [0, 1])
def test_loglike_golden(golden, cfgname, mode):
    cfg, like = _make_like(golden, cfgname, mode)
    g = golden.like
    P, ref = g[cfgname + "_P"], g[cfgname + "_lnlike"]
    ll, st = like.evaluate(P)
    assert np.array_equal(st == 1, np.isneginf(ref))
    assert np.array_equal(np.isneginf(ll), np.isneginf(ref))
    assert (st <= 1).all()
    fin = np.isfinite(ref)
    assert relerr(ll[fin], ref[fin]).max() < TOL
    # emcee's calling convention: one row -> python float
    one = like(P[6])
    assert isinstance(one, float) and abs(one - ref[6]) <= TOL * abs(ref[6])
    assert like(P[0]) == float("-inf")


@pytest.mark.parametrize("mode", [0, 1])
def test_loglike_extra_golden(golden, mode):
    from mbb_emcee_b200 import likelihood
    g = golden.like
    like = likelihood(wavenorm=500.0, device=0)
    like.math_mode = mode
    like.set_phot(g["extra_wave"], g["extra_flux"], g["extra_unc"])
    like.set_cov(g["extra_cov"])
    like.set_gaussian_prior('beta', 1.8, 0.3)
    like.set_uplim('lambda_peak', 300.0)
    like.set_gaussian_prior('lambda_peak', 250.0, 40.0)
    ll = like(g["extra_P"])
    assert relerr(ll, g["extra_lnlike"]).max() < TOL
    assert abs(ll[0] - (-30.13025033346131)) < 1e-11


@pytest.mark.parametrize("cfgname,n", [("cfg1", 4000), ("cfg2", 1500), ("cfg3", 700)])
def test_loglike_vs_oracle_random(golden, oracle, cfgname, n):
    """Seeded random walkers around the truth (and well away from it)."""
    from mbb_emcee_b200 import synthetic
    cfg, like = _make_like(golden, cfgname, 1)
    rng = np.random.RandomState(9000 + n)
    up = np.where(like.has_uplims[:5], like.uplims[:5], np.inf)
    P = synthetic.walker_cloud(cfg["truth"], n, rng, like.lowlims, up,
                               sigma=3.0 * synthetic.P0_SIGMA)
    want = oracle.loglike_batch(_oracle_spec(oracle, like), P)
    for mode in (0, 1):
        like.math_mode = mode
        got = like(P)
        assert relerr(got, want).max() < TOL, mode
    # ragged tail: a batch that is not a multiple of the tile / block size
    got = like(P[:n - 37])
    assert relerr(got, want[:n - 37]).max() < TOL
    # SoA layout gives the same numbers as AoS
    from mbb_emcee_b200 import _native
    soa, st = like.context.loglike(np.ascontiguousarray(P.T), layout=_native.SOA)
    assert np.array_equal(soa, like(P))


def test_all_variants_tabulated_vs_oracle(oracle):
    """Every (thin|thick)x(alpha|noalpha) variant through the warp kernel."""
    from mbb_emcee_b200 import likelihood, synthetic
    rng = np.random.RandomState(77)
    bands = ["PACS_70um", "PACS_160um", "SPIRE_350um", "ALMA_alma_230", "Y_delta_500um"]
    for name, opthin, noalpha in VARIANTS:
        like = likelihood(wavenorm=350.0, opthin=opthin, noalpha=noalpha, response=True, device=0)
        like.set_phot(bands, [40.0, 90.0, 50.0, 3.0, 28.0], [4.0, 9.0, 5.0, 1.0, 3.0])
        P = synthetic.walker_cloud((14.0, 1.8, 300.0, 3.0, 30.0), 300, rng, like.lowlims)
        want = oracle.loglike_batch(_oracle_spec(oracle, like), P)
        for mode in (0, 1, 2):                 # FAITHFUL, FAST, FAST_GAUSS (wavenorm 350: other L' tables)
            like.math_mode = mode
            assert relerr(like(P), want).max() < TOL, (name, mode)


def test_nodes_kernels_small_and_large_batches(oracle):
    """Tabulated passbands, FAST: batches of <= 4096 evaluations give a CTA to every
    evaluation (loglike_nodes_small_kernel: the half-ensemble of a single-source fit), larger
    ones a warp (the persistent nodes kernel).  Both against the oracle, for every model
    variant, with diagonal errors and with a full covariance (explicit inverse and Cholesky
    factor); and the same rows through either kernel agree to rounding."""
    from mbb_emcee_b200 import likelihood, synthetic
    rng = np.random.RandomState(78)
    bands = ["PACS_100um", "PACS_160um", "SPIRE_250um", "SPIRE_350um", "SPIRE_500um", "SCUBA2_850um"]
    flux = np.array([35.0, 80.0, 70.0, 45.0, 22.0, 5.0])
    unc = np.array([4.0, 8.0, 7.0, 5.0, 3.0, 1.0])
    cov = synthetic.cfg3_covariance(flux, unc, spire_idx=(2, 3, 4))
    for name, opthin, noalpha in VARIANTS:
        for solver in (None, "inverse", "cholesky"):
            like = likelihood(wavenorm=500.0, opthin=opthin, noalpha=noalpha, response=True, device=0)
            like.set_phot(bands, flux, unc)
            if solver is not None:
                like.set_cov(cov)
                like.cov_solver = solver
            big = synthetic.walker_cloud((14.0, 1.8, 400.0, 3.0, 30.0), 6000, rng, like.lowlims)
            big[::501, 0] = 0.5                       # below the lower limit: -inf through both kernels
            got_big = like(big)                       # 6000 > 4096: warp per evaluation
            got_small = like(big[:125])               # CTA per evaluation
            pick = np.r_[0:125, rng.randint(125, 6000, 175)]
            want = oracle.loglike_batch(_oracle_spec(oracle, like), big[pick])
            fin = np.isfinite(want)
            assert np.array_equal(np.isneginf(got_big[pick]), np.isneginf(want)) and (~fin).sum() >= 1
            assert relerr(got_big[pick][fin], want[fin]).max() < TOL, (name, solver)
            assert relerr(got_small[fin[:125]], want[:125][fin[:125]]).max() < TOL, (name, solver)
            assert np.array_equal(np.isneginf(got_small), np.isneginf(got_big[:125]))
            f2 = np.isfinite(got_small)
            assert relerr(got_small[f2], got_big[:125][f2]).max() < 1e-13, (name, solver)


def test_whole_wheel_global_table_path(oracle):
    """All 18 shipped filters (4677 nodes): the node table no longer fits the
    shared-memory budget, so the kernel reads it from global memory."""
    from mbb_emcee_b200 import likelihood, response_set, synthetic
    names = list(response_set().keys())
    like = likelihood(response=True, device=0)
    rng = np.random.RandomState(5)
    like.set_phot(names, rng.uniform(5, 50, len(names)), rng.uniform(1, 5, len(names)))
    P = synthetic.walker_cloud((20.0, 1.6, 200.0, 2.5, 30.0), 96, rng, like.lowlims)
    want = oracle.loglike_batch(_oracle_spec(oracle, like), P)
    assert relerr(like(P), want).max() < TOL
    like.math_mode = 2                          # 18 compressed rules in shared memory, full tables in global
    assert relerr(like(P), want).max() < TOL


def test_multi_source_batch(oracle):
    """Several sources in one call: walkers_per_source and explicit src_index."""
    from mbb_emcee_b200 import _native, synthetic
    rng = np.random.RandomState(11)
    nsrc, nw = 7, 40
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(5, 80, (nsrc, 6))
    unc = rng.uniform(1, 6, (nsrc, 6))
    ctx = _native.Context(0)
    ctx.set_model(500.0, True, True)
    off = np.arange(7, dtype=np.int32)
    ctx.set_bands(off, waves, np.ones(6))
    ctx.set_data(flux, ivar=1.0 / unc**2)
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    P = synthetic.walker_cloud((12.0, 1.8, 1300.0, 4.0, 30.0), nsrc * nw, rng, low)
    got, st = ctx.loglike(P, walkers_per_source=nw)
    want = np.empty(nsrc * nw)
    for s in range(nsrc):
        spec = oracle.LikeSpec(500.0, True, True)
        spec.set_phot(waves, flux[s], unc[s])
        want[s * nw:(s + 1) * nw] = oracle.loglike_batch(spec, P[s * nw:(s + 1) * nw])
    assert relerr(got, want).max() < TOL
    idx = rng.randint(0, nsrc, nsrc * nw).astype(np.int32)
    got2, st = ctx.loglike(P, src_index=idx)
    for i in range(0, nsrc * nw, 13):
        spec = oracle.LikeSpec(500.0, True, True)
        spec.set_phot(waves, flux[idx[i]], unc[idx[i]])
        assert abs(got2[i] - oracle.loglike(spec, P[i])) <= TOL * abs(got2[i])


def test_multi_source_covariance_delta_kernel(oracle):
    """Several sources, each with its own full covariance, delta bands: the delta kernel's
    direct-load path with the quadratic form over Cinv rows (likelihood.py:821-825), for
    implicit and explicit source indices and for the resident sampler's shared-memory staging."""
    from mbb_emcee_b200 import _native, synthetic
    rng = np.random.RandomState(23)
    nsrc, nw = 9, 48
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(5, 80, (nsrc, 6))
    unc = rng.uniform(1, 6, (nsrc, 6))
    cov = np.empty((nsrc, 6, 6))
    for s in range(nsrc):
        cov[s] = synthetic.cfg3_covariance(flux[s], unc[s], spire_idx=(3, 4, 5))
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    for opthin, noalpha in ((True, True), (False, False)):
        ctx = _native.Context(0)
        ctx.set_model(500.0, opthin, noalpha)
        ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
        ctx.set_data(flux, cinv=np.linalg.inv(cov))
        truth = (12.0, 1.8, 1300.0, 4.0, 30.0) if opthin else (14.0, 1.8, 400.0, 3.0, 30.0)
        P = synthetic.walker_cloud(truth, nsrc * nw, rng, low)
        want = np.empty(nsrc * nw)
        specs = []
        for s in range(nsrc):
            spec = oracle.LikeSpec(500.0, noalpha, opthin)
            spec.set_phot(waves, flux[s], unc[s])
            spec.set_cov(cov[s])
            specs.append(spec)
            want[s * nw:(s + 1) * nw] = oracle.loglike_batch(spec, P[s * nw:(s + 1) * nw])
        got, st = ctx.loglike(P, walkers_per_source=nw)
        assert (st == 0).all() and relerr(got, want).max() < TOL
        idx = np.repeat(np.arange(nsrc), nw).astype(np.int32)
        got2, _ = ctx.loglike(P, src_index=idx)
        assert np.array_equal(got2, got)
        # the sampler keeps each source's ensemble resident and evaluates against the same Cinv
        import philox_np
        p0 = P.reshape(nsrc, nw, 5)
        out = ctx.ensemble_fit(p0, 0, 5, seed=3, stats=False)
        r = philox_np.replay(lambda s, Q: oracle.loglike_batch(specs[s], Q), p0, 5, 3)
        assert np.array_equal(out["pos"], r[0]) and relerr(out["lnprob"], r[1]).max() < TOL


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_cholesky_covariance_golden(golden, mode):
    """Cholesky-prefactored covariance (north star (3); mbb_set_data_chol replaces
    likelihood.py:356 + :823): cfg3's golden log-likelihoods (8x8 covariance, the nodes kernel's
    warp path) and the survey's extra vector (6 delta bands, thread kernel: lambda_peak terms)
    within 1e-12, and within cond(C) * 1e-15 of the explicit-inverse path."""
    from mbb_emcee_b200 import likelihood
    g = golden.like
    cfg, like = _make_like(golden, "cfg3", mode)
    P, ref = g["cfg3_P"], g["cfg3_lnlike"]
    ll_inv, _ = like.evaluate(P)
    like.cov_solver = "cholesky"
    ll, st = like.evaluate(P)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(ll), np.isneginf(ref)) and (st <= 1).all()
    assert relerr(ll[fin], ref[fin]).max() < TOL
    assert not np.array_equal(ll[fin], ll_inv[fin])          # a different arithmetic did run
    assert relerr(ll[fin], ll_inv[fin]).max() < 1e-15 * np.linalg.cond(like.data_covmatrix)
    like.cov_solver = "inverse"
    assert np.array_equal(like.evaluate(P)[0][fin], ll_inv[fin])
    if mode == 2:
        return
    like = likelihood(wavenorm=500.0, device=0)
    like.math_mode = mode
    like.cov_solver = "cholesky"
    like.set_phot(g["extra_wave"], g["extra_flux"], g["extra_unc"])
    like.set_cov(g["extra_cov"])
    like.set_gaussian_prior('beta', 1.8, 0.3)
    like.set_uplim('lambda_peak', 300.0)
    like.set_gaussian_prior('lambda_peak', 250.0, 40.0)
    assert relerr(like(g["extra_P"]), g["extra_lnlike"]).max() < TOL


def test_cholesky_multi_source_delta_and_sampler(oracle):
    """Per-source Cholesky factors through the delta kernel (forward substitution in
    registers) and the resident sampler, against the oracle's inv(C) arithmetic; bad factors
    are refused."""
    from mbb_emcee_b200 import _native, synthetic, batch_fitter
    rng = np.random.RandomState(29)
    nsrc, nw = 7, 64
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(5, 80, (nsrc, 6))
    unc = rng.uniform(1, 6, (nsrc, 6))
    cov = np.stack([synthetic.cfg3_covariance(flux[s], unc[s], spire_idx=(3, 4, 5)) for s in range(nsrc)])
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    for opthin, noalpha in ((True, True), (False, False)):
        ctx = _native.Context(0)
        ctx.set_model(500.0, opthin, noalpha)
        ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
        ctx.set_data(flux, chol=np.linalg.cholesky(cov))
        truth = (12.0, 1.8, 1300.0, 4.0, 30.0) if opthin else (14.0, 1.8, 400.0, 3.0, 30.0)
        P = synthetic.walker_cloud(truth, nsrc * nw, rng, low)
        want = np.empty(nsrc * nw)
        for s in range(nsrc):
            spec = oracle.LikeSpec(500.0, noalpha, opthin)
            spec.set_phot(waves, flux[s], unc[s])
            spec.set_cov(cov[s])
            want[s * nw:(s + 1) * nw] = oracle.loglike_batch(spec, P[s * nw:(s + 1) * nw])
        got, st = ctx.loglike(P, walkers_per_source=nw)
        assert (st == 0).all() and relerr(got, want).max() < TOL
        ctx.set_data(flux, cinv=np.linalg.inv(cov))
        inv, _ = ctx.loglike(P, walkers_per_source=nw)
        assert not np.array_equal(inv, got) and relerr(inv, got).max() < 1e-13
    bad = np.linalg.cholesky(cov)
    bad[2, 3, 3] = -1.0
    with pytest.raises(_native.MBBNativeError):
        ctx.set_data(flux, chol=bad)
    # a batch fit on factored covariances reproduces the explicit-inverse fit's posterior
    fits = []
    for cholesky in (False, True):
        bf = batch_fitter(nwalkers=nw, opthin=True, noalpha=True, device=0)
        bf.set_data(waves, flux, covmatrix=cov, cholesky=cholesky)
        p0 = bf.generate_initial_values((15.0, 1.8, 1300.0, 4.0, 30.0), (2, 0.2, 100, 0.3, 5), seed=3)
        fits.append(bf.run(20, 60, p0, seed=11))
    a, b = fits
    assert np.allclose(a.mean, b.mean, rtol=1e-6) and np.allclose(a.best_lnprob, b.best_lnprob, rtol=1e-9)


def test_small_batch_graph_replay(golden, oracle):
    """Host-buffer calls with <= 8192 parameter vectors (the 125-walker half-steps of a
    single-source fit, reference mbb_fit.py:533-542) are replayed as one CUDA graph from the
    third call of a shape on; same numbers as the plain path, and a graph never outlives the
    data / tables / priors it was captured with."""
    from mbb_emcee_b200 import synthetic
    for cfgname in ("cfg1", "cfg2", "cfg3"):
        cfg, like = _make_like(golden, cfgname, 1)
        rng = np.random.RandomState(31)
        P = synthetic.walker_cloud(cfg["truth"], 125, rng, like.lowlims)
        P[3, 0] = 0.2                                   # one walker below the T limit
        ctx = like.context
        first = like(P)                                 # plain path (sizes the buffers)
        n0 = ctx.launch_count()
        second = like(P)                                # captured, then replayed
        per_call = ctx.launch_count() - n0
        third = like(P)
        assert ctx.launch_count() - n0 == 2 * per_call and per_call >= 1
        assert np.array_equal(first, second) and np.array_equal(first, third)
        assert np.isneginf(first[3])
        want = oracle.loglike_batch(_oracle_spec(oracle, like), P)
        fin = np.isfinite(want)
        assert relerr(first[fin], want[fin]).max() < TOL
        # other rows through the same graph
        P2 = synthetic.walker_cloud(cfg["truth"], 125, rng, like.lowlims)
        assert relerr(like(P2), oracle.loglike_batch(_oracle_spec(oracle, like), P2)).max() < TOL
        # new photometry / a new prior: the graph is re-captured, never replayed stale
        cov = np.array(like.data_covmatrix) if like.has_data_covmatrix else None
        like.set_phot(cfg["bands"], like.data_flux * 1.1, like._flux_unc)
        if cov is not None:
            like.set_cov(cov)
        like.set_gaussian_prior('beta', 1.7, 0.4)
        a = like(P2)
        assert relerr(a, oracle.loglike_batch(_oracle_spec(oracle, like), P2)).max() < TOL
        assert np.array_equal(a, like(P2))
        # a single row (emcee's calling convention) has its own shape
        assert like(P2[5]) == a[5] and like(P2[5]) == a[5] and like(P2[5]) == a[5]


def test_delta_kernel_launch_paths(oracle):
    """The persistent delta kernel: TMA-fed full tiles, the direct-load fallback
    (partial tail tile, buffer not 16-byte aligned, odd SoA stride), walkers-per-source
    values that exercise the multiply-shift division, the cold paths (soft upper limit
    exceeded; exponents outside the double range -> saturating code) -- all against
    the oracle, and bitwise against each other where the arithmetic is the same."""
    import torch
    from mbb_emcee_b200 import _native, synthetic
    rng = np.random.RandomState(21)
    dev = torch.device("cuda:0")
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    low = np.array([0.01, 0.1, 1, 0.1, 1e-3])
    for opthin in (True, False):
        for wps in (1, 37, 256, 1000):
            n = 3 * 256 + 77                      # three full tiles + a partial one
            nsrc = (n + wps - 1) // wps
            flux = rng.uniform(5, 80, (nsrc, 6))
            unc = rng.uniform(1, 6, (nsrc, 6))
            ctx = _native.Context(0)
            ctx.set_model(500.0, opthin, True)
            ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
            ctx.set_data(flux, ivar=1.0 / unc**2)
            has_up = [0, 1, 1, 1, 0, 0]
            up = [np.inf, 2.2, 1500.0, 20.0, np.inf, np.inf]
            ctx.set_priors(low, has_up, up, [0] * 6, [0.0] * 6, [1.0] * 6)
            P = synthetic.walker_cloud((12.0, 1.8, 400.0, 4.0, 30.0), n, rng, low)
            P[5, 1] = 2.5                          # beta above its soft upper limit -> penalty term
            P[6, 0] = 0.05                         # h nu / k T ~ 4000: exponent outside the double range
            P[7, 0] = 0.005                        # below the T limit -> -inf
            Pd = torch.as_tensor(P, device=dev)
            out = torch.empty(n, dtype=torch.float64, device=dev)
            st = torch.empty(n, dtype=torch.int32, device=dev)
            ctx.loglike_device(n, Pd.data_ptr(), out.data_ptr(), st.data_ptr(), walkers_per_source=wps)
            ctx.sync()
            got, stat = out.cpu().numpy(), st.cpu().numpy()
            want = np.empty(n)
            for s0 in range(nsrc):
                spec = oracle.LikeSpec(500.0, True, opthin)
                spec.set_phot(waves, flux[s0], unc[s0])
                spec.lowlim = low.copy()
                spec.has_uplim = [bool(x) for x in has_up]
                spec.uplim = np.array(up)
                sl = slice(s0 * wps, min((s0 + 1) * wps, n))
                with np.errstate(all="ignore"):
                    want[sl] = oracle.loglike_batch(spec, P[sl])
            assert np.isneginf(got[7]) and np.isneginf(want[7]) and stat[7] == 1
            assert stat[6] == 0 and np.isfinite(got[6])
            fin = np.isfinite(want)
            assert relerr(got[fin], want[fin]).max() < TOL, (opthin, wps)
            # same rows through a buffer that is only 8-byte aligned (direct loads, no TMA)
            buf = torch.empty(n * 5 + 1, dtype=torch.float64, device=dev)
            buf[1:] = Pd.reshape(-1)
            out2 = torch.empty_like(out)
            ctx.loglike_device(n, buf.data_ptr() + 8, out2.data_ptr(), 0, walkers_per_source=wps)
            # SoA (odd row stride n -> direct loads as well)
            Pt = Pd.t().contiguous()
            out3 = torch.empty_like(out)
            ctx.loglike_device(n, Pt.data_ptr(), out3.data_ptr(), 0, walkers_per_source=wps, layout=_native.SOA)
            ctx.sync()
            assert bool(torch.equal(out, out2)) and bool(torch.equal(out, out3))


def test_host_pipeline_pinned_and_pageable():
    """mbb_loglike(MBB_HOST): the chunked three-slot pipeline gives the bits of one
    device-resident launch -- pageable and page-locked caller buffers, AoS and SoA,
    implicit and explicit source indices, a batch just past the pipelining threshold
    (8 chunks of a multiple of 256 + a ragged tail) and a small env-forced chunk."""
    import torch
    from mbb_emcee_b200 import _native, synthetic
    rng = np.random.RandomState(33)
    nw = 125
    n = (1 << 17) + 1000 - ((1 << 17) + 1000) % nw + nw      # whole sources, not a multiple of 256
    nsrc = n // nw
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(5, 80, (nsrc, 6))
    unc = rng.uniform(1, 6, (nsrc, 6))
    ctx = _native.Context(0)
    ctx.set_model(500.0, True, True)
    ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
    ctx.set_data(flux, ivar=1.0 / unc**2)
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    P = synthetic.walker_cloud((12.0, 1.8, 1300.0, 4.0, 30.0), n, rng, low)
    P[::997, 0] = 0.5
    dev = torch.device("cuda:0")
    Pd = torch.as_tensor(P, device=dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    ctx.loglike_device(n, Pd.data_ptr(), out.data_ptr(), st.data_ptr(), walkers_per_source=nw)
    ctx.sync()
    want, wst = out.cpu().numpy(), st.cpu().numpy()
    assert np.isneginf(want[::997]).all() and np.isfinite(want).sum() == n - len(want[::997])
    # pageable
    got, gst = ctx.loglike(P, walkers_per_source=nw)
    assert np.array_equal(got, want) and np.array_equal(gst, wst)
    # page-locked, used in place
    Pp = torch.empty((n, 5), dtype=torch.float64).pin_memory()
    Op = torch.empty(n, dtype=torch.float64).pin_memory()
    Sp = torch.empty(n, dtype=torch.int32).pin_memory()
    Pp.numpy()[...] = P
    Op.fill_(7.0)
    ctx.loglike_into(Pp.numpy(), Op.numpy(), Sp.numpy(), walkers_per_source=nw)
    assert np.array_equal(Op.numpy(), want) and np.array_equal(Sp.numpy(), wst)
    # SoA + explicit (shuffled) source indices, pageable and page-locked
    idx = rng.randint(0, nsrc, n).astype(np.int32)
    src = torch.as_tensor(idx, device=dev)
    ctx.loglike_device(n, Pd.data_ptr(), out.data_ptr(), 0, src_index_ptr=src.data_ptr())
    ctx.sync()
    want2 = out.cpu().numpy()
    assert not np.array_equal(want2, want)
    Pt = np.ascontiguousarray(P.T)
    got2, _ = ctx.loglike(Pt, src_index=idx, layout=_native.SOA)
    assert np.array_equal(got2, want2)
    Tp = torch.empty((5, n), dtype=torch.float64).pin_memory()
    Ip = torch.empty(n, dtype=torch.int32).pin_memory()
    Tp.numpy()[...] = Pt
    Ip.numpy()[...] = idx
    ctx.loglike_into(Tp.numpy(), Op.numpy(), None, src_index=Ip.numpy(), layout=_native.SOA)
    assert np.array_equal(Op.numpy(), want2)
    with pytest.raises(RuntimeError):
        bad = idx.copy()
        bad[12345] = nsrc
        ctx.loglike(P, src_index=bad)
    with pytest.raises(TypeError):
        ctx.loglike_into(P.astype(np.float32), Op.numpy())


def test_host_pipeline_fixed_columns():
    """mbb_set_fixed_params: SoA host batches whose fixed columns (mbb_fit.py:440-443 -- one
    value in every walker) are not copied to the device give the bits of the plain call;
    the caller's fixed columns are really not read (poisoned here); pageable and page-locked
    buffers, a ragged tail, a change of the fixed values, AoS calls in between and clearing
    the promise."""
    import torch
    from mbb_emcee_b200 import _native, synthetic
    rng = np.random.RandomState(34)
    nw = 125
    n = (1 << 17) + 2000 - ((1 << 17) + 2000) % nw + nw
    nsrc = n // nw
    waves = [70.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    flux = rng.uniform(5, 80, (nsrc, 6))
    unc = rng.uniform(1, 6, (nsrc, 6))
    ctx = _native.Context(0)
    ctx.set_model(500.0, False, False)          # thick + alpha: every column enters the model
    ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
    ctx.set_data(flux, ivar=1.0 / unc**2)
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    P = synthetic.walker_cloud((14.0, 1.8, 400.0, 3.0, 30.0), n, rng, low)
    for fixed, vals in (((0, 0, 1, 1, 0), (0, 0, 380.0, 2.5, 0)), ((0, 1, 0, 0, 0), (0, 1.7, 0, 0, 0))):
        Q = P.copy()
        for i in range(5):
            if fixed[i]:
                Q[:, i] = vals[i]
        want, wst = ctx.loglike(Q, walkers_per_source=nw)                # AoS, no promise in effect
        Qt = np.ascontiguousarray(Q.T)
        for i in range(5):
            if fixed[i]:
                Qt[i] = np.nan                                           # must not be read
        ctx.set_fixed_params(fixed, vals)
        got, gst = ctx.loglike(Qt, walkers_per_source=nw, layout=_native.SOA)
        assert np.array_equal(got, want) and np.array_equal(gst, wst)
        l0 = ctx.launch_count()
        got, _ = ctx.loglike(Qt, walkers_per_source=nw, layout=_native.SOA)   # columns already filled
        assert np.array_equal(got, want)
        l1 = ctx.launch_count()
        chk, _ = ctx.loglike(Q, walkers_per_source=nw)                   # AoS in between reuses the slots
        assert np.array_equal(chk, want)
        Tp = torch.empty((5, n), dtype=torch.float64).pin_memory()
        Op = torch.empty(n, dtype=torch.float64).pin_memory()
        Tp.numpy()[...] = Qt
        ctx.loglike_into(Tp.numpy(), Op.numpy(), None, walkers_per_source=nw, layout=_native.SOA)
        assert np.array_equal(Op.numpy(), want)
        assert l1 - l0 < ctx.launch_count() - l1                         # the refill after AoS launched fills
    ctx.set_fixed_params(None)
    got, _ = ctx.loglike(np.ascontiguousarray(P.T), walkers_per_source=nw, layout=_native.SOA)
    want, _ = ctx.loglike(P, walkers_per_source=nw)
    assert np.array_equal(got, want)
    with pytest.raises(RuntimeError):
        ctx.set_fixed_params((1, 0, 0, 0, 0), (np.inf, 0, 0, 0, 0))


def test_contexts_in_concurrent_threads():
    """SURVEY 8b threading contract: a context serialises its own calls, separate contexts
    may be driven from separate threads (ctypes drops the GIL for the call).  Two
    contexts with different models and band tables run host-path batches at the same
    time; each reproduces its single-threaded bits."""
    import threading
    from mbb_emcee_b200 import _native, likelihood, synthetic
    rng = np.random.RandomState(8)
    low = np.array([1, 0.1, 1, 0.1, 1e-3])
    a = _native.Context(0)
    a.set_model(500.0, True, True)
    a.set_bands(np.arange(7, dtype=np.int32), [70.0, 100.0, 160.0, 250.0, 350.0, 500.0], np.ones(6))
    a.set_data(rng.uniform(5, 80, (1, 6)), ivar=1.0 / rng.uniform(1, 6, (1, 6))**2)
    Pa = synthetic.walker_cloud((12.0, 1.8, 1300.0, 4.0, 30.0), 400_000, rng, low)
    cfg = synthetic.CONFIGS["cfg2"]
    like = likelihood(wavenorm=500.0, response=True)
    like.set_phot(cfg["bands"], [40.0, 70.0, 75.0, 50.0, 28.0, 6.0], [4.0, 7.0, 6.0, 5.0, 4.0, 1.5])
    Pb = synthetic.walker_cloud(cfg["truth"], 4000, rng, low)
    wa, _ = a.loglike(Pa)
    wb = like(Pb)
    got = {}
    errs = []

    def run(name, fn, reps):
        try:
            for _ in range(reps):
                got[name] = fn()
        except Exception as e:              # surfaced below: a thread must not fail silently
            errs.append((name, e))
    ta = threading.Thread(target=run, args=("a", lambda: a.loglike(Pa)[0], 6))
    tb = threading.Thread(target=run, args=("b", lambda: like(Pb), 6))
    ta.start(); tb.start(); ta.join(); tb.join()
    assert not errs, errs
    assert np.array_equal(got["a"], wa) and np.array_equal(got["b"], wb)


@pytest.mark.parametrize("cfgname", ["cfg2", "cfg3"])
def test_gauss_mode_golden_and_oracle(golden, oracle, cfgname):
    """MBB_MATH_FAST_GAUSS (tabulated passbands through their 32-point Gauss rules where the
    per-walker bound allows): the reference's golden lnlike values and the oracle on a random
    cloud, at the same 1e-12 bar; and agreement with the full-table FAST mode to 1e-13 over a
    much wider cloud that exercises the fallback (cold, steep, merge point inside a band)."""
    from mbb_emcee_b200 import _native, synthetic
    cfg, like = _make_like(golden, cfgname, _native.MATH_FAST_GAUSS)
    g = golden.like
    P, ref = g[cfgname + "_P"], g[cfgname + "_lnlike"]
    ll, st = like.evaluate(P)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isneginf(ll), np.isneginf(ref))
    assert relerr(ll[fin], ref[fin]).max() < TOL
    rng = np.random.RandomState(77)
    Pr = synthetic.walker_cloud(cfg["truth"], 600, rng, like.lowlims)
    want = oracle.loglike_batch(_oracle_spec(oracle, like), Pr)
    got, st = like.evaluate(Pr)
    fin = np.isfinite(want)
    assert relerr(got[fin], want[fin]).max() < TOL
    n = 60000                      # enough evaluations for the thread-per-evaluation kernel to be chosen
    Pw = np.stack([10**rng.uniform(np.log10(3), np.log10(80), n), rng.uniform(0.1, 9, n),
                   10**rng.uniform(1, 3.17, n), rng.uniform(0.5, 10, n), 10**rng.uniform(0, 2.5, n)], axis=1)
    comp, st3 = like.evaluate(Pw)
    like.math_mode = _native.MATH_FAST
    full, st1 = like.evaluate(Pw)
    assert np.array_equal(st1, st3)
    ok = np.isfinite(full) & (st1 == 0)
    assert ok.sum() > 0.5 * n
    assert relerr(comp[ok], full[ok]).max() < 1e-13
    assert not np.array_equal(comp[ok], full[ok])          # the compressed rules were actually used


@pytest.mark.parametrize("opthin", [True, False])
def test_gauss_thread_kernel_paths(oracle, opthin):
    """MBB_MATH_FAST_GAUSS without the power-law join runs one thread per evaluation
    (mbb_gausskernel.cuh): rule bands, bands that fail the per-walker bound (cold walker ->
    full table from global memory), exponents outside the double range (saturating code),
    the -inf gate, several sources, a partial last tile, soft-limit penalties -- all
    against the oracle; SoA input bitwise equal to AoS."""
    import torch
    from mbb_emcee_b200 import _native, likelihood, synthetic
    cfg = synthetic.CONFIGS["cfg2"]
    rng = np.random.RandomState(31)
    like = likelihood(wavenorm=500.0, opthin=opthin, noalpha=True, response=True, device=0)
    nb = len(cfg["bands"])
    nsrc, wps = 1600, 37           # 59 200 evaluations: the launcher picks the thread kernel from 148*384 up
    n = nsrc * wps
    ncheck = 21 * wps              # oracle-checked prefix (the rest is compared with FAST on the device)
    like.set_phot(cfg["bands"], np.full(nb, 30.0), np.full(nb, 3.0))
    like.set_lowlim('T', 0.01)
    flux = rng.uniform(5, 80, (nsrc, nb))
    unc = rng.uniform(1, 6, (nsrc, nb))
    P = synthetic.walker_cloud((14.0, 1.8, 400.0, 3.0, 30.0), n, rng, like.lowlims)
    P[3] = [2.0, 1.8, 400.0, 3.0, 30.0]            # cold: PACS bands fail the bound -> full tables
    P[4] = [0.05, 1.8, 400.0, 3.0, 30.0]           # h nu / k T ~ 4000 -> saturating path
    P[5] = [0.001, 1.8, 400.0, 3.0, 30.0]          # below the T limit -> -inf
    P[6, 1] = 25.0                                 # beta above its soft limit (and a steep thick factor)
    like._stage()
    ctx = like.context
    ctx.set_data(flux, ivar=1.0 / unc**2)
    ctx.set_math_mode(_native.MATH_FAST)
    full, st1 = ctx.loglike(P, walkers_per_source=wps)
    ctx.set_math_mode(_native.MATH_FAST_GAUSS)
    got, st = ctx.loglike(P, walkers_per_source=wps)
    assert np.array_equal(st, st1)
    ok = np.isfinite(full)
    assert relerr(got[ok], full[ok]).max() < 1e-13 and not np.array_equal(got[ok], full[ok])
    spec0 = _oracle_spec(oracle, like)
    want = np.empty(ncheck)
    for s0 in range(ncheck // wps):
        spec = oracle.LikeSpec(500.0, True, opthin)
        spec.set_phot(spec0.bands, flux[s0], unc[s0])
        spec.lowlim, spec.has_uplim, spec.uplim = spec0.lowlim, spec0.has_uplim, spec0.uplim
        with np.errstate(all="ignore"):
            want[s0 * wps:(s0 + 1) * wps] = oracle.loglike_batch(spec, P[s0 * wps:(s0 + 1) * wps])
    assert np.isneginf(got[5]) and np.isneginf(want[5]) and st[5] == 1
    assert st[3] == 0 and st[4] == 0 and st[6] == 0
    fin = np.isfinite(want)
    assert relerr(got[:ncheck][fin], want[fin]).max() < TOL
    dev = torch.device("cuda:0")
    Pt = torch.as_tensor(P, device=dev).t().contiguous()
    out = torch.empty(n, dtype=torch.float64, device=dev)
    ctx.loglike_device(n, Pt.data_ptr(), out.data_ptr(), 0, walkers_per_source=wps, layout=_native.SOA)
    ctx.sync()
    assert np.array_equal(out.cpu().numpy(), got)


def test_error_statuses():
    """Failures that make the reference raise surface as the same exception types."""
    from mbb_emcee_b200 import likelihood
    like = likelihood(device=0)
    like.set_phot([250.0, 350.0, 500.0], [30.0, 40.0, 30.0], [3.0, 4.0, 3.0])
    like.set_lowlim('alpha', -5.0)
    with pytest.raises(ValueError):
        like([10.0, 2.0, 100.0, -1.0, 30.0])          # alpha <= 0
    like.set_lowlim('beta', -5.0)
    with pytest.raises(ValueError):
        like([10.0, -1.0, 100.0, 2.0, 30.0])          # beta < 0
    ll, st = like.evaluate(np.array([[10.0, 2.0, 100.0, 2.0, 30.0], [np.nan, 2.0, 100.0, 2.0, 30.0]]))
    assert st[0] == 0 and np.isfinite(ll[0]) and st[1] == 9 and np.isnan(ll[1])
    with pytest.raises(ValueError):
        like(np.zeros((3, 4)))


# ------------------------------------------------------------ full-size properties
def test_full_size_properties(oracle):
    """BASELINE configs[4]: 1e5 sources x 512 walkers in one device-resident
    launch.  Properties that do not need the oracle at full size, plus an
    oracle check of a random sample of the 5.12e7 results."""
    import torch
    from mbb_emcee_b200 import _native, synthetic
    cfg = synthetic.CONFIGS["cfg5"]
    nsrc, nw = cfg["nsources"], cfg["nwalkers"]
    n = nsrc * nw
    rng = np.random.RandomState(cfg["seed"])
    waves = np.array(cfg["bands"])
    flux = rng.uniform(5.0, 80.0, (nsrc, 6))
    unc = np.maximum(0.1 * flux, 1.0)
    ctx = _native.Context(0)
    ctx.set_model(cfg["wavenorm"], cfg["opthin"], cfg["noalpha"])
    ctx.set_bands(np.arange(7, dtype=np.int32), waves, np.ones(6))
    ctx.set_data(flux, ivar=1.0 / unc**2)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    truth = torch.tensor([12.0, 1.8, 1300.0, 4.0, 30.0], dtype=torch.float64, device=dev)
    sig = torch.tensor(synthetic.P0_SIGMA, dtype=torch.float64, device=dev)
    P = truth + sig * torch.randn((n, 5), dtype=torch.float64, device=dev, generator=g)
    P[:, 0].clamp_(min=2.0)
    P[:, 1].clamp_(min=0.2)
    P[:, 4].clamp_(min=0.5)
    P[::1000, 0] = 0.5                                   # sprinkle -inf gates
    out = torch.empty(n, dtype=torch.float64, device=dev)
    st = torch.empty(n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ctx.loglike_device(n, P.data_ptr(), out.data_ptr(), st.data_ptr(), walkers_per_source=nw)
    ctx.sync()
    # (1) gate: exactly the sprinkled rows are -inf, nothing is NaN
    assert bool(torch.isneginf(out[::1000]).all())
    assert int(torch.isneginf(out).sum()) == (n + 999) // 1000
    assert not bool(torch.isnan(out).any())
    assert int((st > 1).sum()) == 0
    # (2) determinism: a second launch is bit-identical
    out2 = torch.empty_like(out)
    ctx.loglike_device(n, P.data_ptr(), out2.data_ptr(), 0, walkers_per_source=nw)
    ctx.sync()
    assert bool(torch.equal(out, out2))
    # (3) layout invariance: SoA input gives bit-identical results
    Pt = P.t().contiguous()
    out3 = torch.empty_like(out)
    ctx.loglike_device(n, Pt.data_ptr(), out3.data_ptr(), 0, walkers_per_source=nw, layout=_native.SOA)
    ctx.sync()
    assert bool(torch.equal(out, out3))
    # (4) explicit source indices reproduce the implicit i // nw mapping
    src = (torch.arange(n, device=dev) // nw).to(torch.int32)
    out4 = torch.empty_like(out)
    ctx.loglike_device(n, P.data_ptr(), out4.data_ptr(), 0, src_index_ptr=src.data_ptr())
    ctx.sync()
    assert bool(torch.equal(out, out4))
    # (5) the pipelined host path (pageable numpy in/out) gives the same bits
    sl = slice(0, 5_000_000)
    host, hst = ctx.loglike(P[sl].cpu().numpy(), walkers_per_source=nw)
    assert np.array_equal(host, out[sl].cpu().numpy())
    # (6) a random sample of the full-size results against the oracle
    pick = rng.randint(0, n, 400)
    Ph = P[torch.as_tensor(pick, device=dev)].cpu().numpy()
    got = out[torch.as_tensor(pick, device=dev)].cpu().numpy()
    for j, i in enumerate(pick):
        spec = oracle.LikeSpec(cfg["wavenorm"], cfg["noalpha"], cfg["opthin"])
        spec.set_phot(waves, flux[i // nw], unc[i // nw])
        want = oracle.loglike(spec, Ph[j])
        assert (np.isneginf(want) and np.isneginf(got[j])) or abs(got[j] - want) <= TOL * abs(want)


# ------------------------------------------------------------------ chain identity
@pytest.mark.parametrize("cfgname,nwalk,nsteps", [("cfg1", 250, 120), ("cfg2", 60, 40), ("cfg3", 40, 25)])
def test_chain_bit_identity(golden, oracle, cfgname, nwalk, nsteps):
    """Same RNG seed, same emcee-2.2 move sequence: the chain driven by the
    device log-probability is bit-identical to the one driven by the CPU
    oracle, and the stored log-probabilities agree to 1e-12."""
    from mbb_emcee_b200 import mbb_fitter, synthetic
    cfg, like = _make_like(golden, cfgname, 1)
    fit = mbb_fitter(nwalkers=nwalk, wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"],
                     opthin=cfg["opthin"], response=cfg["response"], device=0)
    g = golden.like
    fit.set_data(cfg["bands"], g[cfgname + "_flux"], g[cfgname + "_unc"],
                 covmatrix=g[cfgname + "_cov"] if cfgname + "_cov" in g else None)
    for nm, v in cfg.get("uplims", []):
        fit.set_uplim(nm, v)
    for nm, m, s in cfg.get("gpriors", []):
        fit.set_gaussian_prior(nm, m, s)
    if cfg["noalpha"]:
        fit.fix_param('alpha')
    np.random.seed(cfg["seed"])
    p0 = fit.generate_initial_values(list(cfg["truth"]), [2, 0.2, 100, 0.3, 5.0])
    fit.sampler.random_state = np.random.RandomState(2024).get_state()
    fit.sampler.reset()
    fit.sampler.run_mcmc(p0, nsteps)
    spec = _oracle_spec(oracle, fit.like)
    chain, lnp, nacc = oracle.stretch_chain(lambda P: oracle.loglike_batch(spec, P), p0, nsteps,
                                            np.random.RandomState(2024))
    assert np.array_equal(fit.sampler.chain, chain)
    assert np.array_equal(fit.sampler.naccepted, nacc)
    assert relerr(fit.sampler.lnprobability, lnp).max() < TOL
    assert 0.05 < np.mean(fit.sampler.acceptance_fraction) < 0.95
    if cfg["noalpha"]:
        assert (fit.sampler.chain[:, :, 3] == cfg["truth"][3]).all()     # fixed stays fixed


def test_fitter_run_end_to_end(golden, oracle):
    """mbb_fitter.run (burn-in, reset, main chain) and mbb_results statistics."""
    from mbb_emcee_b200 import mbb_fitter, mbb_results, synthetic
    cfg = synthetic.CONFIGS["cfg1"]
    g = golden.like
    fit = mbb_fitter(nwalkers=100, opthin=True, noalpha=True, device=0)
    fit.set_data(cfg["bands"], g["cfg1_flux"], g["cfg1_unc"])
    fit.fix_param('alpha')
    np.random.seed(1)
    p0 = fit.generate_initial_values([12.0, 1.8, 2500.0, 4.0, 30.0], [2, 0.2, 100, 0.3, 5.0])
    fit.run(30, 150, p0)
    assert fit.sampled and fit.sampler.chain.shape == (100, 150, 5)
    ch, lnp = fit.sampler.chain, fit.sampler.lnprobability
    # every recorded log-probability is the oracle's for the recorded position
    spec = _oracle_spec(oracle, fit.like)
    pick = np.random.RandomState(2).randint(0, 100 * 150, 300)
    want = oracle.loglike_batch(spec, ch.reshape(-1, 5)[pick])
    assert relerr(lnp.reshape(-1)[pick], want).max() < TOL
    acc = fit.sampler.acceptance_fraction
    assert acc.shape == (100,) and 0.15 < acc.mean() < 0.8
    res = mbb_results(fit=fit, redshift=2.0, lumdist=1.6e4)
    # statistics: the reference's formulas (results.py:314-369, 399-431) on the same chain
    for name, col in (("T", 0), ("beta", 1), ("fnorm", 4)):
        cen = res.par_cen(name)
        x = ch[:, :, col].ravel()
        lo, hi = np.percentile(x, [0.5 * (100 - 68.3), 100 - 0.5 * (100 - 68.3)])
        assert np.allclose(cen, [x.mean(), hi - x.mean(), x.mean() - lo], rtol=1e-12, atol=0)
    w, t = np.unravel_index(np.argmax(lnp), lnp.shape)
    assert np.array_equal(res.best_fit[0], ch[w, t]) and res.best_fit[1] == lnp[w, t]
    assert res.best_fit_chisq == -2.0 * lnp[w, t]
    assert abs(res.par_cen('T')[0] - 12.0) < 3.0
    res.compute_dustmass()
    res.compute_peaklambda()
    assert np.isfinite(res.dustmass).all() and np.isfinite(res.peaklambda).all()
    # ancillaries of individual samples against the oracle (thin, no alpha; the reference builds the
    # peak-wavelength SED thick with alpha whatever the fit used, results.py:574-580)
    dc = oracle.dustmass_consts(2.0, 500.0, 125.0, 1.6e4)
    for (w, t) in ((0, 0), (17, 80), (99, 149)):
        step = ch[w, t]
        assert abs(res.peaklambda[w, t] - oracle.peaklambda_step(step)) <= TOL * res.peaklambda[w, t]
        dm = oracle.dustmass_step(step, 2.64, 500.0, True, *dc)
        assert abs(res.dustmass[w, t] - dm) <= TOL * abs(dm)
    assert "ChiSquare" in str(res)


# ------------------------------------------------------------- chain post-processing
def test_cli_end_to_end(tmp_path):
    """The command line front end on a small photometry file: fit, ancillaries, save/load."""
    from mbb_emcee_b200 import mbb_results, run_mbb_emcee as cli
    phot = tmp_path / "phot.txt"
    phot.write_text("# wave flux err\n100 20 3\n160 60 6\n250 80 6\n350 55 5\n500 30 4\n850 6 1.5\n")
    out = tmp_path / "res.npz"
    res = cli.main([str(phot), str(out), "-n", "60", "-N", "30", "-b", "10", "--initT", "12", "--initLambda0", "300",
                    "--initFnorm", "30", "--seed", "3", "-z", "2.0", "--lumdist", "16000", "--get_peaklambda",
                    "--get_lir", "--get_dustmass", "--priorBeta", "1.8", "0.3"])
    assert res.chain.shape == (60, 30, 5) and np.isfinite(res.lnprobability).all()
    assert res.has_peaklambda and res.has_lir and res.has_dustmass
    # what the command line computed is what the classes compute from the same file and settings:
    # the recorded log-probabilities are the likelihood of the recorded positions (beta prior incl.)
    from mbb_emcee_b200 import likelihood
    like = likelihood(photfile=str(phot), device=0)
    like.set_gaussian_prior("beta", 1.8, 0.3)
    flat, fl = res.chain.reshape(-1, 5), res.lnprobability.reshape(-1)
    pick = np.random.RandomState(1).randint(0, flat.shape[0], 200)
    assert relerr(like(flat[pick]), fl[pick]).max() < 1e-12
    # the same command with the stretch move on the device: same posterior within Monte-Carlo error
    out2 = tmp_path / "res_dev.npz"
    res2 = cli.main([str(phot), str(out2), "-n", "60", "-N", "300", "-b", "100", "--initT", "12", "--initLambda0", "300",
                     "--initFnorm", "30", "--seed", "3", "--sampler", "device", "--priorBeta", "1.8", "0.3"])
    res1 = cli.main([str(phot), str(tmp_path / "res_host.npz"), "-n", "60", "-N", "300", "-b", "100", "--initT", "12",
                     "--initLambda0", "300", "--initFnorm", "30", "--seed", "3", "--priorBeta", "1.8", "0.3"])
    assert res2.chain.shape == (60, 300, 5)
    for col in (0, 1, 4):
        a, b = res1.chain[:, ::30, col].ravel(), res2.chain[:, ::30, col].ravel()
        assert abs(a.mean() - b.mean()) < 5.0 * np.hypot(a.std(), b.std()) / np.sqrt(a.size / 2.0)
    # L_IR of a sample = the SED class's own integral (8-1000 um rest frame, z = 2)
    from mbb_emcee_b200 import modified_blackbody
    s0 = res.chain[3, 11]
    m = modified_blackbody(s0[0], s0[1], s0[2], s0[3], s0[4], wavenorm=500.0)
    lir = 3.11749657e4 * 16000.0**2 * m.freq_integrate(8.0 * 3.0, 1000.0 * 3.0)
    assert abs(res.lir[3, 11] - lir) <= 5e-9 * lir
    back = mbb_results.load(str(out), device=0)
    assert np.array_equal(back.chain, res.chain)
    assert np.array_equal(back.lir_chain, res.lir_chain)


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
def test_chain_post_golden(golden, observed, name, opthin, noalpha):
    from mbb_emcee_b200 import mbb_results, synthetic
    cfg = synthetic.CONFIGS["cfg4"]
    g = golden.results
    chain = g[name + "_chain"]
    res = mbb_results.from_chain(chain, wavenorm=cfg["wavenorm"], noalpha=noalpha, opthin=opthin,
                                 redshift=cfg["z"], lumdist=cfg["lumdist"], device=0)
    res.compute_peaklambda()
    assert relerr(res.peaklambda, g[name + "_peaklambda"]).max() < TOL
    res.compute_dustmass(kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
    assert relerr(res.dustmass, g[name + "_dustmass"]).max() < TOL
    # L_IR, default method: replay of the reference's scipy.integrate.quad call.
    # Same subdivision sequence => 1e-15 agreement.  The adaptive algorithm is
    # chaotic at the ulp level, though: where a termination / ordering test of
    # QUADPACK is decided by the last bit of an integrand value (libdevice pow
    # vs glibc pow), the two runs subdivide differently and differ by the
    # quadrature's own error.  Bar: >= 99.5 % of the samples within 1e-12, the rest
    # within 5e-9.  Observed on B200 with the rule's products and sums rounded separately
    # (mbb_quadpack.cuh q_mul / q_add; a contracted abscissa moved 1 sample in 100 before): all 720
    # golden samples of each variant <= 9e-16, and 3000 of 3000 random chain samples <= 9e-16
    # against scipy (60 % of them bit-identical).
    res.compute_lir(wavemin=cfg["lir"][0], wavemax=cfg["lir"][1])
    dev = relerr(res.lir, g[name + "_lir"])
    observed.record("L_IR quadpack replay vs golden, max [%s]" % name, dev.max())
    observed.record("L_IR quadpack replay vs golden, fraction within 1e-12 [%s]" % name, np.mean(dev < TOL))
    assert dev.max() < 5e-9
    assert np.mean(dev < TOL) >= 0.995, np.mean(dev < TOL)
    lir_q = res.lir.copy()
    # L_IR, fixed-rule quadrature: the true integral (1e-13 vs 40-digit mpmath in
    # test_device_logic_cpu.py::test_freq_integrate); the reference's own quad error
    # reaches 2.7e-8 with the merge kink inside the range
    res.compute_lir(wavemin=cfg["lir"][0], wavemax=cfg["lir"][1], method="gauss")
    assert relerr(res.lir, lir_q).max() < 1e-7
    res.lir = lir_q
    # the allclose-dedupe: step 5 of walker 0 is within 3e-6 of step 4
    assert res.peaklambda[0, 5] == res.peaklambda[0, 4]
    assert res.dustmass[0, 5] == res.dustmass[0, 4]
    assert res.lir[0, 5] == res.lir[0, 4]


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
def test_freq_integrate_golden(golden, oracle, name, opthin, noalpha):
    """modified_blackbody.freq_integrate on the device: the QUADPACK replay
    reproduces the reference to 1e-12; the fixed-rule quadrature agrees with the
    CPU emulation of the same rule to 1e-13 (and with the truth, see CPU tests)."""
    from mbb_emcee_b200 import modified_blackbody
    import hostemu_lib as emu
    g = golden.sed
    tag = name + "_wn250"
    P = g[tag + "_P"][:12]
    ref = g[tag + "_freqint"][:12]
    want, st = emu.lir(opthin, noalpha, P, 250.0, 24.0, 3000.0)
    nclose = 0
    for i in range(len(P)):
        m = modified_blackbody(P[i, 0], P[i, 1], P[i, 2], P[i, 3], P[i, 4], wavenorm=250.0,
                               noalpha=noalpha, opthin=opthin)
        got = m.freq_integrate(24.0, 3000.0)                          # QUADPACK replay
        nclose += abs(got - ref[i]) <= TOL * abs(ref[i])
        assert abs(got - ref[i]) <= 5e-9 * abs(ref[i])
        got = m.freq_integrate(24.0, 3000.0, method="gauss")
        assert abs(got - ref[i]) <= 1e-7 * abs(ref[i])
        assert abs(got - want[i]) <= 1e-13 * abs(want[i])
        assert abs(m.max_wave() - g[tag + "_maxwave"][i]) <= TOL * g[tag + "_maxwave"][i]
    assert nclose >= len(P) - 1        # see test_chain_post_golden on the rare subdivision flips


def test_predict_flux_vs_oracle(oracle):
    """mbb_results._predict_flux (results.py:895-944): a wavelength and a passband."""
    from mbb_emcee_b200 import mbb_fitter, mbb_results, synthetic
    rng = np.random.RandomState(8)
    chain = synthetic.random_walk_chain((14.0, 1.8, 400.0, 3.0, 30.0), 4, 12, rng)
    fit = mbb_fitter(nwalkers=12, response=True, device=0)
    fit.set_data(["SPIRE_250um", "SPIRE_500um"], [30.0, 20.0], [3.0, 2.0])
    res = mbb_results.from_chain(chain, device=0)
    res._response_integrate = True
    res._responsewheel = fit.like._responsewheel
    got_w = res._predict_flux(350.0)
    got_b = res._predict_flux("SPIRE_250um")
    band = oracle.band_from_response(fit.like._responsewheel["SPIRE_250um"])
    for w in range(4):
        for t in range(12):
            s = oracle.make_sed(*chain[w, t])                    # default wavenorm, thick + alpha
            assert abs(got_w[w, t] - oracle.sed_call(s, 350.0)[0]) <= TOL * got_w[w, t]
            assert abs(got_b[w, t] - oracle.band_flux(s, band)) <= TOL * got_b[w, t]
    cen = res.predflux_cen(350.0)
    assert cen.shape == (3,) and cen[0] > 0


def test_chain_stats_equal_numpy(ctx):
    """mbb_chain_stats (what mbb_results' par_cen / *_cen / par_lowlim / par_uplim take from
    numpy.mean and numpy.percentile, reference results.py:314-431): percentiles equal numpy's bit
    for bit (exact order statistics + numpy's interpolation), means to 1e-14, on a 5-column chain
    block with repeats, negative values and ties, with clipping, for a device-resident block;
    NaN and empty selections behave like numpy / the reference."""
    import torch
    from mbb_emcee_b200 import _native, mbb_results, results, synthetic
    rng = np.random.RandomState(11)
    n = 400000
    X = np.array([14.0, 1.8, 400.0, 3.0, 30.0]) + rng.normal(size=(n, 5)) * [2, 0.2, 100, 0.3, 5]
    X[:, 3] = np.round(X[:, 3], 2)                    # heavy ties
    X[::7, 4] = X[3::7, 4][:X[::7, 4].size]           # exact repeats, as a chain has
    X[:, 1] -= 1.8                                    # both signs
    pcs = [15.85, 84.15, 50.0, 99.9]
    mean, count, perc = ctx.chain_stats(X, pcs)
    assert np.array_equal(count, np.full(5, n))
    import math
    exact = np.array([math.fsum(X[:, i]) / n for i in range(5)])
    assert relerr(mean, exact).max() < 4e-16                     # (the second column's mean is ~1e-4 of its spread)
    flat_means = np.array([X[:, i].copy().mean() for i in range(5)])      # the reference's flatten().mean()
    assert np.abs(mean - flat_means).max() < 1e-13
    assert np.array_equal(perc, np.percentile(X, pcs, axis=0).T)
    assert np.array_equal(ctx.chain_stats(X, [0.0, 100.0])[2], np.stack([X.min(0), X.max(0)], 1))
    # clipping per column, 1-D input
    lo, hi = 12.5, 16.0
    m1, c1, p1 = ctx.chain_stats(X[:, 0].copy(), [2.5, 97.5], lowlim=lo, uplim=hi)
    kept = X[:, 0][(X[:, 0] >= lo) & (X[:, 0] <= hi)]
    assert c1[0] == kept.size and abs(m1[0] - math.fsum(kept) / kept.size) < 4e-16 * kept.mean()
    assert np.array_equal(p1[0], np.percentile(kept, [2.5, 97.5]))
    # device-resident block
    T = torch.from_numpy(X).cuda()
    m2, c2, p2 = ctx.chain_stats(T.data_ptr(), pcs, where=_native.DEVICE, nrows=n, ncols=5)
    assert np.array_equal(p2, perc) and np.array_equal(m2, mean)
    # NaN in a column -> NaN out for that column only; nothing kept -> count 0
    Y = X.copy()
    Y[5, 2] = np.nan
    m3, c3, p3 = ctx.chain_stats(Y, [50.0])
    assert np.isnan(m3[2]) and np.isnan(p3[2, 0]) and np.array_equal(p3[[0, 1, 3, 4], 0], np.percentile(X, 50.0, axis=0)[[0, 1, 3, 4]])
    assert ctx.chain_stats(X[:, 0].copy(), [50.0], lowlim=1e9)[1][0] == 0
    # mbb_results routes big chains through it: same triples as numpy on the flattened chain
    assert n >= results.DEVICE_STATS_MIN
    chain = X.reshape(400, 1000, 5).copy()
    chain[:, :, 1] += 1.8
    res = mbb_results.from_chain(chain, device=0)
    for i, name in enumerate(("T", "beta", "lambda0", "alpha", "fnorm")):
        flat = chain[:, :, i].ravel()
        mn = flat.mean()
        pl, ph = np.percentile(flat, [15.85, 84.15])
        got = res.par_cen(name)
        assert abs(got[0] - mn) < 1e-14 * abs(mn)
        assert abs((got[0] + got[1]) - ph) < 1e-13 * abs(ph) and abs((got[0] - got[2]) - pl) < 1e-13 * abs(pl)
        assert np.allclose(res.par_central_values[i], got, rtol=0, atol=1e-12 * abs(mn))
        assert res.par_uplim(name, 95.0) == np.percentile(flat, 95.0)
        assert res.par_lowlim(name, 95.0) == np.percentile(flat, 5.0)
    with pytest.raises(Exception, match="No elements survive"):
        res.par_cen("T", lowlim=1e9)


def test_results_shard_rows_over_devices():
    """mbb_results(devices=[...]): walker rows cut into one shard per context / host thread, the
    outputs landing in disjoint rows of one page-locked array; identical to the single-context
    result (the repeat scan never crosses a walker)."""
    from mbb_emcee_b200 import mbb_results, synthetic
    chain = synthetic.random_walk_chain((14.0, 1.8, 400.0, 3.0, 30.0), 11, 70, np.random.RandomState(8))
    kw = dict(wavenorm=500.0, noalpha=False, opthin=False, redshift=2.0, lumdist=1.6e4)
    one = mbb_results.from_chain(chain, device=0, **kw)
    many = mbb_results.from_chain(chain, devices=[0, 0, 0], **kw)
    for r in (one, many):
        r.compute_peaklambda()
        r.compute_lir()
        r.compute_dustmass()
    assert np.array_equal(one.peaklambda, many.peaklambda)
    assert np.array_equal(one.lir, many.lir)
    assert np.array_equal(one.dustmass, many.dustmass)
    quad = one.lir.copy()
    for r in (one, many):
        r.compute_lir(method="gauss")             # the method reaches every shard's context
    assert np.array_equal(one.lir, many.lir) and not np.array_equal(many.lir, quad)
    assert np.median(relerr(many.lir, quad)) < 1e-8
    assert np.array_equal(many.lir_cen(), many._parcen_internal(many.lir.flatten(), 68.3))


def test_chain_dedupe_rule_on_adversarial_chains(oracle):
    """The allclose-dedupe of results._map_chain (results.py:553-566) is sequential -- a step is
    compared with the last KEPT step -- while the device cuts every walker into segments at steps
    that are certainly new.  Chains built to stress the difference: slow drifts (every step within
    the tolerance of its predecessor, but not of the kept one), jumps between one and two
    tolerances (new, yet not a certain segment start), runs of more than 6 such steps, exact
    repeats, constants.  Every output must be the value computed for the owner that a numpy replay
    of the sequential rule assigns."""
    from mbb_emcee_b200 import mbb_results
    rng = np.random.RandomState(12)
    nw, ns = 7, 700
    base = np.array([14.0, 1.8, 400.0, 3.0, 30.0])
    chain = np.empty((nw, ns, 5))
    for w in range(nw):
        cur = base * (1.0 + 0.05 * rng.standard_normal(5))
        for t in range(ns):
            kind = w if w < 6 else rng.randint(0, 6)
            if kind == 0:                                   # ordinary chain: 35 % accepted moves, else exact repeats
                if rng.uniform() < 0.35:
                    cur = cur + np.array([0.3, 0.03, 10.0, 0.05, 0.8]) * rng.standard_normal(5)
            elif kind == 1:                                 # slow drift, 4e-6 per step
                cur = cur * (1.0 + 4e-6)
            elif kind == 2:                                 # jumps of 1.5 tolerances, alternating sign
                cur = cur * (1.0 + (1.5e-5 if t % 2 else -1.5e-5))
            elif kind == 3:                                 # constant
                pass
            elif kind == 4:                                 # a run of steps each 1.2 tolerances from the previous one
                cur = cur * (1.0 + 1.2e-5)
            else:                                           # one component moves at a time, just above / below tol
                k = t % 5
                cur = cur.copy()
                cur[k] *= 1.0 + (1.02e-5 if (t // 5) % 2 else 0.98e-5)
            chain[w, t] = np.maximum(cur, [3.0, 0.3, 30.0, 0.6, 1.0])
    owner = np.empty((nw, ns), dtype=int)
    for w in range(nw):
        prev, pt = None, 0
        for t in range(ns):
            if prev is None or not np.allclose(prev, chain[w, t]):
                prev, pt = chain[w, t], t
            owner[w, t] = pt
    frac_new = np.mean(owner == np.arange(ns)[None, :])
    assert 0.2 < frac_new < 0.9 and (np.diff(owner[1]) == 0).any() and (owner[1] > 0).any()
    res = mbb_results.from_chain(chain, wavenorm=500.0, noalpha=False, opthin=False, redshift=2.0,
                                 lumdist=1.6e4, device=0)
    res.compute_dustmass()
    res.compute_peaklambda()
    dc = oracle.dustmass_consts(2.0, 500.0, 125.0, 1.6e4)
    uniq = {}
    for w in range(nw):
        for t in range(ns):
            o = owner[w, t]
            if (w, o) not in uniq:
                uniq[(w, o)] = oracle.dustmass_step(chain[w, o], 2.64, 500.0, False, *dc)
            want = uniq[(w, o)]
            assert abs(res.dustmass[w, t] - want) <= 1e-13 * want, (w, t, o)
            assert res.peaklambda[w, t] == res.peaklambda[w, o]
    # a relative change of 4e-6 in T moves the dust mass by far more than 1e-13: a wrong owner shows
    assert abs(uniq[(1, 0)] / uniq[(1, owner[1, -1])] - 1.0) > 1e-9


def test_chain_post_host_path_equals_device_path():
    """mbb_chain_post(MBB_HOST) on a chain of >= 2^20 samples works through 8 chunks of walker rows
    on three streams (copies of one chunk behind the kernels of its neighbours); every output equals
    the single-launch device-resident call bit for bit, with page-locked and pageable host arrays
    and a walker count that the chunks do not divide."""
    import ctypes
    import torch
    from mbb_emcee_b200 import _native, synthetic
    nw, ns = 19, 56000                                   # 1 064 000 samples
    chain = synthetic.random_walk_chain((14.0, 1.8, 400.0, 3.0, 30.0), nw, ns, np.random.RandomState(9))
    ctx = _native.Context(0)
    ctx.set_model(500.0, False, False)
    dev = torch.device("cuda:0")
    ch = torch.as_tensor(chain, device=dev)
    outs = [torch.empty((nw, ns), dtype=torch.float64, device=dev) for _ in range(3)]
    st = torch.empty((nw, ns), dtype=torch.int32, device=dev)
    vp = ctypes.c_void_p
    rc = ctx._lib.mbb_chain_post(ctx._h, nw, ns, vp(ch.data_ptr()), 7, 2.0, 1.6e4, 8.0, 1000.0, 2.64, 125.0,
                                 vp(outs[0].data_ptr()), vp(outs[1].data_ptr()), vp(outs[2].data_ptr()),
                                 vp(st.data_ptr()), 1)
    assert rc == 0
    ctx.sync()
    want = [o.cpu().numpy() for o in outs] + [st.cpu().numpy()]
    for pinned in (False, True):
        alloc = _native.pinned_empty if pinned else (lambda shape, dt=np.float64: np.empty(shape, dtype=dt))
        cin = alloc((nw, ns, 5))
        cin[...] = chain
        got = [alloc((nw, ns)) for _ in range(3)] + [alloc((nw, ns), np.int32)]
        for g in got:
            g[...] = -7
        ctx.chain_post_into(cin, 7, peak=got[0], lir=got[1], dustmass=got[2], status=got[3], z=2.0, dl_mpc=1.6e4)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    assert np.isfinite(want[1]).all() and (want[3] == 0).all()


def test_chain_post_full_size_properties(oracle):
    """BASELINE configs[3]: a 10^7-sample chain (500 walkers x 20000 steps, 35% of
    the steps new).  Size-independent properties + an oracle-checked sample."""
    import hostemu_lib as emu
    from mbb_emcee_b200 import mbb_results, synthetic
    cfg = synthetic.CONFIGS["cfg4"]
    rng = np.random.RandomState(cfg["seed"])
    nw, ns = 500, 20000
    chain = synthetic.random_walk_chain(cfg["truth"], nw, ns, rng)
    res = mbb_results.from_chain(chain, wavenorm=cfg["wavenorm"], redshift=cfg["z"],
                                 lumdist=cfg["lumdist"], device=0)
    res.compute_peaklambda()
    res.compute_dustmass(kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
    res.compute_lir(*cfg["lir"], method="gauss")
    lir_gauss = res.lir.copy()
    res.compute_lir(*cfg["lir"])                      # default: QUADPACK replay
    # the replay inherits the reference's quadrature error; where that error is
    # largest it must still BE the reference's number (checked against scipy below)
    dev = relerr(res.lir, lir_gauss)
    assert np.median(dev) < 1e-8
    pref0 = oracle.LIR_PREFAC * cfg["lumdist"]**2
    for flat in np.argsort(dev, axis=None)[-4:]:
        w, t = np.unravel_index(flat, dev.shape)
        ref = pref0 * oracle.lir_step(chain[w, t], cfg["z"], 8.0, 1000.0, False, False)
        assert abs(res.lir[w, t] - ref) <= 5e-9 * ref, (dev[w, t], res.lir[w, t], ref)
    for arr in (res.peaklambda, res.dustmass, res.lir):
        assert arr.shape == (nw, ns) and np.isfinite(arr).all() and (arr > 0).all()
    # (1) idempotence of the dedupe: exact repeats carry their predecessor's value
    same = np.all(chain[:, 1:, :] == chain[:, :-1, :], axis=2)
    assert 0.5 < same.mean() < 0.8
    for arr in (res.peaklambda, res.dustmass, res.lir):
        assert np.array_equal(arr[:, 1:][same], arr[:, :-1][same])
    # (2) linearity in fnorm: dust mass and L_IR scale with it, the peak does not
    chain2 = chain.copy()
    chain2[:, :, 4] *= 2.0
    res2 = mbb_results.from_chain(chain2[:40], wavenorm=cfg["wavenorm"], redshift=cfg["z"],
                                  lumdist=cfg["lumdist"], device=0)
    res2.compute_peaklambda()
    res2.compute_dustmass(kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
    res2.compute_lir(*cfg["lir"], method="gauss")
    new = ~np.concatenate([np.zeros((40, 1), bool), same[:40]], axis=1)
    assert relerr(res2.dustmass[new], 2.0 * res.dustmass[:40][new]).max() < 1e-14
    assert relerr(res2.lir[new], 2.0 * lir_gauss[:40][new]).max() < 1e-13
    assert np.array_equal(res2.peaklambda[new], res.peaklambda[:40][new])
    # (3) a random sample of new steps against the oracle / the CPU emulation
    wi = rng.randint(0, nw, 60)
    ti = rng.randint(1, ns, 60)
    c = oracle.dustmass_consts(cfg["z"], cfg["wavenorm"], cfg["kappa_wave"], cfg["lumdist"])
    pref = oracle.LIR_PREFAC * cfg["lumdist"]**2
    nchk = nclose = 0
    for w, t in zip(wi, ti):
        while same[w, t - 1] and t > 1:       # walk back to the step that was actually computed
            t -= 1
        if np.allclose(chain[w, t - 1], chain[w, t]):
            continue                          # moved by < 1e-5: reference reuses the predecessor
        st = chain[w, t]
        assert abs(res.peaklambda[w, t] - oracle.peaklambda_step(st)) <= TOL * res.peaklambda[w, t]
        dm = oracle.dustmass_step(st, cfg["kappa"], cfg["wavenorm"], False, *c)
        assert abs(res.dustmass[w, t] - dm) <= TOL * dm
        lir_emu, _ = emu.lir(False, False, st, 500.0, 8.0 * 3.0, 1000.0 * 3.0, prefac=pref)
        assert abs(lir_gauss[w, t] - lir_emu[0]) <= 1e-13 * lir_emu[0]
        lir_ref = pref * oracle.lir_step(st, cfg["z"], 8.0, 1000.0, False, False)
        nchk += 1
        nclose += abs(res.lir[w, t] - lir_ref) <= TOL * lir_ref
        assert abs(res.lir[w, t] - lir_ref) <= 5e-9 * lir_ref
    assert nclose >= nchk - 1, (nclose, nchk)
