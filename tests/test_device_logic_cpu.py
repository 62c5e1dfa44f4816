"""CPU: the device model source (csrc/mbb_model.cuh), compiled for the host,
against the golden vectors of the executed reference.

This checks the LOGIC the kernels are built from -- per-walker setup incl. the
Newton (thin) and Brent (thick) merge-point solves, the four f_nu variants in
both arithmetic modes, max_wave, limits, priors, chi-square with and without
covariance -- before any GPU time is spent.  Tolerance: 1e-12 relative, the
north-star bar (observed: ~1e-15)."""
import numpy as np
import pytest

import hostemu_lib as emu
from conftest import VARIANTS, relerr

TOL = 1e-12


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
@pytest.mark.parametrize("wavenorm", [500.0, 250.0])
def test_setup_constants(golden, name, opthin, noalpha, wavenorm):
    g = golden.sed
    tag = "%s_wn%d" % (name, int(wavenorm))
    out, st = emu.consts(opthin, noalpha, g[tag + "_P"], wavenorm)
    assert (st == 0).all()
    assert relerr(out[:, 0], g[tag + "_normfac"]).max() < TOL
    if not noalpha:
        assert relerr(out[:, 1], g[tag + "_xmerge"]).max() < TOL
        assert relerr(out[:, 2], g[tag + "_kappa"]).max() < TOL
    if not opthin:
        assert relerr(out[:, 3], g[tag + "_x0"]).max() == 0.0


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
def test_max_wave(golden, name, opthin, noalpha):
    g = golden.sed
    tag = name + "_wn500"
    out, st = emu.consts(opthin, noalpha, g[tag + "_P"], 500.0, want_peak=True)
    assert (st == 0).all()
    assert relerr(out[:, 5], g[tag + "_maxwave"]).max() < TOL


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
@pytest.mark.parametrize("wavenorm", [500.0, 250.0])
@pytest.mark.parametrize("fast", [0, 1, 2])
def test_fnu(golden, name, opthin, noalpha, wavenorm, fast):
    g = golden.sed
    tag = "%s_wn%d" % (name, int(wavenorm))
    freq = 299792458e-3 / g["waves"]
    arr = emu.fnu(opthin, noalpha, g[tag + "_P"], wavenorm, freq, scalar_path=False, fast=fast)
    assert relerr(arr, g[tag + "_fnu_array"]).max() < TOL
    sc = emu.fnu(opthin, noalpha, g[tag + "_P"], wavenorm, freq, scalar_path=True, fast=fast)
    assert relerr(sc, g[tag + "_fnu_scalar"]).max() < TOL


def _setup(golden, cfgname):
    from mbb_emcee_b200 import likelihood, synthetic
    cfg = synthetic.CONFIGS[cfgname]
    g = golden.like
    like = likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"], opthin=cfg["opthin"],
                      response=cfg["response"])
    like.set_phot(cfg["bands"], g[cfgname + "_flux"], g[cfgname + "_unc"])
    if cfgname + "_cov" in g:
        like.set_cov(g[cfgname + "_cov"])
    for nm, v in cfg.get("uplims", []):
        like.set_uplim(nm, v)
    for nm, m, s in cfg.get("gpriors", []):
        like.set_gaussian_prior(nm, m, s)
    return cfg, like


@pytest.fixture(params=[6, 8], ids=["tab64", "tab256"])
def tabbits(request):
    """Table flavour of the emulated specialised node code: 64 entries (the delta kernels) or
    256 entries (nodes kernel, Gauss-rule kernels; per-walker constants and L' times 4)."""
    emu.set_tab_bits(request.param)
    yield request.param
    emu.set_tab_bits(6)


def _emu_like(like, P, fast):
    off, wave, weight, scalar = like.band_tables()
    ep = emu.priors_struct(like.lowlims, like.has_uplims, like.uplims, like.has_gpriors,
                           like.gprior_means, like.gprior_ivars)
    return emu.loglike(like.opthin, like.noalpha, fast, P, like.wavenorm, ep, off, wave, weight, scalar,
                       like.data_flux, ivar=None if like.has_data_covmatrix else like._ivar,
                       cinv=like.data_invcovmatrix)


@pytest.mark.parametrize("cfgname", ["cfg1", "cfg2", "cfg3"])
@pytest.mark.parametrize("fast", [0, 1, 2])
def test_loglike(golden, cfgname, fast, tabbits):
    cfg, like = _setup(golden, cfgname)
    g = golden.like
    assert np.array_equal(like.uplims, g[cfgname + "_uplim"])
    P, ref = g[cfgname + "_P"], g[cfgname + "_lnlike"]
    ll, st = _emu_like(like, P, fast)
    assert np.array_equal(st == 1, np.isneginf(ref))
    assert np.array_equal(np.isneginf(ll), np.isneginf(ref))
    assert (st <= 1).all()
    fin = np.isfinite(ref)
    assert relerr(ll[fin], ref[fin]).max() < TOL


@pytest.mark.parametrize("fast", [0, 1, 2])
def test_loglike_extra(golden, fast, tabbits):
    from mbb_emcee_b200 import likelihood
    g = golden.like
    like = likelihood(wavenorm=500.0)
    like.set_phot(g["extra_wave"], g["extra_flux"], g["extra_unc"])
    like.set_cov(g["extra_cov"])
    like.set_gaussian_prior('beta', 1.8, 0.3)
    like.set_uplim('lambda_peak', 300.0)
    like.set_gaussian_prior('lambda_peak', 250.0, 40.0)
    ll, st = _emu_like(like, g["extra_P"], fast)
    assert (st == 0).all()
    assert relerr(ll, g["extra_lnlike"]).max() < TOL


def test_status_codes():
    ep = emu.priors_struct([0, -1, 0, -1, 0], [0] * 6, [np.inf] * 6, [0] * 6, [0] * 6, [1] * 6)
    P = np.array([[10.0, 2.0, 100.0, -1.0, 5.0],      # alpha <= 0  -> ValueError
                  [10.0, -0.5, 100.0, 2.0, 5.0],      # beta < 0    -> ValueError
                  [np.nan, 2.0, 100.0, 2.0, 5.0]])    # NaN
    off = np.array([0, 1], dtype=np.int32)
    ll, st = emu.loglike(False, False, False, P, 500.0, ep, off, [250.0], [1.0], [0], [3.0], ivar=[1.0])
    assert list(st) == [2, 3, 9]
    assert np.isnan(ll).all()


def _mp_truth(oracle, s, minwave, maxwave):
    """40-digit integral of the reference's f_nu, split at the merge point."""
    import mpmath as mp
    mp.mp.dps = 40
    hok = mp.mpf(oracle.H) / (mp.mpf(oracle.K) * mp.mpf(s.T)) * mp.mpf(1e9)
    nf = mp.mpf(s.normfac)

    def grey(x):
        if s.opthin:
            return x**(3 + mp.mpf(s.beta)) / mp.expm1(x)
        return -mp.expm1(-(x / mp.mpf(s.x0))**mp.mpf(s.beta)) * x**3 / mp.expm1(x)

    def f(nu):
        x = hok * nu
        if (not s.noalpha) and x > mp.mpf(s.xmerge):
            return nf * mp.mpf(s.kappa) * x**(-mp.mpf(s.alpha))
        return nf * grey(x)

    f1, f2 = mp.mpf(oracle.UM_TO_GHZ) / maxwave, mp.mpf(oracle.UM_TO_GHZ) / minwave
    pts = [f1, f2]
    if not s.noalpha:
        fm = mp.mpf(s.xmerge) / hok
        if f1 < fm < f2:
            pts = [f1, fm, f2]
    grid = []
    for a, b in zip(pts[:-1], pts[1:]):
        grid += [a + (b - a) * mp.mpf(i) / 8 for i in range(8)]
    grid.append(pts[-1])
    return mp.quad(f, grid) * mp.mpf(10)**-17


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
def test_freq_integrate(golden, oracle, name, opthin, noalpha):
    """Fixed-rule quadrature vs the reference's adaptive quad (2e-8: the
    reference itself is only that good with the merge kink inside the range)
    and vs a 40-digit truth (1e-13)."""
    g = golden.sed
    tag = name + "_wn500"
    ref = g[tag + "_freqint"]
    m = np.isfinite(ref)
    P = g[tag + "_P"][m]
    got, st = emu.lir(opthin, noalpha, P, 500.0, 24.0, 3000.0)
    assert (st == 0).all()
    assert relerr(got, ref[m]).max() < 2e-8      # quad epsrel = 1.49e-8
    for i in range(0, len(P), 5):
        s = oracle.make_sed(*P[i], noalpha=noalpha, opthin=opthin)
        truth = float(_mp_truth(oracle, s, 24.0, 3000.0))
        assert abs(got[i] - truth) <= 1e-13 * abs(truth)


def test_fastmath(tabbits):
    """The lean exp family of csrc/mbb_fastmath.cuh against mpmath, in ulps: saturating
    (CLAMP) and unclamped instantiations, the scaled 1-exp(-t) path and the product
    reduction the node loops use."""
    import mpmath as mp
    mp.mp.dps = 40
    rng = np.random.RandomState(3)
    x = np.concatenate([rng.uniform(-60, 60, 3000), rng.uniform(-1, 1, 3000),
                        rng.uniform(-690, 690, 500), [0.0, 1e-300, -1e-300, 1e-17, 0.34657, -0.34658,
                                                      0.0054, -0.0054, 0.0108, -0.0109]])
    N = 1 << emu.tab_bits()                      # 64 or 256 (modes 4-7 follow the flavour)
    c64 = mp.mpf(N) / mp.log(2)
    band = 2.1e-14 * N / 64

    def worst_ulps(mode, xs, fn):
        got = emu.fastmath(mode, xs)
        worst = 0.0
        for xi, gi in zip(xs, got):
            t = fn(mp.mpf(float(xi)))
            if t == 0:
                assert gi == 0.0
                continue
            ulp = abs(float(t)) * 2.0**-52
            worst = max(worst, abs(float(mp.mpf(float(gi)) - t)) / ulp)
        return worst

    # exp <= 1.5 ulp.  expm1 <= 3 ulp for |x| < ln2/2N (the result is f*g(f) with f = x*N/ln2
    # and g(0) = ln2/N) and for |x| > 0.35 (m != 0); in between the rounding of the table
    # entry T_j shows: relative error <= 2^-53 T_j/|T_j - 1| <= 2.1e-14 for N = 64 (documented
    # in the header), 4x that for N = 256
    small = np.abs(x) < 0.0054 * 64 / N
    large = np.abs(x) > 0.35
    for mode, fn in ((0, mp.exp), (4, mp.exp)):
        w = worst_ulps(mode, x, fn)
        assert w < 1.5, (mode, w)
    for mode in (1, 5):
        # (N = 256: the degree-3 g is good to 4.8e-18 absolute, i.e. 3.5e-15 of its own value)
        assert worst_ulps(mode, x[small], mp.expm1) < (3.0 if N == 64 else 20.0)
        assert worst_ulps(mode, x[large], mp.expm1) < 3.0
        mid = np.concatenate([x[~small & ~large], rng.uniform(-0.35, 0.35, 4000)])
        assert worst_ulps(mode, mid, mp.expm1) * 2.0**-52 < band
    # 1 - exp(-t) with t scaled by the double 64/ln2 inside: the argument is t*C_hi/C exactly
    t = np.concatenate([rng.uniform(0, 40, 2000), 10.0**rng.uniform(-12, 0, 2000), [0.0, 699.0, 5000.0, 1e300]])
    chi = mp.mpf(emu.c64_hi())
    w = worst_ulps(6, t, lambda v: -mp.expm1(-min(v, mp.mpf(700)) * chi / c64))
    assert w * 2.0**-52 < band, w
    # exp(x * 0.7) through the product reduction
    w = worst_ulps(7, rng.uniform(-900, 900, 3000), lambda v: mp.exp(v * mp.mpf("0.7")))
    assert w < 1.5, w
    # saturation instead of garbage outside the double range (CLAMP instantiations)
    big = emu.fastmath(0, np.array([800.0, 1e6, -800.0, -1e6]))
    assert big[0] > 1e300 and big[1] > 1e300 and 0.0 <= big[2] < 1e-300 and 0.0 <= big[3] < 1e-300
    em = emu.fastmath(1, np.array([800.0, -800.0, -50.0]))
    assert em[0] > 1e300 and em[1] == -1.0 and em[2] == -1.0
    # reciprocal and division helpers
    b = np.concatenate([rng.uniform(0.5, 2.0, 2000), 10.0**rng.uniform(-200, 200, 2000)])
    assert np.max(np.abs(emu.fastmath(2, b) * b - 1.0)) < 4e-16
    assert np.max(np.abs(emu.fastmath(8, b) * b - 1.0)) < 4e-16
    # lean log: absolute error relative to max(1, |log x|), incl. arguments next to 1 and next to
    # powers of two (where the table term and e*ln2 cancel)
    xs = np.concatenate([10.0**rng.uniform(-300, 300, 3000), rng.uniform(0.5, 2.0, 3000),
                         1.0 + rng.uniform(-1e-3, 1e-3, 1000), 2.0**rng.randint(-50, 50, 500) * (1 - 1e-9)])
    got = emu.fastmath(9, xs)
    err = max(abs((mp.mpf(float(g)) - mp.log(mp.mpf(float(x)))) / max(1, abs(mp.log(mp.mpf(float(x))))))
              for g, x in zip(got, xs))
    assert err < 2.5e-16, err       # ~1 ulp of a result in [1, 2)


@pytest.mark.parametrize("opthin", [True, False])
def test_grouped_nodes_bit_identical(opthin):
    """The breadth-first multi-node form used by the delta kernel performs the same
    operations per node as node_acc: bit-identical values; and the `safe` flag is set for
    ordinary walkers and cleared when an exponent could leave the double range."""
    from mbb_emcee_b200 import synthetic
    rng = np.random.RandomState(11)
    P = synthetic.walker_cloud([14.0, 1.8, 400.0, 3.0, 30.0], 2000, rng, [1, 0.1, 1, 0.1, 1e-3])
    P = np.vstack([P, [[0.05, 1.8, 400.0, 3.0, 30.0], [14.0, 290.0, 400.0, 3.0, 30.0]]])
    wave = [20.0, 100.0, 160.0, 250.0, 350.0, 500.0]
    a, b, safe = emu.grey_nodes(opthin, P, 500.0, wave, [1.0, 0.9, 1.1, 1.0, 1.2, 0.8])
    assert safe[:2000].all() and not safe[2000:].any()
    assert np.array_equal(a[safe], b[safe])
    assert np.isfinite(a[safe]).all() and (a[safe] > 0).all()


def test_gauss_rule_of_passbands():
    """csrc/mbb_gaussrule.h: the 32-point Gauss rule of a passband's discrete measure
    reproduces the moments of the full table (exactness degree 63) and Planck-like sums
    to rounding; measures it cannot handle are refused."""
    import mpmath as mp
    from mbb_emcee_b200.response import response_set
    mp.mp.dps = 40
    wheel = response_set()
    for name in ("PACS_100um", "SPIRE_350um", "SCUBA2_850um"):
        wv, wt, _ = wheel[name].node_table()
        nu = 299792.458 / np.asarray(wv)
        wt = np.asarray(wt)
        xs, ws = emu.gauss_rule(nu, wt, 32)
        assert (ws > 0).all() and xs.min() >= nu.min() and xs.max() <= nu.max()
        mid, half = (nu.max() + nu.min()) / 2, (nu.max() - nu.min()) / 2
        mass = float(wt.sum())
        for j in (0, 1, 2, 17, 40, 63):
            a = sum(mp.mpf(float(w)) * ((mp.mpf(float(x)) - mid) / half)**j for x, w in zip(nu, wt))
            b = sum(mp.mpf(float(w)) * ((mp.mpf(float(x)) - mid) / half)**j for x, w in zip(xs, ws))
            assert abs(float(a - b)) <= 2e-16 * mass, (name, j)
        for T in (6.0, 14.0, 40.0):
            f = lambda v: v**4.8 / np.expm1(0.0479924 * v / T)
            assert abs(np.sum(ws * f(xs)) / np.sum(wt * f(nu)) - 1.0) < 1e-15
    assert emu.gauss_rule(np.linspace(1, 2, 100), -np.ones(100), 32) is None          # not positive
    assert emu.gauss_rule(np.linspace(1, 2, 40), np.ones(40), 32) is None             # too few points


@pytest.mark.parametrize("cfgname", ["cfg2", "cfg3"])
@pytest.mark.parametrize("opthin,noalpha", [(False, False), (True, True), (False, True), (True, False)])
def test_gauss_mode_matches_full_tables(cfgname, opthin, noalpha, tabbits):
    """MBB_MATH_FAST_GAUSS (emulated): over a very wide parameter cloud the compressed
    rules change lnlike by < 1e-13 relative, most (walker, band) pairs use them, and the
    gate sends the delicate ones (merge point in the band, cold or steep walkers) to the
    full table."""
    from mbb_emcee_b200 import likelihood, synthetic
    cfg = synthetic.CONFIGS[cfgname]
    rng = np.random.RandomState(5)
    like = likelihood(wavenorm=500.0, noalpha=noalpha, opthin=opthin, response=True)
    nb = len(cfg["bands"])
    like.set_phot(cfg["bands"], np.full(nb, 1e3), np.full(nb, 1.0))
    n = 1500
    P = np.stack([10**rng.uniform(np.log10(3), np.log10(80), n), rng.uniform(0.1, 9, n),
                  10**rng.uniform(1, 3.17, n), rng.uniform(0.5, 10, n), 10**rng.uniform(0, 2.5, n)], axis=1)
    P[0] = [13.74, 21.5, 419.25, 3.49, 23.06]        # steep beta: 1 - exp(-t) switches inside a band
    P[1] = [2.0, 1.8, 400.0, 3.0, 30.0]              # cold: x-range over PACS_100 ~ 50
    full, st2 = _emu_like(like, P, 2)
    comp, st3 = _emu_like(like, P, 3)
    ncomp = emu.last_compressed()
    assert np.array_equal(st2, st3)
    ok = np.isfinite(full)
    assert ok.sum() > 0.9 * n
    assert relerr(comp[ok], full[ok]).max() < 1e-13
    frac = ncomp / (ok.sum() * nb)
    assert frac > (0.7 if cfgname == "cfg2" else 0.4), frac
    if cfgname == "cfg2" and not opthin:
        # the two delicate walkers alone: most of their bands must go through the full table
        _emu_like(like, P[:2], 3)
        assert emu.last_compressed() < 2 * nb - 3


@pytest.mark.parametrize("opthin,noalpha", [(False, False), (True, True), (False, True), (True, False)])
def test_gauss_mode_every_shipped_filter(opthin, noalpha, tabbits):
    """MBB_MATH_FAST_GAUSS (emulated), one band at a time over the whole filter wheel:
    with the data flux at zero lnlike = -m^2/2, so the comparison measures the band flux m
    itself (no chi-square cancellation) -- the compressed rule and its gate hold for the
    narrow (SCUBA2, dln nu = 0.17) and the wide (MIPS/PACS, 0.85) filters alike, for a
    normalisation wavelength inside, blueward and redward of the bands."""
    from mbb_emcee_b200 import likelihood
    from mbb_emcee_b200.response import response_set
    rng = np.random.RandomState(21)
    n = 400
    P = np.stack([10**rng.uniform(np.log10(3), np.log10(80), n), rng.uniform(0.1, 9, n),
                  10**rng.uniform(1, 3.17, n), rng.uniform(0.5, 10, n), 10**rng.uniform(0, 2.5, n)], axis=1)
    used = 0
    for wavenorm in (500.0, 70.0):
        for name in response_set().keys():
            like = likelihood(wavenorm=wavenorm, noalpha=noalpha, opthin=opthin, response=True)
            like.set_phot([name], [0.0], [1.0])
            full, st2 = _emu_like(like, P, 2)
            comp, st3 = _emu_like(like, P, 3)
            used += emu.last_compressed()
            assert np.array_equal(st2, st3)
            ok = np.isfinite(full) & (full != 0)
            assert ok.sum() > 0.8 * n
            assert relerr(comp[ok], full[ok]).max() < 5e-14, (name, wavenorm)
    assert used > 0.5 * 2 * 17 * n          # GISMO_2mm (53 nodes) has no rule; the rest mostly use theirs


@pytest.mark.parametrize("opthin,noalpha", [(False, False), (True, True), (True, False), (False, True)])
def test_fast_paths_agree_with_faithful_over_extreme_parameters(opthin, noalpha, tabbits):
    """Log-uniform parameters over the whole range the limits allow (T 1-200 K, beta and alpha
    0.1-20, lambda0 1-3000 um, fnorm 1e-3-1e4 mJy), bands from 12 um to 3 mm: the saturating
    FAST code and the unclamped specialisation (either table flavour) return the FAITHFUL
    status for every walker and the FAITHFUL value to 1e-13."""
    from mbb_emcee_b200 import likelihood
    rng = np.random.RandomState(77)
    n = 1500
    P = np.stack([10**rng.uniform(0.0, 2.3, n), 10**rng.uniform(-1, 1.3, n), 10**rng.uniform(0, 3.47, n),
                  10**rng.uniform(-1, 1.3, n), 10**rng.uniform(-3, 4, n)], axis=1)
    waves = [12.0, 24.0, 70.0, 160.0, 350.0, 850.0, 2000.0, 3000.0]
    flux = rng.uniform(1, 100, len(waves))
    like = likelihood(wavenorm=500.0, noalpha=noalpha, opthin=opthin)
    like.set_phot(waves, flux, np.maximum(0.1 * flux, 1.0))
    ref, st_ref = _emu_like(like, P, 0)
    assert (st_ref == 0).all() and np.isfinite(ref).all()
    for fast in (1, 2):
        got, st = _emu_like(like, P, fast)
        assert np.array_equal(st, st_ref)
        assert relerr(got, ref).max() < 1e-13, fast


def test_walkers_per_source_division():
    """The multiply-shift division the kernels use for evaluation index -> source index is
    exact for every 32-bit dividend, for small, large, power-of-two and odd divisors."""
    rng = np.random.RandomState(4)
    g = np.concatenate([rng.randint(0, 2**32, 20000, dtype=np.uint64), [0, 1, 2**32 - 1, 2**31, 2**31 - 1],
                        np.arange(0, 5000)]).astype(np.uint32)
    for wps in (1, 2, 3, 7, 37, 125, 250, 256, 512, 1000, 4096, 65537, 2**31 - 1, 2**31, 2**32 - 1):
        assert np.array_equal(emu.wps_division(g, wps), (g.astype(np.uint64) // wps).astype(np.uint32)), wps


@pytest.mark.parametrize("name,opthin,noalpha", [v for v in VARIANTS if not v[2]])
def test_fast_merge_point_matches_reference(golden, name, opthin, noalpha):
    """FAST setup: the merge point from the safeguarded Newton iteration in log x (thick) or on
    the Lambert-W equation (thin) equals the reference's brentq / lambertw value far inside
    brentq's own 2e-12 tolerance; extreme slopes converge too."""
    g = golden.sed
    tag = name + "_wn500"
    out, st = emu.fast_setup(opthin, noalpha, g[tag + "_P"], 500.0)
    assert (st == 0).all()
    assert relerr(out[:, 0], g[tag + "_xmerge"]).max() < 5e-13      # brentq itself stops at ~1e-13
    rng = np.random.RandomState(12)
    n = 20000
    P = np.stack([10**rng.uniform(0.2, 2.3, n), rng.uniform(0.1, 19.9, n), 10**rng.uniform(0.1, 3.17, n),
                  rng.uniform(0.11, 19.9, n), 10**rng.uniform(-2, 3, n)], axis=1)
    out, st = emu.fast_setup(opthin, noalpha, P, 500.0)
    assert (st == 0).all() and np.isfinite(out[:, :3]).all()
    ref, st0 = emu.consts(opthin, noalpha, P, 500.0, want_peak=False)      # FAITHFUL: scipy-identical Brent
    assert (st0 == 0).all()
    assert relerr(out[:, 0], ref[:, 1]).max() < 5e-13


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10, for the C++ generator
    the device sampler uses and for the numpy twin the replay tests use."""
    import philox_np
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd))]
    for ctr, key, want in kat:
        out, u = emu.philox(ctr, key)
        assert tuple(int(x) for x in out) == want
        got = philox_np.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want
        assert u == philox_np.u53(got[0], got[1])[0] and 0.0 < u < 1.0
    # the bit-assembled uniforms of the device equal the (m + 0.5) 2^-k form of the numpy twin
    rng = np.random.default_rng(5)
    words = rng.integers(0, 2**32, size=(2000, 3), dtype=np.uint64)
    words[:4] = [[0, 0, 0], [2**32 - 1] * 3, [1, 0xFFF, 1], [2**32 - 1, 0, 0]]
    for w0, w1, w3 in words:
        uz, ua = emu.philox_uniforms(w0, w1, w3)
        assert uz == philox_np.u52w(np.uint64(w0), np.uint64(w1) >> np.uint64(12))
        assert ua == philox_np.u43(np.uint64(w3), np.uint64(w1) & np.uint64(0x7FF))
        assert 0.0 < uz < 1.0 and 0.0 < ua < 1.0


def test_qags_matches_scipy(golden):
    """The QUADPACK replay (csrc/mbb_quadpack.cuh) against scipy.integrate.quad:
    identical result / error estimate / neval on integrands that exercise the
    bisection order and the epsilon-algorithm extrapolation, and the
    reference's freq_integrate values to 1e-15."""
    import warnings
    from scipy.integrate import quad
    fs = [np.sqrt, np.log, lambda x: 1 / np.sqrt(x), lambda x: np.sqrt(abs(x - 0.3)),
          lambda x: np.cos(50 * x), lambda x: np.exp(-x * x), lambda x: abs(x - 1 / 3.),
          lambda x: 1 / (1 + 1000 * (x - .5)**2)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for kind, f in enumerate(fs):
            for a, b in ((0.0, 1.0), (0.0, 2.5), (1e-3, 7.0)):
                res, err, info = quad(f, a, b, full_output=1)[:3]
                got = emu.qags_test(kind, a, b)
                assert abs(got[0] - res) <= 4e-16 * abs(res), (kind, a, b)
                assert abs(got[1] - err) <= 1e-3 * abs(err) + 1e-18
                assert got[2] == info["neval"]
    g = golden.sed
    for name, opthin, noalpha in VARIANTS:
        for wn in (500.0, 250.0):
            tag = "%s_wn%d" % (name, int(wn))
            ref = g[tag + "_freqint"]
            m = np.isfinite(ref)
            got, nev, st = emu.qags(opthin, noalpha, g[tag + "_P"][m], wn, 24.0, 3000.0)
            assert (st == 0).all()
            assert relerr(got, ref[m]).max() < 1e-15


def test_cholesky_quadratic_form_vs_exact():
    """chi-square of a full covariance: the device's two forms (explicit inverse like the
    reference's likelihood.py:356/823; forward substitution with the Cholesky factor staged by
    mbb_set_data_chol) against the exact value from 50-digit arithmetic, on cfg3's 8x8
    covariance and on an ill-conditioned one (cond ~1e9): either form is good to a small multiple
    of cond(C) * 2^-53 (the factorisation / inversion of the host carries that much), far inside
    1e-12 for covariances like cfg3's."""
    import mpmath as mp
    from mbb_emcee_b200 import synthetic
    mp.mp.dps = 50
    rng = np.random.RandomState(5)
    covs = [np.asarray(synthetic.sample_problem("cfg3", 1)[3], dtype=np.float64)]
    a = rng.normal(size=(8, 8))
    covs.append(a @ np.diag(10.0 ** np.linspace(0, 9, 8)) @ a.T)
    for cov in covs:
        cond = np.linalg.cond(cov)
        nb = cov.shape[0]
        L = np.linalg.cholesky(cov)
        staged = np.tril(L, -1) + np.diag(1.0 / np.diag(L))
        cinv = np.linalg.inv(cov)
        Cm = mp.matrix(cov.tolist())
        worst_ch = worst_inv = 0.0
        for _ in range(20):
            diff = rng.normal(size=nb) * np.sqrt(np.diag(cov))
            exact = (mp.matrix(diff.tolist()).T * mp.lu_solve(Cm, mp.matrix(diff.tolist())))[0]
            ch = emu.quad_form(staged, diff, True)
            iv = emu.quad_form(cinv, diff, False)
            worst_ch = max(worst_ch, float(abs(ch - exact) / exact))
            worst_inv = max(worst_inv, float(abs(iv - exact) / exact))
        assert worst_ch < 2e-16 * max(cond, 8.0), (cond, worst_ch)
        assert worst_inv < 2e-16 * max(cond, 8.0), (cond, worst_inv)
        if cond < 1e4:
            assert worst_ch < 1e-13 and worst_inv < 1e-13


def test_segmented_dedupe_equals_the_sequential_rule():
    """csrc/mbb_kernels.cuh chain_dedupe_kernel cuts the sequential np.allclose rule of
    results._map_chain (reference results.py:553-566) into independent segments at steps that
    differ from their predecessor by more than both tolerances.  The same algorithm in numpy,
    against the sequential rule, on walks built around the tolerance (steps of 0 - 3 tolerances,
    single components, drifts, repeats, sign changes, values near zero where the absolute term
    of the tolerance dominates)."""
    rng = np.random.RandomState(77)

    def tol(x):
        return 1e-8 + 1e-5 * np.abs(x)

    def sequential(ch):
        owner = np.empty(len(ch), dtype=int)
        prev, pt = None, 0
        for t, c in enumerate(ch):
            if prev is None or not np.all(np.abs(prev - c) <= tol(c)):
                prev, pt = c, t
            owner[t] = pt
        return owner

    def segmented(ch):
        n = len(ch)
        hard = np.ones(n, dtype=bool)
        d = np.abs(ch[:-1] - ch[1:])
        hard[1:] = np.any(d > 1.001 * (tol(ch[1:]) + tol(ch[:-1])), axis=1)
        owner = np.full(n, -1)
        for t in np.nonzero(hard)[0]:                  # every segment start independently
            prev, pt = ch[t], t
            owner[t] = t
            j = t + 1
            while j < n and not hard[j]:
                if not np.all(np.abs(prev - ch[j]) <= tol(ch[j])):
                    prev, pt = ch[j], j
                owner[j] = pt
                j += 1
        return owner

    nbad = 0
    for trial in range(60):
        n = 400
        scale = rng.choice([1e-9, 1e-6, 1e-3, 1.0, 30.0, 1e4], size=5)      # incl. |x| << atol / rtol
        cur = scale * (1.0 + rng.uniform(-0.5, 0.5, 5))
        ch = np.empty((n, 5))
        for t in range(n):
            kind = rng.randint(0, 7)
            step = np.zeros(5)
            if kind == 1:
                step = tol(cur) * rng.uniform(0, 3, 5) * rng.choice([-1, 1], 5)
            elif kind == 2:
                k = rng.randint(0, 5)
                step[k] = tol(cur)[k] * rng.choice([0.98, 0.999, 1.001, 1.02, 1.98, 2.0, 2.02])
            elif kind == 3:
                step = cur * 4e-6
            elif kind == 4:
                step = scale * rng.standard_normal(5) * 0.1
            elif kind == 5:
                step = -2.0 * cur * (rng.uniform(size=5) < 0.3)               # sign flips
            cur = cur + step
            ch[t] = cur
        a, b = sequential(ch), segmented(ch)
        nbad += int(not np.array_equal(a, b))
        assert (b >= 0).all()
    assert nbad == 0
