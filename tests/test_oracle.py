"""CPU: pin the oracle (oracle/mbb_oracle.py) to the reference.

(a) the reference's own known-answer tests
    (mbb_emcee/tests/test_modified_blackbody.py, test_response.py);
(b) golden vectors produced by executing the unmodified reference
    (tests/golden/make_golden.py).
Same libm calls in the same order => the bar here is 0 ulp / 1e-15.
"""
import math
import os

import numpy as np
import pytest
from numpy.testing import assert_allclose

from conftest import VARIANTS, relerr


# ---- (a) the reference's own known answers --------------------------------
def test_kat_thick(oracle):                      # test_modified_blackbody.py:6-20
    s = oracle.make_sed(10.0, 2.0, 800.0, 2.0, 45.0)
    assert_allclose(oracle.sed_call(s, 500), 45.0, atol=1e-4)
    wave = np.array([250.0, 350.0, 500.0, 850.0])
    assert_allclose(oracle.sed_call(s, wave),
                    [21.96268738, 39.53249977, 45.0, 22.06274444], rtol=1e-4)


def test_kat_thin(oracle):                       # test_modified_blackbody.py:23-36
    s = oracle.make_sed(15.0, 1.8, 200.0, 3.0, 50.0, opthin=True)
    wave = np.array([250.0, 350.0, 500.0, 850.0])
    assert_allclose(oracle.sed_call(s, wave),
                    [178.34976, 111.03026, 50.0, 10.880588], rtol=1e-4)


def test_kat_thinthick(oracle):                  # test_modified_blackbody.py:39-47
    a = oracle.make_sed(15.0, 1.8, 5.0, 3.0, 50.0, opthin=True)
    b = oracle.make_sed(15.0, 1.8, 5.0, 3.0, 50.0, opthin=False)
    wave = np.array([500.0, 850.0, 1100.0, 2500.0])
    assert_allclose(oracle.sed_call(a, wave), oracle.sed_call(b, wave), rtol=1e-3)


def test_kat_merge(oracle):                      # test_modified_blackbody.py:50-69
    mk = oracle.make_sed
    assert oracle.wavemerge(mk(20.0, 1.9, None, 3.5, 50.0, noalpha=True, opthin=True)) is None
    assert_allclose(oracle.wavemerge(mk(20.0, 1.9, None, 3.5, 50.0, opthin=True)), 85.66065, rtol=1e-3)
    assert_allclose(oracle.wavemerge(mk(35.0, 2.2, None, 2.8, 50.0, opthin=True)), 51.40211, rtol=1e-3)
    assert oracle.wavemerge(mk(20.0, 1.9, 250.0, 3.5, 50.0, noalpha=True)) is None
    assert_allclose(oracle.wavemerge(mk(20.0, 1.9, 250.0, 3.5, 50.0)), 109.5506829, rtol=1e-3)
    assert_allclose(oracle.wavemerge(mk(40.0, 1.5, 600.0, 3.0, 50.0)), 60.10021595, rtol=1e-3)


# ---- (b) golden vectors from the executed reference -------------------------
@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
@pytest.mark.parametrize("wavenorm", [500.0, 250.0])
def test_sed_golden(oracle, golden, name, opthin, noalpha, wavenorm):
    g = golden.sed
    tag = "%s_wn%d" % (name, int(wavenorm))
    P = g[tag + "_P"]
    waves = g["waves"]
    for i in range(len(P)):
        s = oracle.make_sed(*P[i], wavenorm=wavenorm, noalpha=noalpha, opthin=opthin)
        assert s.normfac == g[tag + "_normfac"][i]
        if not noalpha:
            assert s.xmerge == g[tag + "_xmerge"][i]
            assert s.kappa == g[tag + "_kappa"][i]
        if not opthin:
            assert s.x0 == g[tag + "_x0"][i]
        assert np.array_equal(oracle.sed_call(s, waves), g[tag + "_fnu_array"][i])
        fs = np.array([oracle.sed_call(s, float(w))[0] for w in waves])
        assert np.array_equal(fs, g[tag + "_fnu_scalar"][i])
        if i < 24:
            assert oracle.max_wave(s) == g[tag + "_maxwave"][i]
        if i < 6:
            assert oracle.freq_integrate(s, 24.0, 3000.0) == g[tag + "_freqint"][i]


def test_c_port_matches_ref_fnu(oracle, golden):
    """oracle/fnu_port.c (plain C) == the reference's compiled fnu.pyx, bit for bit.
    Against the golden arrays always; against oracle/_ref directly when built."""
    g = golden.sed
    for name, opthin, noalpha in VARIANTS:
        tag = name + "_wn500"
        P = g[tag + "_P"]
        for i in range(0, len(P), 3):
            s = oracle.make_sed(*P[i], noalpha=noalpha, opthin=opthin)
            f = oracle.UM_TO_GHZ / g["waves"]
            port = oracle.fnu_native(s, f, impl="port")
            assert np.array_equal(port, g[tag + "_fnu_array"][i])
            if oracle.native_kind() == "reference":
                assert np.array_equal(port, oracle.fnu_native(s, f, impl="reference"))


def _spec_for(oracle, golden, cfgname):
    from mbb_emcee_b200 import synthetic
    from mbb_emcee_b200.response import response_set
    cfg = synthetic.CONFIGS[cfgname]
    g = golden.like
    spec = oracle.LikeSpec(cfg["wavenorm"], cfg["noalpha"], cfg["opthin"])
    if cfg["response"]:
        wheel = response_set()
        bands = []
        for nm in cfg["bands"]:
            if nm not in wheel:
                wheel.add_special(nm)
            bands.append(oracle.band_from_response(wheel[nm]))
        spec.set_phot(bands, g[cfgname + "_flux"], g[cfgname + "_unc"])
    else:
        spec.set_phot(cfg["bands"], g[cfgname + "_flux"], g[cfgname + "_unc"])
    spec.auto_lambda0_uplim(g[cfgname + "_data_wave"].max())
    if cfgname + "_cov" in g:
        spec.set_cov(g[cfgname + "_cov"])
    idx = {"T": 0, "beta": 1, "lambda0": 2, "alpha": 3, "fnorm": 4, "lambda_peak": 5}
    for nm, v in cfg.get("uplims", []):
        spec.set_uplim(idx[nm], v)
    for nm, m, sg in cfg.get("gpriors", []):
        spec.set_gprior(idx[nm], m, sg)
    return spec


@pytest.mark.parametrize("cfgname", ["cfg1", "cfg2", "cfg3"])
def test_like_golden(oracle, golden, cfgname):
    spec = _spec_for(oracle, golden, cfgname)
    g = golden.like
    assert np.array_equal(spec.uplim, g[cfgname + "_uplim"])
    P = g[cfgname + "_P"]
    ll = oracle.loglike_batch(spec, P)
    ref = g[cfgname + "_lnlike"]
    assert np.array_equal(np.isinf(ll), np.isinf(ref))
    assert np.isinf(ref).sum() == 2
    fin = np.isfinite(ref)
    assert relerr(ll[fin], ref[fin]).max() <= 1e-15


def test_like_extra(oracle, golden):
    g = golden.like
    spec = oracle.LikeSpec()
    spec.set_phot(g["extra_wave"], g["extra_flux"], g["extra_unc"])
    spec.auto_lambda0_uplim(g["extra_wave"].max())
    spec.set_cov(g["extra_cov"])
    spec.set_gprior(1, 1.8, 0.3)
    spec.set_uplim(5, 300.0)
    spec.set_gprior(5, 250.0, 40.0)
    ll = oracle.loglike_batch(spec, g["extra_P"])
    assert relerr(ll, g["extra_lnlike"]).max() <= 1e-15
    assert abs(ll[0] - (-30.13025033346133)) < 1e-12


@pytest.mark.parametrize("name,opthin,noalpha", VARIANTS)
def test_results_golden(oracle, golden, name, opthin, noalpha):
    from mbb_emcee_b200 import synthetic
    cfg = synthetic.CONFIGS["cfg4"]
    g = golden.results
    chain = g[name + "_chain"]
    pk = oracle.map_chain(chain, oracle.peaklambda_step)
    assert np.array_equal(pk, g[name + "_peaklambda"])
    c = oracle.dustmass_consts(cfg["z"], cfg["wavenorm"], cfg["kappa_wave"], cfg["lumdist"])
    dm = oracle.map_chain(chain, lambda st: oracle.dustmass_step(
        st, cfg["kappa"], cfg["wavenorm"], opthin, *c))
    assert relerr(dm, g[name + "_dustmass"]).max() <= 1e-15
    sub = chain[:2, :8]
    lir = oracle.LIR_PREFAC * cfg["lumdist"]**2 * oracle.map_chain(
        sub, lambda st: oracle.lir_step(st, cfg["z"], 8.0, 1000.0, opthin, noalpha))
    assert relerr(lir, g[name + "_lir"][:2, :8]).max() <= 1e-15
    # the allclose-dedupe (results.py:561): step 5 of walker 0 is within 3e-6
    # of step 4 and must reuse its value
    assert g[name + "_peaklambda"][0, 5] == g[name + "_peaklambda"][0, 4]
