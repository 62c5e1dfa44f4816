"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (aconley/mbb_emcee at /root/reference) is imported through
oracle/ref_harness.py (compatibility shims + stub imports only; the two
documented one-line fixes for ALMA/GHz-delta specials).  /root/reference does
not exist on the GPU box, so its outputs travel as these small fixtures.

Files:
  golden_sed.npz       modified_blackbody: constants, f_nu (array=Cython path,
                       scalar=numpy path), wavemerge, max_wave, freq_integrate
                       for the four (thin|thick)x(alpha|noalpha) variants.
  golden_response.npz  response tables: normfac / effective wave / sums and a
                       SHA-256 of every node array for the 18 shipped filters
                       and 10 specials.
  golden_like.npz      likelihood.__call__ on cfg1/cfg2/cfg3-style setups
                        incl. covariance, soft upper limits,
                       Gaussian priors, lambda_peak terms, -inf gate.
  golden_results.npz   mbb_results.compute_peaklambda / compute_lir /
                       compute_dustmass on a small chain with repeats.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_harness  # noqa: E402

ref = ref_harness.import_reference()

# load the synthetic-config module without importing the product package
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "_synthetic", os.path.join(ROOT, "mbb_emcee_b200", "synthetic.py"))
synthetic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synthetic)

WAVES = np.array([24.0, 40.0, 60.0, 70.0, 100.0, 160.0, 250.0, 350.0, 500.0,
                  850.0, 1100.0, 2000.0, 3000.0])
VARIANTS = [("thin_noalpha", True, True), ("thin_alpha", True, False),
            ("thick_noalpha", False, True), ("thick_alpha", False, False)]


def sed_params(rng, n):
    P = np.empty((n, 5))
    P[:, 0] = rng.uniform(3.0, 80.0, n)         # T/(1+z); low T puts wavenorm on the power law
    P[:, 1] = rng.uniform(0.1, 3.5, n)          # beta
    P[:, 2] = np.exp(rng.uniform(np.log(2.0), np.log(3000.0), n))   # lambda0
    P[:, 3] = rng.uniform(0.15, 8.0, n)         # alpha
    P[:, 4] = np.exp(rng.uniform(np.log(0.01), np.log(500.0), n))   # fnorm
    # the reference's own test points (test_modified_blackbody.py)
    fixed = np.array([[10.0, 2.0, 800.0, 2.0, 45.0], [15.0, 1.8, 200.0, 3.0, 50.0],
                      [15.0, 1.8, 5.0, 3.0, 50.0], [20.0, 1.9, 250.0, 3.5, 50.0],
                      [35.0, 2.2, 250.0, 2.8, 50.0], [40.0, 1.5, 600.0, 3.0, 50.0],
                      [12.0, 1.8, 150.0, 3.0, 30.0], [14.0, 1.8, 400.0, 3.0, 30.0]])
    P[:len(fixed)] = fixed
    return P


def gen_sed():
    rng = np.random.RandomState(7001)
    n = 160
    out = {"waves": WAVES}
    for name, opthin, noalpha in VARIANTS:
        for wavenorm in (500.0, 250.0):
            tag = "%s_wn%d" % (name, int(wavenorm))
            P = sed_params(rng, n)
            normfac = np.empty(n)
            xmerge = np.full(n, np.nan)
            kappa = np.full(n, np.nan)
            x0 = np.full(n, np.nan)
            wmerge = np.full(n, np.nan)
            fa = np.empty((n, len(WAVES)))
            fs = np.empty((n, len(WAVES)))
            mw = np.empty(n)
            nint = 40
            fi = np.full(n, np.nan)
            for i in range(n):
                m = ref.modified_blackbody(P[i, 0], P[i, 1], P[i, 2], P[i, 3], P[i, 4],
                                           wavenorm=wavenorm, noalpha=noalpha, opthin=opthin)
                normfac[i] = m._normfac
                if not noalpha:
                    xmerge[i] = m._xmerge
                    kappa[i] = m._kappa
                    wmerge[i] = m.wavemerge
                if not opthin:
                    x0[i] = m._x0
                fa[i] = m(WAVES)                               # array -> Cython
                fs[i] = [m(float(w))[0] for w in WAVES]        # scalar -> numpy twin
                mw[i] = m.max_wave()
                if i < nint:
                    fi[i] = m.freq_integrate(8.0 * 3.0, 1000.0 * 3.0)
            out.update({tag + "_P": P, tag + "_normfac": normfac, tag + "_xmerge": xmerge,
                        tag + "_kappa": kappa, tag + "_x0": x0, tag + "_wavemerge": wmerge,
                        tag + "_fnu_array": fa, tag + "_fnu_scalar": fs,
                        tag + "_maxwave": mw, tag + "_freqint": fi})
    np.savez_compressed(os.path.join(HERE, "golden_sed.npz"), **out)
    print("golden_sed.npz", len(out), "arrays")


SPECIALS = ["ZSpec_box_1050um_100", "ALMA_alma_230", "ALMA_alma_345", "SMA_dsb_230_8_2",
            "PdBI_box_135_3.6", "X_gauss_300um_30", "Y_delta_500um", "X_delta_880",
            "Q_alma_100", "W_dsb_850um_50_10"]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def gen_response():
    wheel = ref.response_set()
    for sp in SPECIALS:
        wheel.add_special(sp)
    names = sorted(wheel.keys())
    rows = []
    for k in names:
        r = wheel[k]
        rows.append((k, r._nresp, r._normfac, r._effective_wave, r._effective_freq,
                     float(np.ravel(r(lambda x: 1))[0]) if not r._isdelta else 1.0,
                     _sha(r._wave), _sha(r._freq), _sha(r._resp),
                     _sha(r._sedmult) if not r._isdelta else "",
                     float(r._sedmult.sum()) if not r._isdelta else 0.0))
    out = {
        "names": np.array([r[0] for r in rows]),
        "nresp": np.array([r[1] for r in rows]),
        "normfac_raw": np.array([r[2] for r in rows]),
        "eff_wave": np.array([r[3] for r in rows]),
        "eff_freq": np.array([r[4] for r in rows]),
        "flat_response": np.array([r[5] for r in rows]),
        "sha_wave": np.array([r[6] for r in rows]),
        "sha_freq": np.array([r[7] for r in rows]),
        "sha_resp": np.array([r[8] for r in rows]),
        "sha_sedmult": np.array([r[9] for r in rows]),
        "sum_sedmult": np.array([r[10] for r in rows]),
    }
    np.savez_compressed(os.path.join(HERE, "golden_response.npz"), **out)
    print("golden_response.npz", len(names), "passbands")


def _like_for(cfgname, cfg):
    """Build the reference likelihood for a synthetic config; returns
    (like, flux, unc, cov_or_None)."""
    rng = np.random.RandomState(cfg["seed"])
    like = ref.likelihood(wavenorm=cfg["wavenorm"], noalpha=cfg["noalpha"],
                          opthin=cfg["opthin"], response=cfg["response"])
    t = cfg["truth"]
    truth_sed = ref.modified_blackbody(t[0], t[1], t[2], t[3], t[4], wavenorm=cfg["wavenorm"],
                                       noalpha=cfg["noalpha"], opthin=cfg["opthin"])
    if cfg["response"]:
        # need the responses to evaluate the truth: a throw-away set_phot does that
        like.set_phot(cfg["bands"], np.ones(len(cfg["bands"])), np.ones(len(cfg["bands"])))
        model = np.array([float(np.ravel(r(truth_sed))[0]) for r in like._responses])
    else:
        model = truth_sed(np.asarray(cfg["bands"]))
    flux, unc = synthetic.noisy_photometry(model, rng)
    like.set_phot(cfg["bands"], flux, unc)
    cov = None
    if cfgname == "cfg3":
        cov = synthetic.cfg3_covariance(flux, unc)
        like.set_cov(cov)
    for name, val in cfg.get("uplims", []):
        like.set_uplim(name, val)
    for name, mean, sig in cfg.get("gpriors", []):
        like.set_gaussian_prior(name, mean, sig)
    return like, flux, unc, cov, rng


def gen_like():
    out = {}
    n = 96
    for cfgname in ("cfg1", "cfg2", "cfg3"):
        cfg = synthetic.CONFIGS[cfgname]
        like, flux, unc, cov, rng = _like_for(cfgname, cfg)
        uplim_eff = np.where(like._has_uplim[:5], like._uplim[:5], np.inf)
        P = synthetic.walker_cloud(cfg["truth"], n, rng, like._lowlim, uplim_eff)
        # exercise the gates: below a lower limit, above soft upper limits
        P[0, 0] = 0.5                      # T below limit -> -inf
        P[1, 4] = 1e-4                     # fnorm below limit -> -inf
        P[2, 1] = 21.5                     # beta above its soft limit of 20
        P[3, 3] = 20.7                     # alpha above 20
        P[4, 2] = like._uplim[2] * 1.02    # lambda0 above 3*max(wave)
        P[5, 0] = 70.0                     # T above cfg3's limit of 60
        P[6] = cfg["truth"]
        P[7, 0] = 4.0                      # cold: wavenorm on the power-law side
        ll = np.array([like(P[i]) for i in range(n)])
        out.update({cfgname + "_flux": flux, cfgname + "_unc": unc, cfgname + "_P": P,
                    cfgname + "_lnlike": ll, cfgname + "_lowlim": like._lowlim.copy(),
                    cfgname + "_has_uplim": np.array(like._has_uplim),
                    cfgname + "_uplim": like._uplim.copy(),
                    cfgname + "_data_wave": np.asarray(like._wave, dtype=np.float64)})
        if cov is not None:
            out[cfgname + "_cov"] = cov
            out[cfgname + "_invcov"] = like._invcovmatrix
    # the survey's extra vectors (SURVEY.md 8c last row): delta bands with
    # covariance + beta prior + lambda_peak limit and prior
    like = ref.likelihood(wavenorm=500.0)
    wv = [100.0, 160.0, 250.0, 350.0, 500.0, 850.0]
    fl = np.array([20.0, 60.0, 80.0, 55.0, 30.0, 6.0])
    un = np.array([3.0, 6.0, 6.0, 5.0, 4.0, 1.5])
    like.set_phot(wv, fl, un)
    cov = np.diag(un**2)
    cov[2, 3] = cov[3, 2] = 10.0
    cov[3, 4] = cov[4, 3] = 6.0
    like.set_cov(cov)
    like.set_gaussian_prior('beta', 1.8, 0.3)
    like.set_uplim('lambda_peak', 300.0)
    like.set_gaussian_prior('lambda_peak', 250.0, 40.0)
    rng = np.random.RandomState(7003)
    P = synthetic.walker_cloud((12.0, 1.8, 150.0, 3.0, 30.0), 64, rng, like._lowlim)
    P[0] = (12.0, 1.8, 150.0, 3.0, 30.0)
    out.update({"extra_wave": np.array(wv), "extra_flux": fl, "extra_unc": un,
                "extra_cov": cov, "extra_P": P,
                "extra_lnlike": np.array([like(P[i]) for i in range(len(P))])})
    np.savez_compressed(os.path.join(HERE, "golden_like.npz"), **out)
    print("golden_like.npz", len(out), "arrays; cfg1 lnlike[6] =", out["cfg1_lnlike"][6],
          "extra[0] =", out["extra_lnlike"][0])


def gen_results():
    cfg = synthetic.CONFIGS["cfg4"]
    out = {}
    for name, opthin, noalpha in VARIANTS:
        rng = np.random.RandomState(cfg["seed"] + len(name))
        chain = synthetic.random_walk_chain(cfg["truth"], 6, 30, rng)
        # a nearly-equal (allclose but not identical) step, to pin the dedupe rule
        chain[0, 5, :] = chain[0, 4, :] * (1.0 + 3e-6)
        fit = ref.mbb_fitter(nwalkers=6, noalpha=noalpha, opthin=opthin,
                             wavenorm=cfg["wavenorm"])
        fit.like.set_phot([250.0, 350.0, 500.0], [30.0, 40.0, 30.0], [3.0, 4.0, 3.0])
        fit.sampler.chain = chain
        fit.sampler.lnprobability = np.zeros(chain.shape[:2])
        res = ref.mbb_results(fit=fit, redshift=cfg["z"], lumdist=cfg["lumdist"])
        res.compute_peaklambda()
        res.compute_lir(wavemin=cfg["lir"][0], wavemax=cfg["lir"][1])
        res.compute_dustmass(kappa=cfg["kappa"], kappa_wave=cfg["kappa_wave"])
        out.update({name + "_chain": chain, name + "_peaklambda": res.peaklambda,
                    name + "_lir": res.lir, name + "_dustmass": res.dustmass})
    np.savez_compressed(os.path.join(HERE, "golden_results.npz"), **out)
    print("golden_results.npz", len(out), "arrays")


if __name__ == "__main__":
    gen_sed()
    gen_response()
    gen_like()
    gen_results()
