"""Command line front end: ``python -m mbb_emcee_b200.run_mbb_emcee photfile outfile [...]``.

Accepts the flags of the reference's ``run_mbb_emcee.py`` (reference
mbb_emcee/run_mbb_emcee.py:66-185) with the same defaults and the same order of
operations (:187-339): build the fitter, fix / limit / prior the parameters,
draw the initial ensemble, burn in and sample, derive the requested
ancillaries, write the results.  It is thin wiring over the package API, so
every number comes from the CUDA path.

Differences forced by this image: results are written
with ``mbb_results.save`` (``.npz`` carrying the reference's HDF5 key names)
because h5py is not installed; a covariance file may be ``.npy``/``.txt`` as
well as FITS (FITS needs astropy); ``--threads`` > 1 is refused (a CUDA
context cannot be shared with forked pool workers -- one launch evaluates a
whole half-ensemble anyway); ``--math`` selects the arithmetic mode.
"""
from __future__ import print_function

import argparse
import os.path
import sys

import numpy as np

# (parameter, flag stem): the order is the parameter order T, beta, lambda0, alpha, fnorm
_PARAMS = (("T", "T"), ("beta", "Beta"), ("lambda0", "Lambda0"), ("alpha", "Alpha"), ("fnorm", "Fnorm"))
_INIT = {"T": 10.0, "beta": 2.0, "lambda0": 2500.0, "alpha": 4.0, "fnorm": 40.0}
_INIT_SIGMA = (2.0, 0.2, 100.0, 0.3, 5.0)


def build_parser():
    p = argparse.ArgumentParser(
        prog="run_mbb_emcee",
        description="Fit a modified blackbody to user provided photometry with an MCMC "
                    "(observer frame; wavelengths in um, fluxes in mJy).",
        epilog="Parameters, in order: T [K] (= T_rest/(1+z)), beta, lambda0 [um] (observer frame), "
               "alpha (blue-side power law), fnorm [mJy] at --wavenorm.  alpha and beta always have "
               "soft upper limits of 20, lambda0 of 3x the longest wavelength; 'lambda_peak' is a "
               "ghost parameter that can carry an upper limit and a prior.")
    a = p.add_argument
    a("photfile", help="text file: wavelength [um] (or passband name with --response), flux, error [mJy]")
    a("outfile", help="file to write the results to: .h5/.hdf5 (the reference's HDF5 file; needs h5py) or .npz")
    a("-b", "--burn", type=int, default=50, help="burn-in steps (def: 50)")
    a("-c", "--covfile", default=None, help="covariance matrix file [mJy^2]: FITS (needs astropy), .npy or text")
    a("-C", "--cosmotype", default="WMAP9", help="astropy.cosmology name, used only without --lumdist")
    a("-e", "--covextn", type=int, default=0, help="FITS extension of the covariance matrix (def: 0)")
    for name, stem in _PARAMS:
        a("--fix" + stem, action="store_true", default=None, help="fix %s to its initial value" % name)
        a("--init" + stem, type=float, default=_INIT[name], help="initial %s (def: %g)" % (name, _INIT[name]))
        a("--low" + stem, type=float, default=None, help="lower limit on %s" % name)
        a("--up" + stem, type=float, default=None, help="soft upper limit on %s" % name)
        a("--prior" + stem, nargs=2, type=float, default=None, metavar=("MEAN", "SIGMA"),
          help="Gaussian prior on %s" % name)
    a("--upLambdaPeak", type=float, default=None, help="soft upper limit on the SED peak wavelength [um]")
    a("--priorLambdaPeak", nargs=2, type=float, default=None, metavar=("MEAN", "SIGMA"),
      help="Gaussian prior on the SED peak wavelength")
    a("--kappa", nargs=2, type=float, default=(2.64, 125.0), help="dust opacity [m^2/kg] and its wavelength [um]")
    a("--get_dustmass", action="store_true", help="dust mass [1e8 Msun]; needs the redshift")
    a("--get_lir", action="store_true", help="rest frame L_IR; needs the redshift")
    a("--get_peaklambda", action="store_true", help="observer frame SED peak wavelength")
    a("--lir_range", nargs=2, type=float, default=(8.0, 1000.0), help="rest frame L_IR range [um]")
    a("--lumdist", type=float, default=None, help="luminosity distance [Mpc] (def: from z, needs astropy)")
    a("-p", "--photdir", default=None, help="directory of photfile / covfile")
    a("--maxidx", type=int, default=None, help="(reference flag; dead code there, refused here)")
    a("-n", "--nwalkers", type=int, default=250, help="walkers (def: 250)")
    a("-N", "--nsteps", type=int, default=250, help="steps per walker (def: 250)")
    a("--noalpha", action="store_true", help="no blue-side power law")
    a("--opthin", action="store_true", help="optically thin model")
    a("-r", "--response", action="store_true", help="integrate over the passband responses")
    a("--responsefile", default=None, help="response specification file")
    a("--responsedir", default=None, help="response specification directory")
    a("-t", "--threads", type=int, default=1, help="must be 1 on the device path")
    a("--math", default="fast", choices=["fast", "faithful", "gauss"], help="device arithmetic mode")
    a("--covsolver", default="inverse", choices=["inverse", "cholesky"],
      help="chi-square of a covariance matrix: explicit inverse like the reference (def), or its Cholesky factor")
    a("--sampler", default="host", choices=["host", "device"],
      help="host: emcee's stretch move on the host, one device call per half-ensemble (def); "
           "device: the whole burn-in and main run on the GPU")
    a("--version", action="version", version="mbb_emcee_b200 " + __import__("mbb_emcee_b200").__version__)
    a("--seed", type=int, default=None, help="seed of the initial ensemble and of the sampler")
    a("-v", "--verbose", action="store_true", help="print status messages")
    a("-w", "--wavenorm", type=float, default=500.0, help="normalisation wavelength [um] (def: 500)")
    a("-z", "--redshift", type=float, default=None, help="redshift of the object")
    return p


def configure_fit(fit, args):
    """Apply the fix / limit / prior flags to a fitter (reference :222-287); the same
    parameters are skipped for --opthin (lambda0) and --noalpha (alpha)."""
    skip = set()
    if args.opthin:
        skip.add("lambda0")
    if args.noalpha:
        skip.add("alpha")
    for name, stem in _PARAMS:
        if getattr(args, "fix" + stem) or (name == "alpha" and args.noalpha):
            fit.fix_param(name)
    for name, stem in _PARAMS:
        v = getattr(args, "low" + stem)
        if v is not None and name not in skip:
            fit.set_lowlim(name, v)
    for name, stem in _PARAMS:
        v = getattr(args, "up" + stem)
        if v is not None and name not in skip:
            fit.set_uplim(name, v)
    if args.upLambdaPeak is not None:
        fit.set_uplim("lambda_peak", args.upLambdaPeak)
    for name, stem in _PARAMS:
        v = getattr(args, "prior" + stem)
        if v is not None and name not in skip:
            fit.set_gaussian_prior(name, v[0], v[1])
    if args.priorLambdaPeak is not None:
        fit.set_gaussian_prior("lambda_peak", args.priorLambdaPeak[0], args.priorLambdaPeak[1])


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.nwalkers <= 0:
        raise ValueError("Invalid (non-positive) nwalkers: %d" % args.nwalkers)
    if args.maxidx is not None:
        raise ValueError("--maxidx is dead code in the reference (results.py:549-550) and not supported")
    from . import mbb_fitter, mbb_results, _native

    def in_photdir(name):
        return name if (name is None or args.photdir is None) else os.path.join(args.photdir, name)

    if args.seed is not None:
        np.random.seed(args.seed)
    fit = mbb_fitter(nwalkers=args.nwalkers, photfile=in_photdir(args.photfile),
                     covfile=in_photdir(args.covfile), covextn=args.covextn, wavenorm=args.wavenorm,
                     noalpha=args.noalpha, opthin=args.opthin, nthreads=args.threads,
                     response=args.response, responsefile=args.responsefile, responsedir=args.responsedir,
                     sampler=args.sampler, seed=0 if args.seed is None else args.seed + 1)
    fit.like.math_mode = {"faithful": _native.MATH_FAITHFUL, "fast": _native.MATH_FAST,
                          "gauss": _native.MATH_FAST_GAUSS}[args.math]
    fit.like.cov_solver = args.covsolver
    if args.seed is not None and args.sampler == "host" and hasattr(fit.sampler, "random_state"):
        fit.sampler.random_state = np.random.RandomState(args.seed + 1).get_state()
    configure_fit(fit, args)
    p0init = np.array([getattr(args, "init" + stem) for _, stem in _PARAMS])
    p0 = fit.generate_initial_values(p0init, np.array(_INIT_SIGMA))
    fit.run(args.burn, args.nsteps, p0, verbose=args.verbose)
    res = mbb_results(fit=fit, redshift=args.redshift, lumdist=args.lumdist, cosmo_type=args.cosmotype)
    del fit
    if args.get_peaklambda:
        if args.verbose:
            print("Computing peak obs-frame wavelength")
        res.compute_peaklambda()
    if args.get_lir:
        if args.redshift is None:
            raise ValueError("Must provide redshift if computing L_IR")
        if args.verbose:
            print("Computing L_IR (%0.1f-%0.1fum)" % tuple(args.lir_range))
        res.compute_lir(wavemin=args.lir_range[0], wavemax=args.lir_range[1])
    if args.get_dustmass:
        if args.redshift is None:
            raise ValueError("Must provide redshift if computing m_dust")
        if args.verbose:
            print("Computing dust mass")
        res.compute_dustmass(kappa=args.kappa[0], kappa_wave=args.kappa[1])
    if args.verbose:
        print("Fit results:")
        print(res)
        print("Saving results to %s" % args.outfile)
    written = res.save(args.outfile)
    if args.verbose and written != args.outfile:
        print("  (written as %s)" % written)
    return res


if __name__ == "__main__":
    main()
    sys.exit(0)
