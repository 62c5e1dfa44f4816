"""CPU: results / response files in the reference's HDF5 layout (mbb_emcee/results.py:987-1158,
response.py:578-635, 782-802) -- `.npz` carrier, HDF5 carrier through an h5py stand-in, and
interchange with the executed reference's own writer / reader when /root/reference is mounted."""
import os
import sys

import numpy as np
import pytest

import fake_h5py
from mbb_emcee_b200 import mbb_results, response_set


def _results(with_responses=True, with_cov=True):
    rng = np.random.RandomState(4)
    nw, ns = 12, 30
    chain = np.array([14.0, 1.8, 400.0, 3.0, 30.0]) + rng.normal(size=(nw, ns, 5)) * [1, 0.1, 30, 0.2, 2]
    lnp = -0.5 * rng.chisquare(5, size=(nw, ns))
    res = mbb_results.from_chain(chain, lnp, wavenorm=350.0, noalpha=False, opthin=False,
                                 redshift=2.3, lumdist=1.9e4)
    res._cosmo_type = "Planck13"
    res._has_uplim = [True, False, True, False, False, True]
    res._uplim = np.array([60.0, np.inf, 1500.0, np.inf, np.inf, 400.0])
    res._has_gprior = [False, True, False, False, False, True]
    res._gprior_mean = np.array([0, 1.8, 0, 0, 0, 300.0])
    res._gprior_sigma = np.array([0, 0.3, 0, 0, 0, 60.0])
    res._gprior_ivar = np.array([1, 1 / 0.09, 1, 1, 1, 1 / 3600.0])
    res._fixed = [False, False, False, True, False]
    names = ["SPIRE_250um", "SPIRE_350um", "SCUBA2_850um", "PdBI_box_135_3.6"]
    res._ndata = len(names)
    res._data_flux = np.array([80.0, 60.0, 8.0, 1.0])
    res._data_flux_unc = np.array([6.0, 5.0, 1.0, 0.3])
    if with_responses:
        wheel = response_set()
        wheel.add_special("PdBI_box_135_3.6")
        res._response_integrate = True
        res._responsewheel = wheel
        res._data_wave = np.array([wheel[n].effective_wavelength for n in names])
    else:
        res._data_wave = np.array([250.0, 350.0, 850.0, 2200.0])
    if with_cov:
        cov = np.diag(res._data_flux_unc**2)
        cov[0, 1] = cov[1, 0] = 9.0
        res._has_covmatrix = True
        res._covmatrix = cov
        res._invcovmatrix = np.linalg.inv(cov)
    res.lir, res._has_lir, res._lir_min, res._lir_max = rng.uniform(1, 9, (nw, ns)), True, 8.0, 1000.0
    res.dustmass, res._has_dustmass, res._kappa, res._kappa_wave = rng.uniform(1, 9, (nw, ns)), True, 2.64, 125.0
    res.peaklambda, res._has_peaklambda = rng.uniform(80, 120, (nw, ns)), True
    return res


def _same(a, b):
    for name in ("_z", "_noalpha", "_opthin", "_nwalkers", "_wavenorm", "_response_integrate", "_ndata",
                 "_cosmo_type", "_has_lumdist", "_lumdist", "_has_covmatrix", "_has_lir", "_lir_min", "_lir_max",
                 "_has_dustmass", "_kappa", "_kappa_wave", "_has_peaklambda"):
        assert getattr(a, name) == getattr(b, name), name
    for name in ("_lowlim", "_has_uplim", "_uplim", "_has_gprior", "_gprior_mean", "_gprior_sigma", "_gprior_ivar",
                 "_fixed", "_data_wave", "_data_flux", "_data_flux_unc", "chain", "lnprobability",
                 "par_central_values", "lir", "dustmass", "peaklambda"):
        assert np.array_equal(np.asarray(getattr(a, name)), np.asarray(getattr(b, name))), name
    if a._has_covmatrix:
        assert np.array_equal(a._covmatrix, b._covmatrix) and np.array_equal(a._invcovmatrix, b._invcovmatrix)
        assert np.array_equal(b.covmatrix, a._covmatrix)
    assert np.array_equal(a.best_fit[0], b.best_fit[0]) and a.best_fit[1] == b.best_fit[1]
    assert tuple(a.best_fit[2]) == tuple(b.best_fit[2])
    if a._response_integrate:
        assert set(a._responsewheel.keys()) == set(b._responsewheel.keys())
        for k in a._responsewheel.keys():
            ra, rb = a._responsewheel[k], b._responsewheel[k]
            for x, y in zip(ra.node_table(), rb.node_table()):
                assert np.array_equal(x, y), k
            assert ra._normfac == rb._normfac and ra._effective_wave == rb._effective_wave
            assert ra._normtype == rb._normtype and ra._normparam == rb._normparam and ra._name == rb._name
            assert rb(lambda w: 1.0 + 0 * w) == ra(lambda w: 1.0 + 0 * w)
    assert str(a) == str(b)


@pytest.mark.parametrize("with_responses,with_cov", [(True, True), (False, False)])
def test_npz_round_trip_restores_everything(tmp_path, with_responses, with_cov):
    """save -> load leaves no state behind: covariance, ResponseIntegrate with the response
    tables themselves, cosmology choice, limits, priors, ancillary chains."""
    res = _results(with_responses, with_cov)
    written = res.save(str(tmp_path / "fit"))              # no suffix given
    assert written.endswith("fit.npz") and os.path.exists(written)
    assert not os.path.exists(written + ".npz")
    _same(res, mbb_results.load(written))
    assert res.save(str(tmp_path / "other.npz")) == str(tmp_path / "other.npz")
    _same(res, mbb_results(h5file=str(tmp_path / "other.npz")))
    back = mbb_results.load(written)
    if with_responses:
        assert back.response_integrate and "SPIRE_350um" in back._responsewheel
    else:
        assert back.covmatrix is None and not back.response_integrate


def test_response_set_file_round_trip(tmp_path):
    wheel = response_set()
    wheel.add_special("ALMA_alma_345")
    wheel.add_special("SMA_dsb_230_8_2")
    name = wheel.save(str(tmp_path / "wheel"))
    back = response_set.load(name)
    assert set(back.keys()) == set(wheel.keys())
    for k in wheel.keys():
        for x, y in zip(wheel[k].node_table(), back[k].node_table()):
            assert np.array_equal(x, y)
        assert back[k].isdelta == wheel[k].isdelta and str(back[k]) == str(wheel[k])


def test_hdf5_needs_h5py_and_says_so(tmp_path):
    fake_h5py.uninstall()
    if "h5py" in sys.modules:
        pytest.skip("a real h5py is installed")
    with pytest.raises(ImportError, match="h5py"):
        _results().save(str(tmp_path / "fit.h5"))


def test_hdf5_carrier_round_trip(tmp_path):
    """The HDF5 writer / reader, through a stand-in with h5py's API (tests/fake_h5py.py)."""
    fake_h5py.install()
    try:
        res = _results()
        name = res.writeToHDF5(str(tmp_path / "fit.h5"))
        _same(res, mbb_results.load(name))
        other = mbb_results()
        other.readFromHDF5(name)
        _same(res, other)
    finally:
        fake_h5py.uninstall()


def test_hdf5_interchange_with_the_reference(tmp_path):
    """Files written here are read by the REFERENCE's mbb_results.readFromHDF5, and files written
    by the reference's writeToHDF5 are read here -- same h5py calls on both sides (stand-in)."""
    import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("/root/reference is not mounted (build container only)")
    ref = ref_harness.import_reference()
    fake_h5py.install()
    try:
        ref_results = sys.modules["mbb_emcee.results"].mbb_results
        res = _results()
        name = res.save(str(tmp_path / "ours.h5"))
        theirs = ref_results(h5file=name)                       # the reference's reader on our file
        assert theirs._fitset and theirs.redshift == res._z and theirs._wavenorm == res._wavenorm
        assert np.array_equal(theirs.chain, res.chain) and np.array_equal(theirs.lir, res.lir)
        assert np.array_equal(theirs._covmatrix, res._covmatrix)
        assert theirs._cosmo_type == res._cosmo_type and theirs._lumdist.value == res._lumdist
        assert set(theirs._responsewheel.keys()) == set(res._responsewheel.keys())
        r0 = theirs._responsewheel["SPIRE_250um"]
        assert r0(lambda w: 1.0 + 0 * w) == res._responsewheel["SPIRE_250um"](lambda w: 1.0 + 0 * w)
        assert abs(theirs.par_cen('beta')[0] - res.par_cen('beta')[0]) < 1e-12
        back_name = str(tmp_path / "theirs.h5")
        theirs.writeToHDF5(back_name)                           # the reference's writer
        _same(res, mbb_results.load(back_name))
    finally:
        fake_h5py.uninstall()
