"""Instrument passbands: node tables built once on the host.

Host-side mirror of the reference's ``response`` / ``response_set``
(reference mbb_emcee/response.py:51-640, 642-840).  Table construction is
setup-time work (run once, a few hundred microseconds) and stays in numpy,
with the reference's operation order so that ``wavelength``, ``frequency``,
``response``, the trapezoid weights and ``normfac`` come out bit-for-bit the
same.  What is hot -- the weighted sum of the SED over the
nodes for every walker -- is fused into the CUDA log-likelihood kernel; the
tables built here are what gets staged into device/shared memory
(``likelihood._stage_device_tables``).

Deliberate differences from the reference:
  * ``alma`` specials work (reference response.py:477-487 breaks on
    numpy>=1.16 because it indexes with an index *array*);
  * frequency-unit deltas other than THz work (reference response.py:352-361
    forgets ``_normwave``);
  * HDF5 (de)serialisation is out of scope (h5py is not a dependency).
"""
import math
import os
import re

import numpy

from .utility import read_text_table

__all__ = ["response", "response_set"]

special_types = ["delta", "box", "gauss", "dsb", "alma"]

_C_UM_GHZ = 299792458e-3          # c in um*GHz (reference response.py:228)
_H = 6.6260693e-34                # J s  (reference response.py:44)
_K = 1.3806505e-23                # J/K  (reference response.py:45)

_WAVE_TO_UM = {"angstroms": 1e-4, "a": 1e-4, "microns": None, "um": None,
               "meters": 1e6, "m": 1e6}
_FREQ_TO_GHZ = {"hz": 1e-9, "mhz": 1e-3, "ghz": None, "thz": 1e3}

_PACKED = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                       "resources", "filterwheel.npz")


def response_bb(freq, temperature):
    """Un-normalised Planck f_nu at ``freq`` [GHz] (reference response.py:25-48)."""
    hokt = 1e9 * _H / (_K * float(temperature))
    return freq**3 / numpy.expm1(hokt * freq)


def _scaled(vals, factor):
    # the reference multiplies as ``factor * x`` and leaves x untouched when
    # the unit is already the working one (response.py:216-246)
    return vals if factor is None else factor * vals


def _dsb_edges_ghz(nodes, xtype, xunit):
    """The four sideband edges in GHz (reference response.py:408-431).

    Note the narrower unit vocabulary of the reference's dsb/alma helpers
    ('um', 'a', 'm' are not accepted here although setup() accepts them).
    """
    if xtype == "wave":
        num = {"angstroms": 2997924580.0, "microns": 299792458e-3,
               "meters": 299792458.0e-9}.get(xunit)
        if num is None:
            raise ValueError("Unrecognized wavelength unit {:s}".format(xunit))
        return num / nodes
    if xtype == "freq":
        if xunit not in _FREQ_TO_GHZ:
            raise ValueError("Unrecognized frequency unit {:s}".format(xunit))
        fac = _FREQ_TO_GHZ[xunit]
        return nodes if fac is None else nodes * fac
    raise ValueError("Unknown unit type {:s}".format(xtype))


def _two_sideband(lo0, lo1, hi0, hi1):
    """13 + 3 + 13 node layout shared by dsb and alma (response.py:434-439)."""
    xv = numpy.concatenate((numpy.linspace(lo0, lo1, 13),
                            numpy.linspace(lo1 + 0.0001, hi0 - 0.0001, 3),
                            numpy.linspace(hi0, hi1, 13)))
    rs = numpy.concatenate((numpy.ones(13), numpy.zeros(3), numpy.ones(13)))
    return xv, rs


# ALMA 2SB bands 3, 4, 6, 7, 8 (reference response.py:474-476)
_ALMA_LOW = (92.0, 125.0, 221.0, 283.0, 385.0)
_ALMA_HIGH = (108.0, 163.0, 265.0, 365.0, 500.0)
_ALMA_IF_BOTTOM = (4.0, 4.0, 6.0, 4.0, 4.0)
_ALMA_SIDEBAND = 3.75


def _as_str(v):
    if isinstance(v, bytes):
        return v.decode()
    return str(v)


class response(object):
    """Response of one instrument passband to an SED."""

    def __init__(self, name):
        self._name = str(name)
        self._data_read = False

    # ------------------------------------------------------------------ setup
    def setup(self, inputspec, xtype='wave', xunits='microns',
              senstype='energy', normtype='power', xnorm=250.0,
              normparam=-1.0, dir=None):
        """Build the node table from a text file or a special spec.

        Same arguments and conventions as reference response.py:68-134:
        ``delta_v``, ``box_c_w`` (11 nodes), ``gauss_c_fwhm`` (43 nodes,
        +-3 FWHM), ``dsb_c_w_gap`` and ``alma_c`` (29 nodes) are synthesised;
        anything else is read as a two-column text table.
        """
        if not isinstance(inputspec, str):
            raise TypeError("filename must be string-like")
        ntyp, xtyp = normtype.lower(), xtype.lower()
        xun, styp = xunits.lower(), senstype.lower()
        self._isdelta = False

        parts = inputspec.split('_')
        kind = parts[0].lower()
        if kind == "delta":
            if len(parts) < 2:
                raise ValueError("delta needs central frequency")
            self._make_delta(float(parts[1]), xtyp, xun)
            return
        if kind == "box":
            if len(parts) < 3:
                raise ValueError("box car needs 2 params in {:s}".format(inputspec))
            cent, width = float(parts[1]), float(parts[2])
            xvals = numpy.linspace(cent - 0.5 * width, cent + 0.5 * width, 11)
            resp = numpy.ones(11)
        elif kind == "gauss":
            if len(parts) < 3:
                raise ValueError("gaussian needs 2 params in {:s}".format(inputspec))
            cent, fwhm = float(parts[1]), float(parts[2])
            sig = fwhm / math.sqrt(8 * math.log(2))
            xvals = numpy.linspace(cent - 3.0 * fwhm, cent + 3.0 * fwhm, 43)
            resp = numpy.exp(-0.5 * ((xvals - cent) / sig)**2)
        elif kind == "dsb":
            if len(parts) < 4:
                raise ValueError("dsb needs 3 params in {:s}".format(inputspec))
            cent, width, gap = (float(p) for p in parts[1:4])
            edges = _dsb_edges_ghz(numpy.array([cent - width / 2, cent - gap / 2,
                                                cent + gap / 2, cent + width / 2]),
                                   xtyp, xun)
            edges.sort()
            xvals, resp = _two_sideband(edges[0], edges[1], edges[2], edges[3])
        elif kind == "alma":
            if len(parts) < 2:
                raise ValueError("alma needs 1 params in {:s}".format(inputspec))
            cen = float(_dsb_edges_ghz(numpy.array([float(parts[1])]), xtyp, xun)[0])
            band = [i for i in range(5) if _ALMA_LOW[i] <= cen <= _ALMA_HIGH[i]]
            if not band:
                raise ValueError("Unable to identify ALMA band with central "
                                 "freq {:0.1f}".format(cen))
            bot = _ALMA_IF_BOTTOM[band[0]]
            xvals, resp = _two_sideband(cen - bot - _ALMA_SIDEBAND, cen - bot,
                                        cen + bot, cen + bot + _ALMA_SIDEBAND)
        else:
            if dir is None:
                path = inputspec
            elif dir == '!package-dir!':
                self._from_packed(inputspec, xtyp, xun, styp, ntyp, xnorm, normparam)
                return
            else:
                path = os.path.join(dir, inputspec)
            rows = read_text_table(path)
            if len(rows) == 0:
                raise IOError("No data read from {:s}".format(path))
            xvals = numpy.asarray([r[0] for r in rows])
            resp = numpy.asarray([r[1] for r in rows])
        self._build(xvals, resp, xtyp, xun, styp, ntyp, normtype, senstype,
                    xnorm, normparam)

    def _from_packed(self, filename, xtyp, xun, styp, ntyp, xnorm, normparam):
        """Shipped transmission table, looked up by its original file name."""
        with numpy.load(_PACKED) as pk:
            files = [str(f) for f in pk["wheel_files"]]
            if filename not in files:
                raise IOError("No packaged response table {:s}".format(filename))
            i = files.index(filename)
            xvals = numpy.array(pk["x_%d" % i], dtype=numpy.float64)
            resp = numpy.array(pk["r_%d" % i], dtype=numpy.float64)
        self._build(xvals, resp, xtyp, xun, styp, ntyp, ntyp, styp, xnorm, normparam)

    def _build(self, xvals, resp, xtyp, xun, styp, ntyp, normtype, senstype,
               xnorm, normparam):
        """Node table -> weights and pipeline normalisation.

        Follows reference response.py:205-334 operation by operation: the
        integration is a trapezoid sum in frequency, stored as per-node
        weights ``_sedmult`` (negative: frequency descends as wavelength
        ascends), and ``_normfac`` = 1 / (same sum over the calibration SED).
        """
        if xvals.min() <= 0:
            raise ValueError("Non-positive x value encountered")
        if resp.min() < 0:
            raise ValueError("Negative response encountered")
        if xnorm <= 0:
            raise ValueError("Non-positive xnorm")

        if xtyp == 'wave':
            if xun not in _WAVE_TO_UM:
                raise ValueError("Unrecognized wavelength unit {:s}".format(xun))
            wave = _scaled(xvals, _WAVE_TO_UM[xun])
            self._normwave = _scaled(xnorm, _WAVE_TO_UM[xun])
            freq = _C_UM_GHZ / wave
            self._normfreq = _C_UM_GHZ / self._normwave
        elif xtyp == 'freq':
            if xun not in _FREQ_TO_GHZ:
                raise ValueError("Unrecognized frequency unit {:s}".format(xun))
            freq = _scaled(xvals, _FREQ_TO_GHZ[xun])
            self._normfreq = _scaled(xnorm, _FREQ_TO_GHZ[xun])
            wave = _C_UM_GHZ / freq
            self._normwave = _C_UM_GHZ / self._normfreq
        else:
            raise ValueError("Unknown x type {:s}".format(xtyp))

        order = wave.argsort()                       # ascending wavelength
        self._wave = wave[order]
        self._freq = freq[order]
        rs = numpy.array(resp[order], dtype=numpy.float64)
        n = self._nresp = len(rs)
        rs /= rs.max()
        self._resp = rs

        if styp not in ("energy", "counts"):
            raise ValueError("Unknown sensitivity type {:s}".format(senstype))
        self._sens_energy = styp == "energy"

        self._dnu = self._freq[1:n] - self._freq[0:n - 1]
        w = numpy.empty(n)
        w[0:n - 1] = 0.5 * self._dnu
        w[n - 1] = 0.5 * self._dnu[n - 2]
        w[1:n - 1] += 0.5 * self._dnu[0:n - 2]
        w *= rs
        if not self._sens_energy:
            mid = self._freq[n // 2] if n > 1 else self._freq[0]
            w *= (mid / self._freq)
        self._sedmult = w

        self._normtype = str(ntyp)
        if ntyp == "none":
            self._normparam = None
            self._normfac = -1.0
            eff = (self._freq * w).sum() / w.sum()
        else:
            if ntyp == "power":
                self._normparam = float(normparam)
                cal = (self._freq / self._normfreq)**self._normparam
            elif ntyp == "flat":
                self._normparam = None
                cal = numpy.ones(n)
            elif ntyp == "bb":
                self._normparam = float(normparam)
                if self._normparam <= 0.0:
                    raise ValueError("Invalid (non-positive) blackbody "
                                     "temperature {:f}".format(self._normparam))
                cal = response_bb(self._freq, self._normparam) / \
                    response_bb(self._normfreq, self._normparam)
            else:
                raise ValueError("Unknown normalization type {:s}".format(normtype))
            self._normfac = 1.0 / (cal * w).sum()
            eff = (self._freq * cal * w).sum() * self._normfac
        self._effective_freq = eff
        self._effective_wave = _C_UM_GHZ / eff
        self._data_read = True

    def _make_delta(self, val, xtyp, xun):
        """Single-node passband (reference response.py:336-374)."""
        if val <= 0:
            raise ValueError("Non-positive value")
        if xtyp == 'wave':
            if xun not in _WAVE_TO_UM:
                raise ValueError("Unrecognized wavelength type {:s}".format(xun))
            self._normwave = _scaled(val, _WAVE_TO_UM[xun])
            self._normfreq = _C_UM_GHZ / self._normwave
        elif xtyp == 'freq':
            if xun not in _FREQ_TO_GHZ:
                raise ValueError("Unrecognized frequency unit {:s}".format(xun))
            self._normfreq = _scaled(val, _FREQ_TO_GHZ[xun])
            self._normwave = _C_UM_GHZ / self._normfreq
        else:
            raise ValueError("Unknown x type {:s}".format(xtyp))
        self._isdelta = True
        self._effective_wave = self._normwave
        self._effective_freq = _C_UM_GHZ / self._effective_wave
        self._wave = numpy.array([self._normwave])
        self._freq = numpy.array([self._effective_freq])
        self._resp = numpy.array([1.0])
        self._nresp = 1
        self._normtype = "delta"
        self._normparam = None
        self._normfac = 1.0
        self._sens_energy = True
        self._data_read = True

    # ------------------------------------------------------------- properties
    @property
    def name(self):
        return self._name

    @property
    def data_read(self):
        return self._data_read

    @property
    def isdelta(self):
        return self._data_read and self._isdelta

    @property
    def wavelength(self):
        """Node wavelengths in microns (ascending)."""
        return self._wave if self._data_read else None

    @property
    def frequency(self):
        """Node frequencies in GHz."""
        return self._freq if self._data_read else None

    @property
    def response(self):
        """Peak-normalised transmission at the nodes."""
        return self._resp if self._data_read else None

    @property
    def effective_wavelength(self):
        return self._effective_wave if self._data_read else None

    @property
    def effective_frequency(self):
        return self._effective_freq if self._data_read else None

    @property
    def normfac(self):
        """Pipeline normalisation, sign flipped to be positive
        (reference response.py:536-542)."""
        return -1.0 * self._normfac if self._data_read else None

    # ----------------------------------------------------------- device tables
    def node_table(self):
        """(wave_um[n], weight[n], is_delta) as the CUDA kernels consume them.

        ``weight = sedmult * normfac`` (both negative -> positive), so that
        band flux = sum_i f_nu(wave_i) * weight_i.  For a delta band the weight
        is 1 and the single node is ``_normwave``.
        """
        if not self._data_read:
            raise Exception("Data not read yet")
        if self._isdelta:
            return numpy.array([self._normwave]), numpy.array([1.0]), True
        return self._wave, self._sedmult * self._normfac, False

    # ------------------------------------------------------------------- call
    def __call__(self, fnufunc, freq=False):
        """Band flux of an SED given as a callable (reference response.py:544-576).

        This is the single-SED convenience path; ensembles of SEDs go through
        ``likelihood.__call__`` where the same sum is fused into the kernel.
        """
        if not self._data_read:
            raise Exception("Data not read yet, can't get response")
        if self._isdelta:
            return fnufunc(self._normfreq if freq else self._normwave)
        x = self._freq if freq else self._wave
        return (fnufunc(x) * self._sedmult).sum() * self._normfac

    # ---------------------------------------------------------- serialisation
    _ATTRS = (("Name", "_name"), ("IsDelta", "_isdelta"), ("NormWave", "_normwave"),
              ("NormFreq", "_normfreq"), ("NormType", "_normtype"), ("NormFac", "_normfac"),
              ("SensEnergy", "_sens_energy"), ("EffectiveFreq", "_effective_freq"),
              ("EffectiveWave", "_effective_wave"), ("NResp", "_nresp"))
    _DATA = (("Wave", "_wave"), ("Freq", "_freq"), ("Response", "_resp"))

    def to_tree(self):
        """Attributes and datasets under the reference's HDF5 names (reference
        response.py:578-606); the carrier is chosen by treeio."""
        from .treeio import new_tree
        if not self._data_read:
            raise ValueError("Data must be read to write the response")
        t = new_tree()
        for key, attr in self._ATTRS:
            t["attrs"][key] = getattr(self, attr)
        t["attrs"]["WaveUnits"] = "microns"
        t["attrs"]["FreqUnits"] = "GHz"
        if self._normparam is not None:
            t["attrs"]["NormParam"] = self._normparam
        for key, attr in self._DATA:
            t["data"][key] = numpy.asarray(getattr(self, attr))
        if not self._isdelta:
            t["data"]["Dnu"] = numpy.asarray(self._dnu)
            t["data"]["Sedmult"] = numpy.asarray(self._sedmult)
        return t

    def from_tree(self, t):
        """Inverse of to_tree (reference response.py:608-635): tables are taken
        as stored, nothing is recomputed."""
        a, d = t["attrs"], t["data"]
        self._name = _as_str(a["Name"])
        self._isdelta = bool(a["IsDelta"])
        self._normwave = float(a["NormWave"])
        self._normfreq = float(a["NormFreq"])
        self._normtype = _as_str(a["NormType"])
        self._normparam = a["NormParam"] if "NormParam" in a else None
        if isinstance(self._normparam, numpy.generic):
            self._normparam = self._normparam.item()
        self._normfac = float(a["NormFac"])
        self._sens_energy = bool(a["SensEnergy"])
        self._effective_freq = float(a["EffectiveFreq"])
        self._effective_wave = float(a["EffectiveWave"])
        self._nresp = int(a["NResp"])
        self._wave = numpy.array(d["Wave"], dtype=numpy.float64, ndmin=1)
        self._freq = numpy.array(d["Freq"], dtype=numpy.float64, ndmin=1)
        self._resp = numpy.array(d["Response"], dtype=numpy.float64, ndmin=1)
        for key, attr in (("Dnu", "_dnu"), ("Sedmult", "_sedmult")):
            if key in d:
                setattr(self, attr, numpy.array(d[key], dtype=numpy.float64, ndmin=1))
            elif hasattr(self, attr):
                delattr(self, attr)
        self._data_read = True
        return self

    def writeToHDF5(self, handle):
        """Writes the response to an open h5py file or group (reference :578)."""
        from .treeio import _h5_write
        _h5_write(handle, self.to_tree())

    def readFromHDF5(self, handle):
        """Reads the response from an open h5py file or group (reference :608)."""
        from .treeio import _h5_read
        self.from_tree(_h5_read(handle))

    def __str__(self):
        return "{0:s} lambda_eff: {1:0.1f} [um]".format(self._name,
                                                       self._effective_wave)


class response_set(object):
    """A named collection of passbands (reference response.py:642-840)."""

    def __init__(self, inputfile=None, dir=None):
        self._responses = {}
        self.read(inputfile=inputfile, dir=dir)

    def read(self, inputfile=None, dir=None):
        """(Re)load the set; ``inputfile=None`` loads the shipped wheel."""
        if inputfile is None:
            self._responses.clear()
            with numpy.load(_PACKED) as pk:
                spec = [(str(pk["wheel_names"][i]), str(pk["wheel_files"][i]),
                         str(pk["wheel_xtype"][i]), str(pk["wheel_xunits"][i]),
                         str(pk["wheel_senstype"][i]), str(pk["wheel_normtype"][i]),
                         float(pk["wheel_xnorm"][i]), float(pk["wheel_normparam"][i]))
                        for i in range(len(pk["wheel_names"]))]
            for row in spec:
                self.add(*row, dir='!package-dir!')
            return
        if not isinstance(inputfile, str):
            raise TypeError("filename must be string-like")
        if dir is None:
            path, indir = inputfile, None
        else:
            if not isinstance(dir, str):
                raise TypeError("dir must be string-like")
            path, indir = os.path.join(dir, inputfile), dir
        rows = read_text_table(path)
        if len(rows) == 0:
            raise IOError("No data read from {:s}".format(inputfile))
        self._responses.clear()
        for r in rows:
            self.add(r[0], r[1], r[2].lower(), r[3].lower(), r[4].lower(),
                     r[5].lower(), float(r[6]), float(r[7]), dir=indir)

    def add(self, name, spec, xtype, xunits, senstype, normtype, xnorm,
            normparam, dir=None):
        resp = response(name)
        resp.setup(spec, xtype=xtype, xunits=xunits, senstype=senstype,
                   normtype=normtype, xnorm=xnorm, normparam=normparam, dir=dir)
        self._responses[name] = resp

    def add_special(self, name):
        """``Inst_type_v1[um|ghz]_v2...`` shortcut (reference response.py:722-780):
        always energy sensitivity and flat normalisation, GHz unless 'um'."""
        parts = name.split('_')
        if len(parts) < 2 or parts[1].lower() not in special_types:
            raise ValueError("Unknown 'special' response type in {:s}".format(name))
        kind = parts[1].lower()
        if len(parts) < 3:
            raise ValueError("Special type has no numerical spec")
        first = parts[2].lower()
        numpat = re.compile(r'\d*\.\d+|\d+')
        nums = numpat.findall(first)
        if not nums:
            raise ValueError("Special type needs numeric specification")
        unit = numpat.sub("", first)
        if unit in ("", "ghz"):
            xtype, xunit = 'freq', 'ghz'
        elif unit == "um":
            xtype, xunit = 'wave', 'microns'
        else:
            raise ValueError("Unable to understand unit specification "
                             "{:s}".format(unit))
        spec = '_'.join([kind, nums[0]] + parts[3:])
        resp = response(name)
        resp.setup(spec, xtype=xtype, xunits=xunit, senstype='energy',
                   normtype='flat', xnorm=float(nums[0]), normparam=0)
        self._responses[name] = resp

    # ---------------------------------------------------------- serialisation
    def to_tree(self):
        """One group per response (reference response.py:782-790)."""
        from .treeio import new_tree
        t = new_tree()
        for name, resp in self._responses.items():
            t["groups"][name] = resp.to_tree()
        return t

    def from_tree(self, t):
        """Replaces the set's contents (reference response.py:792-802)."""
        self._responses.clear()
        for name, sub in t["groups"].items():
            self._responses[name] = response(name).from_tree(sub)
        return self

    def writeToHDF5(self, handle):
        from .treeio import _h5_write
        _h5_write(handle, self.to_tree())

    def readFromHDF5(self, handle):
        from .treeio import _h5_read
        self.from_tree(_h5_read(handle))

    def save(self, filename):
        """The whole set to a file ('.h5' needs h5py, anything else is '.npz')."""
        from .treeio import write_tree
        return write_tree(filename, self.to_tree())

    @classmethod
    def load(cls, filename):
        from .treeio import read_tree
        self = cls.__new__(cls)          # (the constructor would read the shipped wheel first)
        self._responses = {}
        return self.from_tree(read_tree(filename))

    def __getitem__(self, name):
        return self._responses[name]

    def keys(self):
        return self._responses.keys()

    def __contains__(self, val):
        return val in self._responses

    def items(self):
        return self._responses.items()

    def values(self):
        return self._responses.values()

    def __delitem__(self, val):
        del self._responses[val]

    def __str__(self):
        return '\n'.join(str(r) for r in self._responses.values())
