"""TEST INFRASTRUCTURE: ctypes wrapper of the host-emulation build of the
device model code (tests/hostemu/emu.cpp).  Never used by the product."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_hostemu", "libmbb_hostemu.so")
_lib = None


class EmuPriors(ctypes.Structure):
    _fields_ = [("lowlim", ctypes.c_double * 5), ("uplim", ctypes.c_double * 6),
                ("gmean", ctypes.c_double * 6), ("givar", ctypes.c_double * 6),
                ("has_uplim", ctypes.c_ubyte * 6), ("has_gprior", ctypes.c_ubyte * 6)]


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(HERE, "hostemu"), "-s"])
        _lib = ctypes.CDLL(_SO)
        _lib.emu_c64_hi.restype = ctypes.c_double
    return _lib


def set_tab_bits(bits):
    """Table flavour of the emulated specialised node code: 6 = 64 entries (delta kernels),
    8 = 256 entries (nodes kernel, Gauss-rule kernels)."""
    lib().emu_set_tab_bits(int(bits))


def tab_bits():
    return int(lib().emu_tab_bits())


def c64_hi():
    return float(lib().emu_c64_hi())


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _c(a, dt=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def priors_struct(lowlim, has_uplim, uplim, has_gprior, gmean, givar):
    ep = EmuPriors()
    for i in range(5):
        ep.lowlim[i] = float(lowlim[i])
    for i in range(6):
        ep.uplim[i] = float(uplim[i])
        ep.gmean[i] = float(gmean[i])
        ep.givar[i] = float(givar[i])
        ep.has_uplim[i] = 1 if has_uplim[i] else 0
        ep.has_gprior[i] = 1 if has_gprior[i] else 0
    return ep


def loglike(opthin, noalpha, fast, pars, wavenorm, ep, band_off, wave, weight, scalar_path,
            flux, ivar=None, cinv=None, wps=None):
    P = _c(pars).reshape(-1, 5)
    n = P.shape[0]
    out = np.empty(n)
    st = np.empty(n, dtype=np.int32)
    off = _c(band_off, np.int32)
    wv, wt = _c(wave), _c(weight)
    sp = _c(scalar_path, np.uint8)
    fl, iv, ci = _c(flux), _c(ivar), _c(cinv)
    lib().emu_loglike(int(opthin), int(not noalpha), int(fast), ctypes.c_longlong(n), _p(P),
                      ctypes.c_double(wavenorm), ctypes.byref(ep), int(off.size - 1), _p(off), _p(wv),
                      _p(wt), _p(sp), _p(fl), _p(iv), _p(ci),
                      ctypes.c_longlong(n if wps is None else wps), _p(out), _p(st))
    return out, st


def quad_form(m, diff, chol):
    """diff' inv(C) diff from the staged matrix, as the device forms it (mbb_model.cuh quad_form)."""
    m, diff = _c(m), _c(diff)
    f = lib().emu_quad_form
    f.restype = ctypes.c_double
    return float(f(_p(m), _p(diff), int(diff.size), int(bool(chol))))


def last_compressed():
    """(walker, band) pairs the last loglike(fast=3) call evaluated with a compressed rule."""
    f = lib().emu_last_compressed
    f.restype = ctypes.c_longlong
    return int(f())


def consts(opthin, noalpha, pars, wavenorm, want_peak=True):
    P = _c(pars).reshape(-1, 5)
    n = P.shape[0]
    out = np.empty((n, 6))
    st = np.empty(n, dtype=np.int32)
    lib().emu_consts(int(opthin), int(not noalpha), ctypes.c_longlong(n), _p(P),
                     ctypes.c_double(wavenorm), int(want_peak), _p(out), _p(st))
    return out, st


def fnu(opthin, noalpha, pars, wavenorm, freq, scalar_path=False, fast=False):
    P = _c(pars).reshape(-1, 5)
    f = _c(freq)
    out = np.empty((P.shape[0], f.size))
    lib().emu_fnu(int(opthin), int(not noalpha), ctypes.c_longlong(P.shape[0]), _p(P),
                  ctypes.c_double(wavenorm), int(f.size), _p(f), int(scalar_path), int(fast), _p(out))
    return out


def lir(opthin, noalpha, pars, wavenorm, minwave, maxwave, prefac=1.0):
    """freq_integrate over [minwave, maxwave] um (observer frame), times prefac."""
    P = _c(pars).reshape(-1, 5)
    n = P.shape[0]
    out = np.empty(n)
    st = np.empty(n, dtype=np.int32)
    fmin, fmax = 299792458e-3 / maxwave, 299792458e-3 / minwave
    lib().emu_lir(int(opthin), int(not noalpha), ctypes.c_longlong(n), _p(P), ctypes.c_double(wavenorm),
                  ctypes.c_double(fmin), ctypes.c_double(fmax), ctypes.c_double(prefac), _p(out), _p(st))
    return out, st


def fastmath(mode, x):
    x = _c(x)
    out = np.empty_like(x)
    lib().emu_fastmath(int(mode), ctypes.c_longlong(x.size), _p(x), _p(out))
    return out


def grey_nodes(opthin, pars, wavenorm, wave, weight):
    """(node_acc values, grey_nodes_n values, safe flags) for 6 nodes per parameter vector."""
    P = _c(pars).reshape(-1, 5)
    n = P.shape[0]
    wv, wt = _c(wave), _c(weight)
    assert wv.size == 6 and wt.size == 6
    a = np.zeros((n, 6))
    b = np.zeros((n, 6))
    safe = np.zeros(n, dtype=np.int32)
    lib().emu_grey_nodes(int(opthin), ctypes.c_longlong(n), _p(P), ctypes.c_double(wavenorm), _p(wv), _p(wt),
                         _p(a), _p(b), _p(safe))
    return a, b, safe.astype(bool)


def gauss_rule(x, w, n):
    """n-point Gauss rule of the discrete measure {x, w} (csrc/mbb_gaussrule.h), or None."""
    x, w = _c(x), _c(w)
    xs, ws = np.empty(n), np.empty(n)
    ok = lib().emu_gauss_rule(int(x.size), _p(x), _p(w), int(n), _p(xs), _p(ws))
    return (xs, ws) if ok else None


def wps_division(g, wps):
    """floor(g / wps) through the launcher's multiply-shift constants (csrc/mbb_hostutil.h)."""
    g = np.ascontiguousarray(g, dtype=np.uint32)
    q = np.empty_like(g)
    lib().emu_wps_division(ctypes.c_longlong(g.size), _p(g), ctypes.c_uint(int(wps)), _p(q))
    return q


def fast_setup(opthin, noalpha, pars, wavenorm):
    """(xmerge, amp_grey, amp_pow, safe)[n, 4] and status[n] of the FAST per-walker setup."""
    P = _c(pars).reshape(-1, 5)
    out = np.empty((P.shape[0], 4))
    st = np.empty(P.shape[0], dtype=np.int32)
    lib().emu_fast_setup(int(opthin), int(not noalpha), ctypes.c_longlong(P.shape[0]), _p(P),
                         ctypes.c_double(wavenorm), _p(out), _p(st))
    return out, st


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    u = ctypes.c_double()
    lib().emu_philox(_p(c), _p(k), _p(out), ctypes.byref(u))
    return out, u.value


def philox_uniforms(w0, w1, w3):
    """(stretch uniform, acceptance uniform) the sampler cuts out of one Philox block."""
    uz, ua = ctypes.c_double(), ctypes.c_double()
    lib().emu_philox_uniforms(ctypes.c_uint(int(w0)), ctypes.c_uint(int(w1)), ctypes.c_uint(int(w3)),
                              ctypes.byref(uz), ctypes.byref(ua))
    return uz.value, ua.value


def qags(opthin, noalpha, pars, wavenorm, minwave, maxwave, prefac=1.0):
    """freq_integrate by the QUADPACK replay (csrc/mbb_quadpack.cuh)."""
    P = _c(pars).reshape(-1, 5)
    n = P.shape[0]
    out = np.empty(n)
    nev = np.empty(n, dtype=np.int32)
    st = np.empty(n, dtype=np.int32)
    fmin, fmax = 299792458e-3 / maxwave, 299792458e-3 / minwave
    lib().emu_qags(int(opthin), int(not noalpha), ctypes.c_longlong(n), _p(P), ctypes.c_double(wavenorm),
                   ctypes.c_double(fmin), ctypes.c_double(fmax), ctypes.c_double(prefac), _p(out), _p(nev),
                   _p(st))
    return out, nev, st


def qags_test(kind, a, b, epsabs=1.49e-8, epsrel=1.49e-8):
    r, e = ctypes.c_double(), ctypes.c_double()
    nv, ier = ctypes.c_int(), ctypes.c_int()
    lib().emu_qags_test(int(kind), ctypes.c_double(a), ctypes.c_double(b), ctypes.c_double(epsabs),
                        ctypes.c_double(epsrel), ctypes.byref(r), ctypes.byref(e), ctypes.byref(nv),
                        ctypes.byref(ier))
    return r.value, e.value, nv.value, ier.value
