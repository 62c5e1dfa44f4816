"""ctypes binding of libmbb_b200.so (the C ABI in include/mbb_b200.h).

There is no CPU fallback anywhere in this package: if the shared library has
not been built, or no CUDA device is usable, the first call that needs the
device raises ``MBBNativeError`` -- loudly, never silently.
"""
import ctypes
import os
import threading

import numpy as np

__all__ = ["MBBNativeError", "Context", "library_path", "load_library",
           "default_context", "raise_for_status", "STATUS_NAMES", "pinned_empty", "host_register"]

_HERE = os.path.dirname(os.path.abspath(__file__))
# MBB_B200_LIB: developer knob for A/B-timing alternative builds of the same library
_LIBNAME = os.environ.get("MBB_B200_LIB") or os.path.join(_HERE, "csrc", "libmbb_b200.so")

AOS, SOA = 0, 1
HOST, DEVICE = 0, 1
MATH_FAITHFUL, MATH_FAST, MATH_FAST_GAUSS = 0, 1, 2
LIR_QUADPACK, LIR_GAUSS = 0, 1

STATUS_NAMES = {0: "ok", 1: "below lower limit", 2: "bad alpha", 3: "bad beta",
                4: "bracket low", 5: "bracket high", 6: "no convergence",
                7: "overflow", 8: "peak bracket", 9: "non-finite"}


class MBBNativeError(RuntimeError):
    """The CUDA library is missing, failed to load, or reported an error."""


_lib = None
_lib_lock = threading.Lock()

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_c_uint8_p = ctypes.POINTER(ctypes.c_uint8)


def library_path():
    return _LIBNAME


def load_library():
    """Load libmbb_b200.so (once) and declare every entry point of the header."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(_LIBNAME):
            raise MBBNativeError(
                "CUDA extension not built: %s is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C mbb_emcee_b200/csrc` (needs nvcc, sm_100a). "
                "This package has no CPU fallback." % _LIBNAME)
        try:
            lib = ctypes.CDLL(_LIBNAME)
        except OSError as exc:
            raise MBBNativeError("could not load %s: %s" % (_LIBNAME, exc))
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
        sig = {
            "mbb_version": (i32, []),
            "mbb_last_error": (ctypes.c_char_p, []),
            "mbb_device_count": (i32, []),
            "mbb_ctx_create": (i32, [i32, ctypes.POINTER(vp)]),
            "mbb_ctx_destroy": (i32, [vp]),
            "mbb_sync": (i32, [vp]),
            "mbb_launch_count": (i64, [vp]),
            "mbb_stream_handle": (ctypes.c_uint64, [vp]),
            "mbb_last_kernel_ms": (i32, [vp, ctypes.POINTER(ctypes.c_float)]),
            "mbb_set_model": (i32, [vp, dbl, i32, i32]),
            "mbb_set_math_mode": (i32, [vp, i32]),
            "mbb_set_lir_method": (i32, [vp, i32]),
            "mbb_set_bands": (i32, [vp, i32, vp, vp, vp, vp]),
            "mbb_set_data": (i32, [vp, i32, i32, vp, vp, vp]),
            "mbb_set_data_chol": (i32, [vp, i32, i32, vp, vp]),
            "mbb_set_priors": (i32, [vp, vp, vp, vp, vp, vp, vp]),
            "mbb_set_fixed_params": (i32, [vp, vp, vp]),
            "mbb_loglike": (i32, [vp, i64, vp, i32, vp, i64, vp, vp, i32]),
            "mbb_fnu": (i32, [vp, i64, vp, i32, i32, vp, i32, i32, vp, vp, i32]),
            "mbb_sed_consts": (i32, [vp, i64, vp, i32, i32, vp, vp, i32]),
            "mbb_chain_post": (i32, [vp, i64, i64, vp, i32, dbl, dbl, dbl, dbl, dbl, dbl,
                                     vp, vp, vp, vp, i32]),
            "mbb_chain_flux": (i32, [vp, i64, i64, vp, i32, vp, vp, i32]),
            "mbb_chain_stats": (i32, [vp, i64, i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32]),
            "mbb_ensemble_run": (i32, [vp, i64, i32, i64, dbl, ctypes.c_uint64, ctypes.c_uint64, vp, vp,
                                       i32, vp, vp, vp, vp, i32, i32]),
            "mbb_ensemble_fit": (i32, [vp, i64, i32, i64, i64, dbl, ctypes.c_uint64, ctypes.c_uint64, i64,
                                       vp, vp, i32, vp, vp, vp, vp, vp, i64, i32, i32]),
            "mbb_host_alloc": (i32, [ctypes.c_size_t, ctypes.POINTER(vp)]),
            "mbb_host_free": (i32, [vp]),
            "mbb_host_register": (i32, [vp, ctypes.c_size_t]),
            "mbb_host_unregister": (i32, [vp]),
            "mbb_fp64_peak": (i32, [vp, i32, ctypes.POINTER(dbl)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)      # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


EXPORTED_SYMBOLS = ["mbb_version", "mbb_last_error", "mbb_device_count", "mbb_ctx_create",
                    "mbb_ctx_destroy", "mbb_sync", "mbb_launch_count", "mbb_stream_handle",
                    "mbb_last_kernel_ms", "mbb_set_model", "mbb_set_math_mode", "mbb_set_lir_method", "mbb_set_bands",
                    "mbb_set_data", "mbb_set_data_chol", "mbb_set_priors", "mbb_set_fixed_params", "mbb_loglike", "mbb_fnu", "mbb_sed_consts",
                    "mbb_chain_post", "mbb_chain_flux", "mbb_chain_stats", "mbb_ensemble_run", "mbb_ensemble_fit", "mbb_host_alloc", "mbb_host_free", "mbb_host_register",
                    "mbb_host_unregister", "mbb_fp64_peak"]

# layout of one row of mbb_ensemble_fit's per-source summary (include/mbb_b200.h MBB_FS_*)
FIT_NSTATS = 28
FS_N, FS_MEAN, FS_M2, FS_MIN, FS_MAX, FS_BESTLNP, FS_BEST, FS_ACC = 0, 1, 6, 11, 16, 21, 22, 27


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (mbb_host_alloc): the MBB_HOST calls copy from / to
    such buffers directly, without staging.  The memory is released when the last view dies."""
    import weakref
    lib = load_library()
    dt = np.dtype(dtype)
    n = int(np.prod(shape))
    nbytes = max(n * dt.itemsize, 1)
    p = ctypes.c_void_p()
    if lib.mbb_host_alloc(nbytes, ctypes.byref(p)) != 0:
        raise MBBNativeError(lib.mbb_last_error().decode())
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    weakref.finalize(buf, lib.mbb_host_free, p)      # numpy views keep `buf` alive through .base
    return np.frombuffer(buf, dtype=dt, count=n).reshape(shape)


def host_register(arr):
    """Page-lock a numpy array in place (cudaHostRegister); returns a callable that undoes it.
    False when registration is refused (the array then simply behaves as pageable memory)."""
    lib = load_library()
    if lib.mbb_host_register(ctypes.c_void_p(arr.ctypes.data), arr.nbytes) != 0:
        return False
    addr = ctypes.c_void_p(arr.ctypes.data)
    return lambda: lib.mbb_host_unregister(addr)


def raise_for_status(status, pars=None):
    """Map device status codes to the exceptions the reference raises
    (SURVEY.md 8b): ValueError for bad alpha/beta and bracket failures
    (modified_blackbody.py:219-224, 294-317), RuntimeError for a non-converged
    root solve, OverflowError for kappa (:326-328), Exception for the
    lambda_peak bracket (:612-630)."""
    status = np.asarray(status)
    bad = np.nonzero(status > 1)[0]
    if bad.size == 0:
        return
    i = int(bad[0])
    code = int(status.flat[i])
    where = "" if pars is None else " for parameters %s" % (np.asarray(pars).reshape(-1, 5)[i],)
    msg = {2: "alpha must be positive", 3: "beta must be non-negative",
           4: "Couldn't bracket low alpha merge point",
           5: "Couldn't bracket high alpha merge point",
           6: "root solve failed to converge",
           7: "overflow computing the merge constant",
           8: "Couldn't bracket maximum", 9: "non-finite parameters or result"}[code] + where
    if code in (2, 3, 4, 5, 9):
        raise ValueError(msg)
    if code == 6:
        raise RuntimeError(msg)
    if code == 7:
        raise OverflowError(msg)
    raise Exception(msg)


class Context(object):
    """One device + one stream + the staged tables (``mbb_ctx``)."""

    def __init__(self, device=0):
        self._lib = load_library()
        handle = ctypes.c_void_p()
        rc = self._lib.mbb_ctx_create(int(device), ctypes.byref(handle))
        if rc != 0:
            raise MBBNativeError(self._lib.mbb_last_error().decode())
        self._h = handle
        self.device = int(device)
        self.model = (500.0, False, False)
        self.math_mode = MATH_FAST

    def _ck(self, rc):
        if rc != 0:
            raise MBBNativeError(self._lib.mbb_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mbb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration -----------------------------------------------------
    def set_model(self, wavenorm, opthin, noalpha):
        self._ck(self._lib.mbb_set_model(self._h, float(wavenorm), int(bool(opthin)),
                                         int(bool(noalpha))))
        self.model = (float(wavenorm), bool(opthin), bool(noalpha))

    def set_math_mode(self, mode):
        self._ck(self._lib.mbb_set_math_mode(self._h, int(mode)))
        self.math_mode = int(mode)

    def set_lir_method(self, method):
        """'quadpack' (default; the reference's number) or 'gauss' (the true integral)."""
        m = {"quadpack": LIR_QUADPACK, "gauss": LIR_GAUSS}.get(method, method)
        self._ck(self._lib.mbb_set_lir_method(self._h, int(m)))

    def set_bands(self, band_off, node_wave, node_weight, scalar_path=None):
        off = np.ascontiguousarray(band_off, dtype=np.int32)
        wv, wt = _f64(node_wave), _f64(node_weight)
        nb = off.size - 1
        if wv.size != off[-1] or wt.size != off[-1]:
            raise ValueError("node arrays do not match band_off")
        sp = None if scalar_path is None else np.ascontiguousarray(scalar_path, dtype=np.uint8)
        self._ck(self._lib.mbb_set_bands(self._h, nb, _ptr(off), _ptr(wv), _ptr(wt), _ptr(sp)))
        self.nbands = nb

    def set_data(self, flux, ivar=None, cinv=None, chol=None):
        """Photometry of nsrc sources with exactly one of: inverse variances ``ivar`` [nsrc, nb],
        inverse covariances ``cinv`` [nsrc, nb, nb], or lower Cholesky factors ``chol``
        [nsrc, nb, nb] of the covariances (mbb_set_data_chol)."""
        flux = np.atleast_2d(_f64(flux))
        nsrc, nb = flux.shape
        if sum(a is not None for a in (ivar, cinv, chol)) != 1:
            raise ValueError("give exactly one of ivar / cinv / chol")
        if chol is not None:
            ch = _f64(chol).reshape(nsrc, nb, nb)
            self._ck(self._lib.mbb_set_data_chol(self._h, nsrc, nb, _ptr(flux), _ptr(ch)))
        else:
            iv = None if ivar is None else np.atleast_2d(_f64(ivar))
            ci = None if cinv is None else _f64(cinv).reshape(nsrc, nb, nb)
            self._ck(self._lib.mbb_set_data(self._h, nsrc, nb, _ptr(flux), _ptr(iv), _ptr(ci)))
        self.nsrc = nsrc

    def set_priors(self, lowlim, has_uplim, uplim, has_gprior, gmean, givar):
        lo = _f64(lowlim)
        hu = np.ascontiguousarray(has_uplim, dtype=np.uint8)
        up = _f64(uplim)
        hg = np.ascontiguousarray(has_gprior, dtype=np.uint8)
        gm, gi = _f64(gmean), _f64(givar)
        if lo.size != 5 or any(a.size != 6 for a in (hu, up, hg, gm, gi)):
            raise ValueError("limits/priors arrays have the wrong length")
        self._ck(self._lib.mbb_set_priors(self._h, _ptr(lo), _ptr(hu), _ptr(up), _ptr(hg),
                                          _ptr(gm), _ptr(gi)))

    def set_fixed_params(self, fixed=None, values=None):
        """Promise that column i of every parameter block equals values[i] where fixed[i]
        (mbb_fitter.fix_param, reference mbb_fit.py:183-200, 440-443): the host path then skips
        those columns of SoA batches on the way to the device.  ``fixed=None`` clears it."""
        if fixed is None:
            self._ck(self._lib.mbb_set_fixed_params(self._h, None, None))
            return
        fx = np.ascontiguousarray(fixed, dtype=np.int32)
        va = _f64(values)
        if fx.size != 5 or va.size != 5:
            raise ValueError("fixed / values must have 5 entries")
        self._ck(self._lib.mbb_set_fixed_params(self._h, _ptr(fx), _ptr(va)))

    # -- compute -----------------------------------------------------------
    def loglike(self, pars, src_index=None, walkers_per_source=None, layout=AOS):
        """Host arrays in, host arrays out: (lnlike[n], status[n])."""
        P = _f64(pars)
        n = P.size // 5
        out = np.empty(n, dtype=np.float64)
        st = np.empty(n, dtype=np.int32)
        si = None if src_index is None else np.ascontiguousarray(src_index, dtype=np.int32)
        wps = n if walkers_per_source is None else int(walkers_per_source)
        self._ck(self._lib.mbb_loglike(self._h, n, _ptr(P), layout, _ptr(si), max(wps, 1),
                                       _ptr(out), _ptr(st), HOST))
        return out, st

    def loglike_into(self, pars, out, status=None, src_index=None, walkers_per_source=None, layout=AOS):
        """Host path with caller-owned buffers, used as they are (no copies on the Python
        side): when they are page-locked (``torch.Tensor.pin_memory().numpy()``,
        ``cudaHostRegister``) the library copies from / to them directly."""
        for name, a, dt in (("pars", pars, np.float64), ("out", out, np.float64), ("status", status, np.int32),
                            ("src_index", src_index, np.int32)):
            if a is not None and not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
                raise TypeError("%s must be a C-contiguous %s array" % (name, np.dtype(dt).name))
        n = pars.size // 5
        if pars.size != 5 * n or out.size != n or (status is not None and status.size != n) or \
                (src_index is not None and src_index.size != n):
            raise ValueError("pars[5n] / out[n] / status[n] / src_index[n] sizes disagree")
        wps = n if walkers_per_source is None else int(walkers_per_source)
        self._ck(self._lib.mbb_loglike(self._h, n, _ptr(pars), layout, _ptr(src_index), max(wps, 1),
                                       _ptr(out), _ptr(status), HOST))

    def loglike_device(self, n, pars_ptr, out_ptr, status_ptr=0, src_index_ptr=0,
                       walkers_per_source=None, layout=AOS):
        """Raw device pointers (e.g. torch.Tensor.data_ptr()); asynchronous."""
        wps = n if walkers_per_source is None else int(walkers_per_source)
        self._ck(self._lib.mbb_loglike(self._h, int(n), ctypes.c_void_p(pars_ptr), layout,
                                       ctypes.c_void_p(src_index_ptr) if src_index_ptr else None,
                                       max(wps, 1), ctypes.c_void_p(out_ptr),
                                       ctypes.c_void_p(status_ptr) if status_ptr else None, DEVICE))

    def fnu(self, pars, freq_ghz, scalar_path=False, math_mode=MATH_FAITHFUL):
        """f_nu[n][nfreq]; math_mode MATH_FAITHFUL (default: the reference's formulas in the
        reference's order) or MATH_FAST (the likelihood kernels' arithmetic)."""
        P = _f64(pars).reshape(-1, 5)
        f = _f64(freq_ghz).ravel()
        out = np.empty((P.shape[0], f.size), dtype=np.float64)
        st = np.empty(P.shape[0], dtype=np.int32)
        self._ck(self._lib.mbb_fnu(self._h, P.shape[0], _ptr(P), AOS, f.size, _ptr(f),
                                   int(bool(scalar_path)), int(math_mode), _ptr(out), _ptr(st), HOST))
        return out, st

    def sed_consts(self, pars, want_peak=False):
        P = _f64(pars).reshape(-1, 5)
        out = np.empty((P.shape[0], 6), dtype=np.float64)
        st = np.empty(P.shape[0], dtype=np.int32)
        self._ck(self._lib.mbb_sed_consts(self._h, P.shape[0], _ptr(P), AOS, int(bool(want_peak)),
                                          _ptr(out), _ptr(st), HOST))
        return out, st

    def chain_post(self, chain, which, z=0.0, dl_mpc=1.0, lir_min=8.0, lir_max=1000.0,
                   kappa=2.64, kappa_wave=125.0):
        chain = _f64(chain)
        nw, ns = chain.shape[0], chain.shape[1]
        pk = np.empty((nw, ns)) if which & 1 else None
        lir = np.empty((nw, ns)) if which & 2 else None
        dm = np.empty((nw, ns)) if which & 4 else None
        st = np.empty((nw, ns), dtype=np.int32)
        self._ck(self._lib.mbb_chain_post(self._h, nw, ns, _ptr(chain), int(which), float(z),
                                          float(dl_mpc), float(lir_min), float(lir_max),
                                          float(kappa), float(kappa_wave), _ptr(pk), _ptr(lir),
                                          _ptr(dm), _ptr(st), HOST))
        return pk, lir, dm, st

    def chain_post_into(self, chain, which, peak=None, lir=None, dustmass=None, status=None, z=0.0,
                        dl_mpc=1.0, lir_min=8.0, lir_max=1000.0, kappa=2.64, kappa_wave=125.0):
        """mbb_chain_post(MBB_HOST) on caller-owned arrays (e.g. the rows of one shard inside shared
        page-locked outputs): chain[nw][ns][5] C-contiguous, outputs [nw][ns] for the bits of
        ``which`` (1 peak, 2 L_IR, 4 dust mass)."""
        nw, ns = chain.shape[0], chain.shape[1]
        for name, a, dt, need in (("chain", chain, np.float64, True), ("peak", peak, np.float64, which & 1),
                                  ("lir", lir, np.float64, which & 2), ("dustmass", dustmass, np.float64, which & 4),
                                  ("status", status, np.int32, False)):
            if a is None:
                if need:
                    raise ValueError("%s is required" % name)
                continue
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
                raise TypeError("%s must be a C-contiguous %s array" % (name, np.dtype(dt).name))
            if name != "chain" and a.shape != (nw, ns):
                raise ValueError("%s must have shape (nwalkers, nsteps)" % name)
        self._ck(self._lib.mbb_chain_post(self._h, nw, ns, _ptr(chain), int(which), float(z), float(dl_mpc),
                                          float(lir_min), float(lir_max), float(kappa), float(kappa_wave),
                                          _ptr(peak), _ptr(lir), _ptr(dustmass), _ptr(status), HOST))

    def chain_flux(self, chain, band=0):
        chain = _f64(chain)
        nw, ns = chain.shape[0], chain.shape[1]
        out = np.empty((nw, ns))
        st = np.empty((nw, ns), dtype=np.int32)
        self._ck(self._lib.mbb_chain_flux(self._h, nw, ns, _ptr(chain), int(band), _ptr(out), _ptr(st), HOST))
        return out, st

    def chain_stats(self, x, percentiles=(), lowlim=None, uplim=None, where=HOST, nrows=None, ncols=None):
        """Mean and percentiles of the columns of x[n][ncols] (or x[n]) on the device
        (mbb_chain_stats), clipped to [lowlim, uplim] per column.  Returns (mean[ncols],
        count[ncols], perc[ncols][len(percentiles)]) with the percentiles formed from the two
        bracketing order statistics exactly as numpy.percentile's default method does
        (numpy/lib/_function_base_impl.py _lerp), so they equal numpy's bit for bit.
        ``where=DEVICE``: x is a device pointer (int) and nrows / ncols give its shape."""
        if where == HOST:
            x = _f64(x)
            if x.ndim == 1:
                x = x.reshape(-1, 1)
            n, nc = x.shape
            xp = _ptr(x)
        else:
            n, nc, xp = int(nrows), int(ncols), ctypes.c_void_p(int(x))
        q = np.true_divide(_f64(np.atleast_1d(percentiles)), 100)
        nq = int(q.size)

        def lim(v, fill):
            if v is None:
                return None
            a = np.full(nc, fill, dtype=np.float64)
            vv = np.atleast_1d(v)
            for i in range(nc):
                if vv[min(i, vv.size - 1)] is not None:
                    a[i] = float(vv[min(i, vv.size - 1)])
            return a
        lo, hi = lim(lowlim, -np.inf), lim(uplim, np.inf)
        mean = np.empty(nc)
        count = np.empty(nc, dtype=np.int64)
        qlo, qhi, gam = (np.empty((nc, nq)) for _ in range(3))
        self._ck(self._lib.mbb_chain_stats(self._h, n, nc, xp, _ptr(lo), _ptr(hi), nq, _ptr(q), _ptr(mean),
                                           _ptr(count), _ptr(qlo), _ptr(qhi), _ptr(gam), where))
        # numpy's _lerp: a + (b - a) t, and b - (b - a)(1 - t) where t >= 0.5
        diff = qhi - qlo
        perc = qlo + diff * gam
        with np.errstate(invalid="ignore"):
            perc = np.where(gam >= 0.5, qhi - diff * (1 - gam), perc)
        return mean, count, perc

    def ensemble_run(self, pos, nsteps, seed=0, step0=0, a=2.0, lnprob=None):
        """Device-resident stretch-move sampler (host arrays in/out).
        pos[nsrc][nwalkers][5]; returns (pos, lnprob, naccept, status)."""
        out = self.ensemble_fit(pos, 0, nsteps, seed=seed, step0=step0, a=a, lnprob=lnprob, stats=False)
        return out["pos"], out["lnprob"], out["naccept"], out["status"]

    def ensemble_fit(self, pos, nburn, nsteps, seed=0, step0=0, src0=0, a=2.0, lnprob=None, stats=True,
                     chain=False, thin=1):
        """mbb_ensemble_fit with host arrays: burn-in, main run, per-source posterior summary and
        (optionally) the recorded chain.  Returns a dict: pos, lnprob, naccept, status, and when
        asked for stats[nsrc][FIT_NSTATS], chain[nrec][nsrc][nw][5], chain_lnprob[nrec][nsrc][nw]."""
        pos = np.array(pos, dtype=np.float64, order="C")
        nsrc, nw = pos.shape[0], pos.shape[1]
        have = lnprob is not None
        lnp = np.array(lnprob, dtype=np.float64, order="C") if have else np.empty((nsrc, nw))
        nacc = np.zeros((nsrc, nw), dtype=np.int32)
        st = np.zeros((nsrc, nw), dtype=np.int32)
        thin = max(int(thin), 1)
        nrec = int(nsteps) // thin
        stt = np.zeros((nsrc, FIT_NSTATS)) if stats else None
        ch = np.empty((nrec, nsrc, nw, 5)) if chain else None
        chl = np.empty((nrec, nsrc, nw)) if chain else None
        self._ck(self._lib.mbb_ensemble_fit(self._h, nsrc, nw, int(nburn), int(nsteps), float(a), int(seed),
                                            int(step0), int(src0), _ptr(pos), _ptr(lnp), int(have), _ptr(nacc),
                                            _ptr(st), _ptr(stt), _ptr(ch), _ptr(chl), 0, thin, HOST))
        out = {"pos": pos, "lnprob": lnp, "naccept": nacc, "status": st}
        if stats:
            out["stats"] = stt
        if chain:
            out["chain"], out["chain_lnprob"] = ch, chl
        return out

    def ensemble_run_device(self, nsrc, nwalkers, nsteps, pos_ptr, lnprob_ptr, have_lnprob=False,
                            seed=0, step0=0, a=2.0, naccept_ptr=0, status_ptr=0, chain_ptr=0,
                            chain_lnprob_ptr=0, thin=1):
        """Raw device pointers; asynchronous (call sync())."""
        def vp(x):
            return ctypes.c_void_p(x) if x else None
        self._ck(self._lib.mbb_ensemble_run(self._h, int(nsrc), int(nwalkers), int(nsteps), float(a),
                                            int(seed), int(step0), vp(pos_ptr), vp(lnprob_ptr),
                                            int(bool(have_lnprob)), vp(naccept_ptr), vp(status_ptr),
                                            vp(chain_ptr), vp(chain_lnprob_ptr), int(thin), DEVICE))

    def ensemble_fit_into(self, pos, lnprob, nburn, nsteps, naccept=None, status=None, stats=None,
                          chain=None, chain_lnprob=None, chain_nsrc=0, have_lnprob=False, seed=0, step0=0,
                          src0=0, a=2.0, thin=1):
        """mbb_ensemble_fit(MBB_HOST) on caller-owned arrays, used in place: pos[nsrc][nw][5] and
        lnprob[nsrc][nw] are updated; the optional outputs are filled.  ``chain`` / ``chain_lnprob``
        may be views [:, lo:] of arrays holding ``chain_nsrc`` sources per record (their first
        element is where this call's source 0 goes).  Page-locked arrays (``pinned_empty``) are
        DMA'd directly."""
        def chk(name, arr, dt, contiguous=True):
            if arr is None:
                return
            if not isinstance(arr, np.ndarray) or arr.dtype != dt or (contiguous and not arr.flags.c_contiguous):
                raise TypeError("%s must be a C-contiguous %s array" % (name, np.dtype(dt).name))
        chk("pos", pos, np.float64)
        chk("lnprob", lnprob, np.float64)
        chk("naccept", naccept, np.int32)
        chk("status", status, np.int32)
        chk("stats", stats, np.float64)
        chk("chain", chain, np.float64, False)
        chk("chain_lnprob", chain_lnprob, np.float64, False)
        nsrc, nw = pos.shape[0], pos.shape[1]
        if pos.shape != (nsrc, nw, 5) or lnprob.shape != (nsrc, nw):
            raise ValueError("pos[nsrc][nw][5] / lnprob[nsrc][nw] shapes disagree")
        for name, arr, shp in (("naccept", naccept, (nsrc, nw)), ("status", status, (nsrc, nw)),
                               ("stats", stats, (nsrc, FIT_NSTATS))):
            if arr is not None and arr.shape != shp:
                raise ValueError("%s must have shape %s" % (name, shp))
        thin = max(int(thin), 1)
        cn = int(chain_nsrc) if chain_nsrc else nsrc
        for name, arr, tail in (("chain", chain, (nw, 5)), ("chain_lnprob", chain_lnprob, (nw,))):
            if arr is None:
                continue
            inner = int(np.prod(tail)) * 8
            if arr.shape[0] != int(nsteps) // thin or arr.shape[1] < nsrc or arr.shape[2:] != tail or \
                    arr.strides[1] != inner or (arr.shape[0] > 1 and arr.strides[0] != cn * inner) or \
                    not arr[0, 0].flags.c_contiguous:
                raise ValueError("%s must be (a [:, lo:] view of) a C-contiguous [nsteps//thin][chain_nsrc]%s array"
                                 % (name, list(tail)))
        self._ck(self._lib.mbb_ensemble_fit(self._h, nsrc, nw, int(nburn), int(nsteps), float(a), int(seed),
                                            int(step0), int(src0), _ptr(pos), _ptr(lnprob), int(bool(have_lnprob)),
                                            _ptr(naccept), _ptr(status), _ptr(stats), _ptr(chain),
                                            _ptr(chain_lnprob), cn, thin, HOST))

    def ensemble_fit_device(self, nsrc, nwalkers, nburn, nsteps, pos_ptr, lnprob_ptr, have_lnprob=False,
                            seed=0, step0=0, src0=0, a=2.0, naccept_ptr=0, status_ptr=0, stats_ptr=0,
                            chain_ptr=0, chain_lnprob_ptr=0, thin=1):
        """Raw device pointers; asynchronous (call sync())."""
        def vp(x):
            return ctypes.c_void_p(x) if x else None
        self._ck(self._lib.mbb_ensemble_fit(self._h, int(nsrc), int(nwalkers), int(nburn), int(nsteps),
                                            float(a), int(seed), int(step0), int(src0), vp(pos_ptr),
                                            vp(lnprob_ptr), int(bool(have_lnprob)), vp(naccept_ptr),
                                            vp(status_ptr), vp(stats_ptr), vp(chain_ptr),
                                            vp(chain_lnprob_ptr), 0, int(thin), DEVICE))

    def sync(self):
        self._ck(self._lib.mbb_sync(self._h))

    def launch_count(self):
        return int(self._lib.mbb_launch_count(self._h))

    def stream_handle(self):
        return int(self._lib.mbb_stream_handle(self._h))

    def last_kernel_ms(self):
        ms = ctypes.c_float()
        self._ck(self._lib.mbb_last_kernel_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def fp64_peak(self, iters=20000):
        t = ctypes.c_double()
        self._ck(self._lib.mbb_fp64_peak(self._h, int(iters), ctypes.byref(t)))
        return float(t.value)


_default_ctx = {}
_default_lock = threading.Lock()


def default_device():
    """CUDA ordinal used when none is given: MBB_B200_DEVICE, else LOCAL_RANK
    (one process per GPU under torchrun), else 0."""
    device = int(os.environ.get("MBB_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    n = load_library().mbb_device_count()
    if n > 0:
        device %= n
    return device


def default_context(device=None):
    """Shared context for single-SED conveniences (modified_blackbody objects)."""
    if device is None:
        device = default_device()
    with _default_lock:
        ctx = _default_ctx.get(device)
        if ctx is None:
            ctx = Context(device)
            _default_ctx[device] = ctx
        return ctx
