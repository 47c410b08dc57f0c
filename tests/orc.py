"""ctypes binding of the CPU oracle (oracle/_ref/liboracle.so).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from grtrans_b200._abi import GrtArgs, args_from_kwargs  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
        _lib = C.CDLL(_SO)
        for f in ("orc_rf", "orc_rd"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_double] * 3
        _lib.orc_rc.restype = C.c_double
        _lib.orc_rc.argtypes = [C.c_double] * 2
        _lib.orc_rj.restype = C.c_double
        _lib.orc_rj.argtypes = [C.c_double] * 4
        _lib.orc_besselk.restype = C.c_double
        _lib.orc_besselk.argtypes = [C.c_int, C.c_double]
        _lib.orc_bnu.restype = C.c_double
        _lib.orc_bnu.argtypes = [C.c_double, C.c_double]
    return _lib


class options(object):
    """with orc.options(name=value, ...): test-only switches of the oracle (orc_set_option, oracle/capi.cc): constants and
    code variants of older reference commits, noise injections for conditioning measurements.  Reset on exit."""

    def __init__(self, **opts):
        self.opts = opts

    def __enter__(self):
        L = lib()
        L.orc_set_option.argtypes = [C.c_char_p, C.c_double]
        assert L.orc_set_option(b"reset", 0.) == 0
        for k, v in self.opts.items():
            assert L.orc_set_option(k.encode(), float(v)) == 0, k

    def __exit__(self, *a):
        lib().orc_set_option(b"reset", 0.)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def sncndn(u, emmc):
    sn, cn, dn = C.c_double(), C.c_double(), C.c_double()
    lib().orc_sncndn(C.c_double(u), C.c_double(emmc), C.byref(sn), C.byref(cn), C.byref(dn))
    return sn.value, cn.value, dn.value


def run(nthreads=0, **kw):
    """Full oracle run.  Returns dict(ivals[nvals,npix,ncams] (float64 of float32), ab[2,npix], freqs,
    seconds, npoints)."""
    a = kw.pop("_args", None) or args_from_kwargs(**kw)
    npix = a.nn[0] * a.nn[1]
    ncams = a.nmdot * a.nfreq * a.nmu * a.nt
    nv = a.nvals + ((19 if a.nn[2] > 1 else 7) if a.extra == 1 else 0)   # extra=1: 19 diagnostics follow the Stokes values
    pixvals = np.zeros((ncams, npix, nv), np.float32)
    pixloc = np.zeros((npix, 2), np.float32)
    freqs = np.zeros(a.nfreq)
    secs = C.c_double()
    npts = C.c_longlong()
    rc = lib().orc_main(C.byref(a), C.c_int(nthreads), _p(pixvals), _p(pixloc), _p(freqs), C.byref(secs), C.byref(npts))
    assert rc == 0
    return dict(ivals=np.ascontiguousarray(pixvals.transpose(2, 1, 0)).astype(np.float64),
                ab=np.ascontiguousarray(pixloc.T).astype(np.float64), freqs=freqs, seconds=secs.value,
                npoints=npts.value)


def ray_dump(i1, i2, cap=None, **kw):
    a = kw.pop("_args", None) or args_from_kwargs(**kw)
    n = i2 - i1 + 1
    nup = a.nn[2]
    if cap is None:
        if a.standard == 1 and nup > 1:
            cap = nup + int(np.ceil(60. * nup / 400.)) - 2 * int(np.ceil(3. * nup / 400.))
        else:
            cap = nup
    d = {}
    ints = ["npts", "status", "ncase"]
    for k in ints:
        d[k] = np.zeros(n, np.int32)
    for k in ["lambda", "t", "r", "th", "ph", "kt", "kr", "kth", "kph"]:
        d[k] = np.zeros((n, cap))
    for k in ["tpm", "tpr"]:
        d[k] = np.zeros((n, cap), np.int32)
    for k in ["rho", "p", "bmag"]:
        d[k] = np.zeros((n, cap), np.float32)
    d["u"] = np.zeros((n, cap, 4)); d["b"] = np.zeros((n, cap, 4))
    for k in ["s2xi", "c2xi", "ang", "rshift", "cosne", "ncgs", "ncgsnth", "bcgs", "tcgs"]:
        d[k] = np.zeros((n, cap))
    d["e_last"] = np.zeros((n, cap, 4)); d["K_last"] = np.zeros((n, cap, 7)); d["tau"] = np.zeros((n, cap))
    d["stokes"] = np.zeros((n, a.nfreq, a.nvals))
    order = ["npts", "status", "ncase", "lambda", "t", "r", "th", "ph", "kt", "kr", "kth", "kph", "tpm", "tpr",
             "rho", "p", "bmag", "u", "b", "s2xi", "c2xi", "ang", "rshift", "cosne", "ncgs", "ncgsnth", "bcgs",
             "tcgs", "e_last", "K_last", "tau", "stokes"]
    rc = lib().orc_ray_dump(C.byref(a), C.c_int(i1), C.c_int(i2), C.c_int(cap), *[_p(d[k]) for k in order])
    assert rc == 0, rc
    d["cap"] = cap
    return d


def rt_integrate(iflag, s, j, K, tau, max_tau=10., hmax=0.1, oatol=1e-8, ortol=1e-6, thin=1e-2, maxsteps=100000,
                 as_written=True):
    s = np.ascontiguousarray(s, np.float64); j = np.ascontiguousarray(j, np.float64)
    K = np.ascontiguousarray(K, np.float64); tau = np.ascontiguousarray(tau, np.float64)
    npts = s.shape[0]
    neq = j.shape[1] if j.ndim > 1 else 1
    out = np.zeros((npts, neq))
    nsteps = C.c_int()
    nout = lib().orc_rt_integrate(C.c_int(iflag), C.c_int(neq), C.c_int(npts), _p(s), _p(j), _p(K), _p(tau), _p(out),
                                  C.c_double(max_tau), C.c_double(hmax), C.c_double(oatol), C.c_double(ortol),
                                  C.c_double(thin), C.c_int(maxsteps), C.c_int(1 if as_written else 0), C.byref(nsteps))
    return out, nout, nsteps.value


def polsynchth(nu, n, b, t, th):
    arrs = [np.ascontiguousarray(np.broadcast_to(x, np.broadcast(nu, n, b, t, th).shape), np.float64).ravel()
            for x in (nu, n, b, t, th)]
    e = np.zeros((arrs[0].size, 11))
    lib().orc_polsynchth(C.c_int(arrs[0].size), *[_p(x) for x in arrs], _p(e))
    return e


def polsynchpl(nu, nnth, b, th, p, gmin, gmax):
    arrs = [np.ascontiguousarray(np.broadcast_to(x, np.broadcast(nu, nnth, b, th, p, gmin).shape), np.float64).ravel()
            for x in (nu, nnth, b, th, p, gmin)]
    e = np.zeros((arrs[0].size, 11))
    lib().orc_polsynchpl(C.c_int(arrs[0].size), *[_p(x) for x in arrs], C.c_double(gmax), _p(e))
    return e


def synchemis(nu, n, b, t):
    arrs = [np.ascontiguousarray(np.broadcast_to(x, np.broadcast(nu, n, b, t).shape), np.float64).ravel()
            for x in (nu, n, b, t)]
    e = np.zeros((arrs[0].size, 11))
    lib().orc_synchemis(C.c_int(arrs[0].size), *[_p(x) for x in arrs], _p(e))
    return e


def synchbinemis(nu, nbins, gmin, gmax, b, th):
    """nbins[npt, nbin] = dn/dgamma per bin.  Returns (e[npt, 11], gammas, dgammas)."""
    nbins = np.ascontiguousarray(nbins, np.float64)
    npt, nb = nbins.shape
    arrs = [np.ascontiguousarray(np.broadcast_to(x, (npt,)), np.float64) for x in (nu, b, th)]
    e = np.zeros((npt, 11)); g = np.zeros(nb); dg = np.zeros(nb)
    lib().orc_synchbinemis(C.c_int(npt), _p(arrs[0]), _p(nbins), C.c_int(nb), C.c_double(gmin), C.c_double(gmax), _p(arrs[1]),
                           _p(arrs[2]), _p(e), _p(g), _p(dg))
    return e, g, dg
