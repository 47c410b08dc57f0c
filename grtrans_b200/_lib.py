"""ctypes loader for libgrtrans_b200.so (the C ABI in include/grtrans_b200.h).

There is NO CPU fallback: if the shared library is missing, or a compute entry point reports that no
CUDA device is usable, a GrtError is raised.
"""
import ctypes as C
import os

import numpy as np

from ._abi import GrtArgs, args_from_kwargs, args_from_inputs, Inputs  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("GRT_B200_LIB", os.path.join(HERE, "libgrtrans_b200.so"))   # (override: kernel-variant experiments)

ERRORS = {1: "no usable CUDA device (GRT_ERR_NO_DEVICE)", 2: "unsupported argument (GRT_ERR_BAD_ARG)",
          3: "input/output file error (GRT_ERR_IO)", 4: "module data state error (GRT_ERR_STATE)",
          5: "CUDA runtime error (GRT_ERR_CUDA)"}

# every symbol include/grtrans_b200.h declares
SYMBOLS = ["grt_args_defaults", "grt_main", "grt_get_ivals", "grt_get_ab", "grt_get_freqs", "grt_data_allocated",
           "grt_del_data", "grt_run", "grt_program", "grt_read_inputs", "grt_write_output", "grt_geokerr_camera",
           "grt_geokerr_rays", "grt_ray_stage_dump", "grt_geodebug", "grt_polsynchth", "grt_polsynchpl", "grt_synchemis",
           "grt_rt_integrate", "grt_session_create", "grt_session_destroy", "grt_session_launch", "grt_session_sync", "grt_session_launch_fetch", "grt_session_upload", "grt_session_run_timed",
           "grt_session_time", "grt_session_launch_count", "grt_session_counts",
           "grt_set_devices", "grt_main_info", "grt_release_cache", "grt_shard_rays", "grt_nccl_unique_id",
           "grt_session_comm_init", "grt_session_launch_gather", "grt_session_run_timed_gather",
           "grt_session_last_launches", "grt_session_tile", "grt_session_full_image", "grt_session_fetch_full",
           "grt_session_fetch", "grt_session_ncams", "grt_session_nvals", "grt_version", "grt_device_count", "grt_dfma_peak", "grt_math_probe"]


class GrtError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise GrtError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % SO_PATH)
        _lib = C.CDLL(SO_PATH)
        _lib.grt_version.restype = C.c_char_p
        _lib.grt_shard_rays.restype = C.c_longlong
        _lib.grt_release_cache.restype = None
        _lib.grt_del_data.restype = None
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise GrtError("%s failed: %s" % (what or "grtrans_b200 call", ERRORS.get(rc, "error %d" % rc)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    return lib().grt_device_count()


def dfma_peak(device=0):
    t = C.c_double()
    check(lib().grt_dfma_peak(C.c_int(device), C.byref(t)), "grt_dfma_peak")
    return t.value


def set_devices(devices=None):
    """Devices grt_main uses (None / empty: GRT_DEVICES, GRT_NGPU or every visible device)."""
    d = list(devices or [])
    arr = (C.c_int * max(1, len(d)))(*d)
    check(lib().grt_set_devices(arr, C.c_int(len(d))), "grt_set_devices")


def main_info():
    nd = C.c_int(); h2d = C.c_longlong(); d2h = C.c_longlong(); nl = C.c_int()
    check(lib().grt_main_info(C.byref(nd), C.byref(h2d), C.byref(d2h), C.byref(nl)), "grt_main_info")
    return dict(ndevices=nd.value, h2d_bytes=h2d.value, d2h_bytes=d2h.value, launches=nl.value)


def release_cache():
    lib().grt_release_cache()


def shard_rays(i1, i2, rank, nranks):
    """1-based geodesic indices of `rank`'s rays (grt_shard_rays: host logic, works without a GPU)."""
    n = lib().grt_shard_rays(C.c_int(i1), C.c_int(i2), C.c_int(rank), C.c_int(nranks), None)
    idx = np.zeros(n, np.int32)
    lib().grt_shard_rays(C.c_int(i1), C.c_int(i2), C.c_int(rank), C.c_int(nranks), _p(idx))
    return idx


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    check(lib().grt_nccl_unique_id(buf), "grt_nccl_unique_id")
    return buf.raw


def main_view(a):
    """grt_main, returning zero-copy views of the module arrays (valid until del_data()): ivals [ncams, npix, nv]."""
    L = lib()
    check(L.grt_main(C.byref(a)), "grt_main")
    p = C.POINTER(C.c_double)()
    dims = (C.c_int * 3)()
    check(L.grt_get_ivals(C.byref(p), dims), "grt_get_ivals")
    return np.ctypeslib.as_array(p, shape=(dims[2], dims[1], dims[0]))


def del_data():
    lib().grt_del_data()


def maxpts_of(a):
    nup = a.nn[2]
    if a.standard == 1 and nup > 1:
        return nup + int(np.ceil(60. * nup / 400.)) - 2 * int(np.ceil(3. * nup / 400.))
    return nup


def main(a):
    """grt_main + copy-out of the module arrays; returns dict(ivals[nvals,npix,ncams], ab[2,npix], freqs)."""
    L = lib()
    check(L.grt_main(C.byref(a)), "grt_main")
    try:
        p = C.POINTER(C.c_double)()
        dims = (C.c_int * 3)()
        check(L.grt_get_ivals(C.byref(p), dims), "grt_get_ivals")
        nv, npix, ncams = dims[0], dims[1], dims[2]
        iv = np.ctypeslib.as_array(p, shape=(ncams, npix, nv)).copy()
        d2 = (C.c_int * 2)()
        check(L.grt_get_ab(C.byref(p), d2), "grt_get_ab")
        ab = np.ctypeslib.as_array(p, shape=(npix, 2)).copy()
        nf = C.c_int()
        check(L.grt_get_freqs(C.byref(p), C.byref(nf)), "grt_get_freqs")
        fr = np.ctypeslib.as_array(p, shape=(nf.value,)).copy()
    finally:
        L.grt_del_data()
    return dict(ivals=np.ascontiguousarray(iv.transpose(2, 1, 0)), ab=np.ascontiguousarray(ab.T), freqs=fr)


def geokerr_rays(a, i1, i2):
    L = lib()
    n = i2 - i1 + 1
    cap = maxpts_of(a)
    d = {k: np.zeros((n, cap)) for k in ["lambda", "t", "r", "th", "ph", "kt", "kr", "kth", "kph"]}
    d["npts"] = np.zeros(n, np.int32)
    d["tpm"] = np.zeros((n, cap), np.int32)
    d["tpr"] = np.zeros((n, cap), np.int32)
    mp = C.c_int()
    gv = (C.c_double * 4)(*a.gridvals)
    check(L.grt_geokerr_rays(C.c_int(a.standard), C.c_double(a.mumin), C.c_double(a.phi0), C.c_double(a.spin),
                             C.c_double(a.uout), C.c_double(a.rcut), C.c_int(a.nrotype), gv, C.c_int(a.nn[0]),
                             C.c_int(a.nn[1]), C.c_int(a.nn[2]), C.c_int(i1), C.c_int(i2), C.byref(mp), _p(d["npts"]),
                             *[_p(d[k]) for k in ["lambda", "t", "r", "th", "ph", "kt", "kr", "kth", "kph", "tpm", "tpr"]],
                             C.c_int(cap)), "grt_geokerr_rays")
    d["cap"] = cap
    return d


def ray_stage_dump(a, i1, i2):
    L = lib()
    n = i2 - i1 + 1
    cap = maxpts_of(a)
    d = {k: np.zeros((n, cap)) for k in ["lambda", "r", "th", "ph", "s2xi", "c2xi", "ang", "rshift", "cosne", "ncgs",
                                          "ncgsnth", "bcgs", "tcgs"]}
    for k in ["rho", "p", "bmag"]:
        d[k] = np.zeros((n, cap), np.float32)
    d["u"] = np.zeros((n, cap, 4)); d["b"] = np.zeros((n, cap, 4))
    d["npts"] = np.zeros(n, np.int32)
    mp = C.c_int()
    order = ["lambda", "r", "th", "ph", "rho", "p", "bmag", "u", "b", "s2xi", "c2xi", "ang", "rshift", "cosne", "ncgs",
             "ncgsnth", "bcgs", "tcgs"]
    check(L.grt_ray_stage_dump(C.byref(a), C.c_int(i1), C.c_int(i2), C.c_int(cap), C.byref(mp), _p(d["npts"]),
                               *[_p(d[k]) for k in order]), "grt_ray_stage_dump")
    d["cap"] = cap
    return d


def math_probe(which, x):
    """exp_fast (0) / log_ge1 (1) of grt_math.cuh evaluated on the device."""
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    check(lib().grt_math_probe(C.c_int(which), C.c_int(x.size), _p(x), _p(out)), "grt_math_probe")
    return out


def geodebug(a, pixel, path="geodebug.out"):
    """debug = 1 dump of one ray (grtrans_driver.f90:86-110, 341-427) in the layout read_geodebug_file.py parses."""
    check(lib().grt_geodebug(C.byref(a), C.c_int(pixel), path.encode()), "grt_geodebug")
    return path


# field name per line of geodebug.out, in file order (read_geodebug_file.py:16-83; `b` is assigned twice there -- the
# fluid-frame field strength, then the cgs field -- and keeps the second)
GEODEBUG_FIELDS = ["header", "lam", "time", "r", "th", "phi", "kt", "kr", "kth", "kphi", "tpr", "tpm", "rho", "p", "bmag",
                   "ut", "ur", "uth", "uphi", "bt", "br", "bth", "bphi", "n", "t", "b", "g", "incang", "ji", "ki", "jq", "jv",
                   "kq", "kv", "rhoq", "rhov", "ii", "iq", "iu", "iv", "taui", "ju", "ku", "rhou", "s2xi", "c2xi", "ang", "g2",
                   "s2xi2", "c2xi2", "aahat", "aat", "aar", "aath", "aaph", "kht", "khr", "khth", "khph", "bht", "bhr", "bhth",
                   "bhph", "fac"]


def read_geodebug(path="geodebug.out"):
    """The reference's reader (read_geodebug_file.py) as a function: one array per line of the file."""
    with open(path) as f:
        lines = f.readlines()
    return {k: np.array(l.split(), dtype=float) for k, l in zip(GEODEBUG_FIELDS, lines)}


def _bc(*xs):
    shape = np.broadcast(*xs).shape
    return [np.ascontiguousarray(np.broadcast_to(x, shape), np.float64).ravel() for x in xs]


def polsynchth(nu, n, b, t, th):
    arrs = _bc(nu, n, b, t, th)
    e = np.zeros((arrs[0].size, 11))
    check(lib().grt_polsynchth(C.c_int(arrs[0].size), *[_p(x) for x in arrs], _p(e)), "grt_polsynchth")
    return e


def polsynchpl(nu, nnth, b, th, p, gmin, gmax):
    arrs = _bc(nu, nnth, b, th, p, gmin)
    e = np.zeros((arrs[0].size, 11))
    check(lib().grt_polsynchpl(C.c_int(arrs[0].size), *[_p(x) for x in arrs], C.c_double(gmax), _p(e)), "grt_polsynchpl")
    return e


def synchemis(nu, n, b, t):
    arrs = _bc(nu, n, b, t)
    e = np.zeros((arrs[0].size, 11))
    check(lib().grt_synchemis(C.c_int(arrs[0].size), *[_p(x) for x in arrs], _p(e)), "grt_synchemis")
    return e


def rt_integrate(iflag, s, j, K, tau, max_tau=10., hmax=0.1, oatol=1e-8, ortol=1e-6, thin=1e-2, maxsteps=100000):
    """s[nrays,npts], j[nrays,npts,neq], K[nrays,npts,nk], tau[nrays,npts] -> intensity[nrays,npts,neq], nptsout."""
    s = np.ascontiguousarray(s, np.float64); j = np.ascontiguousarray(j, np.float64)
    K = np.ascontiguousarray(K, np.float64); tau = np.ascontiguousarray(tau, np.float64)
    if s.ndim == 1:
        s, j, K, tau = s[None], j[None], K[None], tau[None]
    if j.ndim == 2:
        j = j[..., None]; K = K[..., None]
    nrays, npts = s.shape
    neq = j.shape[2]
    out = np.zeros((nrays, npts, neq))
    nout = np.zeros(nrays, np.int32)
    check(lib().grt_rt_integrate(C.c_int(iflag), C.c_int(neq), C.c_int(nrays), C.c_int(npts), _p(s), _p(j), _p(K), _p(tau),
                                 _p(out), _p(nout), C.c_double(max_tau), C.c_double(hmax), C.c_double(oatol),
                                 C.c_double(ortol), C.c_double(thin), C.c_int(maxsteps)), "grt_rt_integrate")
    return out, nout


class Session(object):
    """Resident device session (camera constants, fluid snapshot and image stay in HBM)."""

    def __init__(self, a, device=0, rank=0, nranks=1):
        self.a = a
        self._h = C.c_void_p()
        check(lib().grt_session_create(C.byref(a), C.c_int(device), C.c_int(rank), C.c_int(nranks), C.byref(self._h)),
              "grt_session_create")
        self.ncams = lib().grt_session_ncams(self._h)
        self.nvals = lib().grt_session_nvals(self._h)   # values per pixel (nvals, + 19 diagnostics when extra = 1)

    def launch(self):
        check(lib().grt_session_launch(self._h), "grt_session_launch")

    def sync(self):
        check(lib().grt_session_sync(self._h), "grt_session_sync")

    def launch_fetch(self):
        """Launch and stream the image tile by tile into the session's pinned host buffer.  Returns (image view
        [ncams, nlocal, nvals] float32 -- complete after sync(), owned by the session -- , bytes)."""
        p = C.POINTER(C.c_float)(); b = C.c_longlong()
        check(lib().grt_session_launch_fetch(self._h, C.byref(p), C.byref(b)), "grt_session_launch_fetch")
        nl = self.nlocal()
        if nl == 0:
            return np.zeros((self.ncams, 0, self.nvals), np.float32), 0
        return np.ctypeslib.as_array(p, shape=(self.ncams, nl, self.nvals)), b.value

    def upload(self):
        b = C.c_longlong()
        check(lib().grt_session_upload(self._h, C.byref(b)), "grt_session_upload")
        return b.value

    def run_timed(self, nsteps):
        ms = C.c_float()
        check(lib().grt_session_run_timed(self._h, C.c_int(nsteps), C.byref(ms)), "grt_session_run_timed")
        return ms.value

    def run_timed_gather(self, nsteps):
        ms = C.c_float()
        check(lib().grt_session_run_timed_gather(self._h, C.c_int(nsteps), C.byref(ms)), "grt_session_run_timed_gather")
        return ms.value

    def last_launches(self):
        return lib().grt_session_last_launches(self._h)

    def tile(self):
        return lib().grt_session_tile(self._h)

    def time(self):
        tot = C.c_float()
        st = (C.c_float * 4)()
        check(lib().grt_session_time(self._h, C.byref(tot), st), "grt_session_time")
        return tot.value, [st[i] for i in range(4)]

    def launch_count(self):
        return lib().grt_session_launch_count(self._h)

    def counts(self):
        nr = C.c_longlong(); npnt = C.c_longlong()
        check(lib().grt_session_counts(self._h, C.byref(nr), C.byref(npnt)), "grt_session_counts")
        return nr.value, npnt.value

    def nlocal(self):
        nr = C.c_longlong()
        check(lib().grt_session_counts(self._h, C.byref(nr), None), "grt_session_counts")
        return nr.value

    def comm_init(self, uid):
        """Collective: join the NCCL communicator of the job (uid: the 128 bytes of nccl_unique_id() from rank 0)."""
        check(lib().grt_session_comm_init(self._h, C.c_char_p(uid)), "grt_session_comm_init")

    def launch_gather(self):
        """Collective: launch; finished tiles go to rank 0's full device image behind the following tiles' kernels."""
        check(lib().grt_session_launch_gather(self._h), "grt_session_launch_gather")

    def fetch_full(self):
        """Rank 0: the gathered image [ncams, npix, nvals] float32."""
        npix = self.a.nn[0] * self.a.nn[1]
        img = np.zeros((self.ncams, npix, self.nvals), np.float32)
        check(lib().grt_session_fetch_full(self._h, _p(img)), "grt_session_fetch_full")
        return img

    def fetch(self):
        nl = self.nlocal()
        img = np.zeros((self.ncams, nl, self.nvals), np.float32)
        idx = np.zeros(nl, np.int32)
        check(lib().grt_session_fetch(self._h, _p(img), _p(idx)), "grt_session_fetch")
        return img, idx

    def close(self):
        if self._h:
            lib().grt_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
