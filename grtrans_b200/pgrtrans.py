"""Drop-in for the reference's f2py module `pgrtrans` (Makefile:185-186; pgrtrans.f90).

`from grtrans_b200.pgrtrans import pgrtrans` gives an object with the module's surface:
    pgrtrans.grtrans_main(<92 positional arguments in the order of pgrtrans.f90:12-19>)
    pgrtrans.ivals  (nvals, npix, ncams)   pgrtrans.ab (2, npix)   pgrtrans.freqs (nfreq)   [None when unallocated]
    pgrtrans.del_pgrtrans_data()
so that `len(np.shape(pgrtrans.ivals)) > 0` is the "data already exist" guard of grtrans_batch.py:399-401,416-420.
The computation runs on the GPU through libgrtrans_b200.so (no CPU fallback)."""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import GrtArgs

_ARG_ORDER = ["standard", "mumin", "mumax", "nmu", "phi0", "spin", "uout", "uin", "rcut", "nrotype", "gridvals", "nn", "i1",
              "i2", "fname", "dt", "nt", "nload", "nmdot", "mdotmin", "mdotmax", "ename", "mbh", "nfreq", "fmin", "fmax",
              "muval", "gmin", "gmax", "p1", "p2", "jetalpha", "stype", "use_geokerr", "nvals", "iname", "cflag", "extra",
              "debug", "outfile", "fdfile", "fhfile", "fgfile", "fsim", "fnt", "findf", "fnfiles", "fjonfix", "fnw",
              "fnfreq_tab", "fnr", "foffset", "fdindf", "fmagcrit", "frspot", "fr0spot", "fn0spot", "ftscl", "frscl",
              "fwmin", "fwmax", "ffmin", "ffmax", "frmax", "fsigt", "ffcol", "fmdot", "fnscl", "fnnthscl", "fnnthp",
              "fbeta", "fbl06", "fnp", "ftp", "frin", "frout", "fthin", "fthout", "fphiin", "fphiout", "fscalefac",
              "sigcut", "epcoefindx", "epotherargs", "nepotherargs"]


class _PgrtransModule(object):
    def __init__(self):
        self.ivals = None
        self.ab = None
        self.freqs = None

    def grtrans_main(self, *args):
        if len(args) != len(_ARG_ORDER):
            raise TypeError("grtrans_main takes %d positional arguments (%d given)" % (len(_ARG_ORDER), len(args)))
        if self.ivals is not None:
            raise _lib.GrtError("pgrtrans data already exist (call del_pgrtrans_data first)")
        a = GrtArgs()
        for name, v in zip(_ARG_ORDER, args):
            if name in ("gridvals", "nn", "epcoefindx"):
                arr = getattr(a, name)
                for i in range(len(arr)):
                    arr[i] = v[i]
            elif name == "epotherargs":
                for i in range(min(8, len(v))):
                    a.epotherargs[i] = float(v[i])
            elif name == "use_geokerr":
                a.use_geokerr = 1 if (v is True or v == 1 or str(v).upper().startswith('T')) else 0
            elif isinstance(v, str):
                setattr(a, name, v.strip().encode())
            elif isinstance(getattr(a, name), int):
                setattr(a, name, int(v))
            else:
                setattr(a, name, v)
        L = _lib.lib()
        _lib.check(L.grt_main(C.byref(a)), "grt_main")
        p = C.POINTER(C.c_double)()
        dims = (C.c_int * 3)()
        _lib.check(L.grt_get_ivals(C.byref(p), dims), "grt_get_ivals")
        nv, npix, ncams = dims[0], dims[1], dims[2]
        # Fortran ivals(nv,npix,ncams) == C-order [cam][pix][nv]
        self.ivals = np.ctypeslib.as_array(p, shape=(ncams, npix, nv)).copy().transpose(2, 1, 0)
        d2 = (C.c_int * 2)()
        _lib.check(L.grt_get_ab(C.byref(p), d2), "grt_get_ab")
        self.ab = np.ctypeslib.as_array(p, shape=(npix, 2)).copy().T
        nf = C.c_int()
        _lib.check(L.grt_get_freqs(C.byref(p), C.byref(nf)), "grt_get_freqs")
        self.freqs = np.ctypeslib.as_array(p, shape=(nf.value,)).copy()

    def del_pgrtrans_data(self):
        _lib.lib().grt_del_data()
        self.ivals = None
        self.ab = None
        self.freqs = None


pgrtrans = _PgrtransModule()
