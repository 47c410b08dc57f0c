"""ctypes mirror of include/grtrans_b200.h (struct grt_args) and the keyword -> argument mapping of the
reference's Python front-end.

`GrtArgs` has exactly the fields of `pgrtrans.grtrans_main`'s argument list (pgrtrans.f90:12-58).
`inputs_defaults()` / `args_from_kwargs()` reproduce `grtrans_batch.py:grtrans_inputs` (defaults
:24-150, i1/i2 clamping :155-164, nweights/delta/nep :165-170) and the positional mapping used by
`grtrans.run_pgrtrans` (grtrans_batch.py:406), e.g. ftscl<-ntscl, frscl<-nrscl, fmdot<-tmdot, frin<-srin.
"""
import ctypes as C

STRLEN, PATHLEN = 100, 300


class GrtArgs(C.Structure):
    _fields_ = [
        ("standard", C.c_int), ("mumin", C.c_double), ("mumax", C.c_double), ("nmu", C.c_int),
        ("phi0", C.c_double), ("spin", C.c_double), ("uout", C.c_double), ("uin", C.c_double),
        ("rcut", C.c_double), ("nrotype", C.c_int), ("gridvals", C.c_double * 4), ("nn", C.c_int * 3),
        ("i1", C.c_int), ("i2", C.c_int), ("fname", C.c_char * STRLEN), ("dt", C.c_double),
        ("nt", C.c_int), ("nload", C.c_int), ("nmdot", C.c_int), ("mdotmin", C.c_double),
        ("mdotmax", C.c_double), ("ename", C.c_char * STRLEN), ("mbh", C.c_double), ("nfreq", C.c_int),
        ("fmin", C.c_double), ("fmax", C.c_double), ("muval", C.c_double), ("gmin", C.c_double),
        ("gmax", C.c_double), ("p1", C.c_double), ("p2", C.c_double), ("jetalpha", C.c_double),
        ("stype", C.c_char * STRLEN), ("use_geokerr", C.c_int), ("nvals", C.c_int),
        ("iname", C.c_char * STRLEN), ("cflag", C.c_int), ("extra", C.c_int), ("debug", C.c_int),
        ("outfile", C.c_char * PATHLEN), ("fdfile", C.c_char * PATHLEN), ("fhfile", C.c_char * PATHLEN),
        ("fgfile", C.c_char * PATHLEN), ("fsim", C.c_char * STRLEN),
        ("fnt", C.c_int), ("findf", C.c_int), ("fnfiles", C.c_int), ("fjonfix", C.c_int), ("fnw", C.c_int),
        ("fnfreq_tab", C.c_int), ("fnr", C.c_int), ("foffset", C.c_int), ("fdindf", C.c_int),
        ("fmagcrit", C.c_int),
        ("frspot", C.c_double), ("fr0spot", C.c_double), ("fn0spot", C.c_double), ("ftscl", C.c_double),
        ("frscl", C.c_double), ("fwmin", C.c_double), ("fwmax", C.c_double), ("ffmin", C.c_double),
        ("ffmax", C.c_double), ("frmax", C.c_double), ("fsigt", C.c_double), ("ffcol", C.c_double),
        ("fmdot", C.c_double), ("fnscl", C.c_double), ("fnnthscl", C.c_double), ("fnnthp", C.c_double),
        ("fbeta", C.c_double), ("fbl06", C.c_int),
        ("fnp", C.c_double), ("ftp", C.c_double), ("frin", C.c_double), ("frout", C.c_double),
        ("fthin", C.c_double), ("fthout", C.c_double), ("fphiin", C.c_double), ("fphiout", C.c_double),
        ("fscalefac", C.c_double), ("sigcut", C.c_double),
        ("epcoefindx", C.c_int * 7), ("nepotherargs", C.c_int), ("epotherargs", C.c_double * 8),
    ]


def inputs_defaults():
    """grtrans_batch.py:24-150 (grtrans_inputs.init)."""
    d = dict(
        cflag=1, standard=2, mumin=.1, mumax=1., phi0=-0.5, nmu=10, spin=0.998, uout=0.04, uin=1., rcut=1.,
        nrotype=2, gridvals=[-25., 25., -25., 25.], nn=[100, 100, 1], i1=-1, i2=-1, fname='THINDISK', dt=5.,
        nt=1, nload=1, ename='BB', mbh=10., nmdot=1, mdotmin=1.5e15, mdotmax=1.5e15, nfreq=1, fmin=1.e17,
        fmax=3.e19, muval=1. / 4., sigcut=1.e10, gmin=100, gmax=1e5, p1=3.5, p2=3.5, epotherargs=[1., 1.],
        epcoefindx=[1, 1, 1, 1, 1, 1, 1], jetalpha=0.02, stype='const', use_geokerr='T', nvals=1,
        iname='lsoda', tmdot=0.1, pnw=500, pwmin=1.e-4, pwmax=1.e4, pnfreq_tab=100, prmax=1.e4, pnr=500,
        psigt=0.4, pfcol=1.7, jdfile='m87bl09rfp10xi5a998fluidvars.bin', ndfile='phatdiskm8st25.bin',
        ntscl=1., nrscl=6., snscl=3e7, snnthscl=8e4, snnthp=2.9, sbeta=10., snp=0., snt=0., sbl06=0,
        srin=-1., srout=1e8, sthin=-10., sthout=10., sphiin=0., sphiout=1e4, hrspot=1.5, hr0spot=6.0,
        hn0spot=1e4, ofile='grtrans.out', outfile='', ifile='inputs.in', hhfile='dump040', hdfile='dump',
        hnt=1, hindf=40, tdfile='fieldline', tgfile='dump0000rr2.bin', tnt=1, tnfiles=1, tindf=5206,
        tjonfix=1, toff=0, tsim='thickdiskrr2', tdindf=1, tmagcrit=0, extra=0, debug=0,
        fdfile='fieldline', fgfile='dump0000rr2.bin', fhfile='dump040', fnt=1, fnfiles=1, findf=5206,
        fjonfix=1, foffset=0, fsim='thickdiskrr2', fdindf=1, fmagcrit=0, fscalefac=1., fnw=500, fwmin=1e-4,
        fwmax=1e4, fnfreq_tab=100, ffmin=3.33e16, ffmax=9e19, frmax=1e4, fnr=500, fsigt=0.4, ffcol=1.7,
        frspot=1.5, fr0spot=6.0, fn0spot=1e4, ftscl=1., frscl=6., fmdot=0.1)
    d['pfmin'] = d['fmin'] / 3.
    d['pfmax'] = d['fmax'] * 3.
    return d


class Inputs(object):
    """Drop-in for grtrans_batch.grtrans_inputs (same attributes, same clamping)."""

    def __init__(self, **kwargs):
        self.__dict__.update(inputs_defaults())
        self.__dict__.update(kwargs)
        npix = self.nn[0] * self.nn[1]
        if self.i1 < 0:
            self.i1 = 1
        if self.i2 < 0:
            self.i2 = npix
        if self.i1 > npix - 1:
            print('fixing input i1 out of bounds: ', self.i1, npix)
            self.i1 = npix - 1
        if self.i2 > npix:
            print('fixing input i2 out of bounds: ', self.i2, npix)
            self.i2 = npix
        if self.ename == 'MAXCOMP':
            self.nweights = len(self.epotherargs) - 2
        else:
            self.nweights = len(self.epotherargs) - 1
        self.nep = len(self.epotherargs)
        self.delta = self.epotherargs[0]


def _s(x):
    return x.encode() if isinstance(x, str) else x


def args_from_inputs(inp, file_mode=False):
    """Build the C struct in the positional order of grtrans_batch.py:406 (pgrtrans mode).

    file_mode=True applies the extra per-model copies that `grtrans_inputs.write` does before it
    writes inputs.in (grtrans_batch.py:219-326): FFJET takes fdfile<-jdfile, HARM takes
    fdfile<-hdfile, fhfile<-hhfile, fnt<-hnt, findf<-hindf."""
    a = GrtArgs()
    a.standard = int(inp.standard); a.mumin = inp.mumin; a.mumax = inp.mumax; a.nmu = int(inp.nmu)
    a.phi0 = inp.phi0; a.spin = inp.spin; a.uout = inp.uout; a.uin = inp.uin; a.rcut = inp.rcut
    a.nrotype = int(inp.nrotype)
    for i in range(4):
        a.gridvals[i] = float(inp.gridvals[i])
    for i in range(3):
        a.nn[i] = int(inp.nn[i])
    a.i1 = int(inp.i1); a.i2 = int(inp.i2)
    a.fname = _s(inp.fname); a.dt = inp.dt; a.nt = int(inp.nt); a.nload = int(inp.nload)
    a.nmdot = int(inp.nmdot); a.mdotmin = inp.mdotmin; a.mdotmax = inp.mdotmax
    a.ename = _s(inp.ename); a.mbh = inp.mbh; a.nfreq = int(inp.nfreq); a.fmin = inp.fmin; a.fmax = inp.fmax
    a.muval = inp.muval; a.gmin = float(inp.gmin); a.gmax = float(inp.gmax); a.p1 = inp.p1; a.p2 = inp.p2
    a.jetalpha = inp.jetalpha; a.stype = _s(inp.stype)
    ug = inp.use_geokerr
    a.use_geokerr = 1 if (ug is True or str(ug).upper().startswith('T') or ug == 1) else 0
    a.nvals = int(inp.nvals); a.iname = _s(inp.iname); a.cflag = int(inp.cflag); a.extra = int(inp.extra)
    a.debug = int(inp.debug); a.outfile = _s(inp.outfile)
    fdfile, fhfile, fgfile, fnt, findf = inp.fdfile, inp.fhfile, inp.fgfile, inp.fnt, inp.findf
    if file_mode:
        fdfile, fgfile, fhfile, fnt, findf = inp.tdfile, inp.tgfile, inp.hhfile, inp.tnt, inp.tindf
        if inp.fname in ('HARM', 'HARM3D'):
            fdfile, fhfile, fnt, findf = inp.hdfile, inp.hhfile, inp.hnt, inp.hindf
            if inp.fname == 'HARM3D' and hasattr(inp, 'hgfile'):
                fgfile = inp.hgfile
        elif inp.fname == 'FFJET':
            fdfile = inp.jdfile
    a.fdfile = _s(fdfile); a.fhfile = _s(fhfile); a.fgfile = _s(fgfile); a.fsim = _s(inp.fsim)
    a.fnt = int(fnt); a.findf = int(findf); a.fnfiles = int(inp.fnfiles); a.fjonfix = int(inp.fjonfix)
    a.fnw = int(inp.pnw); a.fnfreq_tab = int(inp.pnfreq_tab); a.fnr = int(inp.pnr); a.foffset = int(inp.foffset)
    a.fdindf = int(inp.fdindf); a.fmagcrit = int(inp.fmagcrit)
    a.frspot = inp.hrspot; a.fr0spot = inp.hr0spot; a.fn0spot = inp.hn0spot; a.ftscl = inp.ntscl
    a.frscl = inp.nrscl; a.fwmin = inp.pwmin; a.fwmax = inp.pwmax; a.ffmin = inp.pfmin; a.ffmax = inp.pfmax
    a.frmax = inp.prmax; a.fsigt = inp.psigt; a.ffcol = inp.pfcol; a.fmdot = inp.tmdot; a.fnscl = inp.snscl
    a.fnnthscl = inp.snnthscl; a.fnnthp = inp.snnthp; a.fbeta = inp.sbeta; a.fbl06 = int(inp.sbl06)
    a.fnp = inp.snp; a.ftp = inp.snt; a.frin = inp.srin; a.frout = inp.srout; a.fthin = inp.sthin
    a.fthout = inp.sthout; a.fphiin = inp.sphiin; a.fphiout = inp.sphiout; a.fscalefac = inp.fscalefac
    a.sigcut = inp.sigcut
    for i in range(7):
        a.epcoefindx[i] = int(inp.epcoefindx[i])
    a.nepotherargs = int(inp.nep)
    for i in range(min(8, len(inp.epotherargs))):
        a.epotherargs[i] = float(inp.epotherargs[i])
    return a


def args_from_kwargs(file_mode=False, **kwargs):
    return args_from_inputs(Inputs(**kwargs), file_mode=file_mode)
