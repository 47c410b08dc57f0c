"""Drop-in for the reference's grtrans_batch.py (class grtrans / grtrans_inputs): same methods, same attribute names,
same array orderings (file mode ivals[npix,nvals,ncams], module mode ivals[nvals,npix,ncams]; A19 in SURVEY.md).

`run_grtrans` calls the C ABI `grt_program('files.in')` in-process where the reference shells out to `./grtrans`
(grtrans_batch.py:357); `run_pgrtrans` goes through the pgrtrans shim with the reference's positional argument list
(grtrans_batch.py:406).  Plotting helpers (matplotlib) are out of scope."""
from __future__ import print_function

import ctypes as C
import os

import numpy as np

from . import _lib, fitsio as fits
from . import namelist as nm
from ._abi import Inputs
from .pgrtrans import pgrtrans

pcG = 6.67259e-8
pcc2 = 8.98755179e20
msun = 1.99e33


class grtrans_inputs(Inputs):
    """grtrans_batch.py:22-330"""

    def write(self, fname):
        names = ['geodata', 'fluiddata', 'emisdata', 'general', 'harm', 'analytic']
        args = ['standard', 'mumin', 'mumax', 'nmu', 'phi0', 'spin', 'uout', 'uin', 'rcut', 'nrotype', 'gridvals', 'nn', 'i1',
                'i2', 'extra', 'debug']
        nargs = [len(args)]
        args1 = ['fname', 'dt', 'nt', 'nload', 'nmdot', 'mdotmin', 'mdotmax', 'sigcut']
        args2 = ['ename', 'mbh', 'nfreq', 'fmin', 'fmax', 'muval', 'gmin', 'gmax', 'p1', 'p2', 'jetalpha', 'stype', 'delta',
                 'nweights', 'coefindx']
        args3 = ['use_geokerr', 'nvals', 'iname', 'cflag']
        args4 = ['fdfile', 'fgfile', 'fhfile', 'fnt', 'fnfiles', 'findf', 'fjonfix', 'foffset', 'fsim', 'fdindf', 'fmagcrit',
                 'fscalefac']
        args5 = ['fnw', 'fwmin', 'fwmax', 'fnfreq_tab', 'ffmin', 'ffmax', 'frmax', 'fnr', 'fsigt', 'ffcol', 'frspot', 'fr0spot',
                 'fn0spot', 'ftscl', 'frscl', 'fmdot', 'fnscl', 'fnnthscl', 'fnnthp', 'fbeta', 'fbl06', 'fnp', 'ftp', 'frin',
                 'frout', 'fthin', 'fthout', 'fphiin', 'fphiout']
        for extra in (args1, args2, args3, args4, args5):
            nargs.append(len(extra))
            args.extend(extra)
        q = lambda s: "'" + s + "'"  # noqa: E731
        cs = lambda v: ','.join(str(x) for x in v)  # noqa: E731
        vals = [self.standard, self.mumin, self.mumax, self.nmu, self.phi0, self.spin, self.uout, self.uin, self.rcut,
                self.nrotype, cs(self.gridvals), cs(self.nn), self.i1, self.i2, self.extra, self.debug]
        vals.extend([q(self.fname), self.dt, self.nt, self.nload, self.nmdot, self.mdotmin, self.mdotmax, self.sigcut])
        vals.extend([q(self.ename), self.mbh, self.nfreq, self.fmin, self.fmax, self.muval, self.gmin, self.gmax, self.p1,
                     self.p2, self.jetalpha, q(self.stype), self.delta, self.nweights, cs(self.epcoefindx)])
        vals.extend([self.use_geokerr, self.nvals, q(self.iname), self.cflag])
        # per-model copies the reference makes before writing (grtrans_batch.py:219-326)
        kw_f = set(getattr(self, '_user_keys', ()))
        fdfile, fgfile, fhfile = self.tdfile, self.tgfile, self.hhfile
        fnt, fnfiles, findf = self.tnt, self.tnfiles, self.tindf
        if self.fname in ('HARM', 'HARM3D', 'KORAL', 'KORAL3D', 'KORALNTH'):
            fdfile, fhfile, fnt, findf = self.hdfile, self.hhfile, self.hnt, self.hindf
        elif self.fname == 'FFJET':
            fdfile = self.jdfile
        # explicit f* keywords win (convenience for scripts that already use the pgrtrans names)
        if 'fdfile' in kw_f: fdfile = self.fdfile  # noqa: E701
        if 'fhfile' in kw_f: fhfile = self.fhfile  # noqa: E701
        if 'fgfile' in kw_f: fgfile = self.fgfile  # noqa: E701
        if 'findf' in kw_f: findf = self.findf  # noqa: E701
        if 'fnt' in kw_f: fnt = self.fnt  # noqa: E701
        vals.extend([q(fdfile), q(fgfile), q(fhfile), fnt, fnfiles, findf, self.tjonfix, self.toff, q(self.tsim), self.tdindf,
                     self.tmagcrit, self.fscalefac])
        vals.extend([self.pnw, self.pwmin, self.pwmax, self.pnfreq_tab, self.pfmin, self.pfmax, self.prmax, self.pnr,
                     self.psigt, self.pfcol, self.hrspot, self.hr0spot, self.hn0spot, self.ntscl, self.nrscl, self.tmdot,
                     self.snscl, self.snnthscl, self.snnthp, self.sbeta, self.sbl06, self.snp, self.snt, self.srin, self.srout,
                     self.sthin, self.sthout, self.sphiin, self.sphiout])
        nm.write_namelist(fname, names, args, vals, nargs)


class grtrans(object):
    """run & analyze grtrans on the GPU (interface of grtrans_batch.py:332-545)"""

    ofile = 'grtrans.out'
    ifile = 'inputs.in'

    def run_grtrans(self, **kwargs):
        if os.path.exists(self.ofile):
            os.remove(self.ofile)
        _lib.check(_lib.lib().grt_program(b'files.in'), "grt_program")

    def write_grtrans_inputs(self, ifile, oname='grtrans.out', **kwargs):
        self.inputs = grtrans_inputs(**kwargs)
        self.inputs._user_keys = tuple(kwargs.keys())
        self.inputs.write(ifile)
        self.set_grtrans_input_file(ifile, oname)

    def set_grtrans_input_file(self, fname, oname):
        nm.write_namelist('files.in', ['files'], ['ifile', 'ofile'], ['"' + fname + '"', '"' + oname + '"'], [2])
        self.ifile = fname
        self.ofile = oname

    def run_pgrtrans(self, **kwargs):
        if len(np.shape(pgrtrans.ivals)) > 0:
            print('ERROR in run_pgrtrans: data already exist!')
            return
        self.inputs = grtrans_inputs(**kwargs)
        i = self.inputs
        pgrtrans.grtrans_main(i.standard, i.mumin, i.mumax, i.nmu, i.phi0, i.spin, i.uout, i.uin, i.rcut, i.nrotype, i.gridvals,
                              i.nn, i.i1, i.i2, i.fname, i.dt, i.nt, i.nload, i.nmdot, i.mdotmin, i.mdotmax, i.ename, i.mbh,
                              i.nfreq, i.fmin, i.fmax, i.muval, i.gmin, i.gmax, i.p1, i.p2, i.jetalpha, i.stype, i.use_geokerr,
                              i.nvals, i.iname, i.cflag, i.extra, i.debug, i.outfile, i.fdfile, i.fhfile, i.fgfile, i.fsim,
                              i.fnt, i.findf, i.fnfiles, i.fjonfix, i.pnw, i.pnfreq_tab, i.pnr, i.foffset, i.fdindf,
                              i.fmagcrit, i.hrspot, i.hr0spot, i.hn0spot, i.ntscl, i.nrscl, i.pwmin, i.pwmax, i.pfmin, i.pfmax,
                              i.prmax, i.psigt, i.pfcol, i.tmdot, i.snscl, i.snnthscl, i.snnthp, i.sbeta, i.sbl06, i.snp,
                              i.snt, i.srin, i.srout, i.sthin, i.sthout, i.sphiin, i.sphiout, i.fscalefac, i.sigcut,
                              i.epcoefindx, i.epotherargs, i.nep)
        self.ivals = pgrtrans.ivals.copy()
        self.ab = pgrtrans.ab.copy()
        self.nu = pgrtrans.freqs.copy()
        self.nx = int(i.nn[0])
        self.ny = int(i.nn[1])
        self.nvals = int(i.nvals)
        self.del_pgrtrans_data()

    def del_pgrtrans_data(self):
        if len(np.shape(pgrtrans.ivals)) == 0:
            print('ERROR in del_pgrtrans_data: no data to delete!')
            return
        pgrtrans.del_pgrtrans_data()

    def read_grtrans_output(self, bin=0):
        if bin == 0:
            hdu = fits.open(self.ofile)
            n = len(hdu)
            ab = np.reshape(hdu[0].data, (int(len(hdu[0].data) / 2), 2))
            nu = np.empty(n - 1); nx = np.empty(n - 1); ny = np.empty(n - 1)
            for i in range(n - 1):
                nx[i] = hdu[i + 1].header.get(6)
                ny[i] = hdu[i + 1].header.get(7)
                nu[i] = hdu[i + 1].header.get(8)
            nx = int(nx[0]); ny = int(ny[0])
            nvals = int(len(hdu[1].data) / nx / ny)
            ivals = np.empty((int(nx * ny), nvals, n - 1))
            for i in range(n - 1):
                ivals[:, :, i] = np.reshape(hdu[i + 1].data, (int(nx * ny), int(nvals)))
        else:
            # Fortran sequential unformatted (camera.f90:322-341), all cameras
            with open(self.ofile, 'rb') as f:
                raw = f.read()
            pos = 0
            ims = []; nus = []

            def rec(dtype, count):
                nonlocal pos
                nb = int(np.frombuffer(raw, 'i4', 1, pos)[0]); pos += 4
                out = np.frombuffer(raw, dtype, count, pos); pos += nb + 4
                return out

            while pos < len(raw):
                nx, ny, nvals = [int(v) for v in rec('i4', 3)]
                nkey = int(rec('i4', 1)[0])
                keys = [float(rec('f4', 1)[0]) for _ in range(nkey)]
                ab = np.reshape(rec('f4', 2 * nx * ny), (nx * ny, 2))
                ims.append(np.reshape(rec('f4', nvals * nx * ny), (nx * ny, nvals)))
                nus.append(keys[2] if nkey >= 3 else 0.)
            ivals = np.stack(ims, 0)
            nu = np.array(nus)
            n = len(ims) + 1
        self.ab = ab
        self.ivals = ivals
        self.nu = nu
        self.nx = nx
        self.ny = ny
        self.nvals = nvals
        if bin == 0:
            self.calc_spec(n - 1)

    def calc_freqs(self, n):
        fmin, fmax = self.inputs.fmin, self.inputs.fmax
        self.freqs = fmin * np.exp(np.arange(n) * np.log(fmax / fmin) / (n - 1)) if n > 1 else np.array([fmin])

    def calc_spec(self, n):
        if self.ny != 1:
            da = self.ab[self.nx, 0] - self.ab[0, 0]
            db = self.ab[1, 1] - self.ab[0, 1]
            spec = np.sum(self.ivals, 0) * da * db
            if self.nvals >= 4:
                self.lp = np.sqrt(spec[1, :] ** 2. + spec[2, :] ** 2.) / spec[0, :]
                self.cp = spec[3, :] / spec[0, :]
                self.lpf = np.sum(np.sqrt(self.ivals[:, 1, :] ** 2. + self.ivals[:, 2, :] ** 2.), 0) * da * db / spec[0, :]
                self.cpf = np.sum(np.abs(self.ivals[:, 3, :]), 0) * da * db / spec[0, :]
        else:
            da = self.ab[1, 0] - self.ab[0, 0]
            db = 0.
            spec = np.empty([n, int(self.inputs.nvals)])
            for i in range(n):
                for j in range(int(self.inputs.nvals)):
                    spec[i, j] = np.sum(self.ivals[:, j, i] * self.ab[:, 0], 0) * da * 2. * 3.14
        self.spec = spec
        self.da = da; self.db = db

    def calc_spec_pgrtrans(self, n):
        if self.ny != 1:
            da = self.ab[0, self.nx] - self.ab[0, 0]
            db = self.ab[1, 1] - self.ab[1, 0]
            spec = np.sum(self.ivals, 1) * da * db
            if self.nvals >= 4:
                self.lp = np.sqrt(spec[1, :] ** 2. + spec[2, :] ** 2.) / spec[0, :]
                self.cp = spec[3, :] / spec[0, :]
                self.lpf = np.sum(np.sqrt(self.ivals[1] ** 2. + self.ivals[2] ** 2.), 0) * da * db / spec[0, :]
                self.cpf = np.sum(np.abs(self.ivals[3]), 0) * da * db / spec[0, :]
        else:
            da = self.ab[0, 1] - self.ab[0, 0]
            db = 0.
            spec = np.empty([n, int(self.inputs.nvals)])
            for i in range(n):
                for j in range(int(self.inputs.nvals)):
                    spec[i, j] = np.sum(self.ivals[j, :, i] * self.ab[0, :], 0) * da * 2. * np.arccos(-1.)
        self.spec = spec

    def convert_to_lum(self):
        lbh = pcG * msun * self.inputs.mbh / pcc2
        fac = (4 * np.pi * lbh ** 2)
        da = self.ab[self.nx, 0] - self.ab[0, 0]
        db = self.ab[1, 1] - self.ab[0, 1]
        self.spec *= fac
        self.ivals *= fac * da * db

    def convert_to_Jy(self, D):
        lbh = pcG * msun * self.inputs.mbh / pcc2
        fac = (lbh ** 2 / D ** 2.) * 1e23
        da = self.ab[self.nx, 0] - self.ab[0, 0]
        db = self.ab[1, 1] - self.ab[0, 1]
        self.ivals *= fac * da * db
        self.spec *= fac
