#!/usr/bin/env python
"""Small runs of every fluid model / integrator for compute-sanitizer (memcheck / initcheck) on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import cases  # noqa: E402
from grtrans_b200 import _lib, args_from_kwargs  # noqa: E402

RUNS = [("thindisk", dict(n=12, nfreq=2)), ("thindisk", dict(n=10, nfreq=2, ename='BBPOL', nvals=4)),
        ("sphacc", dict(n=10, nfreq=2, iname='delo')), ("sphacc", dict(n=8, nfreq=2, iname='lsoda')),
        ("sphacc", dict(n=8, nfreq=2, iname='formal', nvals=1)), ("ffjet", dict(n=10, nfreq=2, iname='formal')),
        ("harm", dict(n=10, nup=100, nfreq=1)), ("harm3d", dict(n=10, nup=100, nfreq=2, nmdot=2, mdotmin=1e19, mdotmax=1e21)),
        ("harm3d", dict(n=8, nup=100, nfreq=2, ename='HYBRIDTHPL', jetalpha=0.05, p1=3.2, gmin=5., magnetized=True)),
        ("iharm", dict(n=10, nup=100, nfreq=2, mmks=1)), ("iharm", dict(n=8, nup=100, nfreq=2, mmks=0, gmin=-1.)),
        ("koral", dict(n=10, nup=100, nfreq=2, threed=True)), ("koral", dict(n=8, nup=100, nfreq=2, threed=False)),
        # time axis: dump slots, arrival-time offsets, movie frames
        ("series", dict(kind='harm2d', n=8, fnt=3, nload=3, nt=2, dt=10.)), ("series", dict(kind='harm3d', n=8, fnt=2, nt=2, dt=20.)),
        # extra = 1 (history scratch + forward pass), the BREMS family, stype = 'tail'
        ("harm3d", dict(n=8, nup=100, nfreq=2, extra=1)), ("harm3d", dict(n=8, nup=100, nfreq=2, extra=1, iname='lsoda')),
        ("sphacc", dict(n=8, nfreq=2, extra=1, nvals=1)), ("harm3d", dict(n=8, nup=100, nfreq=2, ename='SYNCHTHBREMS', nvals=1)),
        ("harm3d", dict(n=8, nup=100, nfreq=2, ename='POLSYNCHPL', stype='tail', p1=3., p2=3., jetalpha=.1)),
        ("harm3d", dict(n=8, nup=100, nfreq=2, ename='POLSYNCHSYMTH')),
        ("harm3d", dict(n=8, nup=100, nfreq=2, ename='MAXJUTT', epotherargs=[2., .6, .3, .1])),
        ("harm3d", dict(n=8, nup=100, nfreq=2, ename='MAXCOMP', epotherargs=[2., 2., .6, .3, .1], extra=1)),
        ("thindisk", dict(n=10, nfreq=2, extra=1)),
        # nfreq a multiple of 32: k_transfer's cp.async record ring in shared memory (16- and 24-double records, DELO / formal)
        ("harm3d", dict(n=8, nup=100, nfreq=32)), ("harm3d", dict(n=8, nup=100, nfreq=64, iname='formal')),
        ("harm3d", dict(n=8, nup=100, nfreq=32, ename='HYBRIDTHPL', jetalpha=0.05, p1=3.2, gmin=5., magnetized=True)),
        ("sphacc", dict(n=8, nfreq=32, nmdot=2, mdotmin=1e19, mdotmax=1e21)),
        # KORALNTH: binned non-thermal electrons (76-double records, the nth array)
        ("koralnth", dict(n=8, nup=100, nfreq=2, ename='BINS')), ("koralnth", dict(n=8, nup=100, nfreq=2, ename='HYBRIDTHBINS', extra=1))]
os.environ["GRT_TILE"] = "64"   # several tiles per run
for cfg, over in RUNS:
    kw = getattr(cases, "cfg_" + cfg)(**over)
    out = _lib.main(args_from_kwargs(**kw))
    s = _lib.Session(args_from_kwargs(**kw))
    v, nb = s.launch_fetch(); s.sync()
    print(cfg, over.get('iname'), over.get('ename'), out['ivals'].shape, float(np.nansum(out['ivals'][0])), nb, flush=True)
    s.close()
a = args_from_kwargs(**cases.cfg_ffjet(n=8, nup=100))
g = _lib.geokerr_rays(a, 1, 64)
d = _lib.ray_stage_dump(a, 1, 64)
_lib.geodebug(args_from_kwargs(**cases.cfg_harm3d(n=8, nup=100, nfreq=1, iname='lsoda')), 30, "/tmp/geodebug_san.out")
x = _lib.math_probe(0, np.linspace(-700., 700., 1000))
print("unit surfaces ok", g['npts'].sum(), d['npts'].sum(), len(_lib.read_geodebug("/tmp/geodebug_san.out")), float(x[500]))
