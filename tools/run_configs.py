#!/usr/bin/env python
"""BASELINE.json configs 1-4 at their full sizes on the GPU box: device time, CPU-oracle time and parity per config.

    python tools/run_configs.py [--out profiles/rNN_configs.json]

The oracle (CPU restatement of the reference) is the checker and the CPU baseline here; it is not used by the product.
cfg4 (HARM 150x150x400) runs the oracle on a sub-range of rows to stay within a minute of CPU time."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import cases  # noqa: E402
import orc  # noqa: E402
from grtrans_b200 import _lib, args_from_kwargs  # noqa: E402

CONFIGS = [
    ("cfg1 THINDISK BB nn=[100,100,1] nfreq=25 nvals=1",
     cases.cfg_thindisk(n=100, nfreq=25, spin=.998, mumin=.5, mumax=.5, gridvals=[-25., 25., -25., 25.]), None),
    ("cfg2 SPHACC POLSYNCHTH nn=[100,100,100] nfreq=25 nvals=4 (delo)",
     cases.cfg_sphacc(n=100, nup=100, nfreq=25, fmin=2.41e10, fmax=6.31e14, iname='delo'), None),
    ("cfg2 SPHACC POLSYNCHTH nn=[100,100,100] nfreq=25 nvals=4 (lsoda -> RK5(4))",
     cases.cfg_sphacc(n=100, nup=100, nfreq=25, fmin=2.41e10, fmax=6.31e14, iname='lsoda'), None),
    ("cfg3 FFJET POLSYNCHPL nn=[100,100,100] nfreq=25 nvals=4 (delo)",
     cases.cfg_ffjet(n=100, nup=100, nfreq=25, iname='delo'), None),
    ("cfg3 FFJET POLSYNCHPL nn=[100,100,100] nfreq=25 nvals=4 (formal)",
     cases.cfg_ffjet(n=100, nup=100, nfreq=25, iname='formal'), None),
    ("cfg3 FFJET POLSYNCHPL nn=[100,100,100] nfreq=25 nvals=4 (lsoda -> RK5(4))",
     cases.cfg_ffjet(n=100, nup=100, nfreq=25, iname='lsoda'), None),
    ("cfg4 HARM(synthetic dump) POLSYNCHTH nn=[150,150,400] nfreq=1 nvals=4 delo",
     cases.cfg_harm(n=150, nup=400, nfreq=1, fmin=2.3e11, fmax=2.3e11, iname='delo'), (9001, 13500)),
    ("next-row IHARM (synthetic FMKS dump, R-beta electrons) POLSYNCHTH nn=[128,128,400] nfreq=8 nvals=4 delo",
     cases.cfg_iharm(n=128, nup=400, nfreq=8, mmks=1, fmin=1e11, fmax=1e13), (8001, 9280)),
    ("next-row KORAL3D (synthetic MKS3 dump, T_e from the dump) POLSYNCHTH nn=[128,128,400] nfreq=8 nvals=4 delo",
     cases.cfg_koral(n=128, nup=400, nfreq=8, threed=True), (8001, 9280)),
    ("next-row time axis: HARM series, 3 dumps resident + t0 offsets, 3 movie frames, nn=[128,128,400] nfreq=2 nvals=4 delo",
     # (no sub-range for the oracle here: the t0 offsets of an nload > 1 run are normalised over the pixels i1..i2)
     cases.cfg_series('harm2d', n=128, nup=400, nfreq=2, fnt=3, nload=3, nt=3, dt=10., fmin=1e12, fmax=3e12), None),
    ("next-row extra=1: HARM3D POLSYNCHTH + 19 diagnostics per pixel, nn=[128,128,400] nfreq=4 nvals=4 delo",
     cases.cfg_harm3d(n=128, nup=400, nfreq=4, fmin=2e11, fmax=2e12, extra=1), (8001, 8640)),
    ("next-row stype='tail': HARM3D HYBRIDTHPL, nn=[128,128,400] nfreq=4 nvals=4 delo",
     cases.cfg_harm3d(n=128, nup=400, nfreq=4, fmin=1e11, fmax=1e13, ename='HYBRIDTHPL', stype='tail', p1=3., p2=3., jetalpha=.05),
     (8001, 8640)),
    ("next-row KORALNTH (synthetic 2-D MKS3 dump with 12 electron bins) HYBRIDTHBINS nn=[128,128,400] nfreq=8 nvals=1",
     cases.cfg_koralnth(n=128, nup=400, nfreq=8, ename='HYBRIDTHBINS', fmin=1e10, fmax=1e14), (8001, 9280)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    args = ap.parse_args()
    rows = []
    for name, kw, sub in CONFIGS:
        a = args_from_kwargs(**kw)
        npix = a.nn[0] * a.nn[1]
        s = _lib.Session(a)
        s.launch(); s.sync()
        ms = s.run_timed(3) / 3.0
        nr, npts = s.counts()
        s.close()
        got = _lib.main(args_from_kwargs(**kw))['ivals']
        i1, i2 = sub if sub else (1, npix)
        t0 = time.time()
        ref = orc.run(_args=args_from_kwargs(i1=i1, i2=i2, **kw))
        cpu_s = ref['seconds']
        g, r = got[:a.nvals, i1 - 1:i2], ref['ivals'][:a.nvals, i1 - 1:i2]   # (extra=1: the Stokes block)
        # Cameras in which the reference algorithm itself is numerically unstable are reported separately: at nu << nu_c the
        # polsynchth fits give |j_V| > j_I, the absorption matrix gets a negative eigenvalue and the transfer solution grows
        # exponentially (I < 0, |I| up to 1e27): rounding differences are amplified without bound there.
        unstable = [k for k in range(r.shape[2]) if np.nanmin(r[0, :, k]) < -1e-3 * np.nanmax(np.abs(r[0, :, k]))]
        stable = [k for k in range(r.shape[2]) if k not in unstable]
        relI, flux, quv, nanok = cases.image_errors(g[:, :, stable], r[:, :, stable])
        l1 = float(np.nansum(np.abs(g - r)) / np.nansum(np.abs(r)))
        l1u = float(max([np.nansum(np.abs(g[:, :, k] - r[:, :, k])) / np.nansum(np.abs(r[:, :, k])) for k in unstable] + [0.]))
        row = {"config": name, "rays": npix, "nfreq": a.nfreq, "points": int(npts), "gpu_ms": ms,
               "gpu_rays_freq_per_s": npix * a.nfreq / (ms * 1e-3),
               "cpu_oracle_rays_freq_per_s": (i2 - i1 + 1) * a.nfreq / cpu_s, "cpu_threads": os.cpu_count(),
               "cpu_rays": i2 - i1 + 1, "max_rel_dI": relI, "flux_err": flux, "max_dQUV_over_Imax": quv, "nan_masks_equal": bool(nanok),
               "L1_rel_err_all_cameras": l1, "cameras_where_reference_is_unstable": unstable,
               "max_L1_rel_err_unstable_cameras": l1u}
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
