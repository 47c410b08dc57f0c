"""Oracle pinning against the reference's own end-to-end goldens (run_grtrans_test_problems_public.py:30-77).

The metric and tolerances are the reference's: sum|d|/sum|ref| < 1e-2 (FFJET), < 1e-1 (SPHACC).  The achieved
residuals are recorded in DESIGN.md: FFJET 5.3e-3 overall with a UNIFORM 1.0058 scale (the golden predates the
current phys_constants) and 6.5e-4 after removing it; SPHACC 3.3e-3 / 5.1e-3.
"""
import os

import numpy as np
import pytest

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
FFJET_BIN = os.path.join(HERE, "data", "m87bl09rfp10xi5a998fluidvars.bin")


@pytest.mark.slow
def test_ffjet_golden():
    g = np.load(os.path.join(HERE, "golden", "ffjet_golden.npz"))["ivals"].astype(np.float64)  # [10000,4]
    r = orc.run(fname='FFJET', fdfile=FFJET_BIN, nfreq=1, nmu=1, fmin=3.45e11, fmax=3.45e11, ename='POLSYNCHPL', nvals=4,
                spin=0.998, standard=1, nn=[100, 100, 400], uout=0.01, mbh=3.4e9, mumin=.906, mumax=.906,
                gridvals=[-40, 20, -20, 40], ntscl=2., nrscl=70.)
    iv = r['ivals'][:, :, 0].T
    terr = np.sum(np.abs(iv - g)) / np.sum(np.abs(g))
    assert terr < 1e-2, terr                                   # the reference's own criterion
    assert iv[:, 0].argmax() == g[:, 0].argmax() == 5838      # landmark
    # the residual is a uniform scale: after removing it the restatement matches to < 1e-3 in I, Q, U
    w = g[:, 0] > 1e-6
    scale = np.median(iv[w, 0] / g[w, 0])
    assert abs(scale - 1.0058) < 5e-4
    assert np.std(iv[w, 0] / g[w, 0]) < 1e-3
    for s in range(3):
        assert np.sum(np.abs(iv[:, s] / scale - g[:, s])) / np.sum(np.abs(g[:, s])) < 1.5e-3
    assert np.sum(np.abs(iv[:, 3] / scale - g[:, 3])) / np.sum(np.abs(g[:, 3])) < 5e-2


@pytest.mark.slow
def test_sphacc_golden():
    g = np.load(os.path.join(HERE, "golden", "sphacc_golden.npz"))
    r = orc.run(fname='SPHACC', nfreq=25, nmu=1, fmin=1e8, fmax=1e15, ename='SYNCHTHAV', nvals=1, spin=0., mbh=1.,
                standard=1, nn=[10000, 1, 100], gridvals=[0., 400., 0., 0.], uout=.0025)
    iv, ab = r['ivals'], r['ab']
    prof = iv[0, :, 14]
    terr = np.sum(np.abs(prof - g["intensity"])) / np.sum(np.abs(g["intensity"]))
    assert terr < 1e-1 and terr < 1e-2, terr
    da = ab[0, 1] - ab[0, 0]
    spec = np.array([np.sum(iv[0, :, i] * ab[0, :]) * da * 2. * np.arccos(-1.) for i in range(25)])
    terr = np.sum(np.abs(spec - g["spectrum"][:, 0])) / np.sum(np.abs(g["spectrum"]))
    assert terr < 1e-1 and terr < 1e-2, terr
    assert abs(prof.argmax() - 131) <= 1
