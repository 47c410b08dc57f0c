"""Oracle pinning against the reference's own end-to-end goldens (run_grtrans_test_problems_public.py:30-77).

The reference's criterion is sum|d|/sum|ref| < 1e-2 (FFJET), < 1e-1 (SPHACC); the oracle meets it as it is.  The FFJET
golden (test_grtrans_ffjet.p) was written by an OLDER commit of the reference than the sources in /root/reference; what
differs was bisected with the oracle's test-only switches (orc_set_option, oracle/capi.cc), each a NAMED substitution:

  G = 6.6726e-8, msun = 1.989e33   phys_constants.f90:3-6 now has 6.67e-8 / 1.998e33: the optically thin image scales with
                                   L_BH = G M msun / c^2.  Uniform scale oracle/golden 1.0060 -> 1.0019.
  polsynchpl without rho_Q         the Faraday-conversion term of the power-law emissivity (`kstarq`, polsynchemis.f90:
                                   581-585, "JAD changing sign here after Monika flagged ...") is younger than the golden:
                                   with rho_Q = 0 Stokes V goes from sum|dV|/sum|V| = 2.6e-2 to 7e-4 and
                                   max|dV| from 1.7e-3 to 2e-5 of I_max.  I, Q, U do not notice.
  ODEPACK LSODA itself             the golden holds LSODA's truncation error.  On this problem (coefficients linear between
                                   400 nodes with the reference's as-written neighbour index, radtrans_integrate.f90:453-465,
                                   i.e. a right-hand side with kinks) LSODA at rtol 1e-6 / atol 1e-8 is off the converged
                                   solution by 4e-4 rms and up to 9e-3 per pixel, and moves by 1e-4 rms (3e-3 max) when the
                                   coefficients are perturbed by 1e-7: measured with scipy.integrate.odeint (the same ODEPACK
                                   routine) on the oracle's own coefficients.  The oracle's RK5(4) is converged to 2e-6.
                                   Where LSODA is converged (|LSODA/RK - 1| < 2e-5, 15 % of the pixels) oracle and golden agree
                                   to 2.5e-4 rms / 9e-4 max in I; over all pixels the rms is 5.7e-4, of which LSODA's error
                                   explains the correlated half (corr(golden - RK, LSODA - RK) = 0.53 over the image, 0.77 on
                                   the worst pixels).
What remains unexplained: a uniform factor 1.0019 (many constant sets of that era fit it, e.g. c = 3e10 or e = 4.8e-10; not
identifiable from one number) and 2.5e-4 rms.  Swapping the as-written neighbour index for the intended one moves pixels
by 1.1e-2 rms, so the golden does pin the restatement down to that quirk; the image at nup = 400 is itself 7e-3 rms away
from the nup = 800 image, so 2.5e-4 is 1/30 of the discretisation error of the test problem.
"""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.integrate

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
FFJET_BIN = os.path.join(HERE, "data", "m87bl09rfp10xi5a998fluidvars.bin")
FFJET_KW = dict(fname='FFJET', fdfile=FFJET_BIN, nfreq=1, nmu=1, fmin=3.45e11, fmax=3.45e11, ename='POLSYNCHPL', nvals=4,
                spin=0.998, standard=1, nn=[100, 100, 400], uout=0.01, mbh=3.4e9, mumin=.906, mumax=.906,
                gridvals=[-40, 20, -20, 40], ntscl=2., nrscl=70.)
# the named substitutions (see the module docstring)
GOLDEN_ERA = {"G": 6.6726e-8, "msun": 1.989e33, "pl_rhoq_scale": 0.0}


class era(object):
    """with era(GOLDEN_ERA): ... runs the oracle with the constants / code variants of the golden's commit."""

    def __init__(self, opts):
        self.opts = opts

    def __enter__(self):
        L = orc.lib()
        L.orc_set_option.argtypes = [C.c_char_p, C.c_double]
        assert L.orc_set_option(b"reset", 0.) == 0
        for k, v in self.opts.items():
            assert L.orc_set_option(k.encode(), float(v)) == 0, k

    def __exit__(self, *a):
        orc.lib().orc_set_option(b"reset", 0.)


def _golden():
    return np.load(os.path.join(HERE, "golden", "ffjet_golden.npz"))["ivals"].astype(np.float64)  # [10000,4]


def _l1(a, b):
    return np.sum(np.abs(a - b)) / np.sum(np.abs(b))


@pytest.mark.slow
def test_ffjet_golden_reference_criterion():
    """The reference's own test (run_grtrans_test_problems_public.py:60-77) with today's sources."""
    g = _golden()
    iv = orc.run(**FFJET_KW)['ivals'][:, :, 0].T
    assert _l1(iv, g) < 1e-2
    assert iv[:, 0].argmax() == g[:, 0].argmax() == 5838      # landmark


@pytest.mark.slow
def test_ffjet_golden_named_substitutions():
    g = _golden()
    with era(GOLDEN_ERA):
        iv = orc.run(**FFJET_KW)['ivals'][:, :, 0].T
    imax = g[:, 0].max()
    w = g[:, 0] > 1e-3 * imax
    ratio = iv[w, 0] / g[w, 0]
    scale = np.median(ratio)
    assert _l1(iv, g) < 1.5e-3                                # 5.3e-3 without the substitutions
    assert abs(scale - 1.0) < 2.2e-3                          # 1.0019 (1.0060 without): the unexplained uniform factor
    assert np.std(ratio / scale) < 6.5e-4                     # LSODA truncation error of the golden, see below
    ivs = iv / scale
    assert _l1(ivs[:, 0], g[:, 0]) < 9e-4
    assert _l1(ivs[:, 1], g[:, 1]) < 1.2e-3 and _l1(ivs[:, 2], g[:, 2]) < 1.2e-3
    assert _l1(ivs[:, 3], g[:, 3]) < 1e-3                     # 2.6e-2 with the rho_Q term
    assert np.abs(ivs[:, 3] - g[:, 3]).max() < 5e-5 * imax    # 1.7e-3 with it; north_star: Q/U/V within 1e-3 of I
    assert np.abs(ivs[:, 1:3] - g[:, 1:3]).max() < 1.5e-3 * imax
    # the rho_Q term really is what V needs: with it the V residual is 30 x larger
    with era({"G": 6.6726e-8, "msun": 1.989e33}):
        iv2 = orc.run(**FFJET_KW)['ivals'][:, :, 0].T
    assert _l1(iv2[:, 3] / scale, g[:, 3]) > 2e-2
    # and the as-written neighbour index of radtrans_aux is what the golden's commit ran (intended index: 1.1e-2 rms)
    os.environ["ORC_LSODA_INTENDED"] = "1"
    try:
        with era(GOLDEN_ERA):
            iv3 = orc.run(**FFJET_KW)['ivals'][:, :, 0].T
    finally:
        del os.environ["ORC_LSODA_INTENDED"]
    r3 = iv3[w, 0] / g[w, 0]
    assert np.std(r3 / np.median(r3)) > 5e-3


def _odepack_lsoda(s0, j, K, tau, rtol=1e-6, atol=1e-8, hmax=0.1, max_tau=10.):
    """radtrans_integrate_lsoda (radtrans_integrate.f90:200-307) with the real ODEPACK LSODA (scipy.integrate.odeint):
    tau cut, zero-emissivity trim, coefficient model radtrans_aux :453-465 as written, user Jacobian (jt = 1),
    itask = 4 with tcrit = max(t), hmax."""
    npts = len(s0)
    lamdex = npts if tau.max() <= max_tau else int(np.searchsorted(tau, max_tau, side='right')) + 1
    i1 = 1
    if j[0, 0] == 0:
        nz = np.nonzero(j[1:, 0])[0]
        if len(nz):
            i1 = nz[0] + 2
    i2 = lamdex
    if j[lamdex - 1, 0] == 0:
        for ii in range(1, lamdex):
            if j[lamdex - ii - 1, 0] != 0:
                i2 = lamdex - ii
                break
    ss = s0[::-1].copy()
    t = s0[i1 - 1:i2][::-1].copy()
    if len(t) < 2:
        return np.zeros(4)

    def coef(lam):
        lindx = min(max(np.searchsorted(ss, lam, side="right"), 1), npts - 1)
        w = (lam - ss[lindx - 1]) / (ss[lindx] - ss[lindx - 1])
        indx = npts - lindx + 1
        u = min(indx + 1, npts)
        return (1 - w) * j[indx - 1] + w * j[u - 1], (1 - w) * K[indx - 1] + w * K[u - 1]

    def mat(k):
        return np.array([[k[0], k[1], k[2], k[3]], [k[1], k[0], k[6], -k[5]], [k[2], -k[6], k[0], k[4]],
                         [k[3], k[5], -k[4], k[0]]])

    def rhs(I, lam):
        jj, k = coef(lam)
        return jj - mat(k).dot(I)

    def jac(I, lam):
        return -mat(coef(lam)[1])

    return scipy.integrate.odeint(rhs, np.zeros(4), t, Dfun=jac, rtol=rtol, atol=atol, hmax=hmax, mxstep=100000,
                                  tcrit=[t.max()])[-1]


@pytest.mark.slow
def test_ffjet_golden_scatter_is_lsoda_truncation_error():
    """The per-pixel scatter of the golden around the oracle is ODEPACK's: integrate the oracle's own coefficients with the
    real LSODA at the reference's tolerances for a fixed sample of pixels and compare the three images there."""
    g = _golden()
    w = np.nonzero(g[:, 0] > 1e-2 * g[:, 0].max())[0]
    pix = np.random.RandomState(5).choice(w, 160, replace=False)
    lso, rk = [], []
    with era(GOLDEN_ERA):
        for p in pix:
            d = orc.ray_dump(int(p) + 1, int(p) + 1, **FFJET_KW)
            n = int(d['npts'][0])
            s0, j, K, tau = d['lambda'][0, :n], d['e_last'][0, :n], d['K_last'][0, :n], d['tau'][0, :n]
            I, nout, _ = orc.rt_integrate(0, s0.copy(), j.copy(), K.copy(), tau.copy())
            rk.append(I[nout - 1])
            lso.append(_odepack_lsoda(s0, j, K, tau))
            # (both are in the driver's fac-normalised units; the image value is rk * fac * LBH = d['stokes'])
            lso[-1] = lso[-1] * (d['stokes'][0, 0, 0] / I[nout - 1, 0])
            rk[-1] = d['stokes'][0, 0].copy()
    lso, rk, gg = np.array(lso), np.array(rk), g[pix]
    scale = 1.0019
    e_lsoda = lso[:, 0] / rk[:, 0] - 1.0                      # LSODA's own error (RK5(4) is converged to 2e-6)
    e_gold = gg[:, 0] * scale / rk[:, 0] - 1.0
    assert 1e-4 < np.std(e_lsoda) < 3e-3                      # ~4e-4 rms: 400 x the tolerance it was asked for
    assert np.corrcoef(e_lsoda, e_gold)[0, 1] > 0.35          # 0.53 over the whole image
    assert np.std(e_gold - e_lsoda) < np.std(e_gold)          # the real LSODA is closer to the golden than the converged RK
    quiet = np.abs(e_lsoda) < 3e-5
    assert quiet.sum() >= 8
    assert np.std(e_gold[quiet]) < 4e-4 and np.abs(e_gold[quiet]).max() < 1.5e-3   # 2.5e-4 / 9e-4 over all quiet pixels


def _sphacc():
    r = orc.run(fname='SPHACC', nfreq=25, nmu=1, fmin=1e8, fmax=1e15, ename='SYNCHTHAV', nvals=1, spin=0., mbh=1.,
                standard=1, nn=[10000, 1, 100], gridvals=[0., 400., 0., 0.], uout=.0025)
    iv, ab = r['ivals'], r['ab']
    da = ab[0, 1] - ab[0, 0]
    spec = np.array([np.sum(iv[0, :, i] * ab[0, :]) * da * 2. * np.arccos(-1.) for i in range(25)])
    return iv[0, :, 14], spec


@pytest.mark.slow
def test_sphacc_golden():
    """run_grtrans_test_problems_public.py:30-55 (reference tolerance 1e-1).  The two SPHACC goldens are an independent
    check of the constant substitution found on the FFJET golden: with G, msun of that era both residuals drop 3-4 x."""
    g = np.load(os.path.join(HERE, "golden", "sphacc_golden.npz"))
    gi, gs = g["intensity"], g["spectrum"][:, 0]
    prof, spec = _sphacc()
    t_prof, t_spec = _l1(prof, gi), _l1(spec, gs)
    assert t_prof < 1e-2 and t_spec < 1e-2, (t_prof, t_spec)          # 3.3e-3, 5.1e-3
    assert abs(prof.argmax() - 131) <= 1
    with era({"G": GOLDEN_ERA["G"], "msun": GOLDEN_ERA["msun"]}):
        prof, spec = _sphacc()
    e_prof, e_spec = _l1(prof, gi), _l1(spec, gs)
    assert e_prof < 1.2e-3 and e_spec < 2.5e-3, (e_prof, e_spec)      # 7.9e-4, 1.9e-3
    assert e_prof < 0.4 * t_prof and e_spec < 0.5 * t_spec
