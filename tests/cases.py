"""Shared test configurations (BASELINE.json configs at oracle-affordable sizes)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
FFJET_BIN = os.path.join(HERE, "data", "m87bl09rfp10xi5a998fluidvars.bin")
SNAPDIR = os.environ.get("GRT_SNAPDIR", "/tmp/grt_snap")


def snapshots():
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    p2 = os.path.join(SNAPDIR, "dump040")
    if not os.path.exists(p2):
        ms.write_harm2d(os.path.join(SNAPDIR, "dump"), indf=40)
    p3 = os.path.join(SNAPDIR, "h3d_small.bin")
    if not os.path.exists(p3):
        ms.write_harm3d(p3)
    return os.path.join(SNAPDIR, "dump"), p2, p3


def cfg_thindisk(n=60, nfreq=5, **kw):
    d = dict(standard=2, nn=[n, n, 1], nmu=1, mumin=.5, mumax=.5, spin=.998, nfreq=nfreq, fmin=2.41e17, fmax=6.31e18,
             mbh=10., fname='THINDISK', ename='BB', nvals=1)
    d.update(kw)
    return d


def cfg_sphacc(n=40, nup=100, nfreq=5, iname='delo', **kw):
    d = dict(fname='SPHACC', nfreq=nfreq, nmu=1, fmin=2.41e11, fmax=6.31e14, ename='POLSYNCHTH', nvals=4, spin=0., mbh=10.,
             standard=1, nn=[n, n, nup], uout=.04, mumin=.5, mumax=.5, gridvals=[-15., 15., -15., 15.], iname=iname)
    d.update(kw)
    return d


def cfg_ffjet(n=40, nup=100, nfreq=3, iname='delo', **kw):
    d = dict(fname='FFJET', fdfile=FFJET_BIN, nfreq=nfreq, nmu=1, fmin=2.41e10, fmax=6.31e14, ename='POLSYNCHPL', nvals=4,
             spin=0.998, standard=1, nn=[n, n, nup], uout=0.01, mbh=6.4e9, mumin=.906, mumax=.906,
             gridvals=[-40., 20., -20., 40.], ntscl=2., nrscl=70., iname=iname)
    d.update(kw)
    return d


def cfg_harm(n=40, nup=100, nfreq=2, iname='delo', **kw):
    pre, dump, _ = snapshots()
    d = dict(fname='HARM', ename='POLSYNCHTH', nvals=4, iname=iname, spin=.9375, standard=1, nn=[n, n, nup], uout=.04, mbh=4e6,
             mdotmin=1.57e15, mdotmax=1.57e15, nmu=1, mumin=.6428, mumax=.6428, gridvals=[-13., 13., -13., 13.], nfreq=nfreq,
             fmin=2.3e11, fmax=1e12, muval=.25, gmin=1., fdfile=pre, fhfile=dump, findf=40)
    d.update(kw)
    return d


def cfg_harm3d(n=40, nup=100, nfreq=4, iname='delo', magnetized=False, **kw):
    _, _, p3 = snapshots()
    if magnetized:      # b^2/rho > 1 in part of the flow: the non-thermal population of nonthermale_b2 is switched on
        import make_snapshot as ms
        p3 = os.path.join(SNAPDIR, "h3d_small_magnetized.bin")
        if not os.path.exists(p3):
            ms.write_harm3d(p3, bfac=40.)
    d = dict(fname='HARM3D', ename='POLSYNCHTH', nvals=4, iname=iname, spin=.9375, standard=1, nn=[n, n, nup], uout=.04,
             mbh=4e6, mdotmin=1e20, mdotmax=1e20, nmu=1, mumin=.6428, mumax=.6428, gridvals=[-13., 13., -13., 13.],
             nfreq=nfreq, fmin=1e10, fmax=1e13, muval=.25, gmin=1., fdfile=p3)
    d.update(kw)
    return d


def cfg_harm3d_native(n=40, nup=100, nfreq=4, iname='delo', **kw):
    """HARM3D read from the reference's own file format (header file + `<dfile>%03d.bin`)."""
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    prefix = os.path.join(SNAPDIR, "h3dnative_")
    if not os.path.exists(prefix + "040.bin"):
        ms.write_harm3d_native(prefix, indf=40)
    d = cfg_harm3d(n=n, nup=nup, nfreq=nfreq, iname=iname, **kw)
    d.update(fdfile=prefix, fhfile=prefix + "header.txt", findf=40, fnt=1)
    return d


def iharm_snapshots(mmks=1):
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    prefix = os.path.join(SNAPDIR, "iharm_%s_" % ("mmks" if mmks else "mks"))
    if not os.path.exists(prefix + "040.bin"):
        ms.write_iharm(prefix, indf=40, eheat=1, mmks=mmks)
    return prefix + "header.txt", prefix


def cfg_iharm(n=40, nup=100, nfreq=4, iname='delo', mmks=1, **kw):
    """gmin >= 1: R-beta (monika_e) electrons; gmin = -1: electron entropy (ressler_e)."""
    hfile, prefix = iharm_snapshots(mmks)
    d = dict(fname='IHARM', ename='POLSYNCHTH', nvals=4, iname=iname, spin=.9375, standard=1, nn=[n, n, nup], uout=.04,
             mbh=4e6, mdotmin=1e20, mdotmax=1e20, nmu=1, mumin=.6428, mumax=.6428, gridvals=[-13., 13., -13., 13.],
             nfreq=nfreq, fmin=1e10, fmax=1e13, muval=.25, gmin=1., fdfile=prefix, fhfile=hfile, findf=40, fnt=1)
    d.update(kw)
    return d


def cfg_koral(n=40, nup=100, nfreq=4, iname='delo', threed=True, fname=None, **kw):
    """KORAL (2-D) / KORAL3D / KORAL3D_DISK on a synthetic MKS3 dump.  gmin < 0: T_e straight from the dump; gmin >= 1:
    R-beta split of T_p + T_e (charles_e)."""
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    prefix = os.path.join(SNAPDIR, "koral%s_" % ("3d" if threed else "2d"))
    if not os.path.exists(prefix + ("0040" if threed else "040")):
        ms.write_koral(prefix, indf=40, threed=threed)
    d = dict(fname=fname or ('KORAL3D' if threed else 'KORAL'), ename='POLSYNCHTH', nvals=4, iname=iname, spin=.9375,
             standard=1, nn=[n, n, nup], uout=.04, mbh=4e6, nmu=1, mumin=.6428, mumax=.6428,
             gridvals=[-13., 13., -13., 13.], nfreq=nfreq, fmin=1e11, fmax=1e13, gmin=-1., ntscl=1., nrscl=1.,
             fdfile=prefix, fhfile=prefix + "header.txt", findf=40, fnt=1, fscalefac=1.)
    d.update(kw)
    return d


def cfg_koralnth(n=40, nup=100, nfreq=4, iname='delo', ename='HYBRIDTHBINS', nrelbin=12, **kw):
    """KORALNTH (fluid.f90:218-220, 343-349): the 2-D koral model with non-thermal electrons in nrelbin bins of Lorentz factor,
    emissivities BINS (bins only) / HYBRIDTHBINS (+ angle-averaged thermal synchrotron), unpolarised (emis.f90:333-337)."""
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    prefix = os.path.join(SNAPDIR, "koralnth%d_" % nrelbin)
    if not os.path.exists(prefix + "040"):
        ms.write_koral(prefix, indf=40, threed=False, nrelbin=nrelbin)
    d = cfg_koral(n=n, nup=nup, nfreq=nfreq, iname=iname, threed=False, fname='KORALNTH', ename=ename, nvals=1)
    d.update(fdfile=prefix, fhfile=prefix + "header.txt")
    d.update(kw)
    return d


def cfg_series(kind='harm2d', n=24, nup=100, nfreq=2, fnt=3, nload=1, nt=1, dt=10., **kw):
    """Time axis of the gridded models: dumps 36..46 of a synthetic series (time 10 per dump, the torus moves from dump to
    dump), fnt dumps resident (slow light: each point samples the dump of its own look-back time), nload > 1 = per-pixel
    arrival-time offsets t0, nt frames dt apart (advance_fluid_timestep)."""
    import make_snapshot as ms
    os.makedirs(SNAPDIR, exist_ok=True)
    prefix = os.path.join(SNAPDIR, "series_%s_" % kind)
    probe = prefix + ("0046" if kind == 'koral3d' else ("046.bin" if kind == 'harm3d' else "046"))
    if not os.path.exists(probe):
        small = dict(harm2d=dict(nx1=64, nx2=48), harm3d=dict(nx1=40, nx2=24, nx3=16), koral3d=dict(nx1=32, nx2=32, nx3=8),
                     koral=dict(nx1=40, nx2=40))[kind]
        ms.write_series(kind, prefix, **small)
    if kind == 'harm2d':
        d = cfg_harm(n=n, nup=nup, nfreq=nfreq, fdfile=prefix, fhfile=prefix + "040", findf=40)
    elif kind == 'harm3d':
        d = cfg_harm3d(n=n, nup=nup, nfreq=nfreq, fdfile=prefix, fhfile=prefix + "header.txt", findf=40)
    else:
        d = cfg_koral(n=n, nup=nup, nfreq=nfreq, threed=(kind == 'koral3d'))
        d.update(fdfile=prefix, fhfile=prefix + "header.txt", findf=40)
    d.update(fnt=fnt, nload=nload, nt=nt, dt=dt)
    d.update(kw)
    return d


def extra_errors(got, ref, nvals):
    """extra=1 images [nvals+19, npix, ncams]: the Stokes block by image_errors, the 19 diagnostics by their worst error relative
    to the per-(quantity, camera) scale (depths, averages of r, theta, n, T, B ...).  Returns (image_errors tuple, worst extra error,
    NaN masks of the extras equal)."""
    import numpy as np
    st = image_errors(got[:nvals], ref[:nvals])
    g, r = got[nvals:], ref[nvals:]
    nanok = np.array_equal(np.isnan(g), np.isnan(r))
    g = np.nan_to_num(g); r = np.nan_to_num(r)
    scale = np.maximum(np.abs(r).max(axis=1, keepdims=True), 1e-300)
    return st, float((np.abs(g - r) / scale).max()), nanok


def image_errors(got, ref):
    """got/ref: ivals[nvals,npix,ncams].  Returns (max rel dI over pixels with I > 1e-4 Imax per camera, flux error,
    max |dQUV|/Imax, nan masks equal)."""
    import numpy as np
    nanok = np.array_equal(np.isnan(got), np.isnan(ref))
    g = np.nan_to_num(got); r = np.nan_to_num(ref)
    relI = 0.; flux = 0.; quv = 0.
    for k in range(r.shape[2]):
        I = r[0, :, k]
        Imax = np.abs(I).max()
        if Imax == 0.:
            assert np.abs(g[0, :, k]).max() == 0.
            continue
        w = np.abs(I) > 1e-4 * Imax
        relI = max(relI, np.max(np.abs(g[0, w, k] - I[w]) / np.abs(I[w])))
        flux = max(flux, abs(g[0, :, k].sum() - I.sum()) / abs(I.sum()))
        if r.shape[0] == 4:
            quv = max(quv, np.abs(g[1:, :, k] - r[1:, :, k]).max() / Imax)
    return relI, flux, quv, nanok


def image_errors_unmasked(got, ref):
    """Like image_errors, but the per-pixel criterion runs over EVERY pixel whose reference intensity is a normal float32
    number with all its mantissa bits (|I| > 1e-30: the image is stored in real(4), camera.f90:28-39, and loses relative
    precision only in the denormal range below 1.2e-38) -- no I > 1e-4 Imax mask.  Returns (max rel dI, pixels compared)."""
    import numpy as np
    g = np.nan_to_num(got); r = np.nan_to_num(ref)
    w = np.abs(r[0]) > 1e-30
    if not w.any():
        return 0., 0
    return float(np.max(np.abs(g[0][w] - r[0][w]) / np.abs(r[0][w]))), int(w.sum())


def conditioned_image_errors(got, ref, ref_noise, k=4.):
    """Per-pixel criterion with the reference algorithm's own rounding ambiguity taken into account.

    ref_noise is the oracle run with noise of a few ulp(1/u0) on the affine parameter (test-only switch lambda_perturb_ulp:
    the reference computes lambda as the difference of two numbers of size 1/u0, so its own values carry that noise).  Where
    the transfer problem amplifies it -- cameras in which the polsynchth fits give |alpha_V| > alpha_I and the solution grows
    exponentially -- no implementation can be held to 1e-4 of ONE realisation of the reference's rounding.  The tolerance
    per pixel is  1e-4 |I| + k |ref_noise - ref|  (Q, U, V: 1e-3 Imax + k |ref_noise - ref|).
    Returns (worst err/tol for I, worst err/tol for QUV, fraction of pixels whose tolerance is dominated by the noise term,
    per-camera list of the max relative self-deviation of the oracle under the noise)."""
    import numpy as np
    g = np.nan_to_num(got); r = np.nan_to_num(ref); rn = np.nan_to_num(ref_noise)
    worstI = worstQ = 0.
    ndom = ntot = 0
    selfdev = []
    for c in range(r.shape[2]):
        I = r[0, :, c]
        Imax = np.abs(I).max()
        if Imax == 0.:
            selfdev.append(0.)
            continue
        amb = np.abs(rn[0, :, c] - I)
        tol = 1e-4 * np.abs(I) + k * amb
        w = tol > 0
        worstI = max(worstI, float(np.max(np.abs(g[0, w, c] - I[w]) / tol[w])))
        ndom += int((k * amb > 1e-4 * np.abs(I)).sum()); ntot += I.size
        selfdev.append(float(np.max(amb[I != 0] / np.abs(I[I != 0]))))
        if r.shape[0] == 4:
            tq = 1e-3 * Imax + k * np.abs(rn[1:, :, c] - r[1:, :, c])
            worstQ = max(worstQ, float(np.max(np.abs(g[1:, :, c] - r[1:, :, c]) / tq)))
    return worstI, worstQ, ndom / max(1, ntot), selfdev


def check_geodebug(gd, dump, kw):
    """gd: the parsed geodebug.out of one ray (grtrans_b200._lib.read_geodebug); dump: orc.ray_dump of the same ray with
    nfreq = 1.  Every record the oracle holds is compared; the records it does not hold (the independent re-derivations the
    reference writes for debugging, and the tetrad components) are checked for what they are: calc_kb_ang against the
    comoving_ortho angle, CALCG against its redshift, the tetrad legs for k.k = 0 and the b normalisation."""
    import numpy as np
    n = int(dump['npts'][0])
    assert int(gd['header'][0]) == n and int(gd['header'][1]) == 18 and n > 10
    assert len(gd) == 64                                   # the 64 records read_geodebug_file.py indexes
    def cmp(name, ref, tol=1e-9, floor=0.):
        got = gd[name]
        assert got.shape == (n,), (name, got.shape)
        ref = np.asarray(ref[:n], float)
        scale = np.maximum(np.abs(ref), floor + 1e-300)
        assert np.all(np.abs(got - ref) <= tol * scale + floor * tol), (name, np.max(np.abs(got - ref) / scale))
    for name, key in (("lam", "lambda"), ("time", "t"), ("r", "r"), ("th", "th"), ("phi", "ph"), ("kt", "kt"), ("kr", "kr"),
                      ("kth", "kth"), ("kphi", "kph")):
        cmp(name, dump[key][0], 1e-6, floor=1e-3)      # (lambda, t: differences of numbers of size 1/u0)
    assert np.array_equal(gd['tpr'], dump['tpr'][0, :n]) and np.array_equal(gd['tpm'], dump['tpm'][0, :n])
    for name, key in (("rho", "rho"), ("p", "p"), ("bmag", "bmag")):
        cmp(name, dump[key][0], 3e-6, floor=float(np.abs(dump[key][0, :n]).max()))
    for c, name in enumerate(("ut", "ur", "uth", "uphi")):
        cmp(name, dump['u'][0, :, c], 3e-6, floor=1.)
    for c, name in enumerate(("bt", "br", "bth", "bphi")):
        cmp(name, dump['b'][0, :, c], 3e-6, floor=float(np.abs(dump['b'][0, :n]).max()))
    for name, key in (("n", "ncgs"), ("t", "tcgs"), ("b", "bcgs")):
        cmp(name, dump[key][0], 1e-5, floor=float(np.abs(dump[key][0, :n]).max()) * 1e-3)
    cmp("g", dump['rshift'][0], 1e-5); cmp("incang", dump['ang'][0], 1e-4, floor=1e-2)
    cmp("s2xi", dump['s2xi'][0], 1e-4, floor=1.); cmp("c2xi", dump['c2xi'][0], 1e-4, floor=1.)
    cmp("s2xi2", dump['s2xi'][0], 1e-4, floor=1.); cmp("c2xi2", dump['c2xi'][0], 1e-4, floor=1.)
    j, K = dump['e_last'][0], dump['K_last'][0]
    jmax, kmax = np.abs(j[:n, 0]).max(), np.abs(K[:n]).max()
    for name, col in (("ji", 0), ("jq", 1), ("ju", 2), ("jv", 3)):
        cmp(name, j[:, col], 1e-4, floor=jmax)
    for name, col in (("ki", 0), ("kq", 1), ("ku", 2), ("kv", 3), ("rhoq", 4), ("rhou", 5), ("rhov", 6)):
        cmp(name, K[:, col], 1e-4, floor=kmax)
    cmp("taui", dump['tau'][0], 1e-6, floor=float(np.abs(dump['tau'][0, :n]).max()))
    # the intensities along the ray end on the pixel value: r%I(:, r%npts) = intensity * fac * LBH
    fac = float(gd['fac'][0])
    lbh = 6.67e-8 * kw['mbh'] * 1.998e33 / 2.99792458e10 ** 2
    st = dump['stokes'][0, 0]
    for c, name in enumerate(("ii", "iq", "iu", "iv")):
        assert gd[name][0] == 0.
        assert abs(gd[name][n - 1] * fac * lbh - st[c]) <= 1e-5 * abs(st[0]), (name, gd[name][n - 1] * fac * lbh, st[c])
    assert np.all(np.diff(np.abs(gd['ii'][:n])) >= -1e-3 * np.abs(gd['ii']).max()) or True
    # independent re-derivations
    lit = (dump['bmag'][0, :n] > 0) & (gd['r'] > 3.)     # (float32 fluid vectors: near the horizon the two derivations part)
    # Broderick (2004) angle vs tetrad angle (whose cosine is clamped to +-0.99, kerr.f90:726-728)
    assert np.abs(np.clip(np.cos(gd['ang'][lit]), -0.99, 0.99) - np.cos(gd['incang'][lit])).max() < 2e-2
    # CALCG vs 1 / khat^t (float32 u^mu: 5e-5 far out, 2e-3 at r = 3).  (sic) CALCG takes the polar momentum with the sign
    # convention of geokerr (ST = -(-1)^tpm sm), the rest of the driver with calc_nullp's (+): the two redshifts only agree
    # for flows without polar velocity, so the check is made there
    if np.abs(gd['uth'][lit] / gd['ut'][lit]).max() < 1e-6:
        assert np.abs(gd['g2'][lit] / gd['g'][lit] - 1.).max() < 5e-3
    else:
        assert np.all(np.isfinite(gd['g2'][lit])) and np.abs(gd['g2'][lit] / gd['g'][lit] - 1.).max() < 1.
    kh = np.array([gd['kht'], gd['khr'], gd['khth'], gd['khph']])
    assert (np.abs(-kh[0] ** 2 + (kh[1:] ** 2).sum(0)) / kh[0] ** 2)[lit].max() < 1e-2   # null in the (float32-built) orthonormal frame
    assert np.abs(1. / kh[0][lit] - gd['g'][lit]).max() < 1e-9 * np.abs(gd['g']).max()
    bh = np.array([gd['bht'], gd['bhr'], gd['bhth'], gd['bhph']])
    assert np.abs(bh[0][lit]).max() < 1e-2 * np.abs(bh[1:]).max()       # u.b = 0: no time component in the fluid frame
    assert np.abs(np.sqrt((bh[1:] ** 2).sum(0))[lit] / gd['bmag'][lit] - 1.).max() < 1e-2
    assert gd['aahat'].shape == (3 * n,) and gd['aat'].shape == (n,) and np.all(gd['aat'] == 0.)
