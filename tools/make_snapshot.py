#!/usr/bin/env python
"""Seeded synthetic GRMHD-like snapshots in the reference's input formats.

The reference's HARM test input `dump040` is missing from the checkout (.MISSING_LARGE_BLOBS), so
configs 4 and 5 run on analytic torus-like fluids written in

  * the HARM 2-D ASCII dump layout read by fluid_model_harm.f90:275-408 (26-number header; rows of 34
    columns: x1,x2,r,theta = cols 1-4, rho = col 5, u = col 6, u^mu = cols 14-17, b^mu = cols 22-25,
    gdet = col 34), and
  * a small binary container for the 3-D model (harm3d has no bundled reader input): int32 nx1,nx2,nx3;
    float32 asim,gam,h; float32[n] x1,x2,x3,r,th,ph; float64[n] rho,u; float64[4n] ucon, bcon (modified
    Kerr-Schild components), index (ix1*nx2+ix2)*nx3+ix3.

The flow is a valid (u.u=-1, u.b=0) field in Kerr-Schild coordinates: the normal-observer infall plus a
sub-Keplerian rotation, a Gaussian torus in (r, theta) with m=1,2 azimuthal perturbations in 3-D, and a
toroidal + radial magnetic field.  Everything is a closed-form function of (seed, grid), so the same file
is regenerated on the GPU box.
"""
import numpy as np


def _ks_metric(r, th, a):
    s2 = np.sin(th) ** 2
    rho2 = r ** 2 + a ** 2 * np.cos(th) ** 2
    psi = 2. * r / rho2
    g = {}
    g['tt'] = -(1. - psi); g['tr'] = psi; g['tp'] = -a * s2 * psi; g['rr'] = 1. + psi
    g['rp'] = -a * s2 * (1. + psi); g['hh'] = rho2; g['pp'] = s2 * (rho2 + a ** 2 * (1. + psi) * s2)
    return g


def _fluid(r, th, ph, a, seed, threed, bfac=1.0, phase=0.0):
    rng = np.random.RandomState(seed)
    amp1, amp2, ph1, ph2 = 0.3 * rng.rand() + 0.2, 0.2 * rng.rand() + 0.1, 2 * np.pi * rng.rand(), 2 * np.pi * rng.rand()
    g = _ks_metric(r, th, a)
    psi = g['tr']
    vr = -psi / (1. + psi)                       # normal observer dr/dt in KS
    om = 0.8 / (r ** 1.5 + a)                    # sub-Keplerian dphi/dt
    for _ in range(40):                           # shrink the rotation until the 4-velocity is timelike
        norm = g['tt'] + 2 * g['tr'] * vr + 2 * g['tp'] * om + g['rr'] * vr ** 2 + 2 * g['rp'] * vr * om + g['pp'] * om ** 2
        bad = norm > -1e-3
        if not bad.any():
            break
        om = np.where(bad, 0.5 * om, om)
    norm = g['tt'] + 2 * g['tr'] * vr + 2 * g['tp'] * om + g['rr'] * vr ** 2 + 2 * g['rp'] * vr * om + g['pp'] * om ** 2
    ut = 1. / np.sqrt(-norm)
    u = np.stack([ut, ut * vr, np.zeros_like(r), ut * om], -1)
    # density: Gaussian torus + floor, azimuthal m=1,2 modes
    # `phase` moves the torus and rotates the azimuthal pattern: consecutive dumps of a time series differ
    r0, w, hh = 8.0 + 4.0 * phase, 5.0, 0.45 * (1. + 0.5 * phase)
    ph = ph + 2.0 * phase
    rho = np.exp(-((r - r0) / w) ** 2) * np.exp(-(np.cos(th) / hh) ** 2) + 1e-4 * (r / r0) ** -1.5
    if threed:
        rho = rho * (1. + amp1 * np.cos(ph - ph1) + amp2 * np.cos(2. * ph - ph2))
    uu = rho * 0.02 * (r / r0) ** -1.0           # internal energy density (code units)
    # magnetic field: toroidal + radial pieces, b^t from u.b = 0
    b0 = 0.05 * bfac * np.sqrt(rho)
    br = 0.3 * b0 / r
    bp = b0 / r * (1. + (0.2 * np.sin(ph) if threed else 0.))
    bth = 0.1 * b0 / r * np.cos(th)
    u_t = g['tt'] * u[..., 0] + g['tr'] * u[..., 1] + g['tp'] * u[..., 3]
    u_r = g['tr'] * u[..., 0] + g['rr'] * u[..., 1] + g['rp'] * u[..., 3]
    u_h = g['hh'] * u[..., 2]
    u_p = g['tp'] * u[..., 0] + g['rp'] * u[..., 1] + g['pp'] * u[..., 3]
    bt = -(u_r * br + u_h * bth + u_p * bp) / u_t
    b = np.stack([bt, br, bth, bp], -1)
    return rho, uu, u, b


def _to_mks(v, r, dthdx2):
    o = v.copy()
    o[..., 1] = v[..., 1] / r
    o[..., 2] = v[..., 2] / dthdx2
    return o


def harm2d_arrays(nx1=96, nx2=64, a=0.9375, hslope=0.3, rin=1.1, rout=40.0, seed=12345, phase=0.0):
    dx1 = np.log(rout / rin) / nx1
    x1 = np.log(rin) + (np.arange(nx1) + 0.5) * dx1
    x2 = (np.arange(nx2) + 0.5) / nx2
    X1, X2 = np.meshgrid(x1, x2, indexing='ij')          # theta fastest
    r = np.exp(X1)
    th = np.pi * X2 + 0.5 * (1. - hslope) * np.sin(2. * np.pi * X2)
    dthdx2 = np.pi * (1. + (1. - hslope) * np.cos(2. * np.pi * X2))
    rho, uu, u, b = _fluid(r, th, np.zeros_like(r), a, seed, False, phase=phase)
    return dict(nx1=nx1, nx2=nx2, a=a, h=hslope, dx1=dx1, dx2=1. / nx2, startx1=np.log(rin), startx2=0., x1=X1, x2=X2,
                r=r, th=th, rho=rho, u=uu, ucon=_to_mks(u, r, dthdx2), bcon=_to_mks(b, r, dthdx2), gam=4. / 3.)


def write_harm2d(prefix, indf=40, time=0.0, **kw):
    """Writes `<prefix>%03d` (dump, also usable as the header file).  Returns the path."""
    d = harm2d_arrays(**kw)
    path = "%s%03d" % (prefix, indf)
    hdr = np.zeros(26)
    hdr[0] = time; hdr[1] = d['nx1']; hdr[2] = d['nx2']; hdr[3] = d['startx1']; hdr[4] = d['startx2']; hdr[5] = d['dx1']
    hdr[6] = d['dx2']; hdr[9] = d['a']; hdr[10] = d['gam']; hdr[24] = d['h']
    n = d['nx1'] * d['nx2']
    rows = np.zeros((n, 34))
    rows[:, 0] = d['x1'].ravel(); rows[:, 1] = d['x2'].ravel(); rows[:, 2] = d['r'].ravel(); rows[:, 3] = d['th'].ravel()
    rows[:, 4] = d['rho'].ravel(); rows[:, 5] = d['u'].ravel()
    rows[:, 13:17] = d['ucon'].reshape(n, 4); rows[:, 21:25] = d['bcon'].reshape(n, 4)
    rows[:, 33] = (d['r'] ** 2 * np.sin(d['th'])).ravel()
    with open(path, 'w') as f:
        f.write(" ".join("%.17g" % v for v in hdr) + "\n")
        for row in rows:
            f.write(" ".join("%.17g" % v for v in row) + "\n")
    return path


def harm3d_arrays(nx1=64, nx2=32, nx3=32, a=0.9375, rin=1.1, rout=40.0, seed=12345, bfac=1.0, phase=0.0):
    dx1 = np.log(rout / rin) / nx1
    x1 = np.log(rin) + (np.arange(nx1) + 0.5) * dx1
    x2 = (np.arange(nx2) + 0.5) / nx2
    x3 = (np.arange(nx3) + 0.5) * 2. * np.pi / nx3
    X1, X2, X3 = np.meshgrid(x1, x2, x3, indexing='ij')  # phi fastest
    r = np.exp(X1); th = np.pi * X2; ph = X3
    rho, uu, u, b = _fluid(r, th, ph, a, seed, True, bfac, phase=phase)
    dth = np.pi * np.ones_like(r)
    return dict(nx1=nx1, nx2=nx2, nx3=nx3, a=a, x1=X1, x2=X2, x3=X3, r=r, th=th, ph=ph, rho=rho, u=uu,
                ucon=_to_mks(u, r, dth), bcon=_to_mks(b, r, dth), gam=4. / 3.)


def write_harm3d(path, **kw):
    d = harm3d_arrays(**kw)
    with open(path, 'wb') as f:
        np.array([d['nx1'], d['nx2'], d['nx3']], np.int32).tofile(f)
        np.array([d['a'], d['gam'], 0.0], np.float32).tofile(f)
        for k in ['x1', 'x2', 'x3', 'r', 'th', 'ph']:
            d[k].astype(np.float32).ravel().tofile(f)
        d['rho'].astype(np.float64).ravel().tofile(f)
        d['u'].astype(np.float64).ravel().tofile(f)
        d['ucon'].astype(np.float64).ravel().tofile(f)
        d['bcon'].astype(np.float64).ravel().tofile(f)
    return path


def write_harm3d_native(prefix, indf=40, time=0.0, **kw):
    """The reference's own harm3d input: `<prefix>header.txt` (15 numbers, fluid_model_harm3d.f90:337-366) and
    `<prefix>%03d.bin` = a text header line + float32 data(35, n) (:516-584).  Returns (header path, data prefix)."""
    d = harm3d_arrays(**kw)
    n = d['nx1'] * d['nx2'] * d['nx3']
    data = np.zeros((n, 35), np.float32)
    data[:, 3] = d['x1'].ravel(); data[:, 4] = d['x2'].ravel(); data[:, 5] = d['x3'].ravel()
    data[:, 6] = d['r'].ravel(); data[:, 7] = d['th'].ravel(); data[:, 8] = d['ph'].ravel()
    data[:, 9] = d['rho'].ravel(); data[:, 10] = d['u'].ravel()
    data[:, 18:22] = d['ucon'].reshape(n, 4); data[:, 26:30] = d['bcon'].reshape(n, 4)
    x1 = d['x1'][:, 0, 0]; x2 = d['x2'][0, :, 0]; x3 = d['x3'][0, 0, :]
    hdr = [time, d['nx1'], d['nx2'], d['nx3'], x1[0] - 0.5 * (x1[1] - x1[0]), 0.0, 0.0, x1[1] - x1[0], x2[1] - x2[0], x3[1] - x3[0],
           d['a'], d['gam'], 0.9, 1.0, 0.0]
    line = " ".join("%.17g" % v for v in hdr) + "\n"
    with open("%s%03d.bin" % (prefix, indf), 'wb') as f:
        f.write(line.encode())
        data.tofile(f)
    hpath = prefix + "header.txt"
    with open(hpath, 'w') as f:
        f.write(line)
    return hpath, prefix


def iharm_arrays(nx1=64, nx2=32, nx3=32, a=0.9375, hslope=0.3, mmks=1, mks_smooth=0.5, poly_xt=0.82, poly_alpha=14.0,
                 rin=1.1, rout=40.0, seed=12345):
    """(M)MKS grid of iharm3d / ebhlight: x1 = log r, theta(x2[, x1]) per fluid_model_iharm.f90:82-118, x3 = phi."""
    dx1 = np.log(rout / rin) / nx1
    startx1 = np.log(rin)
    x1 = startx1 + (np.arange(nx1) + 0.5) * dx1
    x2 = (np.arange(nx2) + 0.5) / nx2
    x3 = (np.arange(nx3) + 0.5) * 2. * np.pi / nx3
    X1, X2, X3 = np.meshgrid(x1, x2, x3, indexing='ij')  # x3 fastest, x1 slowest

    def theta(x2v, x1v):
        thg = np.pi * x2v + 0.5 * (1. - hslope) * np.sin(2. * np.pi * x2v)
        if not mmks:
            return thg
        D = np.pi / (2. + 2. / (poly_xt ** poly_alpha * (1. + poly_alpha)))
        y = 2. * x2v - 1.
        thj = D * y * (1. + (y / poly_xt) ** poly_alpha / (1. + poly_alpha)) + np.pi / 2.
        return thg + np.exp(-mks_smooth * (x1v - startx1)) * (thj - thg)

    r = np.exp(X1); th = theta(X2, X1); ph = X3
    rho, uu, u, b = _fluid(r, th, ph, a, seed, True)
    e = 1e-5
    dthdx2 = (theta(X2 + e, X1) - theta(X2 - e, X1)) / (2 * e)
    dthdx1 = (theta(X2, X1 + e) - theta(X2, X1 - e)) / (2 * e)

    def to_code(v):  # KS (t, r, theta, phi) components -> (t, x1, x2, x3) components
        o = v.copy()
        o[..., 1] = v[..., 1] / r
        o[..., 2] = (v[..., 2] - dthdx1 * o[..., 1]) / dthdx2
        return o
    # electron entropy such that Theta_e = (m_p/m_e) kel rho^(1/3) ~ 3..30 in the torus
    kel = 10. * (9.10938188e-28 / 1.67262158e-24) / np.maximum(rho, 1e-12) ** (1. / 3.) * (0.3 + np.exp(-((r - 8.) / 6.) ** 2))
    return dict(nx1=nx1, nx2=nx2, nx3=nx3, a=a, hslope=hslope, mmks=int(bool(mmks)), mks_smooth=mks_smooth, poly_xt=poly_xt,
                poly_alpha=poly_alpha, startx1=startx1, rin=rin, rout=rout, x1=X1, x2=X2, x3=X3, rho=rho, u=uu,
                ucon=to_code(u), bcon=to_code(b), kel=kel, gam=4. / 3.)


def write_iharm(prefix, indf=40, eheat=1, **kw):
    """Writes `<prefix>%03d.bin` (float32 data(13+eHEAT, n) stream) and `<prefix>header.txt`
    (fluid_model_iharm.f90:427-470, :508-651).  Returns (header path, data prefix)."""
    d = iharm_arrays(**kw)
    n = d['nx1'] * d['nx2'] * d['nx3']
    dlen = 13 + (1 if eheat else 0)
    data = np.zeros((n, dlen), np.float32)
    data[:, 0] = d['x1'].ravel(); data[:, 1] = d['x2'].ravel(); data[:, 2] = d['x3'].ravel()
    data[:, 3] = d['rho'].ravel(); data[:, 4] = d['u'].ravel()
    data[:, 5:9] = d['ucon'].reshape(n, 4); data[:, 9:13] = d['bcon'].reshape(n, 4)
    if eheat:
        data[:, 13] = d['kel'].ravel()
    data.tofile("%s%03d.bin" % (prefix, indf))
    hdr = [0.0, d['nx1'], d['nx2'], d['nx3'], d['a'], d['hslope'], d['gam'], d['mks_smooth'], d['poly_xt'], d['poly_alpha'],
           d['startx1'], d['mmks'], 1 if eheat else 0, 0, 8 + (2 if eheat else 0), d['rin'], d['rout']]
    hpath = prefix + "header.txt"
    with open(hpath, 'w') as f:
        f.write(" ".join("%.17g" % v for v in hdr) + "\n")
    return hpath, prefix


def koral_arrays(threed=True, nx1=48, nx2=40, nx3=16, a=0.9375, h=0.7, r0=0.0, aa=0.0, bb=0.05, pp=2.0, rin=1.1, rout=40.0,
                 seed=12345, phase=0.0):
    """KORAL MKS3 grid (fluid_model_koral.f90:62-80): x1 = log(r - r0), theta(x2, r), x3 = phi in (-pi, pi); BL vectors."""
    dx1 = (np.log(rout - r0) - np.log(rin - r0)) / nx1
    x1 = np.log(rin - r0) + (np.arange(nx1) + 0.5) * dx1
    x2 = (np.arange(nx2) + 0.5) / nx2
    x3 = -np.pi + (np.arange(nx3) + 0.5) * 2. * np.pi / nx3 if threed else np.zeros(1)
    X3, X1, X2 = np.meshgrid(x3, x1, x2, indexing='ij')   # x2 fastest, then x1, then x3
    r = r0 + np.exp(X1)
    w = aa + 2. ** pp * (bb - aa) / r ** pp
    th = 0.5 * np.pi * (1. + np.tan(h * np.pi * (-0.5 + X2 + (1. - 2. * X2) * w)) / np.tan(0.5 * h * np.pi))
    rho, uu, u, b = _fluid(r, th, X3, a, seed, threed, phase=phase)

    def ks2bl(v):
        delta = r ** 2 - 2. * r + a ** 2
        o = v.copy()
        o[..., 0] = v[..., 0] - 2. * r / delta * v[..., 1]
        o[..., 3] = v[..., 3] - a / delta * v[..., 1]
        return o
    te = 4e10 * (0.5 + np.exp(-((r - 8.) / 6.) ** 2)) * (1. + 0.3 * np.cos(th) ** 2)
    be = -0.05 + 0.3 * np.cos(th) ** 4 * (1. - np.exp(-r / 10.))
    return dict(nx1=nx1, nx2=nx2, nx3=(nx3 if threed else 1), a=a, h=h, r0=r0, aa=aa, bb=bb, pp=pp, x1=X1, x2=X2, x3=X3, r=r, th=th,
                rho=1e-16 * rho, te=te, be=be, ucon=ks2bl(u), bcon=600. * ks2bl(b))


def write_koral(prefix, indf=40, threed=True, time=0.0, nrelbin=0, relgammamin=10., relgammamax=1e5, **kw):
    """ASCII dumps read by fluid_model_koral.f90:318-402 (38 columns) / fluid_model_koral3d.f90:397-514 ("shortfile", 23
    columns): `<prefix>%03d` / `<prefix>%04d` and the header file `<prefix>header.txt`.  Returns (header path, prefix).
    nrelbin > 0 (2-D only): the KORALNTH layout -- a second header line `nrelbin relgammamin relgammamax` and rows of
    41 + nrelbin numbers with the non-thermal electron spectrum dn/dgamma of bin ie in column 43 + ie - 1 (1-based; (sic) the
    reference's nthpos = 43 reads one column to the right of where a 41 + nrelbin row ends, fluid_model_koral.f90:345-355,400,
    so the last bin comes from the first column of the next row, which is 0 here): a power law gamma^-2.5 whose
    normalisation follows the magnetisation."""
    d = koral_arrays(threed=threed, **kw)
    n = d['nx1'] * d['nx2'] * d['nx3']
    dlen = 23 if threed else (41 + nrelbin if nrelbin > 0 else 38)
    rows = np.zeros((n, dlen))
    rows[:, 3] = d['x1'].ravel(); rows[:, 4] = d['x2'].ravel()
    if threed:
        rows[:, 5] = d['x3'].ravel(); rows[:, 6] = d['r'].ravel(); rows[:, 7] = d['th'].ravel(); rows[:, 8] = d['x3'].ravel()
        bpos, ppos = 15, 20
        rows[:, 21] = d['be'].ravel()
    else:
        rows[:, 5] = d['r'].ravel(); rows[:, 6] = d['th'].ravel()
        bpos, ppos = 24, 32
    rows[:, 9] = d['rho'].ravel()
    rows[:, 11:15] = d['ucon'].reshape(n, 4)
    rows[:, bpos:bpos + 4] = d['bcon'].reshape(n, 4)
    rows[:, ppos] = d['te'].ravel()
    if threed:
        hdr = [time, d['nx1'], d['nx2'], d['nx3'], d['a'], 0.0, d['r0'], d['h'], d['aa'], d['bb'], d['pp']]
        path = "%s%04d" % (prefix, indf)
    else:
        hdr = [time, d['nx1'], d['nx2'], d['a'], 0.0, d['r0'], d['h'], d['aa'], d['bb'], d['pp']]
        path = "%s%03d" % (prefix, indf)
    line = " ".join("%.17g" % v for v in hdr) + "\n"
    if nrelbin > 0:
        assert not threed
        # dn/dgamma [cm^-3 per unit gamma] = f_nth n_e (p-1) gmin^(p-1) gamma^-p with a non-thermal fraction that follows the
        # magnetisation; bin ie sits in 0-based column 42 + ie (the last one does not fit into the row, see above)
        edges = relgammamin * (relgammamax / relgammamin) ** (np.arange(nrelbin + 1) / float(nrelbin))
        gam = np.sqrt(edges[:-1] * edges[1:])
        ne = d['rho'].ravel() / 1.6726e-24
        b2 = np.einsum('...i,...i', d['bcon'].reshape(n, 4)[:, 1:], d['bcon'].reshape(n, 4)[:, 1:])
        fnth = 1.0 * b2 / (b2 + 0.3 * b2.max())
        pidx = 2.5
        for ie in range(nrelbin):
            col = 42 + ie
            if col < dlen:
                rows[:, col] = fnth * ne * (pidx - 1.) * relgammamin ** (pidx - 1.) * gam[ie] ** (-pidx)
        line2 = "%d %.17g %.17g\n" % (nrelbin, relgammamin, relgammamax)
        line = line + line2
    with open(path, 'w') as f:
        f.write(line)
        np.savetxt(f, rows, fmt="%.9e")
    hpath = prefix + "header.txt"
    with open(hpath, 'w') as f:
        f.write(line)
    return hpath, prefix


def write_series(kind, prefix, first=36, last=46, tstep=10.0, **kw):
    """A time series of dumps `first..last` (dump k: time k*tstep, torus displaced by phase 0.04*(k-40)) for the slow-light /
    movie paths (fnt > 1, nt > 1).  kind: 'harm2d' | 'harm3d' | 'koral3d' | 'koral'.  Returns what the single writer returns."""
    out = None
    for k in range(first, last + 1):
        a = dict(indf=k, time=k * tstep, phase=0.04 * (k - 40))
        a.update(kw)
        if kind == 'harm2d':
            out = write_harm2d(prefix, **a)
        elif kind == 'harm3d':
            out = write_harm3d_native(prefix, **a)
        elif kind == 'koral3d':
            out = write_koral(prefix, threed=True, **a)
        else:
            out = write_koral(prefix, threed=False, **a)
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 3 and sys.argv[1] == "harm2d":
        print(write_harm2d(sys.argv[2]))
    elif len(sys.argv) >= 3 and sys.argv[1] == "harm3d":
        print(write_harm3d(sys.argv[2]))
    else:
        print("usage: make_snapshot.py harm2d <prefix> | harm3d <path>")
