// grtrans_b200 -- Boyer-Lindquist geometry device functions fused into the per-point kernels.
//
// Replaces kerr.f90: blmetric_cov(:381-400, and its real(4) twin :361-379), calc_nullp (:255-290),
// lnrf_frame / lnrf_frame_inv (real(4) variants :427-449, :476-500), transport_perpk (:502-548),
// comoving_ortho_core (:586-730), calc_polvec (:998-1045), calc_rms (:94-100), krolikc (:102-129)
// and class_four_vector.f90's packed-metric dot product (:115-135).  The BL metric has only five
// non-zero components, so vectors are contracted with (gtt, gtp, grr, gthth, gpp) directly instead
// of the reference's 10-component packed metric stored per vector.
#pragma once
#include "grt_math.cuh"

namespace grt {

struct Metric {  // covariant BL metric
  double tt, tp, rr, hh, pp;
};

GRT_HD Metric bl_metric_cs(double r, double cth, double sth, double a) {  // with cos(theta), sin(theta) supplied
  double delta = r * r - 2.0 * r + a * a;
  double rho2 = r * r + a * a * cth * cth;
  double sigma = sq(r * r + a * a) - a * a * delta * sth * sth;
  Metric g;
  g.tt = -(delta - a * a * sth * sth) / rho2;
  g.tp = -2.0 * a * r * sth * sth / rho2;
  g.rr = rho2 / delta;
  g.hh = rho2;
  g.pp = sigma / rho2 * sth * sth;
  return g;
}
GRT_HD Metric bl_metric(double r, double th, double a) { return bl_metric_cs(r, cos(th), sin(th), a); }

struct MetricF {
  float tt, tp, rr, hh, pp;
};
GRT_HD MetricF bl_metric_f(float r, float th, float a) {
  float cth = r4cos(th), sth = r4sin(th);
  float delta = r * r - 2.f * r + a * a;
  float rho2 = r * r + a * a * cth * cth;
  float sigma = (r * r + a * a) * (r * r + a * a) - a * a * delta * sth * sth;
  MetricF g;
  g.tt = -(delta - a * a * sth * sth) / rho2;
  g.tp = -2.f * a * r * sth * sth / rho2;
  g.rr = rho2 / delta;
  g.hh = rho2;
  g.pp = sigma / rho2 * sth * sth;
  return g;
}

// g_{mu nu} a^mu b^nu, evaluated in the order of the reference's lower()+dot_product
GRT_HD double mdot(const Metric &g, const double *a, const double *b) {
  double d0 = g.tt * a[0] + g.tp * a[3];
  double d1 = g.rr * a[1];
  double d2 = g.hh * a[2];
  double d3 = g.tp * a[0] + g.pp * a[3];
  return d0 * b[0] + d1 * b[1] + d2 * b[2] + d3 * b[3];
}

// contravariant photon momentum from the constants of motion (calc_nullp with rcomp, thcomp)
GRT_HD void null_momentum(double q2, double l, double a, double r, double mu, double su, double smu, double *k) {
  double u = 1.0 / r;
  double rho2 = r * r + a * a * mu * mu;
  double delta = r * r - 2.0 * r + a * a;
  double Mf = q2 + (a * a - q2 - l * l) * mu * mu - a * a * mu * mu * mu * mu;
  Mf = (Mf > 0.0) ? Mf : 0.0;
  double Uf = 1.0 + (a * a - q2 - l * l) * u * u + 2.0 * (sq(a - l) + q2) * u * u * u - a * a * q2 * u * u * u * u;
  Uf = (Uf > 0.0) ? Uf : 0.0;
  k[0] = (-a * (a * (1.0 - mu * mu) - l) + (r * r + a * a) / delta * (r * r + a * a - a * l)) / rho2;
  k[1] = su * (r * r * sqrt(Uf)) / rho2;
  k[2] = 1.0 / rho2 * smu * sqrt(Mf / (1.0 - mu * mu));
  k[3] = (-a + l / (1.0 - mu * mu) + a / delta * (r * r + a * a - a * l)) / rho2;
}

// LNRF 3-velocity -> coordinate 3-velocity (single precision, as the reference stores fluid data)
GRT_HD void lnrf_inv_f(float vr, float vt, float om_l, float r, float a, float th, float &vrc, float &vtc, float &vpc) {
  float mu = r4cos(th);
  float D = r * r - 2.f * r + a * a;
  float AR = (r * r + a * a) * (r * r + a * a) - a * a * D * (1.f - mu * mu);
  float RHO = r * r + a * a * mu * mu;
  float ENU = sqrtf(D * RHO / AR);
  float EMU1 = sqrtf(RHO / D);
  float EMU2 = sqrtf(RHO);
  float EPSI = sqrtf(1.f - mu * mu) * sqrtf(AR / RHO);
  float OM = 2.f * a * r / AR;
  vrc = ENU / EMU1 * vr;
  vtc = ENU / EMU2 * vt;
  vpc = ENU / EPSI * om_l + OM;
  if (!((double)D > 0.0)) vrc = vtc = vpc = 0.f;
}

GRT_HD void lnrf_fwd_f(float vr, float vt, float omega, float r, float a, float th, float &vrl, float &vtl, float &vpl) {
  float mu = r4cos(th);
  float D = r * r - 2.f * r + a * a;
  float AR = (r * r + a * a) * (r * r + a * a) - a * a * D * (1.f - mu * mu);
  float RHO = r * r + a * a * mu * mu;
  float ENU = sqrtf(D * RHO / AR);
  float EMU1 = sqrtf(RHO / D);
  float EMU2 = sqrtf(RHO);
  float EPSI = sqrtf(1.f - mu * mu) * sqrtf(AR / RHO);
  float OM = 2.f * a * r / AR;
  vrl = EMU1 / ENU * vr;
  vtl = EMU2 / ENU * vt;
  vpl = EPSI / ENU * (omega - OM);
  if (!((double)D > 0.0)) vrl = vtl = vpl = 0.f;
}

GRT_HD float isco_radius_f(float a) {
  float z1 = 1.f + r4pow(1.f - a * a, 1.f / 3.f) * (r4pow(1.f + a, 1.f / 3.f) + r4pow(1.f - a, 1.f / 3.f));
  float z2 = sqrtf(3.f * a * a + z1 * z1);
  float s = sqrtf((3.f - z1) * (3.f + z1 + 2.f * z2));
  return 3.f + z2 - (signbit(a) ? -fabsf(s) : fabsf(s));
}

// unit space-like vector in the disk plane, perpendicular to p in the Keplerian frame
// (Agol 1997 convention f^0=0; calc_polvec with polarisation angle psi)
GRT_HD void disk_polvec(double r, double mu, const double *p, double a, double psi, double *f) {
  double delta = r * r - 2.0 * r + a * a;
  double AR = sq(r * r + a * a) - a * a * delta * (1.0 - mu * mu);
  double OM = 2.0 * a * r / AR;
  double rho = r * r + a * a * mu * mu;
  double ptt = r * sqrt(delta / AR) * p[0];
  double prt = r / sqrt(delta) * p[1];
  double ptht = r * p[2];
  double ppht = sqrt(AR) / r * (p[3] - OM * p[0]);
  double vel = 1.0 / (pow(r, 3.0 / 2.0) + a);
  double EPSI = sqrt(1.0 - mu * mu) * sqrt(AR / rho);
  double ENU = sqrt(delta * rho / AR);
  vel = EPSI / ENU * (vel - OM);
  double frl = sqrt(delta) / r * (vel * (ptt - prt * prt / ptt) - ppht);
  double fthl = -vel * prt * ptht / ptt / r;
  double fphl = r * prt / sqrt(AR) * (1.0 - vel * ppht / ptt);
  double frp = sqrt(delta) * ptht * prt / r * (-1.0 + vel * ppht / ptt);
  double fthp = 1.0 / r * (prt * prt + (1.0 + vel * vel) * ppht * ppht - 2.0 * vel * ppht * ptt + vel * ptht * ptht * ppht / ptt);
  double fphp = r * ptht / sqrt(AR) * (-(1.0 + vel * vel) * ppht + vel * ptt + vel * ppht * ppht / ptt);
  double cs = cos(psi), sn = sin(psi);
  f[0] = 0.0;
  f[1] = cs * frl + sn * frp;
  f[2] = cs * fthl + sn * fthp;
  f[3] = cs * fphl + sn * fphp;
  f[0] = 0. * f[1];
  Metric g = bl_metric(r, acos(mu), a);
  double s = sqrt(mdot(g, f, f));
  for (int i = 0; i < 4; i++) f[i] = f[i] / s;
}

// calc_polar_psi (kerr.f90:954-996) + calc_kappapw (:1047-1063): polarisation angle at infinity of a photon emitted with the
// thin-disk polarisation vector (psi = 0) through the Penrose-Walker constant, which is a single-precision complex there
GRT_HD void polar_psi(double r, double muf, double q2, double a, double alpha, double beta, double rshift, double mus,
                      const double *p, double &c2psi, double &s2psi, double &cosne) {
  double f[4];
  disk_polvec(r, muf, p, a, 0.0, f);
  const double sm2 = 1.0 - muf * muf;
  const double al = (p[0] * f[1] - p[1] * f[0]) + a * sm2 * (p[1] * f[3] - p[3] * f[1]);
  const double be = (r * r + a * a) * sqrt(sm2) * (p[3] * f[2] - p[2] * f[3]) - a * sqrt(sm2) * (p[0] * f[2] - p[2] * f[0]);
  const float x1 = (float)al, y1 = (float)(-be), x2 = (float)r, y2 = (float)(-a * muf);
#if defined(__CUDA_ARCH__)
  const float re = __fsub_rn(__fmul_rn(x1, x2), __fmul_rn(y1, y2)), im = __fadd_rn(__fmul_rn(x1, y2), __fmul_rn(y1, x2));
#else
  const float re = x1 * x2 - y1 * y2, im = x1 * y2 + y1 * x2;
#endif
  const double kappa2 = (double)re, kappa1 = -(double)im;
  const double gammac = -alpha - a * (1.0 - mus * mus);
  const double polarpsi = atan2(beta * kappa2 - gammac * kappa1, -beta * kappa1 - gammac * kappa2);
  s2psi = sin(2.0 * polarpsi);
  c2psi = cos(2.0 * polarpsi);
  cosne = rshift * sqrt(q2) / r;
}

struct OrthoOut {
  double s2xi, c2xi, ang, g, cosne;
};

// fluid-frame quantities: polarisation-basis rotation (s2xi,c2xi), angle between k and b, redshift
// g and emission-angle cosine (Kulkarni+2011 tetrad, Walker-Penrose transported basis)
struct OrthoDebug {  // comoving_ortho_debug (kerr.f90:562-584): the tetrad components the debug dump lists
  double aahat[3], aa[4], khat[4], bhat[4];
};
GRT_HD void comoving_ortho(double r, double th, double a, double alpha, double beta, double mus, const double *u,
                           const double *b, const double *k, OrthoOut &o, OrthoDebug *dbg = nullptr) {
  // cos(theta), sin(theta) once for the metric, the polarisation reference vector and the tetrad normalisation (they were
  // three separate FP64 libm evaluations of the same argument)
  const double cth = cos(th), sth = sin(th);
  Metric m = bl_metric_cs(r, cth, sth, a);
  double ut = u[0], ur = u[1], um = u[2], up = u[3];
  double utc = m.tt * ut + m.tp * up, upc = m.pp * up + m.tp * ut, urc = m.rr * ur, umc = um * m.hh;
  double Kap1 = alpha + a * sqrt(1. - mus * mus), Kap2 = -beta;
  double aa[4] = {0.0, 0.0, 0.0, 0.0};
  if (Kap1 != 0.0 || Kap2 != 0.0) {
    double k0 = k[0], k1 = k[1], k2 = k[2], k3 = k[3];
    double gam1 = a * cth * k0 - a * a * cth * sth * sth * k3;
    double gam2 = r * (r * r + a * a) * sth * k3 - a * r * sth * k0;
    double gam3 = a * a * cth * sth * sth * k1 - r * (r * r + a * a) * sth * k2;
    double del1 = r * k0 - r * a * sth * sth * k3;
    double del2 = -a * cth * sth * (r * r + a * a) * k3 + a * a * sth * cth * k0;
    double del3 = r * a * sth * sth * k1 + a * cth * sth * (r * r + a * a) * k2;
    double kp = m.pp * k3 + m.tp * k0;
    double denom = (gam2 * del1 - gam1 * del2) * kp + (gam3 * del2 - gam2 * del3) * m.rr * k1 -
                   (gam3 * del1 - gam1 * del3) * m.hh * k2;
    double f1 = (gam2 * Kap1 - del2 * Kap2) * kp - m.hh * k2 * (gam3 * Kap1 - del3 * Kap2);
    double f2 = (del1 * Kap2 - gam1 * Kap1) * kp + m.rr * k1 * (gam3 * Kap1 - del3 * Kap2);
    double f3 = m.hh * k2 * (gam1 * Kap1 - del1 * Kap2) - m.rr * k1 * (gam2 * Kap1 - del2 * Kap2);
    if (fabs(denom) > 0.0) {
      f1 /= denom;
      f2 /= denom;
      f3 /= denom;
    }
    aa[1] = f1; aa[2] = f2; aa[3] = f3;
  } else {
    aa[3] = 1.0 / sqrt(m.pp);
  }
  double delta = r * r + a * a - 2. * r;
  double w = utc * ut + upc * up;
  double nr2 = -m.rr * w * (1. + umc * um);
  double nm2 = m.hh * (1. + umc * um);
  double np2 = -w * delta * sth * sth;
  double snr = sqrt(nr2), snm = sqrt(nm2), snp = sqrt(np2);
  double ekt[4] = {-1.0 * u[0], -1.0 * u[1], -1.0 * u[2], -1.0 * u[3]};
  double ekr[4] = {(urc * ut) / snr, -w / snr, 0.0, (urc * up) / snr};
  double ekm[4] = {(umc * ut) / snm, (umc * ur) / snm, (1 + umc * um) / snm, (umc * up) / snm};
  double ekp[4] = {upc / snp, 0.0, 0.0, -utc / snp};
  double bhr = mdot(m, ekr, b), bhm = mdot(m, ekm, b), bhp = mdot(m, ekp, b);
  double kht = mdot(m, ekt, k), khr = mdot(m, ekr, k), khm = mdot(m, ekm, k), khp = mdot(m, ekp, k);
  double aht = mdot(m, ekt, aa), ahr = mdot(m, ekr, aa), ahm = mdot(m, ekm, aa), ahp = mdot(m, ekp, aa);
  double bdotb = mdot(m, b, b);
  double a1 = ahr - khr * aht / kht, a2 = ahm - khm * aht / kht, a3 = ahp - khp * aht / kht;
  double knorm = khr * khr + khm * khm + khp * khp;
  double sk = sqrt(knorm);
  double b1 = -(a2 * khp - a3 * khm) / sk, b2 = -(a3 * khr - a1 * khp) / sk, b3 = -(a1 * khm - a2 * khr) / sk;
  double bdotk = bhr * khr + bhm * khm + bhp * khp;
  double angnorm;
  if (bdotb > 0.0) {
    double ad = bhr * a1 + bhm * a2 + bhp * a3;
    double bd = bhr * b1 + bhm * b2 + bhp * b3;
    double den = ad * ad + bd * bd;
    o.s2xi = -2.0 * ad * bd / den;
    o.c2xi = (bd * bd - ad * ad) / den;
    angnorm = bdotk / sk / sqrt(bdotb);
  } else {
    o.s2xi = 0.0;
    o.c2xi = 1.0;
    angnorm = 0.5;
  }
  if (dbg) {
    dbg->aahat[0] = a1; dbg->aahat[1] = a2; dbg->aahat[2] = a3;
    for (int i = 0; i < 4; i++) dbg->aa[i] = aa[i];
    dbg->khat[0] = kht; dbg->khat[1] = khr; dbg->khat[2] = khm; dbg->khat[3] = khp;
    dbg->bhat[0] = mdot(m, ekt, b); dbg->bhat[1] = bhr; dbg->bhat[2] = bhm; dbg->bhat[3] = bhp;
  }
  double t = (angnorm <= 0.99) ? angnorm : 0.99;
  t = (angnorm >= -0.99) ? t : -0.99;
  o.ang = acos(t);
  o.g = 1.0 / kht;
  o.cosne = o.g * sqrt(beta * beta + mus * mus * (alpha * alpha - a * a)) / r;
}

// calc_kb_ang (kerr.f90:220-252): angle between wave vector and field in the fluid frame, Broderick (2004) covariant form --
// the debug dump's independent check of comoving_ortho's angle
GRT_HD double kb_angle(double r, double th, double a, const double *k, const double *b, const double *u) {
  const Metric m = bl_metric(r, th, a);
  const double bdotk = mdot(m, b, k), bdotb = mdot(m, b, b), kdotk = mdot(m, k, k), om = -mdot(m, k, u);
  double cdot2 = bdotk * bdotk / bdotb / (kdotk + om * om);
  cdot2 = (cdot2 <= 1.0) ? ((cdot2 >= 0.0) ? cdot2 : 0.0) : 1.0;
  return acos(((bdotk < 0.0 || (bdotk == 0.0 && signbit(bdotk))) ? -1.0 : 1.0) * sqrt(cdot2));
}

// CALCG (kerr.f90:181-218): redshift from the LNRF velocities of the fluid -- the debug dump's independent check of
// comoving_ortho's g.  vr, vt, om = u^r/u^t, u^th/u^t, u^ph/u^t (lnrf_frame, :402-425, is applied here).  (sic) ST carries
// geokerr's sign convention for the polar momentum, opposite to calc_nullp's: the two redshifts agree only for v^theta = 0.
GRT_HD double calcg_lnrf(double r, double mu, double q2, double l, double a, int tpm, int tpr, double su, double sm, double vr,
                         double vt, double omf) {
  const double u = 1.0 / r;
  const double D = r * r - 2.0 * r + a * a;
  const double AR = sq(r * r + a * a) - a * a * D * (1.0 - mu * mu);
  const double RHO = r * r + a * a * mu * mu;
  const double ENU = sqrt(D * RHO / AR), EMU1 = sqrt(RHO / D), EMU2 = sqrt(RHO), EPSI = sqrt(1.0 - mu * mu) * sqrt(AR / RHO);
  const double OM = 2.0 * a * r / AR;
  double vrl = EMU1 / ENU * vr, vtl = EMU2 / ENU * vt, vpl = EPSI / ENU * (omf - OM);
  if (!(D > 0.0)) vrl = vtl = vpl = 0.0;
  const double SR = ((tpr & 1) ? -1.0 : 1.0) * su, ST = -((tpm & 1) ? -1.0 : 1.0) * sm;
  double OMEGA = ENU / EPSI * vpl + OM;
  if (EPSI == 0.0) OMEGA = 0.0;
  const double GAM = 1.0 / sqrt(1.0 - (vrl * vrl + vtl * vtl + vpl * vpl));
  double RR = -a * a * q2 * u * u * u * u + 2.0 * u * u * u * (q2 + (a - l) * (a - l)) + u * u * (a * a - q2 - l * l) + 1.0;
  double TT = (q2 + mu * mu * (a * a - l * l - q2) - a * a * mu * mu * mu * mu) / (1.0 - mu * mu);
  TT = (TT >= 0.0) ? sqrt(TT) : 0.0;
  RR = (RR >= 0.0) ? sqrt(RR) * r * r : 0.0;
  return ENU / GAM / (1.0 - l * OMEGA - EMU1 * ENU * vrl / RHO * SR * RR - EMU2 * ENU * vtl / RHO * ST * TT);
}

}  // namespace grt
