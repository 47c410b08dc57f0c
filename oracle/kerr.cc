// ORACLE -- TEST INFRASTRUCTURE ONLY (see geokerr.cc header).
// CPU restatement of the geometry pieces of the hot path:
//   /root/reference/kerr.f90 (metric, calc_nullp, LNRF frames, transport_perpk, comoving_ortho_core,
//   calc_polvec, calc_rms), class_four_vector.f90 (packed-metric dot product) and geodesics.f90
//   (initialize_pixels, initialize_geodesic, geokerr_wrapper).
#include <cmath>
#include <cstring>

#include "oracle.h"

namespace orc {

namespace pc {
const double pi = std::acos(-1.0);
// phys_constants.f90:3-6 (current reference values; see oracle.h)
double h = 6.626e-27, k = 1.38e-16, c = 2.99792458e10, e = 4.8032e-10, G = 6.67e-8, m = 9.10938188e-28,
       mp = 1.67262158e-24, c2 = c * c, sigb = 5.6704e-5, msun = 1.998e33, sigt = 6.6523e-25;
}
LegacyOptions legacy;

static inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }

// kerr.f90:381-400  packed order (tt,tr,tth,tph,rr,rth,rph,thth,thph,phph)
void blmetric_cov(double r, double th, double a, double *g) {
  double cth = std::cos(th), sth = std::sin(th);
  double delta = r * r - 2.0 * r + a * a;
  double rho2 = r * r + a * a * cth * cth;
  double sigma = (r * r + a * a) * (r * r + a * a) - a * a * delta * sth * sth;
  for (int i = 0; i < 10; i++) g[i] = 0.0;
  g[0] = -(delta - a * a * sth * sth) / rho2;
  g[3] = -2.0 * a * r * sth * sth / rho2;
  g[4] = rho2 / delta;
  g[7] = rho2;
  g[9] = sigma / rho2 * sth * sth;
}

// kerr.f90:361-379 (single precision twin)
void blmetric_cov_real(float r, float th, float a, float *g) {
  float cth = std::cos(th), sth = std::sin(th);
  float delta = r * r - 2.f * r + a * a;
  float rho2 = r * r + a * a * cth * cth;
  float sigma = (r * r + a * a) * (r * r + a * a) - a * a * delta * sth * sth;
  for (int i = 0; i < 10; i++) g[i] = 0.f;
  g[0] = -(delta - a * a * sth * sth) / rho2;
  g[3] = -2.f * a * r * sth * sth / rho2;
  g[4] = rho2 / delta;
  g[7] = rho2;
  g[9] = sigma / rho2 * sth * sth;
}

// class_four_vector.f90:115-135  a*b = lower(a) . b with a's packed metric
double dot4(const double *m, const double *a, const double *b) {
  double d0 = m[0] * a[0] + m[1] * a[1] + m[2] * a[2] + m[3] * a[3];
  double d1 = m[1] * a[0] + m[4] * a[1] + m[5] * a[2] + m[6] * a[3];
  double d2 = m[2] * a[0] + m[5] * a[1] + m[7] * a[2] + m[8] * a[3];
  double d3 = m[3] * a[0] + m[6] * a[1] + m[8] * a[2] + m[9] * a[3];
  return d0 * b[0] + d1 * b[1] + d2 * b[2] + d3 * b[3];
}

// kerr.f90:255-290 with rcomp and thcomp present (the only way the hot path calls it,
// geodesics.f90:362-364)
void calc_nullp(double q2, double l, double a, double r, double mu, double su, double smu, double *k) {
  double u = 1.0 / r;
  double rho2 = r * r + a * a * mu * mu;
  double delta = r * r - 2.0 * r + a * a;
  double Mfunc = q2 + (a * a - q2 - l * l) * mu * mu - a * a * mu * mu * mu * mu;
  Mfunc = (Mfunc > 0.0) ? Mfunc : 0.0;
  double pmu = 1.0 / rho2 * smu * std::sqrt(Mfunc / (1.0 - mu * mu));
  double Ufunc = 1.0 + (a * a - q2 - l * l) * u * u + 2.0 * ((a - l) * (a - l) + q2) * u * u * u - a * a * q2 * u * u * u * u;
  Ufunc = (Ufunc > 0.0) ? Ufunc : 0.0;
  double Rfunc = r * r * std::sqrt(Ufunc);
  double pu = su * Rfunc / rho2;
  double pt = (-a * (a * (1.0 - mu * mu) - l) + (r * r + a * a) / delta * (r * r + a * a - a * l)) / rho2;
  double pphi = (-a + l / (1.0 - mu * mu) + a / delta * (r * r + a * a - a * l)) / rho2;
  k[0] = pt;
  k[1] = pu;
  k[2] = pmu;
  k[3] = pphi;
}

// kerr.f90:476-500
void lnrf_frame_inv_real(float vr, float vt, float omega, float r, float a, float th, float &vrl, float &vtl,
                         float &vpl) {
  float mu = std::cos(th);
  float D = r * r - 2.f * r + a * a;
  float AR = (r * r + a * a) * (r * r + a * a) - a * a * D * (1.f - mu * mu);
  float RHO = r * r + a * a * mu * mu;
  float ENU = std::sqrt(D * RHO / AR);
  float EMU1 = std::sqrt(RHO / D);
  float EMU2 = std::sqrt(RHO);
  float EPSI = std::sqrt(1.f - mu * mu) * std::sqrt(AR / RHO);
  float OM = 2.f * a * r / AR;
  vrl = ENU / EMU1 * vr;
  vtl = ENU / EMU2 * vt;
  vpl = ENU / EPSI * omega + OM;
  if (!((double)D > 0.0)) {
    vrl = 0.f;
    vtl = 0.f;
    vpl = 0.f;
  }
}

// kerr.f90:427-449
void lnrf_frame_real(float vr, float vt, float omega, float r, float a, float th, float &vrl, float &vtl, float &vpl) {
  float mu = std::cos(th);
  float D = r * r - 2.f * r + a * a;
  float AR = (r * r + a * a) * (r * r + a * a) - a * a * D * (1.f - mu * mu);
  float RHO = r * r + a * a * mu * mu;
  float ENU = std::sqrt(D * RHO / AR);
  float EMU1 = std::sqrt(RHO / D);
  float EMU2 = std::sqrt(RHO);
  float EPSI = std::sqrt(1.f - mu * mu) * std::sqrt(AR / RHO);
  float OM = 2.f * a * r / AR;
  vrl = EMU1 / ENU * vr;
  vtl = EMU2 / ENU * vt;
  vpl = EPSI / ENU * (omega - OM);
  if (!((double)D > 0.0)) {
    vrl = 0.f;
    vtl = 0.f;
    vpl = 0.f;
  }
}

// kerr.f90:94-100 (all single precision)
float calc_rms(float a) {
  float z1 = 1.f + std::pow(1.f - a * a, 1.f / 3.f) * (std::pow(1.f + a, 1.f / 3.f) + std::pow(1.f - a, 1.f / 3.f));
  float z2 = std::sqrt(3.f * a * a + z1 * z1);
  float s = std::sqrt((3.f - z1) * (3.f + z1 + 2.f * z2));
  return 3.f + z2 - (std::signbit(a) ? -std::fabs(s) : std::fabs(s));
}

// kerr.f90:998-1045
void calc_polvec(double r, double mu, const double *p, double a, double psi, double *f) {
  double delta = r * r - 2.0 * r + a * a;
  double AR = (r * r + a * a) * (r * r + a * a) - a * a * delta * (1.0 - mu * mu);
  double OM = 2.0 * a * r / AR;
  double rho = r * r + a * a * mu * mu;
  double ptt = r * std::sqrt(delta / AR) * p[0];
  double prt = r / std::sqrt(delta) * p[1];
  double ptht = r * p[2];
  double ppht = std::sqrt(AR) / r * (p[3] - OM * p[0]);
  double vel = 1.0 / (std::pow(r, 3.0 / 2.0) + a);
  double EPSI = std::sqrt(1.0 - mu * mu) * std::sqrt(AR / rho);
  double ENU = std::sqrt(delta * rho / AR);
  vel = EPSI / ENU * (vel - OM);
  double frl = std::sqrt(delta) / r * (vel * (ptt - prt * prt / ptt) - ppht);
  double fthl = -vel * prt * ptht / ptt / r;
  double fphl = r * prt / std::sqrt(AR) * (1.0 - vel * ppht / ptt);
  double frp = std::sqrt(delta) * ptht * prt / r * (-1.0 + vel * ppht / ptt);
  double fthp = 1.0 / r * (prt * prt + (1.0 + vel * vel) * ppht * ppht - 2.0 * vel * ppht * ptt + vel * ptht * ptht * ppht / ptt);
  double fphp = r * ptht / std::sqrt(AR) * (-(1.0 + vel * vel) * ppht + vel * ptt + vel * ppht * ppht / ptt);
  double fr = std::cos(psi) * frl + std::sin(psi) * frp;
  double fth = std::cos(psi) * fthl + std::sin(psi) * fthp;
  double fph = std::cos(psi) * fphl + std::sin(psi) * fphp;
  f[0] = 0. * fr;
  f[1] = fr;
  f[2] = fth;
  f[3] = fph;
  double metric[10];
  blmetric_cov(r, std::acos(mu), a, metric);
  double normf = dot4(metric, f, f);
  double s = std::sqrt(normf);
  for (int i = 0; i < 4; i++) f[i] = f[i] / s;
}

// kerr.f90:954-996 calc_polar_psi + :1047-1063 calc_kappapw: polarisation angle at infinity of a photon emitted with the
// thin-disk polarisation vector (calc_polvec, psi = 0), by the Penrose-Walker constant.  kappa_pw is a default (single
// precision) complex in the reference.
void calc_polar_psi(double r, double muf, double q2, double a, double alpha, double beta, double rshift, double mus,
                    const double *p, double &c2psi, double &s2psi, double &cosne) {
  double f[4];
  calc_polvec(r, muf, p, a, 0.0, f);
  double al = (p[0] * f[1] - p[1] * f[0]) + a * (1.0 - muf * muf) * (p[1] * f[3] - p[3] * f[1]);
  double be = (r * r + a * a) * std::sqrt(1.0 - muf * muf) * (p[3] * f[2] - p[2] * f[3]) -
              a * std::sqrt(1.0 - muf * muf) * (p[0] * f[2] - p[2] * f[0]);
  // cmplx(alpha,-beta)*cmplx(r,-a*mu) in single precision
  float x1 = (float)al, y1 = (float)(-be), x2 = (float)r, y2 = (float)(-a * muf);
  float re = x1 * x2 - y1 * y2, im = x1 * y2 + y1 * x2;
  double kappa2 = (double)re, kappa1 = -(double)im;
  double gammac = -alpha - a * (1.0 - mus * mus);
  double denom = beta * kappa2 - gammac * kappa1;
  double num = -beta * kappa1 - gammac * kappa2;
  double polarpsi = std::atan2(denom, num);
  s2psi = std::sin(2.0 * polarpsi);
  c2psi = std::cos(2.0 * polarpsi);
  cosne = rshift * std::sqrt(q2) / r;
}

// kerr.f90:502-548
static void transport_perpk(const double *kv, double r, double th, double a, const double *metric, double Kap1,
                            double Kap2, double &f1, double &f2, double &f3) {
  double g03 = metric[3], g11 = metric[4], g22 = metric[7], g33 = metric[9];
  double cth = std::cos(th), sth = std::sin(th);
  double k0 = kv[0], k1 = kv[1], k2 = kv[2], k3 = kv[3];
  double gam1 = a * cth * k0 - a * a * cth * sth * sth * k3;
  double gam2 = r * (r * r + a * a) * sth * k3 - a * r * sth * k0;
  double gam3 = a * a * cth * sth * sth * k1 - r * (r * r + a * a) * sth * k2;
  double del1 = r * k0 - r * a * sth * sth * k3;
  double del2 = -a * cth * sth * (r * r + a * a) * k3 + a * a * sth * cth * k0;
  double del3 = r * a * sth * sth * k1 + a * cth * sth * (r * r + a * a) * k2;
  double denom = (gam2 * del1 - gam1 * del2) * (g33 * k3 + g03 * k0) + (gam3 * del2 - gam2 * del3) * g11 * k1 -
                 (gam3 * del1 - gam1 * del3) * g22 * k2;
  f1 = (gam2 * Kap1 - del2 * Kap2) * (g33 * k3 + g03 * k0) - g22 * k2 * (gam3 * Kap1 - del3 * Kap2);
  f2 = (del1 * Kap2 - gam1 * Kap1) * (g33 * k3 + g03 * k0) + g11 * k1 * (gam3 * Kap1 - del3 * Kap2);
  f3 = g22 * k2 * (gam1 * Kap1 - del1 * Kap2) - g11 * k1 * (gam2 * Kap1 - del2 * Kap2);
  if (std::fabs(denom) > 0.0) {
    f1 = f1 / denom;
    f2 = f2 / denom;
    f3 = f3 / denom;
  }
}

// kerr.f90:586-730
void comoving_ortho(double r, double th, double a, double alpha, double beta, double mus, const double *u,
                    const double *b, const double *k, double &s2xi, double &c2xi, double &ang, double &g,
                    double &cosne) {
  const double angmin = -0.99, angmax = 0.99;
  double metric[10];
  blmetric_cov(r, th, a, metric);
  double gtt = metric[0], gtp = metric[3], grr = metric[4], gmm = metric[7], gpp = metric[9];
  double ut = u[0], ur = u[1], um = u[2], up = u[3];
  double utc = gtt * ut + gtp * up, upc = gpp * up + gtp * ut, urc = grr * ur, umc = um * gmm;
  double Kap1 = alpha + a * std::sqrt(1. - mus * mus);
  double Kap2 = -beta;
  double al1, al2, al3;
  if ((Kap1 != 0.0) || (Kap2 != 0.0)) {
    transport_perpk(k, r, th, a, metric, Kap1, Kap2, al1, al2, al3);
  } else {
    al1 = 0.0;
    al2 = 0.0;
    al3 = 1.0 / std::sqrt(metric[9]);
  }
  double aa[4] = {0.0, al1, al2, al3};
  double delta = r * r + a * a - 2. * r;
  double nr2 = -grr * (utc * ut + upc * up) * (1. + umc * um);
  double nm2 = gmm * (1. + umc * um);
  double np2 = -(utc * ut + upc * up) * delta * std::sin(th) * std::sin(th);
  double ekt[4] = {-1.0 * u[0], -1.0 * u[1], -1.0 * u[2], -1.0 * u[3]};
  double ekr[4] = {(urc * ut) / std::sqrt(nr2), -(utc * ut + upc * up) / std::sqrt(nr2), 0.0, (urc * up) / std::sqrt(nr2)};
  double ekm[4] = {(umc * ut) / std::sqrt(nm2), (umc * ur) / std::sqrt(nm2), (1 + umc * um) / std::sqrt(nm2),
                   (umc * up) / std::sqrt(nm2)};
  double ekp[4] = {upc / std::sqrt(np2), 0.0, 0.0, -utc / std::sqrt(np2)};
  double bhatr = dot4(metric, ekr, b), bhatm = dot4(metric, ekm, b), bhatp = dot4(metric, ekp, b);
  double khatt = dot4(metric, ekt, k), khatr = dot4(metric, ekr, k), khatm = dot4(metric, ekm, k),
         khatp = dot4(metric, ekp, k);
  double ahatt = dot4(metric, ekt, aa), ahatr = dot4(metric, ekr, aa), ahatm = dot4(metric, ekm, aa),
         ahatp = dot4(metric, ekp, aa);
  double bdotb = dot4(metric, b, b);
  double aahat1 = ahatr - khatr * ahatt / khatt;
  double aahat2 = ahatm - khatm * ahatt / khatt;
  double aahat3 = ahatp - khatp * ahatt / khatt;
  double knorm = khatr * khatr + khatm * khatm + khatp * khatp;
  double bbhat1 = -(aahat2 * khatp - aahat3 * khatm) / std::sqrt(knorm);
  double bbhat2 = -(aahat3 * khatr - aahat1 * khatp) / std::sqrt(knorm);
  double bbhat3 = -(aahat1 * khatm - aahat2 * khatr) / std::sqrt(knorm);
  double bdotk = bhatr * khatr + bhatm * khatm + bhatp * khatp;
  double angnorm;
  if (bdotb > 0.0) {
    double aadotbp = bhatr * aahat1 + bhatm * aahat2 + bhatp * aahat3;
    double bpdotbb = bhatr * bbhat1 + bhatm * bbhat2 + bhatp * bbhat3;
    s2xi = -2.0 * aadotbp * bpdotbb / (aadotbp * aadotbp + bpdotbb * bpdotbb);
    c2xi = (bpdotbb * bpdotbb - aadotbp * aadotbp) / (aadotbp * aadotbp + bpdotbb * bpdotbb);
    angnorm = bdotk / std::sqrt(knorm) / std::sqrt(bdotb);
  } else {
    s2xi = 0.0;
    c2xi = 1.0;
    angnorm = 0.5;
  }
  // merge(merge(angnorm,angmax,angnorm<=angmax),angmin,angnorm>=angmin): NaN -> angmin
  double t = (angnorm <= angmax) ? angnorm : angmax;
  t = (angnorm >= angmin) ? t : angmin;
  ang = std::acos(t);
  g = 1.0 / khatt;
  cosne = g * std::sqrt(beta * beta + mus * mus * (alpha * alpha - a * a)) / r;
}

// geodesics.f90:64-111
void initialize_pixels(GeokerrArgs &g, int standard, double mu0, double phi0, double a, double uout, double uin,
                       double rcut, int nrotype, double a1, double a2, double b1, double b2, int nro, int nphi,
                       int nup) {
  int n = nro * nphi;
  g.uf.assign(n, 0); g.muf.assign(n, 0); g.su.assign(n, 0); g.sm.assign(n, 0); g.alpha.assign(n, 0);
  g.beta.assign(n, 0); g.l.assign(n, 0); g.q2.assign(n, 0); g.t0.assign(n, 0); g.tpm.assign(n, 0); g.tpr.assign(n, 0);
  g.a = a;
  g.mu0 = mu0;
  g.phi0 = phi0;
  g.nup = nup;
  g.uout = uout;
  g.uin = uin;
  if (standard == 1) {
    g.usegeor = 0;
    g.phit = 1;
    if (nup > 1) {
      g.mufill = 1;
      g.kext = (int)std::ceil(3.0 * g.nup / 400.0);
      g.next = (int)std::ceil(60.0 * g.nup / 400.0);
    } else {
      g.mufill = 0;
      g.kext = 0;
      g.next = 0;
    }
  } else {
    g.mufill = 0;
    g.phit = 1;
    g.usegeor = 1;
    g.kext = 0;
    g.next = 0;
  }
  double aa = a, mm = mu0;
  initialize_camera_geokerr(standard, a1, a2, b1, b2, rcut, nrotype, nro, nphi, nup, g.u0, g.uout, mm, aa, g.offset,
                            g.uf.data(), g.muf.data(), g.alpha.data(), g.beta.data(), g.q2.data(), g.l.data(),
                            g.sm.data(), g.su.data(), g.tpm.data(), g.tpr.data());
}

// geodesics.f90:206-257 (initialize_geodesic) + :271-377 (geokerr_wrapper)
void initialize_geodesic(Geo &g, const GeokerrArgs &ga, int i, int &status) {
  status = 1;
  int ix = i - 1;
  double mu0 = ga.mu0;
  // g%gk%phi0=acos(-1.)*gargs%phi0 : pi in SINGLE precision (geodesics.f90:218)
  double phi0 = (double)std::acos(-1.f) * ga.phi0;
  double a = ga.a;
  double t0 = ga.t0[ix];
  double u0 = ga.u0, uf = ga.uf[ix], uout = ga.uout;
  double alpha = ga.alpha[ix], beta = ga.beta[ix], su = ga.su[ix], sm = ga.sm[ix], l = ga.l[ix], q2 = ga.q2[ix];
  int tpm = ga.tpm[ix], tpr = ga.tpr[ix];
  int npts = ga.nup + ga.next - 2 * ga.kext;
  int nup = ga.nup;
  double muf = ga.muf[ix];
  std::vector<double> ufi(npts + 2, 0.0), mufi(npts + 2, 0.0), dti(npts + 2, 0.0), dphi(npts + 2, 0.0),
      lambdai(npts + 2, 0.0);
  std::vector<int> tpmi(npts + 2, 0), tpri(npts + 2, 0);
  int ncase = 0;
  geokerr(u0, uf, uout, mu0, muf, a, l, q2, alpha, beta, tpm, tpr, su, sm, nup, ga.offset, ga.phit != 0,
          ga.usegeor != 0, ga.mufill != 0, ncase, ga.kext, ga.next, ufi.data(), mufi.data(), dti.data(), dphi.data(),
          tpmi.data(), tpri.data(), lambdai.data());
  g.npts = nup;
  g.ncase = ncase;
  g.alpha = alpha; g.beta = beta; g.a = a; g.mu0 = mu0; g.q2 = q2; g.l = l; g.su = su; g.sm = sm;
  int i1 = 1, i2 = g.npts;
  if (ga.usegeor == 1 && g.npts > 1) {
    if (ufi[1] < 0.0) {
      for (int ii = 2; ii <= g.npts; ii++)
        if (ufi[ii] >= 0.0) {
          i1 = ii;
          break;
        }
    } else {
      i1 = 1;
    }
    if (ufi[g.npts] < 0.0) {
      for (int ii = 1; ii <= g.npts - 1; ii++)
        if (ufi[g.npts - ii] >= 0.0) {
          i2 = g.npts - ii;
          break;
        }
    } else {
      i2 = g.npts;
    }
    g.npts = i2 - i1 + 1;
  }
  int n = g.npts;
  g.x.assign(n, Vec4());
  g.k.assign(n, Vec4());
  g.lambda.assign(n, 0.0);
  g.tpmarr.assign(n, 0);
  g.tprarr.assign(n, 0);
  for (int j = 0; j < n; j++) {
    g.x[j].d[1] = 1.0 / ufi[i1 + j];
    g.x[j].d[2] = std::acos(mufi[i1 + j]);
    g.x[j].d[3] = phi0 - dphi[i1 + j];
    if (std::fabs(mu0) == 1.0) g.x[j].d[3] = g.x[j].d[3] + fsign(1.0, mu0) * std::atan2(beta, alpha);
    if (n != 1) {
      g.lambda[j] = lambdai[n] - lambdai[i1 + j];  // (sic) LAMBDAI(g%npts) with the trimmed count
      g.x[j].d[0] = dti[1] - dti[i1 + j] - t0;
      g.tpmarr[j] = tpmi[i1 + j];
      g.tprarr[j] = tpri[i1 + j];
    } else {
      g.lambda[j] = lambdai[1];
      g.x[j].d[0] = dti[1];
      g.tpmarr[j] = tpmi[1];
      g.tprarr[j] = tpri[1];
    }
  }
  double sumr = 0.0;
  for (int j = 0; j < n; j++) {
    double r = g.x[j].d[1];
    double sgr = su * ((g.tprarr[j] & 1) ? -1.0 : 1.0);
    double sgm = sm * ((g.tpmarr[j] & 1) ? -1.0 : 1.0);
    calc_nullp(q2, l, a, r, std::cos(g.x[j].d[2]), sgr, sgm, g.k[j].d);
    sumr += r;
  }
  if (sumr < 0.) status = -1;
}

}  // namespace orc
