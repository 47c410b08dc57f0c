// grtrans_b200 -- emissivity / absorption / Faraday coefficients as fused per-(point,frequency)
// device functions.
//
// Replaces emis.f90:461-571 (calc_emissivity dispatch + copy-out), :153-184 (bbemis, fbbpolemis),
// polsynchemis.f90:700-863 (polsynchth), :527-631 (polsynchpl), :633-698 (synchpl), :865-904
// (synchemis), :1014-1032 (bnu), bessel.f90 (K0, K1, K2 by NR polynomials + recurrence) and
// chandra_tab24.f90:29-38.  Output order (jI,jQ,jU,jV, aI,aQ,aU,aV, rhoQ,rhoU,rhoV).
#pragma once
#include "grt_math.cuh"

namespace grt {

enum { EM_POLSYNCHTH = 2, EM_SYNCHPL = 3, EM_POLSYNCHPL = 4, EM_BB = 8, EM_BBPOL = 9, EM_SYNCHTHAV = 13 };

namespace pc {
constexpr double h = 6.626e-27, k = 1.38e-16, c = 2.99792458e10, e = 4.8032e-10, G = 6.67e-8, m = 9.10938188e-28,
                 mp = 1.67262158e-24, c2 = c * c, msun = 1.998e33;
}

struct EmisTables {      // device pointers (log-valued tables, built on the host)
  const double *gx, *gy;  // 181, 3*181
  const double *xv, *ya, *yp, *yv, *yap, *yav;  // 81, 5 x 3*81
  const float *ch_mu, *ch_I, *ch_del;           // 21 each
};

GRT_HD double poly7(double x, const double *c) {
  double p = c[6];
  for (int i = 5; i >= 0; i--) p = x * p + c[i];
  return p;
}
GRT_HD double poly9(double x, const double *c) {
  double p = c[8];
  for (int i = 7; i >= 0; i--) p = x * p + c[i];
  return p;
}
GRT_HD double bess_i0(double x) {
  const double p[7] = {1.00, 3.51562290, 3.08994240, 1.20674920, 0.26597320, 0.360768e-10, 0.45813e-20};
  const double q[9] = {0.398942280,   0.1328592e-10, 0.225319e-20,   -0.157565e-20, 0.916281e-20,
                       -0.2057706e-10, 0.2635537e-10, -0.1647633e-10, 0.392377e-20};
  double ax = fabs(x);
  if (ax < 3.75) return poly7(sq(x / 3.750), p);
  return (exp(ax) / sqrt(ax)) * poly9(3.750 / ax, q);
}
GRT_HD double bess_i1(double x) {
  const double p[7] = {0.50, 0.878905940, 0.514988690, 0.150849340, 0.2658733e-10, 0.301532e-20, 0.32411e-30};
  const double q[9] = {0.398942280,  -0.3988024e-10, -0.362018e-20, 0.163801e-20, -0.1031555e-10,
                       0.2282967e-10, -0.2895312e-10, 0.1787654e-10, -0.420059e-20};
  double ax = fabs(x), r;
  if (ax < 3.75)
    r = ax * poly7(sq(x / 3.750), p);
  else
    r = (exp(ax) / sqrt(ax)) * poly9(3.750 / ax, q);
  return x < 0.0 ? -r : r;
}
GRT_HD double bess_k0(double x) {
  const double p[7] = {-0.57721566, 0.42278420, 0.23069756, 0.3488590e-1, 0.262698e-2, 0.10750e-3, 0.74e-5};
  const double q[7] = {1.25331414, -0.7832358e-1, 0.2189568e-1, -0.1062446e-1, 0.587872e-2, -0.251540e-2, 0.53208e-3};
  if (x <= 2.0) return (-log(x / 2.0) * bess_i0(x)) + poly7(x * x / 4.0, p);
  return (exp(-x) / sqrt(x)) * poly7(2.0 / x, q);
}
GRT_HD double bess_k1(double x) {
  const double p[7] = {1.0, 0.15443144, -0.67278579, -0.18156897, -0.1919402e-1, -0.110404e-2, -0.4686e-4};
  const double q[7] = {1.25331414, 0.23498619, -0.3655620e-1, 0.1504268e-1, -0.780353e-2, 0.325614e-2, -0.68245e-3};
  if (x <= 2.0) return (log(x / 2.0) * bess_i1(x)) + (1.0 / x) * poly7(x * x / 4.0, p);
  return (exp(-x) / sqrt(x)) * poly7(2.0 / x, q);
}

GRT_HD double planck_bnu(double T, double nu) {
  double b;
  if (pc::h * nu / pc::k / T < 1e-6)
    b = 2.0 * nu * nu * pc::k * T / pc::c2;
  else
    b = 2.0 * pc::h * nu / pc::c2 * nu * nu / (exp(pc::h * nu / pc::k / T) - 1.0);
  if (b == 0) b = 2.220446049250313e-16;
  return b;
}

// ---- thermal polarised synchrotron ------------------------------------------------------------------
// Frequency-independent part (per geodesic point) hoisted out of the frequency loop.
struct ThermalPoint {
  double thetae, n, b, t, sinth, costh, tanth, k0, k1ok2, ik2, step, nucfac;
};
GRT_HD void thermal_point(double n, double b, double t, double theta, ThermalPoint &P) {
  P.n = n; P.b = b; P.t = t;
  P.thetae = pc::k * t / pc::m / pc::c / pc::c + 1e-10;
  P.sinth = sin(theta);
  P.costh = cos(theta);
  P.tanth = tan(theta);
  P.nucfac = 3.0 * pc::e * b * P.sinth / 4.0 / GRT_PI / pc::m / pc::c * P.thetae * P.thetae + 1.0;
  if (P.thetae > 1e-2) {
    double x = 1.0 / P.thetae;
    double k0 = bess_k0(x), k1 = bess_k1(x);
    double k2 = k0 + 1.0 * (2.0 / x) * k1;  // besselk(2,x): one step of the upward recurrence
    P.k0 = k0;
    P.k1ok2 = k1 / k2;
    P.ik2 = 1.0 / k2;
    P.step = 0.5 + 0.5 * tanh((P.thetae - 1.0) / 0.05);
  } else {
    P.k0 = P.k1ok2 = P.ik2 = P.step = 0.0;
  }
}

GRT_HD void polsynchth_eval(const ThermalPoint &P, double nu, double *e) {
  const double ec = pc::e, c = pc::c, m = pc::m, pi = GRT_PI;
  double xm = nu / P.nucfac;
  double x13 = pow(xm, 1.0 / 3.0), x23 = pow(xm, 2.0 / 3.0);
  double ex = exp(-1.8899 * x13);
  double iix = 2.5651 * (1.0 + 1.92 / x13 + .9977 / x23) * ex;
  double iqx = 2.5651 * (1.0 + .93193 / x13 + .499873 / x23) * ex;
  double ivx = (1.81384 / xm + 3.42319 / x23 + 0.0292545 / sqrt(xm) + 2.03773 / x13) * ex;
  double th2 = P.thetae * P.thetae;
  double ji = ec * ec / c / sqrt(3.0) / 2.0 * P.n / th2 * nu * iix;
  double jq = ec * ec / c / sqrt(3.0) / 2.0 * P.n / th2 * nu * iqx;
  double jv = 4.0 * ec * ec / c / 3.0 / sqrt(3.0) / P.tanth * P.n / 2.0 / (th2 * P.thetae) * nu * ivx;
  double bn = planck_bnu(P.t, nu);
  double wp2 = 4.0 * pi * P.n * ec * ec / m;
  double omega0 = ec * P.b / m / c;
  double xarg = P.thetae * sqrt(sqrt(2.0) * P.sinth * (1e3 * omega0 / 2.0 / pi / nu));
  double extraterm = (.011 * exp(-xarg / 47.2) - pow(2.0, -1.0 / 3.0) / pow(3.0, 23.0 / 6.0) * pi * 1e4 * pow(xarg + 1e-16, -8.0 / 3.0)) *
                     (0.5 + 0.5 * tanh((log(xarg) - log(120.0)) / 0.1));
  double jff = 2.011 * exp(-pow(xarg, 1.035) / 4.7) - cos(xarg / 2.0) * exp(-pow(xarg, 1.2) / 2.73) - .011 * exp(-xarg / 47.2) +
               extraterm;
  double w = 2.0 * pi * nu, w3 = w * w * w, w4 = w3 * w;
  double eps11m22, eps12;
  if (P.thetae > 1e-2) {
    double shgm = (double)0.43793091f * log(1.0 + (double)0.00185777f * pow(xarg, (double)1.50316886f));
    eps11m22 = jff * wp2 * omega0 * omega0 / w4 * (P.k1ok2 + 6.0 * P.thetae) * P.sinth * P.sinth;
    eps12 = wp2 * omega0 / w3 * (P.k0 - P.step * shgm) * P.ik2 * P.costh;
  } else {
    eps11m22 = jff * wp2 * omega0 * omega0 / w4 * (1.0 + 6.0 * P.thetae) * P.sinth * P.sinth;
    eps12 = wp2 * omega0 / w3 * P.costh;
  }
  e[0] = ji; e[1] = jq; e[2] = 0.0; e[3] = jv;
  e[4] = ji / bn; e[5] = jq / bn; e[6] = 0.0 / bn; e[7] = jv / bn;
  e[8] = 2.0 * pi * nu / 2.0 / c * eps11m22;
  e[9] = e[6];
  e[10] = 2.0 * pi * nu / c * eps12;
}

// ---- power-law synchrotron ------------------------------------------------------------------------
GRT_HD double pl_lookup(double lx, int p_iy, const double *xv, int nx, const double *yv) {
  int ix = (int)((lx - xv[0]) / (xv[nx - 1] - xv[0]) * (nx));
  ix = (ix < (nx - 1)) ? ((ix > 1) ? ix : 1) : (nx - 1);
  int indx = p_iy * nx + ix;
  double yix = yv[indx - 1], yix1 = yv[indx], xix1 = xv[ix], xix = xv[ix - 1];
  return exp(yix + (yix1 - yix) / (xix1 - xix) * (lx - xix));
}

template <bool POL>
GRT_HD void synchpl_eval(const EmisTables &T, double nu, double nnth, double b, double th, double p, double gmin,
                         double gmax, double *e) {
  const double ec = pc::e, c = pc::c, m = pc::m, pi = GRT_PI;
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  double sinth = POL ? sin(th) + 1e-10 : sin(th);
  double nubperp = ec * b / m / c / 2.0 / pi * sinth + 1e-10;
  double nucmin = 3.0 / 2.0 * nubperp * gmin * gmin, nucmax = 3.0 / 2.0 * nubperp * gmax * gmax;
  double omega0 = nubperp * 2.0 * pi, omega = nu * 2.0 * pi;
  double lxmin = log(nu / nucmin), lxmax = log(nu / nucmax);
  double A = (p - 1.0) * nnth / (pow(gmin, 1.0 - p) - pow(gmax, 1.0 - p));
  int iy = (p < 5.0) ? ((p > 3.2) ? 1 : 0) : 2;
  double gxfac = pl_lookup(lxmax, iy, T.gx, 181, T.gy) - pl_lookup(lxmin, iy, T.gx, 181, T.gy);
  double gafac = pl_lookup(lxmax, iy, T.xv, 81, T.ya) - pl_lookup(lxmin, iy, T.xv, 81, T.ya);
  double s3 = POL ? sqrt(3.0) : (double)sqrtf(3.f);
  double jfac = A * ec * ec / c * s3 / 4.0 * pow(3.0 * nubperp / 2.0 / nu, (p - 1.0) / 2.0) * nubperp;
  double afac = pow(2.0 * pi, 3) * A * ec * ec * sqrt(3.0) * omega0 * (p + 2.0) / 32.0 / (pi * pi) / m / c / (omega * omega) *
                pow(2.0 * omega / 3.0 / omega0, -p / 2.0);
  e[0] = jfac * gxfac;
  e[4] = afac * gafac;
  if (POL) {
    double tanth = tan(th) + (signbit(cos(th)) ? -1.0 : 1.0) * 1e-10;
    double gpfac = pl_lookup(lxmax, iy, T.xv, 81, T.yp) - pl_lookup(lxmin, iy, T.xv, 81, T.yp);
    double gapfac = pl_lookup(lxmax, iy, T.xv, 81, T.yap) - pl_lookup(lxmin, iy, T.xv, 81, T.yap);
    double gvfac = pl_lookup(lxmax, iy, T.xv, 81, T.yv) - pl_lookup(lxmin, iy, T.xv, 81, T.yv);
    double gavfac = pl_lookup(lxmax, iy, T.xv, 81, T.yav) - pl_lookup(lxmin, iy, T.xv, 81, T.yav);
    e[1] = jfac * gpfac;
    e[3] = jfac * 4.0 / 3.0 / tanth * sqrt(3.0 * nubperp / 2.0 / nu) * gvfac;
    double alpha = (p - 1) / 2.0;
    double kperp = A * ec * ec / m / c / nubperp;
    double nui = gmin * gmin * nubperp;
    e[8] = 1.0 * kperp * pow(nubperp / nu, 3.0) * pow(gmin, -2.0 * alpha + 1.0) * (1.0 - pow(nui / nu, alpha - 1.0 / 2.0)) *
           pow(alpha - 1.0 / 2.0, -1.0);
    e[10] = 2.0 * (alpha + 3.0 / 2.0) / (alpha + 1) * kperp * pow(nubperp / nu, 2.0) * log(gmin) * pow(gmin, -2 * (alpha + 1)) / tanth;
    e[5] = afac * gapfac;
    e[7] = afac * 4.0 / 3.0 / tanth * gavfac * pow(2.0 * omega / 3.0 / omega0, -1.0 / 2.0);
  }
}

// angle-averaged thermal synchrotron (Mahadevan et al. 1996)
GRT_HD void synchemis_eval(double nu, double n, double B, double T, double *e) {
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  double thetae = pc::k * T / pc::m / pc::c2 + 1e-10;
  double nucrit = 3.0 * pc::e * B / (4 * GRT_PI * pc::m * pc::c) * pow(thetae, 2.0) + 1.0;
  double xm = nu / nucrit;
  double lx = log(xm);
  double ipx = 4.0505 / exp(lx / 6.) * (1 + .40 / sqrt(sqrt(xm)) + 0.5316 / sqrt(xm)) * exp(-1.8899 * exp(lx / 3.));
  double jnu = 4.43e-30 / 2.0 * nu * n * ipx / (thetae * thetae);
  e[0] = jnu;
  e[4] = (fabs(jnu) > 0.) ? jnu / planck_bnu(T, nu) : 0.0;
}

// Chandrasekhar (1960) Table XXIV limb darkening / polarisation degree, float32 linear interpolation
GRT_HD void chandra_interp(const EmisTables &T, float mu, float &I, float &del) {
  const int n = 21;
  int jl = 0, ju = n + 1;
  bool ascnd = T.ch_mu[n - 1] >= T.ch_mu[0];
  while (ju - jl > 1) {
    int jm = (ju + jl) >> 1;
    if (ascnd == (mu >= T.ch_mu[jm - 1]))
      jl = jm;
    else
      ju = jm;
  }
  if (mu == T.ch_mu[0]) jl = 1;
  else if (mu == T.ch_mu[n - 1]) jl = n - 1;
  int j = jl < 1 ? 1 : (jl > n - 1 ? n - 1 : jl);
  float w = (mu - T.ch_mu[j - 1]) / (T.ch_mu[j] - T.ch_mu[j - 1]);
  I = (1.f - w) * T.ch_I[j - 1] + w * T.ch_I[j];
  del = (1.f - w) * T.ch_del[j - 1] + w * T.ch_del[j];
}

}  // namespace grt
