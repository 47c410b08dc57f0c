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

enum { EM_POLSYNCHTH = 2, EM_SYNCHPL = 3, EM_POLSYNCHPL = 4, EM_BB = 8, EM_BBPOL = 9, EM_SYNCHTHAV = 13,
       EM_HYBRIDTHPL = 20, EM_HYBRIDTH = 21, EM_HYBRIDPL = 22,
       // unpolarised thermal models sharing SYNCHTHAV's point record (n, B, T) and kernel slot (em_slot)
       EM_BREMS = 5, EM_SYNCHTHAVNOABS = 14, EM_SYNCHTHBREMS = 27, EM_RHO = 12,
       EM_POLSYNCHSYMTH = 15,   // Pandya et al. (2016) fits (sympolemisth)
       EM_MAXJUTT = 23, EM_MAXCOMP = 24,    // sums of thermal components (calc_maxjutt.f90, calc_maxcomp.f90)
       // non-thermal electrons in bins of Lorentz factor (KORALNTH; emis.f90:520-526) [+ angle-averaged thermal synchrotron]
       EM_BINS = 25, EM_HYBRIDTHBINS = 26 };
#define GRT_MAXJUTT_NCOMP 7   // at most 7 components (8 ep%otherargs: delta [, selection], weights)
#define GRT_MAXBINS 64        // at most 64 Lorentz-factor bins in a point record (12 + 64 doubles)
// kernel / record slot of an emissivity type: the (n, B, T) scalar models run in the SYNCHTHAV instantiations and pick
// their formula from TransferParams::etype at run time (a thread-uniform switch)
GRT_HD constexpr int em_slot(int et) {
  return (et == EM_BREMS || et == EM_SYNCHTHAVNOABS || et == EM_SYNCHTHBREMS || et == EM_RHO) ? EM_SYNCHTHAV
                                                                                           : (et == EM_MAXCOMP ? EM_MAXJUTT : (et == EM_HYBRIDTHBINS ? EM_BINS : et));
}

namespace pc {
constexpr double h = 6.626e-27, k = 1.38e-16, c = 2.99792458e10, e = 4.8032e-10, G = 6.67e-8, m = 9.10938188e-28,
                 mp = 1.67262158e-24, c2 = c * c, msun = 1.998e33;
}

struct EmisTables {      // device pointers (log-valued tables, built on the host)
  const double *gx, *gy;  // 181, 3*181
  const double *xv, *ya, *yp, *yv, *yap, *yav;  // 81, 5 x 3*81
  const float *ch_mu, *ch_I, *ch_del;           // 21 each
  // Faraday-fit table (optional, nullptr = evaluate the fits directly): cubic pieces of jffunc(xarg) and of
  // log(1 + 0.00185777 xarg^1.50316886) on a uniform grid in u = log(xarg); far[k*8 + 0..3] / [4..7] = a,b,c,d of piece k,
  // value = a + w (b + w (c + w d)), w = (u - far_u0) far_ih - k
  const double *far;
  double far_u0, far_ih;
  int far_n;
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
// polsynchth (polsynchemis.f90:700-863) is evaluated for every (geodesic point, frequency) pair, i.e. ~3e10 times for
// the headline workload, so it is split into
//   thermal_point()  per geodesic point: everything that does not depend on frequency, with the redshift g and the
//                    length unit LBH of the invariant scaling (emis.f90:831-838, grtrans_driver.f90:217) folded in;
//   thermal_eval()   per (point, frequency): 11 exp, 1 log, 1 cos, 1 sqrt, 1 reciprocal and ~90 multiply-adds.
// Every power x**p of the reference becomes exp(p*log x) with the logarithm assembled from per-point and per-frequency
// logarithms; divisions by constants-of-the-point are pre-inverted.  The algebra is otherwise the reference's, term by term
// (fits iix/iqx/ivx :831-861, jffunc/shgmfunc :737-761, Kirchhoff alpha=j/B_nu), so results agree to round-off (~1e-13).
//
// Output of thermal_eval are the INVARIANT coefficients of frequency nu_obs: j*g^2 and K/g*LBH evaluated at the comoving
// frequency nu = nu_obs/g (for the unit-level surface g = LBH = 1 gives the plain comoving polsynchth values).
struct ThermalPoint {  // 10 doubles (+ g), stored in the point record
  double xp13;         // (g nu_c')^(-1/3):  (nu/nu_c')^(1/3) = nu_obs^(1/3) xp13
  double ixp13, xpm32; // 1/xp13 and xp13^(-3/2): (nu/nu_c')^(-1/3), (nu/nu_c')^(-1/2) without a reciprocal / square root
  double AI, AV;       // comoving j_{I,Q} = AI nu_obs {iix,iqx},  j_V = AV nu_obs ivx   (the 1/g of nu = nu_obs/g folded in)
  double hkT_g;        // h/(k T g): h nu/(k T) = hkT_g nu_obs
  double RQ;           // rho_Q LBH/g = RQ jff / nu_obs^3
  double RV0, RV1;     // rho_V LBH/g = (RV0 - RV1 shgm) / nu_obs^2
  double g;            // redshift
  double lAx;          // log(Ax), xarg = Ax nu_obs^(-1/2): log(xarg) = lAx - 0.5 log(nu_obs), the abscissa of the Faraday fits
};

// Per observed frequency.  Every x**p of the reference with x = (per-point factor) * nu**q is evaluated as
// (per-point factor)**p * nu**(p q): both powers are computed once (per point in k_fluid_ortho, per frequency here), the
// hot loop only multiplies them.
struct FreqConst {
  double nu, inu2, inu3;
  double n13, in13;             // nu^(1/3), nu^(-1/3)
  double nm12;                  // nu^(-1/2)
  double lnm12;                 // log(nu^(-1/2))
};
GRT_HD FreqConst freq_const(double nu_obs) {
  FreqConst f;
  f.nu = nu_obs;
  double i = 1.0 / nu_obs;
  f.inu2 = i * i;
  f.inu3 = f.inu2 * i;
  const double l = log(nu_obs);
  f.n13 = exp(l * (1.0 / 3.0));
  f.in13 = exp(l * (-1.0 / 3.0));
  f.nm12 = exp(-0.5 * l);
  f.lnm12 = -0.5 * l;
  return f;
}

GRT_HD void thermal_point(double n, double b, double t, double theta, double g, double lbh, ThermalPoint &P) {
  const double ec = pc::e, c = pc::c, m = pc::m, pi = GRT_PI;
  double thetae = pc::k * t / pc::m / pc::c / pc::c + 1e-10;
  double sinth = sin(theta), costh = cos(theta), tanth = tan(theta);
  double nucfac = 3.0 * pc::e * b * sinth / 4.0 / GRT_PI / pc::m / pc::c * thetae * thetae + 1.0;
  double th2 = thetae * thetae;
  const double lxp = -log(g) - log(nucfac);
  P.xp13 = exp(lxp * (1.0 / 3.0));
  P.ixp13 = exp(lxp * (-1.0 / 3.0));
  P.xpm32 = exp(lxp * (-0.5));
  P.AI = ec * ec / c / sqrt(3.0) / 2.0 * n / th2 / g;
  P.AV = 4.0 * ec * ec / c / 3.0 / sqrt(3.0) / tanth * n / 2.0 / (th2 * thetae) / g;
  P.hkT_g = pc::h / pc::k / t / g;
  double wp2 = 4.0 * pi * n * ec * ec / m;
  double omega0 = ec * b / m / c;
  // xarg = thetae*sqrt(sqrt(2) sin(theta) 1e3 omega0/(2 pi nu)),  nu = nu_obs/g
  P.lAx = log(thetae * sqrt(sqrt(2.0) * sinth * (1e3 * omega0 / 2.0 / pi) * g));
  double X, Z0, Z1;
  if (thetae > 1e-2) {
    double x = 1.0 / thetae;
    double k0 = bess_k0(x), k1 = bess_k1(x);
    double k2 = k0 + 1.0 * (2.0 / x) * k1;  // besselk(2,x): one step of the upward recurrence
    X = k1 / k2 + 6.0 * thetae;
    Z0 = k0 / k2;
    Z1 = (0.5 + 0.5 * tanh((thetae - 1.0) / 0.05)) / k2;
  } else {
    X = 1.0 + 6.0 * thetae;
    Z0 = 1.0;
    Z1 = 0.0;
  }
  P.RQ = wp2 * omega0 * omega0 * X * sinth * sinth / (16.0 * pi * pi * pi * c) * g * g * lbh;
  double rvc = wp2 * omega0 * costh / (4.0 * pi * pi * c) * g * lbh;
  P.RV0 = rvc * Z0;
  P.RV1 = rvc * Z1;
  P.g = g;
}

// e = (jI,jQ,jU,jV, aI,aQ,aU,aV, rhoQ,rhoU,rhoV) in invariant form (see above); jraw = comoving j_I.  kcl = c^2 LBH / (2 h).
// The emissivities are formed in the reference's order -- comoving j first, then j g^2 and alpha = j/B_nu scaled by LBH/g --
// so that they underflow (exp(-1.89 x^(1/3)) reaches 1e-320 in cold, weakly magnetised zones) where the reference's do: a
// DELO step divides by alpha_I, and "j_inv == 0 but alpha_I != 0" must stay distinguishable from "both 0".
// Per call: 5 exp, 1 cos, 1 sqrt, 1 reciprocal (+ 1 log where the Faraday-rotation fit is switched on, + 1 division inside
// the 16 < xarg < 6552 window of the high-frequency correction) and ~110 multiply-adds.
// jffunc (with its high-frequency correction) and the logarithm of shgmfunc, evaluated as written (polsynchemis.f90:737-761),
// from lx = log(xarg) (powers xarg^p = exp(p lx)).  Out of line: with the
// Faraday-fit table in place this is the rare path (xarg outside the table, or the table switched off).
GRT_HDN void faraday_fits(double lx, bool want_log, double &jff, double &lt) {
  const double xarg = exp(lx), x1 = exp(1.035 * lx), x2 = exp(1.2 * lx), x3 = exp((double)1.50316886f * lx),
               x4 = exp((-8.0 / 3.0) * lx);
  const double e1 = .011 * exp_fast(xarg * (-1.0 / 47.2));
  jff = 2.011 * exp_fast(x1 * (-1.0 / 4.7)) - cos(0.5 * xarg) * exp_fast(x2 * (-1.0 / 2.73)) - e1;
  // extraterm * (0.5 + 0.5 tanh((log(xarg) - log(120))/0.1)) = extraterm * t^20/(1 + t^20), t = xarg/120; the switch is
  // exactly 0 in FP64 below t = e^-2 and exactly 1 above t = e^4
  const double t = xarg * (1.0 / 120.0);
  if (t > 0.1353352832366127) {
    double sw = 1.0;
    if (t < 54.598150033144236) {
      const double t2 = t * t, t4 = t2 * t2, t5 = t4 * t, t10 = t5 * t5, t20 = t10 * t10;
      sw = t20 / (1.0 + t20);
    }
    const double c83 = 0.7937005259840998 / 67.44733739010397 * GRT_PI * 1e4;  // 2^(-1/3) / 3^(23/6) pi 1e4
    jff += (e1 - c83 * x4) * sw;  // (xarg + 1e-16)^(-8/3): the 1e-16 is below rounding for xarg > 16
  }
  lt = want_log ? log_ge1(1.0 + (double)0.00185777f * x3) : 0.0;
}

GRT_HD void thermal_eval(const ThermalPoint &P, const FreqConst &F, double kcl, double *e, double &jraw,
                         const EmisTables &FT) {
  // Faraday-fit table: the eight coefficients of this (point, frequency) are fetched first, with a clamped index so that the
  // loads need no branch; their latency is hidden behind the synchrotron part below
  bool tabulated = false;
  double tw = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0, q5 = 0.0, q6 = 0.0, q7 = 0.0;
  if (FT.far) {  // (uniform)
    const double u = (P.lAx + F.lnm12 - FT.far_u0) * FT.far_ih;
    tabulated = u >= 0.0 && u < (double)FT.far_n;  // (NaN: direct evaluation, which propagates it)
    const int k = tabulated ? (int)u : 0;
    tw = u - (double)k;
#if defined(__CUDA_ARCH__)
    const double2 *q = reinterpret_cast<const double2 *>(FT.far) + (size_t)k * 4;
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    q0 = a.x; q1 = a.y; q2 = b.x; q3 = b.y; q4 = c.x; q5 = c.y; q6 = d.x; q7 = d.y;
#else
    const double *q = FT.far + (size_t)k * 8;
    q0 = q[0]; q1 = q[1]; q2 = q[2]; q3 = q[3]; q4 = q[4]; q5 = q[5]; q6 = q[6]; q7 = q[7];
#endif
  }
  const double x13 = F.n13 * P.xp13;
  const double i13 = F.in13 * P.ixp13, i23 = i13 * i13;
  const double ex = exp_fast(-1.8899 * x13);
  const double iix = 2.5651 * (1.0 + 1.92 * i13 + .9977 * i23) * ex;
  const double iqx = 2.5651 * (1.0 + .93193 * i13 + .499873 * i23) * ex;
  const double ivx = (1.81384 * (i23 * i13) + 3.42319 * i23 + 0.0292545 * (F.nm12 * P.xpm32) + 2.03773 * i13) * ex;
  const double ji = P.AI * F.nu * iix, jq = P.AI * F.nu * iqx, jv = P.AV * F.nu * ivx;
  jraw = ji;
  // 1/B_nu (bnu, polsynchemis.f90:1014-1032): Rayleigh-Jeans below h nu/kT = 1e-6; B_nu == 0 is replaced by epsilon
  // (the Planck exponential stays inside its branch: most (point, frequency) pairs of a hot flow are in the Rayleigh-Jeans
  // regime h nu / k T < 1e-6 and never evaluate it -- hoisting it next to the first exp cost 20 ms per step)
  const double xh = P.hkT_g * F.nu;
  const double g2 = P.g * P.g;
  double ka;  // alpha_S LBH/g = j_S * ka
  if (xh < 1e-6)
    ka = kcl * g2 * xh * F.inu3;
  else if (xh < 600.0)
    ka = kcl * g2 * (exp_fast(xh) - 1.0) * F.inu3;
  else {  // deep Wien tail: follow the reference literally (B_nu underflowing to 0 is replaced by epsilon)
    double ig = 1.0 / P.g, nuc = F.nu * ig;
    double bn = 2.0 * pc::h * nuc / pc::c2 * nuc * nuc / (exp(xh) - 1.0);
    if (bn == 0.0) bn = 2.220446049250313e-16;
    ka = ig * (kcl * 2.0 * pc::h / pc::c2) / bn;
  }
  e[0] = ji * g2; e[1] = jq * g2; e[2] = 0.0; e[3] = jv * g2;
  e[4] = ji * ka; e[5] = jq * ka; e[6] = 0.0; e[7] = jv * ka;
  // Faraday rotation / conversion (Shcherbakov 2008 fits jffunc, shgmfunc): both are functions of xarg alone
  double jff, lt = 0.0;
  if (tabulated) {
    jff = q0 + tw * (q1 + tw * (q2 + tw * q3));
    lt = q4 + tw * (q5 + tw * (q6 + tw * q7));
  }
  if (!tabulated) faraday_fits(P.lAx + F.lnm12, P.RV1 != 0.0, jff, lt);
  e[8] = P.RQ * jff * F.inu3;
  e[9] = 0.0;
  double rv = P.RV0;
  if (P.RV1 != 0.0) rv -= P.RV1 * ((double)0.43793091f * lt);
  e[10] = rv * F.inu2;
}

// ---- POLSYNCHSYMTH: sympolemisth (polsynchemis.f90:915-1012), Pandya et al. (2016) fits for j_I, j_Q, j_V with the Faraday
// coefficients of polsynchth.  Per point: everything but the powers of x = nu/(2/9 nu_c thetae^2 sin theta).
struct SymPoint {
  double Xs;          // x = nu Xs
  double A;           // n e^2/c nu_c
  double S;           // sqrt(2) pi/27 sin(theta)
  double cq;          // (7 thetae^(24/25) + 35)/(10 thetae^(24/25) + 75) 2^(11/12)
  double v0, v1;      // (37 - 87 sin(theta - 28/25))/100/(thetae + 1);  thetae^(3/5)/25 + 7/10
  double T;           // for B_nu
  double lAx, RQ, RV0, RV1;  // Faraday part as in ThermalPoint (g = LBH = 1)
};
GRT_HD void sym_point(double n, double b, double t, double theta, SymPoint &Q) {
  ThermalPoint P;
  thermal_point(n, b, t, theta, 1.0, 1.0, P);
  Q.lAx = P.lAx; Q.RQ = P.RQ; Q.RV0 = P.RV0; Q.RV1 = P.RV1; Q.T = t;
  const double thetae = pc::k * t / pc::m / (pc::c * pc::c) + 1e-10;
  const double nuc = pc::e * b / pc::m / pc::c / 2.0 / GRT_PI + 1.0;
  const double sinth = sin(theta);
  Q.Xs = 1.0 / (2.0 / 9.0 * nuc * (thetae * thetae) * sinth);
  Q.A = n * (pc::e * pc::e) / pc::c * nuc;
  Q.S = sqrt(2.0) * GRT_PI / 27.0 * sinth;
  const double t2425 = pow(thetae, 24.0 / 25.0);
  Q.cq = (7.0 * t2425 + 35.0) / (10.0 * t2425 + 75.0) * pow(2.0, 11.0 / 12.0);
  Q.v0 = (37.0 - 87.0 * sin(theta - 28.0 / 25.0)) / 100.0 / (thetae + 1.0);
  Q.v1 = pow(thetae, 3.0 / 5.0) / 25.0 + 7.0 / 10.0;
}
// comoving coefficients at frequency nu
GRT_HD void sym_eval(const SymPoint &Q, double nu, const EmisTables &FT, double *e) {
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  const double x = nu * Q.Xs, lx = log(x);
  const double sx = sqrt(x), x16 = exp(lx * (1.0 / 6.0));
  const double ti = sx + pow(2.0, 11.0 / 12.0) * x16, tq = sx + Q.cq * x16;
  const double jis = Q.S * (ti * ti), jqs = -Q.S * (tq * tq);
  const double jvs = -Q.v0 * exp((5.0 / 3.0) * log(1.0 + Q.v1 * exp(lx * (9.0 / 25.0))));
  const double fac = Q.A * exp(-exp(lx * (1.0 / 3.0)));
  const double ji = fac * jis, jq = -fac * jqs, jv = -fac * jvs;
  const double ib = 1.0 / planck_bnu(Q.T, nu);
  e[0] = ji; e[1] = jq; e[3] = jv;
  e[4] = ji * ib; e[5] = jq * ib; e[7] = jv * ib;
  // Faraday coefficients: the fits of polsynchth in xarg = Ax nu^(-1/2) (table or direct, as in thermal_eval)
  const double lxa = Q.lAx - 0.5 * log(nu);
  double jff = 0.0, lt = 0.0;
  bool tabulated = false;
  if (FT.far) {
    const double u = (lxa - FT.far_u0) * FT.far_ih;
    if (u >= 0.0 && u < (double)FT.far_n) {
      const int k = (int)u;
      const double w = u - (double)k;
      const double *q = FT.far + (size_t)k * 8;
      jff = q[0] + w * (q[1] + w * (q[2] + w * q[3]));
      lt = q[4] + w * (q[5] + w * (q[6] + w * q[7]));
      tabulated = true;
    }
  }
  if (!tabulated) faraday_fits(lxa, Q.RV1 != 0.0, jff, lt);
  const double inu = 1.0 / nu;
  e[8] = Q.RQ * jff * (inu * inu * inu);
  double rv = Q.RV0;
  if (Q.RV1 != 0.0) rv -= Q.RV1 * ((double)0.43793091f * lt);
  e[10] = rv * (inu * inu);
}

// plain comoving polsynchth values at frequency nu (the polsynchemis f2py surface): g = LBH = 1
GRT_HD void polsynchth_eval(double nu, double n, double b, double t, double theta, double *e) {
  ThermalPoint P;
  thermal_point(n, b, t, theta, 1.0, 1.0, P);
  double jraw;
  EmisTables none;
  none.far = nullptr; none.far_u0 = none.far_ih = 0.0; none.far_n = 0;
  thermal_eval(P, freq_const(nu), pc::c2 / (2.0 * pc::h), e, jraw, none);
}

// ---- power-law synchrotron ------------------------------------------------------------------------
// polsynchpl / synchpl (polsynchemis.f90:527-631, 633-698) with the tabulated G(x) integrals (:62-92, 107-525).  Like
// the thermal model it is split into per-run constants (p, gmin, gmax), a per-point part (field, angle, density) and the
// per-(point, frequency) part, which is 15 exp (12 of them the log-log table lookups) and ~80 multiply-adds:
//   x**q with x = nu/(1.5 nu_B,perp)  ->  exp(q lr),  lr = log(nu_obs) - [log g + log(1.5 nu_B,perp)]
// and the five 81-point tables share one index / weight computation per abscissa.
struct PlConst {  // per run
  double p, gmin, gmax, cA, l2gmin, l2gmax, hp, am, kq, kv;
  int iy;
};
GRT_HD PlConst pl_const(double p, double gmin, double gmax) {
  PlConst c;
  c.p = p; c.gmin = gmin; c.gmax = gmax;
  c.cA = (p - 1.0) / (pow(gmin, 1.0 - p) - pow(gmax, 1.0 - p));
  c.l2gmin = 2.0 * log(gmin);
  c.l2gmax = 2.0 * log(gmax);
  c.iy = (p < 5.0) ? ((p > 3.2) ? 1 : 0) : 2;
  const double alpha = (p - 1) / 2.0;
  c.hp = (p - 1.0) / 2.0;
  c.am = alpha - 1.0 / 2.0;
  c.kq = pow(gmin, -2.0 * alpha + 1.0) * pow(alpha - 1.0 / 2.0, -1.0);
  c.kv = 2.0 * (alpha + 3.0 / 2.0) / (alpha + 1) * log(gmin) * pow(gmin, -2 * (alpha + 1));
  return c;
}
struct PlPoint {  // per geodesic point
  double nnth, nubperp, L, itanth;  // L = log(g) + log(1.5 nu_B,perp);  itanth = 1/(tan(th) + sign(cos th) 1e-10)
};
template <bool POL>
GRT_HD void pl_point(double nnth, double b, double th, double g, PlPoint &P) {
  const double ec = pc::e, c = pc::c, m = pc::m, pi = GRT_PI;
  double sinth = POL ? sin(th) + 1e-10 : sin(th);
  P.nnth = nnth;
  P.nubperp = ec * b / m / c / 2.0 / pi * sinth + 1e-10;
  P.L = log(g) + log(1.5 * P.nubperp);
  P.itanth = POL ? 1.0 / (tan(th) + (signbit(cos(th)) ? -1.0 : 1.0) * 1e-10) : 0.0;
}

struct PlAbscissa {  // position of one log-abscissa in a table
  int ix;
  double w;
};
GRT_HD PlAbscissa pl_locate(double lx, const double *xv, int nx) {
  PlAbscissa a;
  int ix = (int)((lx - xv[0]) / (xv[nx - 1] - xv[0]) * (nx));
  ix = (ix < (nx - 1)) ? ((ix > 1) ? ix : 1) : (nx - 1);
  a.ix = ix;
  a.w = (lx - xv[ix - 1]) / (xv[ix] - xv[ix - 1]);
  return a;
}
GRT_HD double pl_value(const PlAbscissa &a, int iy, int nx, const double *yv) {
  const double y0 = yv[iy * nx + a.ix - 1], y1 = yv[iy * nx + a.ix];
  return exp_fast(y0 + (y1 - y0) * a.w);
}

// comoving coefficients at frequency nu = nu_obs/g;  lnu_obs = log(nu_obs)
template <bool POL>
GRT_HD void pl_eval(const EmisTables &T, const PlConst &C, const PlPoint &P, double nu, double lnu_obs, double *e) {
  const double ec = pc::e, c = pc::c, m = pc::m, pi = GRT_PI;
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  const double lr = lnu_obs - P.L;  // log(nu / (1.5 nu_B,perp))
  const double A = C.cA * P.nnth;
  const PlAbscissa gmx = pl_locate(lr - C.l2gmax, T.gx, 181), gmn = pl_locate(lr - C.l2gmin, T.gx, 181);
  const PlAbscissa xmx = pl_locate(lr - C.l2gmax, T.xv, 81), xmn = pl_locate(lr - C.l2gmin, T.xv, 81);
  const double gxfac = pl_value(gmx, C.iy, 181, T.gy) - pl_value(gmn, C.iy, 181, T.gy);
  const double gafac = pl_value(xmx, C.iy, 81, T.ya) - pl_value(xmn, C.iy, 81, T.ya);
  const double E1 = exp_fast(-C.hp * lr);  // (3 nu_B,perp / 2 nu)^((p-1)/2)
  const double E2 = exp_fast(-0.5 * lr);   // (3 nu_B,perp / 2 nu)^(1/2)
  const double s3 = POL ? sqrt(3.0) : (double)sqrtf(3.f);
  const double jfac = A * ec * ec / c * s3 / 4.0 * E1 * P.nubperp;
  const double omega0 = P.nubperp * 2.0 * pi, omega = nu * 2.0 * pi;
  const double afac = (2.0 * pi) * (2.0 * pi) * (2.0 * pi) * A * ec * ec * sqrt(3.0) * omega0 * (C.p + 2.0) / 32.0 / (pi * pi) / m / c /
                      (omega * omega) * (E1 * E2);  // (2 omega / 3 omega0)^(-p/2)
  e[0] = jfac * gxfac;
  e[4] = afac * gafac;
  if (POL) {
    const double gpfac = pl_value(xmx, C.iy, 81, T.yp) - pl_value(xmn, C.iy, 81, T.yp);
    const double gapfac = pl_value(xmx, C.iy, 81, T.yap) - pl_value(xmn, C.iy, 81, T.yap);
    const double gvfac = pl_value(xmx, C.iy, 81, T.yv) - pl_value(xmn, C.iy, 81, T.yv);
    const double gavfac = pl_value(xmx, C.iy, 81, T.yav) - pl_value(xmn, C.iy, 81, T.yav);
    e[1] = jfac * gpfac;
    e[3] = jfac * 4.0 / 3.0 * P.itanth * E2 * gvfac;
    const double kperp = A * ec * ec / m / c / P.nubperp;
    const double rb = P.nubperp / nu;  // nu_B,perp / nu
    // (nui/nu)^(alpha-1/2), nui = gmin^2 nu_B,perp:  log(nui/nu) = 2 log gmin - lr - log 1.5
    const double En = exp_fast(C.am * (C.l2gmin - lr - 0.4054651081081644));
    e[8] = 1.0 * kperp * (rb * rb * rb) * C.kq * (1.0 - En);
    e[10] = C.kv * kperp * (rb * rb) * P.itanth;
    e[5] = afac * gapfac;
    e[7] = afac * 4.0 / 3.0 * P.itanth * gavfac * E2;
  }
}

// plain comoving values (the polsynchemis f2py surface)
template <bool POL>
GRT_HD void synchpl_eval(const EmisTables &T, double nu, double nnth, double b, double th, double p, double gmin,
                         double gmax, double *e) {
  PlPoint P;
  pl_point<POL>(nnth, b, th, 1.0, P);
  pl_eval<POL>(T, pl_const(p, gmin, gmax), P, nu, log(nu), e);
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

// synchbinemis (polsynchemis.f90:1036-1150): unpolarised synchrotron emission / absorption of electrons given as dn/dgamma
// in logarithmic bins of Lorentz factor.  fx / k53x are the reference's fits to F(x) = x int_x^inf K_5/3 and to K_5/3(x)
// (:1104-1148; single-precision literals 1./3., -5./3., 1.e-6 promoted as gfortran does).  nb[b] = dn/dgamma of bin b,
// bins[0..nbin) = bin centres, bins[nbin..2 nbin) = bin widths (emis_model_bins, emis.f90:884-908, built on the host).
GRT_HD double bins_fx(double x) {
  if (x >= 1000.) return 0.0;
  const double third = (double)(1.f / 3.f);
  const double sx = sqrt(x), cx = pow(x, third);
  const double d1 = exp(-0.97947838884478688 * x + -0.83333239129525072 * sx + 0.15541796026816246 * cx);
  const double d2 = 1. - exp(-0.0469247165562628882 * x + -0.7005501805646288 * sx + 0.0103876297841949544 * cx);
  return 2.149528241534479 * cx * d1 + 1.2533141373155 * sx * exp(-x) * d2;
}
GRT_HD double bins_k53x(double x) {
  if (x >= 1000.) return 0.0;
  if (x <= (double)1.e-6f) return 6.7e16;
  const double third = (double)(1.f / 3.f);
  const double sx = sqrt(x);
  const double d1 = exp(-1.0194198041210243 * x + 0.28011396300530672 * sx + -0.0771058491739234908 * pow(x, third));
  const double d2 = 1. - exp(-15.761577796582387 * x);
  return 1.433018827689652 * pow(x, (double)(-5.f / 3.f)) * d1 + 1.2533141373155 / sx * exp(-x) * d2;
}
GRT_HDN void bins_eval(double nu, double B, double theta, const double *nb, int nbin, const double *bins, double *e) {
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  const double s3 = (double)sqrtf(3.f), s33 = (double)(3.f * sqrtf(3.f));  // sqrt(3.), 3.*sqrt(3.) are real(4) expressions
  const double sth = sin(theta);
  const double prefj = s3 * (pc::e * pc::e * pc::e) * fabs(B) * sth / (4. * (GRT_PI * GRT_PI) * pc::m * (pc::c * pc::c));
  const double prefa = (fabs(B) > 0.) ? 4. * GRT_PI * pc::e / (s33 * fabs(B) * sth) : 0.0;
  const double nup = (3.0 * pc::e * B * sth) / (4. * GRT_PI * pc::m * pc::c);
  double sj = 0.0, sa = 0.0;
  for (int b = 0; b < nbin; b++) {
    const double gam = bins[b], dg = bins[nbin + b], g2 = gam * gam;
    const double xm = nu / (nup * g2);
    sj += bins_fx(xm) * nb[b] * dg;
    sa += bins_k53x(xm) * nb[b] * dg / (g2 * g2 * gam);
  }
  e[0] = prefj * sj;
  e[4] = prefa * sa;
}

// free-free emission / absorption (brememisHEROIC, emis.f90:188-238); literals without d0 are single precision there
GRT_HD void brems_eval(double nu, double ne, double T, double *e) {
  for (int i = 0; i < 11; i++) e[i] = 0.0;
  const double rho = ne * (double)1.67219e-24f;
  const double sqrtt = sqrt(T);
  const double thetae = pc::k * T / pc::m / pc::c2;
  const double sqrtthetae = sqrt(thetae);
  const double r10 = 1e5 / T, r2 = r10 * r10, r4 = r2 * r2, r8 = r4 * r4;
  const double tempfactor = 1. / (sqrtt + r8 * r2) + 1e-32;
  const double arg = pc::h * nu / (pc::k * T);
  double fei, fee;
  if (thetae < 1.) {
    fei = (double)1.016f * sqrtthetae * (1 + (double)1.781f * pow(thetae, (double)1.34f));
    fee = thetae * sqrtthetae * (1 + (double)1.1f * thetae + thetae * thetae * (1 - 1.25 * sqrtthetae));
  } else {
    fei = (double)1.432f * thetae * (log((double)1.123f * thetae + (double)0.48f) + 1.5);
    fee = (double)1.328f * thetae * (log((double)1.123f * thetae) + (double)1.28f);
  }
  double anu = 0.0;
  if (!(arg > 100.)) {
    const double x = (arg < (double)1.e-8f) ? arg : 1 - exp(-arg);
    anu = (1.10e61 / sqrtt) * rho * rho * fei * x * tempfactor / (nu * nu * nu);
    anu = anu + (1.14e51 / sqrtt / T) * rho * rho * fee * x * tempfactor / (nu * nu);
  }
  e[0] = anu * planck_bnu(T, nu);
  e[4] = anu;
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
