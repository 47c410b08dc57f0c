// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product
// path (grtrans_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it.
//
// CPU restatement (plain sequential C++) of the reference's semi-analytic Kerr null-geodesic
// solver "geokerr" (Dexter & Agol 2009), following the control flow of
//   /root/reference/geokerr_wrapper.f
// routine by routine, including its in-place mutation of inputs (uf, u0, tpr), the FIRSTPT
// caching of per-ray integrals, and the mu-fill output layout.  Each function cites the
// reference lines it restates.  The reference Fortran cannot be compiled in this image (no
// Fortran compiler), so this restatement is pinned instead against scipy's Carlson/Jacobi
// functions and the reference's golden images (see tests/ and DESIGN.md).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>

namespace orc {

typedef std::complex<double> cplx;

static inline double fsign(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
static inline double powm1(int n) { return (n & 1) ? -1.0 : 1.0; }  // (-1)**n for integer n

// ---------------------------------------------------------------- Carlson integrals
// geokerr_wrapper.f:3444-3489
double rf(double x, double y, double z) {
  const double ERRTOL = 0.0025, THIRD = 1.0 / 3.0, a1 = 1.0 / 24.0, C2 = 0.1, C3 = 3.0 / 44.0, C4 = 1.0 / 14.0;
  double xt = x, yt = y, zt = z, ave, delx, dely, delz;
  do {
    double sqrtx = std::sqrt(xt), sqrty = std::sqrt(yt), sqrtz = std::sqrt(zt);
    double alamb = sqrtx * (sqrty + sqrtz) + sqrty * sqrtz;
    xt = 0.25 * (xt + alamb);
    yt = 0.25 * (yt + alamb);
    zt = 0.25 * (zt + alamb);
    ave = THIRD * (xt + yt + zt);
    if (ave == 0.0) {
      delx = dely = delz = 0.0;
    } else {
      delx = (ave - xt) / ave;
      dely = (ave - yt) / ave;
      delz = (ave - zt) / ave;
    }
  } while (std::max(std::max(std::fabs(delx), std::fabs(dely)), std::fabs(delz)) > ERRTOL);
  double e2 = delx * dely - delz * delz;
  double e3 = delx * dely * delz;
  return (1.0 + (a1 * e2 - C2 - C3 * e3) * e2 + C4 * e3) / std::sqrt(ave);
}

// geokerr_wrapper.f:3608-3645
double rc(double x, double y) {
  const double ERRTOL = 0.0012, THIRD = 1.0 / 3.0, a1 = .3, C2 = 1.0 / 7.0, C3 = .375, C4 = 9.0 / 22.0;
  double xt, yt, w, ave, s;
  if (y > 0.) {
    xt = x;
    yt = y;
    w = 1.0;
  } else {
    xt = x - y;
    yt = -y;
    w = std::sqrt(x) / std::sqrt(xt);
  }
  do {
    double alamb = 2.0 * std::sqrt(xt) * std::sqrt(yt) + yt;
    xt = .25 * (xt + alamb);
    yt = .25 * (yt + alamb);
    ave = THIRD * (xt + yt + yt);
    s = (yt - ave) / ave;
  } while (std::fabs(s) > ERRTOL);
  return w * (1.0 + s * s * (a1 + s * (C2 + s * (C3 + s * C4)))) / std::sqrt(ave);
}

// geokerr_wrapper.f:3648-3694
double rd(double x, double y, double z) {
  const double ERRTOL = 0.0015, a1 = 3.0 / 14.0, C2 = 1.0 / 6.0, C3 = 9.0 / 22.0, C4 = 3.0 / 26.0, C5 = 0.25 * C3,
               C6 = 1.5 * C4;
  double xt = x, yt = y, zt = z, sum = 0.0, fac = 1.0, ave, delx, dely, delz;
  do {
    double sqrtx = std::sqrt(xt), sqrty = std::sqrt(yt), sqrtz = std::sqrt(zt);
    double alamb = sqrtx * (sqrty + sqrtz) + sqrty * sqrtz;
    sum = sum + fac / (sqrtz * (zt + alamb));
    fac = 0.25 * fac;
    xt = .25 * (xt + alamb);
    yt = .25 * (yt + alamb);
    zt = .25 * (zt + alamb);
    ave = .2 * (xt + yt + 3.0 * zt);
    delx = (ave - xt) / ave;
    dely = (ave - yt) / ave;
    delz = (ave - zt) / ave;
  } while (std::max(std::max(std::fabs(delx), std::fabs(dely)), std::fabs(delz)) > ERRTOL);
  double ea = delx * dely, eb = delz * delz, ec = ea - eb, ed = ea - 6.0 * eb, ee = ed + ec + ec;
  return 3.0 * sum +
         fac * (1.0 + ed * (-a1 + C5 * ed - C6 * delz * ee) + delz * (C2 * ee + delz * (-C3 * ec + delz * C4 * ea))) /
             (ave * std::sqrt(ave));
}

// geokerr_wrapper.f:3697-3762
double rj(double x, double y, double z, double p) {
  const double ERRTOL = 0.0015, a1 = 3.0 / 14.0, C2 = 1.0 / 3.0, C3 = 3.0 / 22.0, C4 = 3.0 / 26.0, C5 = .75 * C3,
               C6 = 1.5 * C4, C7 = .5 * C2, C8 = C3 + C3;
  double sum = 0.0, fac = 1.0, xt, yt, zt, pt, a = 0, b = 0, rcx = 0;
  if (p > 0.0) {
    xt = x;
    yt = y;
    zt = z;
    pt = p;
  } else {
    xt = std::min(std::min(x, y), z);
    zt = std::max(std::max(x, y), z);
    yt = x + y + z - xt - zt;
    a = 1.0 / (yt - p);
    b = a * (zt - yt) * (yt - xt);
    pt = yt + b;
    double rho = xt * zt / yt;
    double tau = p * pt / yt;
    rcx = rc(rho, tau);
  }
  double ave, delx, dely, delz, delp;
  do {
    double sqrtx = std::sqrt(xt), sqrty = std::sqrt(yt), sqrtz = std::sqrt(zt);
    double alamb = sqrtx * (sqrty + sqrtz) + sqrty * sqrtz;
    double t = pt * (sqrtx + sqrty + sqrtz) + sqrtx * sqrty * sqrtz;
    double alpha = t * t;
    double t2 = pt + alamb;
    double beta = pt * t2 * t2;
    sum = sum + fac * rc(alpha, beta);
    fac = .25 * fac;
    xt = .25 * (xt + alamb);
    yt = .25 * (yt + alamb);
    zt = .25 * (zt + alamb);
    pt = .25 * (pt + alamb);
    ave = .2 * (xt + yt + zt + pt + pt);
    delx = (ave - xt) / ave;
    dely = (ave - yt) / ave;
    delz = (ave - zt) / ave;
    delp = (ave - pt) / ave;
  } while (std::max(std::max(std::fabs(delx), std::fabs(dely)), std::max(std::fabs(delz), std::fabs(delp))) > ERRTOL);
  double ea = delx * (dely + delz) + dely * delz;
  double eb = delx * dely * delz;
  double ec = delp * delp;
  double ed = ea - 3.0 * ec;
  double ee = eb + 2.0 * delp * (ea - ec);
  double r = 3.0 * sum + fac *
                             (1.0 + ed * (-a1 + C5 * ed - C6 * ee) + eb * (C7 + delp * (-C8 + delp * C4)) +
                              delp * ea * (C2 - delp * C3) - C2 * delp * ec) /
                             (ave * std::sqrt(ave));
  if (p <= 0.0) r = a * (b * r + 3.0 * (rcx - rf(xt, yt, zt)));
  return r;
}

// ---------------------------------------------------------------- Jacobi elliptic functions
// geokerr_wrapper.f:3277-3346
void sncndn(double uu, double emmc, double &sn, double &cn, double &dn) {
  const double CA = 3.e-8;
  double em[14], en[14];
  double emc = emmc, u = uu, a, b, c = 0, d = 0;
  int l = 0;
  if (emc != 0.0) {
    bool bo = (emc < 0.0);
    if (bo) {
      d = 1.0 - emc;
      emc = -emc / d;
      d = std::sqrt(d);
      u = d * u;
    }
    a = 1.0;
    dn = 1.0;
    for (int i = 1; i <= 13; i++) {
      l = i;
      em[i] = a;
      emc = std::sqrt(emc);
      en[i] = emc;
      c = 0.5 * (a + emc);
      if (std::fabs(a - emc) <= CA * a) break;
      emc = a * emc;
      a = c;
    }
    u = c * u;
    sn = std::sin(u);
    cn = std::cos(u);
    if (sn != 0.) {
      a = cn / sn;
      c = a * c;
      for (int ii = l; ii >= 1; ii--) {
        b = em[ii];
        a = c * a;
        c = dn * c;
        dn = (en[ii] + a) / (b + a);
        a = c / b;
      }
      a = 1.0 / std::sqrt(c * c + 1.0);
      if (sn < 0.)
        sn = -a;
      else
        sn = a;
      cn = c * sn;
    }
    if (bo) {
      a = dn;
      dn = cn;
      cn = a;
      sn = sn / d;
    }
  } else {
    cn = 1.0 / std::cosh(u);
    dn = cn;
    sn = std::tanh(u);
  }
}

// ---------------------------------------------------------------- polynomial roots
// geokerr_wrapper.f:3533-3606  (A is 1-based A(1..M+1), A(1) constant term)
static void laguer(const cplx *A, int M, cplx &X, int &ITS) {
  const int MR = 8, MT = 10, MAXIT = MT * MR;
  const double EPSS = 1.e-15;
  static const double FRAC[9] = {0., 0.5, 0.25, 0.75, 0.13, 0.38, 0.62, 0.88, 1.0};
  for (int ITER = 1; ITER <= MAXIT; ITER++) {
    ITS = ITER;
    cplx B = A[M + 1];
    double ERR = std::abs(B);
    cplx D(0., 0.), F(0., 0.);
    double ABX = std::abs(X);
    for (int J = M; J >= 1; J--) {
      F = X * F + D;
      D = X * D + B;
      B = X * B + A[J];
      ERR = std::abs(B) + ABX * ERR;
    }
    ERR = EPSS * ERR;
    cplx DX;
    if (std::abs(B) <= ERR) {
      return;
    } else {
      cplx G = D / B;
      cplx G2 = G * G;
      cplx H = G2 - 2.0 * F / B;
      cplx SQ = std::sqrt(double(M - 1) * (double(M) * H - G2));
      cplx GP = G + SQ;
      cplx GM = G - SQ;
      double ABP = std::abs(GP), ABM = std::abs(GM);
      if (ABP < ABM) GP = GM;
      if (std::max(ABP, ABM) > 0.0) {
        DX = double(M) / GP;
      } else {
        std::complex<float> e = std::exp(std::complex<float>((float)std::log(1.0 + ABX), (float)ITER));
        DX = cplx(e.real(), e.imag());
      }
    }
    cplx X1 = X - DX;
    if (X == X1) return;
    if (ITER % MT != 0) {
      X = X1;
    } else {
      X = X - DX * FRAC[ITER / MT];
    }
  }
  // 'Too many iterations'
}

// geokerr_wrapper.f:3348-3441  (A(1..M+1), ROOTS(1..M), 1-based)
void zroots(const cplx *A, int M, cplx *ROOTS, bool POLISH) {
  const double EPS = 1e-6;
  const int MAXM = 7;
  cplx AD[MAXM + 2];
  for (int J = 1; J <= M + 1; J++) AD[J] = A[J];
  cplx XSTART(0., 0.);
  int ncc = 0;
  int ITS;
  for (int J = M; J >= 1; J--) {
    cplx X = XSTART;
    if (J < M) {
      if (ROOTS[J + 1].imag() != 0.0 && ncc == 0) {
        // (sic) the reference builds the "conjugate" start from the REAL part twice
        XSTART = cplx(ROOTS[J + 1].real(), -ROOTS[J + 1].real());
        ncc = 1;
      } else {
        XSTART = cplx(0., 0.);
        ncc = 0;
      }
    } else {
      XSTART = cplx(0., 0.);
      ncc = 0;
    }
    laguer(AD, J, X, ITS);
    if (std::fabs(X.imag()) <= 2.0 * EPS * EPS * std::fabs(X.real())) X = cplx(X.real(), 0.0);
    ROOTS[J] = X;
    cplx B = AD[J + 1];
    for (int JJ = J; JJ >= 1; JJ--) {
      cplx C = AD[JJ];
      AD[JJ] = B;
      B = X * B + C;
    }
  }
  if (POLISH) {
    for (int J = 1; J <= M; J++) laguer(A, M, ROOTS[J], ITS);
  }
  for (int J = 2; J <= M; J++) {
    cplx X = ROOTS[J];
    int I;
    for (I = J - 1; I >= 1; I--) {
      if (ROOTS[I].real() <= X.real()) goto L37;
      ROOTS[I + 1] = ROOTS[I];
    }
    I = 0;
  L37:
    ROOTS[I + 1] = X;
  }
}

// ---------------------------------------------------------------- Carlson-form integrals of Table 1
// p is the 1-based exponent array p(1..5).  geokerr_wrapper.f:1188-1262
static double ellcubicreal(const int *p, double a1, double b1, double a2, double b2, double a3, double b3, double a4,
                           double b4, double &rff, double y, double x) {
  const double one = 1., two = 2., half = 0.5, three = 3.;
  double ellcubic = 0.;
  double d12 = a1 * b2 - a2 * b1, d13 = a1 * b3 - a3 * b1, d14 = a1 * b4 - a4 * b1, d24 = a2 * b4 - a4 * b2,
         d34 = a3 * b4 - a4 * b3;
  double X1 = std::sqrt(a1 + b1 * x), X2 = std::sqrt(a2 + b2 * x), X3 = std::sqrt(a3 + b3 * x),
         X4 = std::sqrt(a4 + b4 * x);
  double Y1 = std::sqrt(a1 + b1 * y), Y2 = std::sqrt(a2 + b2 * y), Y3 = std::sqrt(a3 + b3 * y),
         Y4 = std::sqrt(a4 + b4 * y);
  double U1c = (X1 * Y2 * Y3 + Y1 * X2 * X3) / (x - y);
  double U12 = U1c * U1c;
  double t2 = (X2 * Y1 * Y3 + Y2 * X1 * X3) / (x - y);
  double U22 = t2 * t2;
  double t3 = (X3 * Y1 * Y2 + Y3 * X1 * X2) / (x - y);
  double U32 = t3 * t3;
  if (p[4] == 0) {
    rff = rf(U32, U22, U12);
    ellcubic = two * rff;
  } else {
    double W22 = U12 - b4 * d12 * d13 / d14;
    double q = X4 * Y4 / X1 / Y1;
    double Q22 = q * q * W22;
    double P22 = Q22 + b4 * d24 * d34 / d14;
    double I1c = two * rff;
    if (p[4] == -2) {
      double I3c = two * rc(P22, Q22) - two * b1 * d12 * d13 / three / d14 * rj(U32, U22, U12, W22);
      ellcubic = (b4 * I3c - b1 * I1c) / d14;
    } else {
      double r12 = a1 / b1 - a2 / b2, r13 = a1 / b1 - a3 / b3;
      double r24i = b2 * b4 / (a2 * b4 - a4 * b2), r34i = b3 * b4 / (a3 * b4 - a4 * b3);
      double I2c = two / three * d12 * d13 * rd(U32, U22, U12) + two * X1 * Y1 / U1c;
      double K2c = b2 * b3 * I2c - two * b4 * (X1 * X2 * X3 / (X4 * X4) - Y1 * Y2 * Y3 / (Y4 * Y4));
      double bd = b1 / d14;
      ellcubic = half * b4 / d14 / d24 / d34 * K2c + bd * bd * (one - half * r12 * r13 * r24i * r34i) * I1c;
    }
  }
  return ellcubic;
}

// geokerr_wrapper.f:1264-1340
static double ellcubiccomplex(const int *p, double a1, double b1, double a4, double b4, double f, double g, double h,
                              double &rff, double y, double x) {
  const double one = 1., two = 2., half = 0.5, three = 3., four = 4., six = 6.;
  double ellcubic = 0.;
  double X1 = std::sqrt(a1 + b1 * x), X4 = std::sqrt(a4 + b4 * x);
  double Y1 = std::sqrt(a1 + b1 * y), Y4 = std::sqrt(a4 + b4 * y);
  double d14 = a1 * b4 - a4 * b1;
  double beta1 = g * b1 - two * h * a1;
  double a11 = std::sqrt(two * f * b1 * b1 - two * g * a1 * b1 + two * h * a1 * a1);
  double c44 = std::sqrt(two * f * b4 * b4 - two * g * a4 * b4 + two * h * a4 * a4);
  double a142 = two * f * b1 * b4 - g * (a1 * b4 + a4 * b1) + two * h * a1 * a4;
  double xi = std::sqrt(f + g * x + h * x * x);
  double eta = std::sqrt(f + g * y + h * y * y);
  double xe = xi + eta, xmy = x - y;
  double m = (X1 + Y1) * std::sqrt(xe * xe - h * xmy * xmy) / xmy;
  double M2 = m * m;
  double Lp2 = M2 - beta1 + std::sqrt(two * h) * a11;
  double Lm2 = M2 - beta1 - std::sqrt(two * h) * a11;
  if (p[4] == 0) {
    rff = rf(M2, Lm2, Lp2);
    ellcubic = four * rff;
  } else {
    double I1c = four * rff;
    double U = (X1 * eta + Y1 * xi) / xmy;
    double U2 = U * U;
    double Wp2 = M2 - b1 * (a142 + a11 * c44) / d14;
    double W2 = U2 - a11 * a11 * b4 / two / d14;
    double q = X4 * Y4 / X1 / Y1;
    double Q2 = q * q * W2;
    double P2 = Q2 + c44 * c44 * b4 / two / d14;
    double rho = std::sqrt(two * h) * a11 - beta1;
    if (p[4] == -2) {
      double I3c = (two * a11 / three / c44) *
                       ((-four * b1 / d14) * (a142 + a11 * c44) * rj(M2, Lm2, Lp2, Wp2) - six * rff + three * rc(U2, W2)) +
                   two * rc(P2, Q2);
      ellcubic = (b4 * I3c - b1 * I1c) / d14;
    } else {
      double r24Xr34 = half * c44 * c44 / h / (b4 * b4);
      double r12Xr13 = half * a11 * a11 / h / (b1 * b1);
      double N2c = two / three * std::sqrt(two * h) / a11 * (four * rho * rd(M2, Lm2, Lp2) - six * rff + three / U) +
                   two / X1 / Y1 / U;
      double K2c = half * a11 * a11 * N2c - two * d14 * (xi / X1 / (X4 * X4) - eta / Y1 / (Y4 * Y4));
      double bd = b1 / d14;
      ellcubic = half / d14 / (h * b4 * r24Xr34) * K2c + bd * bd * (one - half * r12Xr13 / r24Xr34) * I1c;
    }
  }
  return ellcubic;
}

// geokerr_wrapper.f:1342-1360
static double asech(double x) {
  if (x > 0.0) return std::log((1.0 + std::sqrt(1.0 - x * x)) / x);
  return 0.0;
}

// geokerr_wrapper.f:1362-1442
static double ellquarticreal(const int *p, double a1, double b1, double a2, double b2, double a3, double b3, double a4,
                             double b4, double a5, double b5, double &rff, double y, double x) {
  const double one = 1., two = 2., half = 0.5, three = 3.;
  double d12 = a1 * b2 - a2 * b1, d13 = a1 * b3 - a3 * b1, d14 = a1 * b4 - a4 * b1, d24 = a2 * b4 - a4 * b2,
         d34 = a3 * b4 - a4 * b3, d15 = a1 * b5 - a5 * b1, d25 = a2 * b5 - a5 * b2, d35 = a3 * b5 - a5 * b3,
         d45 = a4 * b5 - a5 * b4;
  double X1 = std::sqrt(a1 + b1 * x), X2 = std::sqrt(a2 + b2 * x), X3 = std::sqrt(a3 + b3 * x),
         X4 = std::sqrt(a4 + b4 * x);
  double Y1 = std::sqrt(a1 + b1 * y), Y2 = std::sqrt(a2 + b2 * y), Y3 = std::sqrt(a3 + b3 * y),
         Y4 = std::sqrt(a4 + b4 * y);
  double t;
  t = (X1 * X2 * Y3 * Y4 + Y1 * Y2 * X3 * X4) / (x - y);
  double U122 = t * t;
  t = (X1 * X3 * Y2 * Y4 + Y1 * Y3 * X2 * X4) / (x - y);
  double U132 = t * t;
  t = (X1 * X4 * Y2 * Y3 + Y1 * Y4 * X2 * X3) / (x - y);
  double U142 = t * t;
  if (p[5] == 0) {
    rff = rf(U122, U132, U142);
    return two * rff;
  } else {
    double I1 = two * rff;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double W22 = U122 - d13 * d14 * d25 / d15;
    double xy = X1 * Y1;
    double Q22 = X52 * Y52 / (xy * xy) * W22;
    double P22 = Q22 + d25 * d35 * d45 / d15;
    if (p[5] == -2) {
      double I3 = two * d12 * d13 * d14 / three / d15 * rj(U122, U132, U142, W22) + two * rc(P22, Q22);
      return (b5 * I3 - b1 * I1) / d15;
    } else {
      double I2 = two / three * d12 * d13 * rd(U122, U132, U142) + two * X1 * Y1 / X4 / Y4 / std::sqrt(U142);
      double r12 = a1 / b1 - a2 / b2, r13 = a1 / b1 - a3 / b3;
      double r25i = b2 * b5 / (a2 * b5 - a5 * b2), r35i = b3 * b5 / (a3 * b5 - a5 * b3);
      double A111m1m2 = X1 * X2 * X3 / X4 / X52 - Y1 * Y2 * Y3 / Y4 / Y52;
      double bd = b1 / d15;
      return half * b5 * b5 * d24 * d34 / d15 / d25 / d35 / d45 * I2 +
             bd * bd * (one - half * r12 * r13 * r25i * r35i) * I1 - b5 * b5 / d15 / d25 / d35 * A111m1m2;
    }
  }
}

// geokerr_wrapper.f:1444-1529
static double ellquarticcomplex(const int *p, double a1, double b1, double a4, double b4, double a5, double b5, double f,
                                double g, double h, double &rff, double y, double x) {
  const double one = 1., two = 2., half = 0.5, three = 3., four = 4., six = 6.;
  double X1 = std::sqrt(a1 + b1 * x), X4 = std::sqrt(a4 + b4 * x);
  double Y1 = std::sqrt(a1 + b1 * y), Y4 = std::sqrt(a4 + b4 * y);
  double a11 = std::sqrt(two * f * b1 * b1 - two * g * a1 * b1 + two * h * a1 * a1);
  double c44 = std::sqrt(two * f * b4 * b4 - two * g * a4 * b4 + two * h * a4 * a4);
  double a142 = two * f * b1 * b4 - g * (a1 * b4 + a4 * b1) + two * h * a1 * a4;
  double xi = std::sqrt(f + g * x + h * x * x);
  double eta = std::sqrt(f + g * y + h * y * y);
  double xe = xi + eta, xmy = x - y;
  double m = (X1 * Y4 + Y1 * X4) * std::sqrt(xe * xe - h * xmy * xmy) / xmy;
  double M2 = m * m;
  double Lp2 = M2 + a142 + a11 * c44;
  double Lm2 = std::max(M2 + a142 - a11 * c44, 0.0);
  if (p[5] == 0) {
    rff = rf(M2, Lm2, Lp2);
    return four * rff;
  } else {
    double I1 = four * rff;
    double d14 = a1 * b4 - a4 * b1, d15 = a1 * b5 - a5 * b1, d45 = a4 * b5 - a5 * b4;
    double X52 = a5 + b5 * x, Y52 = a5 + b5 * y;
    double c552 = two * f * b5 * b5 - two * g * a5 * b5 + two * h * a5 * a5;
    double c55 = std::sqrt(c552);
    double a152 = two * f * b1 * b5 - g * (a1 * b5 + a5 * b1) + two * h * a5 * a1;
    double U = (X1 * X4 * eta + Y1 * Y4 * xi) / xmy;
    double U2 = U * U;
    double Wp2 = M2 + d14 * (a152 + a11 * c55) / d15;
    double W2 = U2 - a11 * a11 * d45 / two / d15;
    double xy = X1 * Y1;
    double Q2 = X52 * Y52 / (xy * xy) * W2;
    double P2 = Q2 + c55 * c55 * d45 / two / d15;
    if (p[5] == -2) {
      double I3 = (two * a11 / three / c55) *
                      ((four * d14 / d15) * (a152 + a11 * c55) * rj(M2, Lm2, Lp2, Wp2) - six * rff + three * rc(U2, W2)) +
                  two * rc(P2, Q2);
      return (b5 * I3 - b1 * I1) / d15;
    } else {
      double I2 = two * a11 / three / c44 * (four * (a142 + a11 * c44) * rd(M2, Lm2, Lp2) - six * rff) +
                  two * a11 / c44 / U + two * X1 * Y1 / X4 / Y4 / U;
      double A111m1m2 = X1 * xi / X4 / X52 - Y1 * eta / Y4 / Y52;
      return b5 * b5 / (two * d15 * d45) * c44 * c44 / c552 * I2 +
             b1 * b1 / (d15 * d15) * (one - half * b5 * b5 * a11 * a11 / (b1 * b1) / c552) * I1 -
             two * b5 * b5 / d15 / c552 * A111m1m2;
    }
  }
}

// geokerr_wrapper.f:1531-1645
static double elldoublecomplex(const int *p, double f1, double g1, double h1, double f2, double g2, double h2,
                               double a5, double b5, double &rff, double y, double x) {
  const double one = 1., half = 0.5, two = 2., three = 3., four = 4.;
  double ellquartic;
  double xi1 = std::sqrt(f1 + g1 * x + h1 * x * x), xi2 = std::sqrt(f2 + g2 * x + h2 * x * x);
  double eta1 = std::sqrt(f1 + g1 * y + h1 * y * y), eta2 = std::sqrt(f2 + g2 * y + h2 * y * y);
  double theta1 = two * f1 + g1 * (x + y) + two * h1 * x * y;
  double theta2 = two * f2 + g2 * (x + y) + two * h2 * x * y;
  double zeta1 = std::sqrt(2 * xi1 * eta1 + theta1), zeta2 = std::sqrt(2 * xi2 * eta2 + theta2);
  double M = zeta1 * zeta2 / (x - y);
  double M2 = M * M;
  double delta122 = two * f1 * h2 + two * f2 * h1 - g1 * g2;
  double delta112 = four * f1 * h1 - g1 * g1;
  double delta222 = four * f2 * h2 - g2 * g2;
  double Delta = std::sqrt(delta122 * delta122 - delta112 * delta222);
  double Deltap = delta122 + Delta, Deltam = delta122 - Delta;
  double Lp2 = M2 + Deltap, Lm2 = M2 + Deltam;
  if (p[5] == 0) {
    rff = rf(M2, Lm2, Lp2);
    ellquartic = four * rff;
  } else {
    double U = (xi1 * eta2 + eta1 * xi2) / (x - y);
    double U2 = U * U;
    double alpha15 = two * f1 * b5 - g1 * a5, alpha25 = two * f2 * b5 - g2 * a5;
    double beta15 = g1 * b5 - two * h1 * a5, beta25 = g2 * b5 - two * h2 * a5;
    double gamma1 = half * (alpha15 * b5 - beta15 * a5), gamma2 = half * (alpha25 * b5 - beta25 * a5);
    double Lambda = delta112 * gamma2 / gamma1;
    double Omega2 = M2 + Lambda;
    double psi = half * (alpha15 * beta25 - alpha25 * beta15);
    double psi2 = psi * psi;
    double xi5 = a5 + b5 * x, eta5 = a5 + b5 * y;
    double Am111m1 = one / xi1 * xi2 - one / eta1 * eta2;
    double A1111m4 = xi1 * xi2 / (xi5 * xi5) - eta1 * eta2 / (eta5 * eta5);
    double xmy = x - y;
    double XX = xi5 * eta5 * (theta1 * half * Am111m1 - xi5 * eta5 * A1111m4) / (xmy * xmy);
    double S = half * (M2 + delta122) - U2;
    double S2 = S * S;
    double mu = gamma1 * xi5 * eta5 / xi1 / eta1;
    double T = mu * S + two * gamma1 * gamma2;
    double T2 = T * T;
    double V2 = mu * mu * (S2 + Lambda * U2);
    double b2 = Omega2 * Omega2 * (S2 / U2 + Lambda);
    double a2 = b2 + Lambda * Lambda * psi2 / gamma1 / gamma2;
    double H = delta112 * psi * (rj(M2, Lm2, Lp2, Omega2) / three + half * rc(a2, b2)) / (gamma1 * gamma1) -
               XX * rc(T2, V2);
    if (p[5] == -2) {
      ellquartic = -two * (b5 * H + beta15 * rff / gamma1);
    } else {
      double A1111m2 = xi1 * xi2 / xi5 - eta1 * eta2 / eta5;
      double xi1p = half * (g1 + two * h1 * x) / xi1;
      double eta1p = half * (g1 + two * h1 * y) / eta1;
      double B = xi1p * xi2 - eta1p * eta2;
      double G = two / three * Delta * Deltap * rd(M2, Lm2, Lp2) + half * Delta / U +
                 (delta122 * theta1 - delta112 * theta2) / four / xi1 / eta1 / U;
      double Sigma = G - Deltap * rff + B;
      ellquartic = b5 * (beta15 / gamma1 + beta25 / gamma2) * H + beta15 * beta15 * rff / (gamma1 * gamma1) +
                   b5 * b5 * (Sigma - b5 * A1111m2) / gamma1 / gamma2;
    }
  }
  return ellquartic;
}

// ---------------------------------------------------------------- mu integrals (Legendre pieces)
// geokerr_wrapper.f:3137-3172
static void calcimuasym(double a, double mneg, double mpos, double mu0, double muplus, double &imu0, double &imum,
                        double &rf0, double &rfc) {
  double fac = std::fabs(a) * muplus;
  double m1 = mneg / mpos;
  double s0 = std::sqrt((mpos - mu0 * mu0) / (mpos - mneg));
  double phi0 = std::asin(s0);
  double ak = std::sqrt(1.0 - m1);
  double q0 = (1.0 - s0 * ak) * (1.0 + s0 * ak);
  rf0 = rf(1.0 - s0 * s0, q0, 1.0);
  rfc = rf(0.0, m1, 1.0);
  imum = 2.0 * rfc;
  imu0 = s0 * rf0;
  if (std::tan(phi0) < 0.0) imu0 = imum - imu0;
  imu0 = imu0 / fac;
  imum = imum / fac / 2.0;
}
// geokerr_wrapper.f:3175-3203
static void calcimuasymf(double a, double mneg, double mpos, double muf, double muplus, double &imuf, double imum,
                         double &rff) {
  (void)imum;
  double fac = std::fabs(a) * muplus;
  double m1 = mneg / mpos;
  double sf = std::sqrt((mpos - muf * muf) / (mpos - mneg));
  double ak = std::sqrt(1.0 - m1);
  double qf = (1.0 - sf * ak) * (1.0 + sf * ak);
  rff = rf(1.0 - sf * sf, qf, 1.0);
  imuf = sf * rff;
  imuf = imuf / fac;
}
// geokerr_wrapper.f:3206-3241
static void calcimusym(double a, double mneg, double mpos, double mu0, double muplus, double &imu0, double &imum,
                       double &rf0, double &rfc) {
  double fac = std::fabs(a) * std::sqrt(mpos - mneg);
  double m1 = -mneg / (mpos - mneg);
  double phi0 = std::acos(mu0 / muplus);
  double s0 = std::sin(phi0);
  double ak = std::sqrt(1.0 - m1);
  double q0 = (1.0 - s0 * ak) * (1.0 + s0 * ak);
  rf0 = rf(1.0 - s0 * s0, q0, 1.0);
  rfc = rf(0.0, m1, 1.0);
  imum = 2.0 * rfc;
  imu0 = s0 * rf0;
  if (std::tan(phi0) < 0.0) imu0 = imum - imu0;
  imu0 = imu0 / fac;
  imum = imum / fac;
}
// geokerr_wrapper.f:3244-3274
static void calcimusymf(double a, double mneg, double mpos, double muf, double muplus, double &imuf, double imum,
                        double &rff) {
  double fac = std::fabs(a) * std::sqrt(mpos - mneg);
  double m1 = -mneg / (mpos - mneg);
  double phif = std::acos(muf / muplus);
  double sf = std::sin(phif);
  double ak = std::sqrt(1.0 - m1);
  double qf = (1.0 - sf * ak) * (1.0 + sf * ak);
  rff = rf(1.0 - sf * sf, qf, 1.0);
  imuf = sf * rff;
  if (std::tan(phif) < 0.0) imuf = imum * fac - imuf;
  imuf = imuf / fac;
}

// geokerr_wrapper.f:2508-2590
static void ellphitmu(double phi0, double phif, double m1, double n, bool firstpt, double rf0, double rff, double rfc,
                      double &rdc, double &rjc, double &elle0, double &ellef, double &ellec, double &ellpi0,
                      double &ellpif, double &ellpic) {
  const double pi = std::acos(-1.0);
  double ak = std::sqrt(1.0 - m1);
  if (firstpt) {
    double s0 = std::sin(phi0);
    double q0 = (1.0 - s0 * ak) * (1.0 + s0 * ak);
    double rd0 = rd(1.0 - s0 * s0, q0, 1.0);
    rdc = rd(0.0, m1, 1.0);
    elle0 = s0 * (rf0 - ((s0 * ak) * (s0 * ak)) * rd0 / 3.0);
    ellec = 2.0 * (rfc - ak * ak * rdc / 3.0);
    if (n != 1.0) {
      double rj0 = rj(1.0 - s0 * s0, q0, 1.0, 1.0 - n * s0 * s0);
      rjc = rj(0.0, m1, 1.0, 1.0 - n);
      ellpi0 = s0 * (rf0 + n * s0 * s0 * rj0 / 3.0);
      ellpic = 2.0 * (rfc + n * rjc / 3.0);
    } else {
      ellpi0 = 0.0;
      ellpic = 0.0;
    }
    if (phi0 > pi / 2.0) {
      ellpi0 = ellpic - ellpi0;
      elle0 = ellec - elle0;
    }
  }
  double sf = std::sin(phif);
  double qf = (1.0 - sf * ak) * (1.0 + sf * ak);
  double rdf = rd(1.0 - sf * sf, qf, 1.0);
  ellef = sf * (rff - ((sf * ak) * (sf * ak)) * rdf / 3.0);
  if (n != 1.0) {
    double rjf = rj(1.0 - sf * sf, qf, 1.0, 1.0 - n * sf * sf);
    ellpif = sf * (rff + n * sf * sf * rjf / 3.0);
  } else {
    ellpif = 0.0;
  }
  if (phif > pi / 2.0) {
    ellpic = 2.0 * (rfc + n * rjc / 3.0);
    ellec = 2.0 * (rfc - ak * ak * rdc / 3.0);
    ellpif = ellpic - ellpif;
    ellef = ellec - ellef;
  }
}

// geokerr_wrapper.f:2361-2432
static void calcphitmuasym(double a, double mneg, double mpos, double mu0, double muf, double muplus, double &phimu0,
                           double &phimuf, double &phimum, double &tmu0, double &tmuf, double &tmum, double rf0,
                           double rff, double rfc, double &rdc, double &rjc, bool firstpt) {
  double fac = std::fabs(a) * muplus;
  double m1 = mneg / mpos;
  double n, fac2;
  if (mpos != 1.0) {
    n = (mpos - mneg) / (1.0 - mpos);
    fac2 = 1.0 - mpos;
  } else {
    n = -1.0;
    fac2 = 1.0;
  }
  double phi0 = std::asin(std::sqrt((mpos - mu0 * mu0) / (mpos - mneg)));
  double phif = std::asin(std::sqrt((mpos - muf * muf) / (mpos - mneg)));
  double elle0 = 0, ellef = 0, ellec = 0, ellpi0 = 0, ellpif = 0, ellpic = 0;
  ellphitmu(phi0, phif, m1, -n, firstpt, rf0, rff, rfc, rdc, rjc, elle0, ellef, ellec, ellpi0, ellpif, ellpic);
  if (firstpt) {
    tmu0 = fac * elle0;
    tmum = fac * ellec / 2.0;
    phimu0 = ellpi0 / fac2 / fac;
    phimum = ellpic / fac2 / fac / 2.0;
  }
  tmuf = fac * ellef;
  phimuf = ellpif / fac2 / fac;
}

// geokerr_wrapper.f:2435-2504
static void calcphitmusym(double a, double mneg, double mpos, double mu0, double muf, double muplus, double &phimu0,
                          double &phimuf, double &phimum, double &tmu0, double &tmuf, double &tmum, double rf0,
                          double rff, double rfc, double &rdc, double &rjc, bool firstpt) {
  double fac = std::fabs(a) * std::sqrt(mpos - mneg);
  double m1 = -mneg / (mpos - mneg);
  double n, fac2;
  if (mpos != 1.0) {
    n = mpos / (1.0 - mpos);
    fac2 = 1.0 - mpos;
  } else {
    n = -1.0;
    fac2 = 1.0;
  }
  double phi0 = std::acos(mu0 / muplus);
  double phif = std::acos(muf / muplus);
  double elle0 = 0, ellef = 0, ellec = 0, ellpi0 = 0, ellpif = 0, ellpic = 0;
  ellphitmu(phi0, phif, m1, -n, firstpt, rf0, rff, rfc, rdc, rjc, elle0, ellef, ellec, ellpi0, ellpif, ellpic);
  if (firstpt) {
    tmu0 = fac * elle0;
    tmum = fac * ellec;
    phimu0 = ellpi0 / fac2 / fac;
    phimum = ellpic / fac2 / fac;
  }
  tmuf = fac * ellef;
  phimuf = ellpif / fac2 / fac;
}

// geokerr_wrapper.f:2593-2625
static double phifnkerr(double u, double u1, double u2, double l, double a) {
  double up = 1.0 / (1.0 + std::sqrt(1.0 - a * a));
  double umi = 1.0 - std::sqrt(1.0 - a * a);
  double s = std::sqrt(u - u1);
  double su21 = std::sqrt(u2 - u1);
  double sup = std::sqrt(1 - u1 / up);
  double sum = std::sqrt(1 - u1 * umi);
  double fn = (l + 2.0 * a * u2 - 2.0 * l * u2) / (su21 * (u2 * umi - 1.0) * (u2 / up - 1.0)) *
                  std::log((su21 + s) / (su21 - s)) +
              (2.0 * a + l * (-2.0 + umi)) * std::sqrt(umi) *
                  std::log((sum + s * std::sqrt(umi)) / (sum - s * std::sqrt(umi))) /
                  (sum * (u2 * umi - 1.0) * (umi - 1.0 / up)) +
              (2.0 * a + l * (-2.0 + 1.0 / up)) * std::sqrt(1.0 / up) *
                  std::log((sup + std::sqrt(1.0 / up) * s) / (sup - std::sqrt(1.0 / up) * s)) / (1.0 / up - umi) * sup *
                  (u2 / up - 1.0);
  return fn;
}

// geokerr_wrapper.f:2629-2670
static double tfnkerr(double u0, double uf, double u1, double u2, double l, double a) {
  double up = 1.0 / (1.0 + std::sqrt(1.0 - a * a));
  double umi = 1.0 - std::sqrt(1.0 - a * a);
  double sf = std::sqrt(uf - u1), ss = std::sqrt(u0 - u1);
  double su21 = std::sqrt(u2 - u1), su1 = std::sqrt(-u1);
  double sup = std::sqrt(1 - u1 / up), sum = std::sqrt(1 - u1 * umi);
  double squmi = std::sqrt(umi), squp = std::sqrt(1.0 / up);
  double fn = sf / (uf * u1 * u2) - ss / (u0 * u1 * u2) -
              ((-u2 - 2.0 * u1 * (1.0 + u2 * umi + u2 / up)) *
               std::log((sf + su1) * (su1 - ss) / ((-sf + su1) * (ss + su1)))) /
                  (2.0 * std::pow(-u1, 3.0 / 2.0) * u2 * u2) +
              ((1.0 - 2.0 * a * l * u2 * u2 * u2 + a * a * u2 * u2 * (1.0 + 2.0 * u2)) *
               std::log((su21 + sf) * (su21 - ss) / ((su21 - sf) * (su21 + ss)))) /
                  (u2 * u2 * su21 * (u2 * umi - 1.0) * (u2 / up - 1.0)) +
              ((-2.0 * a * l + a * a * (2.0 + umi) + umi * umi * umi) * squmi *
               std::log((sum + sf * squmi) * (sum - ss * squmi) / ((sum - sf * squmi) * (sum + ss * squmi)))) /
                  (sum * (u2 * umi - 1.0) * (umi - 1 / up)) +
              ((-2.0 * a * l + a * a * (2.0 + 1 / up) + 1.0 / (up * up * up)) * squp *
               std::log((sup + sf * squp) * (sup - ss * squp) / ((sup - sf * squp) * (sup + ss * squp))) /
               ((1.0 / up - umi) * sup * (u2 / up - 1.0)));
  return fn;
}

// roots of the biquadratic M(mu), shared text of geomu/geophitime/geor/findmroots
static inline void mroots(double a, double q2, double ql2, double &mneg, double &mpos) {
  double yy = -0.5 * (a * a - ql2 + fsign(1.0, a * a - ql2) * std::sqrt((a * a - ql2) * (a * a - ql2) + 4.0 * q2 * a * a));
  if ((a * a - ql2) < 0.0) {
    mneg = -yy / a / a;
    mpos = q2 / yy;
  } else {
    mneg = q2 / yy;
    mpos = -yy / a / a;
  }
}

// ---------------------------------------------------------------- GEOMU  geokerr_wrapper.f:771-1186
void geomu(double &u0, double &uf, double mu0, double &muf, double a, double l, double &l2, double q2, double &iu,
           int &tpm, int &tpr, double su, double sm, GeoCache &c, bool pht, bool firstpt) {
  const double one = 1., two = 2., third = 0.3333333333333333;
  const double pi = std::acos(-one);
  int p[6] = {0, 0, 0, 0, 0, 0};
  double uplus = one / (one + std::sqrt(one - a * a));
  l2 = l * l;
  double ql2 = q2 + l2;
  double cc = a * a - q2 - l2;
  double dd = two * ((a - l) * (a - l) + q2);
  double ee = -a * a * q2;
  double &u1 = c.u1, &u2 = c.u2, &u3 = c.u3, &u4 = c.u4;
  double iu1 = 0.0, muplus = 0.0, i2mu = 0.0, sn, cn, dn;
  if ((ee == 0.0) && (dd != 0.0)) {
    p[1] = -1; p[2] = -1; p[3] = -1; p[4] = 0; p[5] = 0;
    double qq = cc * cc / dd / dd / 9.0;
    double rr = (two * cc * cc * cc / (dd * dd * dd) + 27.0 / dd) / 54.0;
    double dis = rr * rr - qq * qq * qq;
    if (dis < -1.e-16) {
      double theta = std::acos(rr / std::pow(qq, 1.5));
      u1 = -two * std::sqrt(qq) * std::cos(theta / 3.0) - cc / dd / 3.0;
      u2 = -two * std::sqrt(qq) * std::cos((theta - two * pi) / 3.0) - cc / dd / 3.0;
      u3 = -two * std::sqrt(qq) * std::cos((theta + two * pi) / 3.0) - cc / dd / 3.0;
      iu1 = 0.0;
      for (;;) {  // label 80
        if (u0 <= u2) {
          if (uf > u2) uf = u2;
          c.ncase = 1;
          if (firstpt && (u0 != u2)) {
            c.iu0 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, 0.0, c.rffu0, u0, u2);
          } else if (u0 == u2) {
            c.iu0 = 0.0;
          }
          if (uf != u2) iu1 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, 0.0, c.rffu1, uf, u2);
          iu = su * (c.iu0 - powm1(tpr) * iu1) / std::sqrt(dd);
          break;
        } else if (u0 >= u3) {
          if (uf < u3) uf = u3;
          c.ncase = 2;
          if (firstpt && (u0 != u3)) {
            c.iu0 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, 0.0, c.rffu0, u3, u0);
          } else if (u0 == u3) {
            c.iu0 = 0.0;
          }
          if (uf != u3) iu1 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, 0.0, c.rffu1, u3, uf);
          iu = su * (c.iu0 - powm1(tpr) * iu1) / std::sqrt(dd);
          break;
        } else {
          // 'WARNING - Unphysical Cubic Real. Input modified.'
          if (su == 1)
            u0 = u3;
          else
            u0 = u2;
        }
      }
    } else if (std::fabs(dis) < 1.e-16) {
      u1 = -two * std::sqrt(qq) - cc / dd / 3.0;
      u2 = -two * std::sqrt(qq) * std::cos((two * pi) / 3.0) - cc / dd / 3.0;
      u3 = u2;
      tpr = 0;
      if (u0 <= u2) {
        c.ncase = 1;
        if (uf > u2) {
          uf = u2;
          iu = su * 1.e300;
        } else {
          double farg = (std::sqrt(uf - u1) + std::sqrt(u2 - u1)) / std::fabs(std::sqrt(uf - u1) - std::sqrt(u2 - u1));
          double sarg = (std::sqrt(u0 - u1) + std::sqrt(u2 - u1)) / std::fabs(std::sqrt(u0 - u1) - std::sqrt(u2 - u1));
          iu = su * (std::log(farg) - std::log(sarg)) / std::sqrt((u2 - u1) * dd);
        }
      } else if (u0 >= u2) {
        c.ncase = 2;
        if (uf < u2) {
          uf = u2;
          iu = su * 1.e300;
        } else {
          double farg = (std::sqrt(uf - u1) + std::sqrt(u2 - u1)) / std::fabs(std::sqrt(uf - u1) - std::sqrt(u2 - u1));
          double sarg = (std::sqrt(u0 - u1) + std::sqrt(u2 - u1)) / std::fabs(std::sqrt(u0 - u1) - std::sqrt(u2 - u1));
          iu = -su * (std::log(farg) - std::log(sarg)) / std::sqrt((u2 - u1) * dd);
        }
      }
    } else {
      c.ncase = 3;
      double aa = -fsign(1.0, rr) * std::pow(std::fabs(rr) + std::sqrt(dis), third);
      double bb;
      if (aa != 0.0)
        bb = qq / aa;
      else
        bb = 0.0;
      u1 = (aa + bb) - cc / dd / 3.0;
      double f = -one / dd / u1;
      double g = f / u1;
      if (uf > u0) {
        iu = su * ellcubiccomplex(p, -u1, one, 0.0, 0.0, f, g, one, c.rffu0, u0, uf) / std::sqrt(dd);
      } else if (uf < u0) {
        iu = -su * ellcubiccomplex(p, -u1, one, 0.0, 0.0, f, g, one, c.rffu0, uf, u0) / std::sqrt(dd);
      } else {
        iu = 0.0;
      }
    }
    if (q2 == 0.0) {
      double s1 = fsign(1.0, mu0);
      if (l2 < a * a && mu0 != 0.0) {
        muplus = s1 * std::sqrt(one - l2 / a / a);
        muf = muplus / std::cosh(std::fabs(a * muplus) * iu - s1 * sm * asech(mu0 / muplus));
      } else {
        muf = 0.0;
      }
    } else if (a == 0.0) {
      double a1 = sm;
      muplus = std::sqrt(q2 / ql2);
      if (mu0 > muplus) muplus = mu0;
      c.i1mu = std::acos(mu0 / muplus) / std::sqrt(ql2);
      c.i3mu = pi / std::sqrt(ql2);
      if (sm == 1) {
        tpm = (int)((iu - c.i1mu) / c.i3mu) + (int)((1 + fsign(1.0, (iu - c.i1mu) / c.i3mu)) / 2.0);
      } else {
        tpm = (int)((iu + c.i1mu) / c.i3mu);
      }
      double a2 = sm * powm1(tpm);
      double a3 = two * (int)((two * double(tpm) + 3.0 - sm) / 4.0) - one;
      muf = -muplus * std::cos(std::sqrt(ql2) * (iu - a1 * c.i1mu - a3 * c.i3mu) / a2);
    }
  } else if (ee == 0.0 && dd == 0.0) {
    c.ncase = 4;
    tpr = 0;
    iu = su * (uf - u0);
    muf = 0.0;
  } else {
    p[1] = -1; p[2] = -1; p[3] = -1; p[4] = -1; p[5] = 0;
    if (firstpt) {
      cplx cf[6], root[5];
      cf[1] = cplx(one, 0.0);
      cf[2] = cplx(0.0, 0.0);
      cf[3] = cplx(cc, 0.0);
      cf[4] = cplx(dd, 0.0);
      cf[5] = cplx(ee, 0.0);
      zroots(cf, 4, root, true);
      int nreal = 0;
      for (int i = 1; i <= 4; i++)
        if (root[i].imag() == 0.0) nreal++;
      if (nreal == 2) {
        c.ncase = 5;
        u1 = root[1].real();
        if (root[2].imag() == 0.0)
          u4 = root[2].real();
        else
          u4 = root[4].real();
      } else if (nreal == 0) {
        c.ncase = 6;
      } else {
        u1 = root[1].real();
        u2 = root[2].real();
        u3 = root[3].real();
        u4 = root[4].real();
        for (;;) {  // label 90
          if (u2 > uplus && u3 > uplus) {
            c.ncase = 5;
            break;
          } else if (u0 <= u2) {
            c.ncase = 7;
            break;
          } else if (u0 >= u3) {
            c.ncase = 8;
            break;
          } else {
            // 'WARNING--Unphysical Quartic Real. Inputs modified.'
            if (su == 1)
              u0 = 0.0;
            else
              u0 = uplus;
          }
        }
      }
    }
    if (c.ncase == 5) {
      double qs = fsign(1.0, q2);
      double f = -qs * one / std::fabs(ee) / u1 / u4;
      double g = (u4 + u1) / u1 / u4 * f;
      if (uf > u0) {
        iu = su * ellquarticcomplex(p, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, c.rffu0, u0, uf) /
             std::sqrt(std::fabs(ee));
      } else if (uf < u0) {
        iu = -su * ellquarticcomplex(p, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, c.rffu0, uf, u0) /
             std::sqrt(std::fabs(ee));
      } else {
        iu = 0.0;
      }
    } else if (c.ncase == 6) {
      c.ncase = 6;
      cplx coefs[8], hroots[7];
      coefs[1] = cplx(one, 0.0);
      coefs[2] = cplx(-cc / std::sqrt(ee), 0.0);
      coefs[3] = cplx(-one, 0.0);
      coefs[4] = cplx(std::sqrt(ee) * (two * cc / ee - (dd / ee) * (dd / ee)), 0.0);
      coefs[5] = cplx(-one, 0.0);
      coefs[6] = cplx(-cc / std::sqrt(ee), 0.0);
      coefs[7] = cplx(one, 0.0);
      zroots(coefs, 6, hroots, true);
      int i = 0;
      c.h1 = 0.0;
      do {
        i = i + 1;
        if (i > 6) break;  // guard: the reference would run off the array
        if (hroots[i].imag() == 0.0) c.h1 = hroots[i].real();
      } while (c.h1 == 0.0);
      double h1 = c.h1;
      double h2 = one / h1;
      double g1 = dd / ee / (h2 - h1);
      double g2 = -g1;
      double f1 = one / std::sqrt(ee);
      double f2 = f1;
      if (uf > u0) {
        iu = su * elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, 0.0, c.rffu0, u0, uf) / std::sqrt(std::fabs(ee));
      } else if (uf < u0) {
        iu = -su * elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, 0.0, c.rffu0, uf, u0) / std::sqrt(std::fabs(ee));
      } else {
        iu = 0.0;
      }
    } else {
      if (std::fabs(u3 - u2) > 1.e-12) {
        iu1 = 0.0;
        if (c.ncase == 7) {
          if (uf > u2) uf = u2;
          if (firstpt && (u0 != u2)) {
            c.iu0 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, 0.0, c.rffu0, u0, u2);
          } else if (u0 == u2) {
            c.iu0 = 0.0;
          }
          if (uf != u2) iu1 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, 0.0, c.rffu1, uf, u2);
          iu = su * (c.iu0 - powm1(tpr) * iu1) / std::sqrt(std::fabs(ee));
        } else if (c.ncase == 8) {
          if (uf < u3) uf = u3;
          if (firstpt && (u0 != u3)) {
            c.iu0 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, 0.0, c.rffu0, u3, u0);
          } else if (u0 == u3) {
            c.iu0 = 0.0;
          }
          if (uf != u3) iu1 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, 0.0, c.rffu1, u3, uf);
          iu = su * (c.iu0 - powm1(tpr) * iu1) / std::sqrt(std::fabs(ee));
        }
      } else {
        tpr = 0;
        if (c.ncase == 7) {
          if (uf > u2) uf = u2;
          c.iu0 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, 0.0, c.rffu0, u0, uf);
          iu = su * c.iu0 / std::sqrt(std::fabs(ee));
        } else if (c.ncase == 8) {
          if (uf < u3) uf = u3;
          c.iu0 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, 0.0, c.rffu0, u0, uf);
          iu = su * c.iu0 / std::sqrt(std::fabs(ee));
        } else {
          // 'ERROR - Unphysical Quartic Real'
          c.ncase = 0;
        }
      }
    }
    double mneg, mpos;
    mroots(a, q2, ql2, mneg, mpos);
    if (mpos > 1.0) mpos = 1.0;
    muplus = std::sqrt(mpos);
    if (mneg < 0.0) {
      if (muplus < mu0) {
        muplus = mu0;
        mpos = muplus * muplus;
      }
      double muminus = -muplus;
      if (firstpt) calcimusym(a, mneg, mpos, mu0, muplus, c.i1mu, c.i3mu, c.rffmu1, c.rffmu3);
      double a1 = sm;
      if (sm == 1) {
        tpm = (int)((iu - c.i1mu) / c.i3mu) + (int)((1 + fsign(1.0, (iu - c.i1mu) / c.i3mu)) / 2.0);
      } else {
        tpm = (int)((iu + c.i1mu) / c.i3mu);
      }
      double a2 = sm * powm1(tpm);
      double a3 = two * (int)((two * double(tpm) + 3.0 - sm) / 4.0) - one;
      double iarg = std::fabs(a) / a2 * (iu - a1 * c.i1mu - a3 * c.i3mu);
      double uarg = std::sqrt(mpos - mneg) * iarg;
      double m1 = -mneg / (mpos - mneg);
      sncndn(uarg, m1, sn, cn, dn);
      muf = muminus * cn;
      if (pht) calcimusymf(a, mneg, mpos, muf, muplus, i2mu, c.i3mu, c.rffmu2);
    } else {
      muplus = fsign(1.0, mu0) * muplus;
      double muminus = fsign(1.0, mu0) * std::sqrt(mneg);
      if (std::fabs(muminus) > std::fabs(mu0)) {
        muminus = mu0;
        mneg = muminus * muminus;
      }
      if (std::fabs(mu0) > std::fabs(muplus)) {
        muplus = mu0;
        mpos = muplus * muplus;
      }
      if (firstpt) calcimuasym(a, mneg, mpos, mu0, muplus, c.i1mu, c.i3mu, c.rffmu1, c.rffmu3);
      double a1 = sm;
      if (sm == 1) {
        tpm = (int)((iu - c.i1mu) / c.i3mu) + (int)(std::fabs((1 + fsign(1.0, (iu - c.i1mu) / c.i3mu)) / 2.0));
      } else {
        tpm = (int)((iu + c.i1mu) / c.i3mu);
      }
      double a2 = sm * powm1(tpm);
      double a3 = two * (int)((two * double(tpm) + 3.0 - sm) / 4.0) - one;
      double iarg = std::fabs(a) / a2 * (iu - a1 * c.i1mu - a3 * c.i3mu);
      double uarg = std::fabs(muplus) * iarg;
      double m1 = mneg / mpos;
      sncndn(uarg, m1, sn, cn, dn);
      muf = muminus / dn;
      if (pht) calcimuasymf(a, mneg, mpos, muf, muplus, i2mu, c.i3mu, c.rffmu2);
    }
  }
}

// ---------------------------------------------------------------- GEOPHITIME  geokerr_wrapper.f:1648-2357
void geophitime(double u0, double uf, double mu0, double muf, double a, double l, double l2, double q2, int tpm, int tpr,
                double su, double sm, double iu, double &phimu, double &tmu, double &phiu, double &tu, double &lambda,
                GeoCache &c, bool firstpt) {
  const double one = 1., two = 2.;
  const double pi = std::acos(-one);
  int p[6] = {0, -1, -1, -1, 0, 0};
  double uplus = one / (one + std::sqrt(one - a * a));
  double umi = one - std::sqrt(one - a * a);
  double ur = -one / (two * std::sqrt(one - a * a));
  double qs = fsign(1.0, q2);
  double dd = two * ((a - l) * (a - l) + q2);
  double ee = -a * a * q2;
  double ql2 = q2 + l2;
  double a1 = sm;
  double a2 = sm * powm1(tpm);
  double a3 = two * (int)((two * double(tpm) + 3.0 - sm) / 4.0) - one;
  double &u1 = c.u1, &u2 = c.u2, &u3 = c.u3, &u4 = c.u4;
  double lambdau = 0.0;
  double phiu01 = 0, phiu02 = 0, phiu11 = 0, phiu12 = 0, tu11 = 0, tu12 = 0, tu13 = 0, tu14 = 0;
  double tu0, tu1 = 0, tu2 = 0, tu3 = 0, tu4 = 0, phiu0 = 0, phiu1 = 0, phiu2 = 0, phimu2 = 0, tmu2 = 0;
  // common coefficient groups of Eq. (47)/(48)
  const double ct1 = (umi - one / uplus), ct2 = (umi * umi - one / (uplus * uplus));
  const double ct3 = (2.0 * a * (a - l) + a * a / uplus + one / (uplus * uplus * uplus));
  const double ct4 = (2.0 * a * (a - l) + a * a * umi + umi * umi * umi);
  const double cp1 = (l / uplus + 2.0 * (a - l)), cp2 = (l * umi + 2.0 * (a - l));
  (void)phiu2;
  if (c.ncase == 0) {
    tu = 0.0;
    phiu = 0.0;
    phimu = 0.0;
    tmu = 0.0;
    lambdau = 0.0;
  } else if (c.ncase < 3) {
    p[5] = 0;
    if (std::fabs(u3 - u2) < 1.e-12) {
      tu = su * tfnkerr(u0, uf, u1, u2, l, a) / std::sqrt(dd);
      phiu = su * (phifnkerr(uf, u1, u2, l, a) - phifnkerr(u0, u1, u2, l, a)) / std::sqrt(dd);
      if (u0 >= u3) {
        phiu = -phiu;
        tu = -tu;
      }
    } else if (u0 <= u2) {
      if (firstpt && (u0 != u2)) {
        p[4] = -2;
        phiu01 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, one, -one / uplus, c.rffu0, u0, u2);
        phiu02 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, one, -umi, c.rffu0, u0, u2);
        c.tu02 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, one, c.rffu0, u0, u2);
        c.tu03 = phiu01;
        c.tu04 = phiu02;
        p[4] = -4;
        c.tu01 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, one, c.rffu0, u0, u2);
      } else if (u0 == u2) {
        c.tu01 = c.tu02 = c.tu03 = c.tu04 = 0.0;
        phiu01 = phiu02 = 0.0;
      } else {
        phiu01 = c.tu03;
        phiu02 = c.tu04;
      }
      if (std::fabs(uf - u2) > 1e-16) {
        p[4] = -2;
        phiu11 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, one, -one / uplus, c.rffu1, uf, u2);
        phiu12 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, one, -umi, c.rffu1, uf, u2);
        tu12 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, one, c.rffu1, uf, u2);
        tu13 = phiu11;
        tu14 = phiu12;
        p[4] = -4;
        tu11 = ellcubicreal(p, -u1, one, u2, -one, u3, -one, 0.0, one, c.rffu1, uf, u2);
      } else {
        tu11 = tu12 = tu13 = tu14 = 0.0;
        phiu11 = phiu12 = 0.0;
      }
      tu0 = ct1 * c.tu01 + ct2 * c.tu02 - ct3 * c.tu03 + ct4 * c.tu04;
      tu1 = ct1 * tu11 + ct2 * tu12 - ct3 * tu13 + ct4 * tu14;
      tu = su * (tu0 - powm1(tpr) * tu1) / std::sqrt(dd) * ur;
      phiu0 = -cp1 * phiu01 + cp2 * phiu02;
      phiu1 = -cp1 * phiu11 + cp2 * phiu12;
      phiu = su * (phiu0 - powm1(tpr) * phiu1) / std::sqrt(dd) * ur;
      lambdau = su * (c.tu01 - powm1(tpr) * tu11) / std::sqrt(dd);
    } else if (u0 >= u3) {
      if (firstpt && (u0 != u3)) {
        p[4] = -2;
        phiu01 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, one, -one / uplus, c.rffu0, u3, u0);
        phiu02 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, one, -umi, c.rffu0, u3, u0);
        c.tu02 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, one, c.rffu0, u3, u0);
        c.tu03 = phiu01;
        c.tu04 = phiu02;
        p[4] = -4;
        c.tu01 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, one, c.rffu0, u3, u0);
      } else if (u0 == u3) {
        c.tu01 = c.tu02 = c.tu03 = c.tu04 = 0.0;
        phiu01 = phiu02 = 0.0;
      } else {
        phiu01 = c.tu03;
        phiu02 = c.tu04;
      }
      if (uf != u3) {
        p[4] = -2;
        phiu11 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, one, -one / uplus, c.rffu1, u3, uf);
        phiu12 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, one, -umi, c.rffu1, u3, uf);
        tu12 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, one, c.rffu1, u3, uf);
        tu13 = phiu11;
        tu14 = phiu12;
        p[4] = -4;
        tu11 = -ellcubicreal(p, -u1, one, -u2, one, -u3, one, 0.0, one, c.rffu1, u3, uf);
      } else {
        tu11 = tu12 = tu13 = tu14 = 0.0;
        phiu11 = phiu12 = 0.0;
      }
      tu0 = ct1 * c.tu01 + ct2 * c.tu02 - ct3 * c.tu03 + ct4 * c.tu04;
      tu1 = ct1 * tu11 + ct2 * tu12 - ct3 * tu13 + ct4 * tu14;
      tu = su * (tu0 - powm1(tpr) * tu1) / std::sqrt(dd) * ur;
      phiu0 = -cp1 * phiu01 + cp2 * phiu02;
      phiu1 = -cp1 * phiu11 + cp2 * phiu12;
      phiu = su * (phiu0 - powm1(tpr) * phiu1) / std::sqrt(dd) * ur;
      lambdau = su * (c.tu01 - powm1(tpr) * tu11) / std::sqrt(dd);
    }
  } else if (c.ncase == 3) {
    double f = -one / dd / u1;
    double g = f / u1;
    double h = one;
    if (u0 < uf) {
      p[4] = -4;
      tu1 = ellcubiccomplex(p, -u1, one, 0.0, one, f, g, h, c.rffu0, u0, uf);
      p[4] = -2;
      tu2 = ellcubiccomplex(p, -u1, one, 0.0, one, f, g, h, c.rffu0, u0, uf);
      tu3 = ellcubiccomplex(p, -u1, one, one, -one / uplus, f, g, h, c.rffu0, u0, uf);
      tu4 = ellcubiccomplex(p, -u1, one, one, -umi, f, g, h, c.rffu0, u0, uf);
      tu = su * ur / std::sqrt(dd) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      phiu0 = -tu3;
      phiu1 = -tu4;
      phiu = su * ur / std::sqrt(dd) * (cp1 * phiu0 - cp2 * phiu1);
      lambdau = su * tu1 / std::sqrt(dd);
    } else if (u0 > uf) {
      p[4] = -4;
      tu1 = -ellcubiccomplex(p, -u1, one, 0.0, one, f, g, h, c.rffu0, uf, u0);
      p[4] = -2;
      tu2 = -ellcubiccomplex(p, -u1, one, 0.0, one, f, g, h, c.rffu0, uf, u0);
      tu3 = -ellcubiccomplex(p, -u1, one, one, -one / uplus, f, g, h, c.rffu0, uf, u0);
      tu4 = -ellcubiccomplex(p, -u1, one, one, -umi, f, g, h, c.rffu0, uf, u0);
      tu = su * ur / std::sqrt(dd) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      phiu0 = -tu3;
      phiu1 = -tu4;
      phiu = su * ur / std::sqrt(dd) * (cp1 * phiu0 - cp2 * phiu1);
      lambdau = su * tu1 / std::sqrt(dd);
    } else {
      tu = 0.0;
      phiu = 0.0;
      lambdau = 0.0;
    }
  }
  if (q2 == 0.0) {
    double s1 = fsign(1.0, mu0);
    a1 = s1 * sm;
    a2 = s1 * sm * powm1(tpm + 1);
    if (std::fabs(l) < std::fabs(a)) {
      double muplus = s1 * std::sqrt(one - l2 / a / a);
      double sq = std::sqrt(one - muplus * muplus);
      c.phimu1 = fsign(1.0, muplus * l / a) * (std::atan((muplus - std::tan(std::asin(mu0 / muplus) / two)) / sq) +
                                               std::atan((muplus + std::tan(std::asin(mu0 / muplus) / two)) / sq));
      phimu2 = fsign(1.0, muplus * l / a) * (std::atan((muplus - std::tan(std::asin(muf / muplus) / two)) / sq) +
                                             std::atan((muplus + std::tan(std::asin(muf / muplus) / two)) / sq));
      phimu = a1 * c.phimu1 + a2 * phimu2;
      tmu = std::fabs(muplus * a) * (a2 * std::sqrt(1.0 - muf * muf / (muplus * muplus)) +
                                     a1 * std::sqrt(1.0 - mu0 * mu0 / (muplus * muplus)));
    } else {
      phimu = 0.0;
      tmu = 0.0;
    }
  } else if (a == 0.0) {
    tmu = 0.0;
    double muplus = std::sqrt(q2 / ql2);
    double vfm = (muf - muplus * muplus) / (one - muf) / muplus;
    double vfp = -(muf + muplus * muplus) / (one + muf) / muplus;
    if (std::fabs(vfm) > one) vfm = fsign(1.0, vfm) * one;
    if (std::fabs(vfp) > one) vfp = fsign(1.0, vfp) * one;
    double vsm, vsp;
    if ((mu0 - muplus * muplus) == 0.0)
      vsm = -one;
    else
      vsm = (mu0 - muplus * muplus) / (one - mu0) / muplus;
    if ((mu0 + muplus * muplus) == 0.0)
      vsp = -one;
    else
      vsp = -(mu0 + muplus * muplus) / (one + mu0) / muplus;
    if (std::fabs(vsp) > one) vsp = fsign(1.0, vsp) * one;
    if (std::fabs(vsm) > one) vsm = fsign(1.0, vsm) * one;
    c.phimu1 = pi - std::asin(vsm) + std::asin(vsp);
    phimu2 = std::asin(vfm) - std::asin(vfp) + pi;
    c.phimu3 = two * pi;
    phimu = -l * iu + fsign(1.0, l) * 0.5 * (a1 * c.phimu1 + a2 * phimu2 + a3 * c.phimu3);
  }
  if (c.ncase == 4) {
    phiu = su * ur * (l * std::log((uf / uplus - one) / (u0 / uplus - one)) -
                      l * umi * umi * std::log((uf * umi - one) / (u0 * umi - one)));
    tu = su * ur *
         (ct1 * (one / u0 - one / uf) + ct2 * std::log(uf / u0) +
          (a * a / uplus + one / (uplus * uplus * uplus)) * uplus * std::log((uf / uplus - one) / (u0 / uplus - one)) -
          (a * a * umi + umi * umi * umi) * umi * std::log((uf * umi - one) / (u0 * umi - one)));
    lambdau = su * (one / u0 - one / uf);
  } else if (c.ncase == 5) {
    p[4] = -1;
    double f = -qs * one / std::fabs(ee) / u1 / u4;
    double g = (u4 + u1) / u1 / u4 * f;
    double h = one;
    if (u0 < uf) {
      p[5] = -4;
      tu1 = ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, 0.0, one, f, g, h, c.rffu0, u0, uf);
      p[5] = -2;
      tu2 = ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, 0.0, one, f, g, h, c.rffu0, u0, uf);
      tu3 = ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, one, -one / uplus, f, g, h, c.rffu0, u0, uf);
      tu4 = ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, one, -umi, f, g, h, c.rffu0, u0, uf);
      tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      lambdau = su * tu1 / std::sqrt(std::fabs(ee));
      phiu0 = -tu3;
      phiu1 = -tu4;
      phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
    } else if (u0 > uf) {
      p[5] = -4;
      tu1 = -ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, 0.0, one, f, g, h, c.rffu0, uf, u0);
      p[5] = -2;
      tu2 = -ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, 0.0, one, f, g, h, c.rffu0, uf, u0);
      tu3 = -ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, one, -one / uplus, f, g, h, c.rffu0, uf, u0);
      tu4 = -ellquarticcomplex(p, -u1, one, u4 * qs, -one * qs, one, -umi, f, g, h, c.rffu0, uf, u0);
      tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      phiu0 = -tu3;
      phiu1 = -tu4;
      phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
      lambdau = su * tu1 / std::sqrt(std::fabs(ee));
    } else {
      tu = 0.0;
      phiu = 0.0;
      lambdau = 0.0;
    }
  } else if (c.ncase == 6) {
    p[4] = -1;
    double h1 = c.h1;
    double h2 = one / h1;
    double g1 = dd / ee / (h2 - h1);
    double g2 = -g1;
    double f1 = one / std::sqrt(ee);
    double f2 = f1;
    if (u0 < uf) {
      p[5] = -4;
      tu1 = elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, one, c.rffu0, u0, uf);
      p[5] = -2;
      tu2 = elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, one, c.rffu0, u0, uf);
      tu3 = elldoublecomplex(p, f1, g1, h1, f2, g2, h2, one, -one / uplus, c.rffu0, u0, uf);
      tu4 = elldoublecomplex(p, f1, g1, h1, f2, g2, h2, one, -umi, c.rffu0, u0, uf);
      phiu0 = -tu3;
      phiu1 = -tu4;
      tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
      lambdau = su * tu1 / std::sqrt(-qs * ee);
    } else if (u0 > uf) {
      p[5] = -4;
      tu1 = -elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, one, c.rffu0, uf, u0);
      p[5] = -2;
      tu2 = -elldoublecomplex(p, f1, g1, h1, f2, g2, h2, 0.0, one, c.rffu0, uf, u0);
      tu3 = -elldoublecomplex(p, f1, g1, h1, f2, g2, h2, one, -one / uplus, c.rffu0, uf, u0);
      tu4 = -elldoublecomplex(p, f1, g1, h1, f2, g2, h2, one, -umi, c.rffu0, uf, u0);
      phiu0 = -tu3;
      phiu1 = -tu4;
      tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
      phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
      lambdau = su * tu1 / std::sqrt(-qs * ee);
    } else {
      tu = 0.0;
      phiu = 0.0;
      // (reference leaves lambdau unset here; it is 0 in this restatement)
    }
  } else if (c.ncase > 6) {
    p[4] = -1;
    if (std::fabs(u3 - u2) < 1.e-12) {
      // equal-roots quartic cases (reference :2086-2173; note it combines phiu0/phiu1 as written)
      if (u0 < uf) {
        if (u0 < u2) {
          p[5] = -2;
          phiu1 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -one / uplus, c.rffu0, u0, uf);
          phiu2 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -umi, c.rffu0, u0, uf);
          tu2 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, u0, uf);
          tu3 = phiu1;
          tu4 = phiu2;
          p[5] = -4;
          tu1 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, u0, uf);
        } else if (u0 > u3) {
          // (reference does not reset p(5)=-2 here; p(5) is still 0 from initialisation => RF-only value)
          phiu1 = ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -one / uplus, c.rffu0, u0, uf);
          phiu2 = ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -umi, c.rffu0, u0, uf);
          tu2 = ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, u0, uf);
          tu3 = phiu1;
          tu4 = phiu2;
          p[5] = -4;
          tu1 = ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, u0, uf);
        }
        tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
        phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
        lambdau = tu1 / std::sqrt(-qs * ee);
      } else if (u0 > uf) {
        if (u0 < u2) {
          p[5] = -2;
          phiu1 = -ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -one / uplus, c.rffu0, uf, u0);
          phiu2 = -ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -umi, c.rffu0, uf, u0);
          tu2 = -ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, uf, u0);
          tu3 = -phiu1;
          tu4 = -phiu2;
          p[5] = -4;
          tu1 = -ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, uf, u0);
        } else if (u0 > u3) {
          p[5] = -2;
          phiu1 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -one / uplus, c.rffu0, uf, u0);
          phiu2 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -umi, c.rffu0, uf, u0);
          tu2 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, uf, u0);
          tu3 = -phiu1;
          tu4 = -phiu2;
          p[5] = -4;
          tu1 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, uf, u0);
        }
        tu = su * ur / std::sqrt(-qs * ee) * (ct1 * tu1 + ct2 * tu2 - ct3 * tu3 + ct4 * tu4);
        phiu = su * ur / std::sqrt(-qs * ee) * (cp1 * phiu0 - cp2 * phiu1);
        lambdau = tu1 / std::sqrt(-qs * ee);
      } else {
        tu = 0.0;
        phiu = 0.0;
        lambdau = 0.0;
      }
    } else if (u0 <= u2) {
      if (firstpt && (u0 != u2)) {
        p[5] = -2;
        phiu01 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -one / uplus, c.rffu0, u0, u2);
        phiu02 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -umi, c.rffu0, u0, u2);
        c.tu02 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, u0, u2);
        c.tu03 = phiu01;
        c.tu04 = phiu02;
        p[5] = -4;
        c.tu01 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu0, u0, u2);
      } else if (u0 == u2) {
        c.tu01 = c.tu02 = c.tu03 = c.tu04 = 0.0;
        phiu01 = phiu02 = 0.0;
      } else {
        phiu01 = c.tu03;
        phiu02 = c.tu04;
      }
      if (std::fabs(uf - u2) > 1e-12) {
        p[5] = -2;
        phiu11 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -one / uplus, c.rffu1, uf, u2);
        phiu12 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, one, -umi, c.rffu1, uf, u2);
        tu12 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu1, uf, u2);
        tu13 = phiu11;
        tu14 = phiu12;
        p[5] = -4;
        tu11 = ellquarticreal(p, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, one, c.rffu1, uf, u2);
      } else {
        tu11 = tu12 = tu13 = tu14 = 0.0;
        phiu11 = phiu12 = 0.0;
      }
      tu0 = ct1 * c.tu01 + ct2 * c.tu02 - ct3 * c.tu03 + ct4 * c.tu04;
      tu1 = ct1 * tu11 + ct2 * tu12 - ct3 * tu13 + ct4 * tu14;
      tu = su * (tu0 - powm1(tpr) * tu1) / std::sqrt(std::fabs(ee)) * ur;
      phiu0 = -cp1 * phiu01 + cp2 * phiu02;
      phiu1 = -cp1 * phiu11 + cp2 * phiu12;
      phiu = su * (phiu0 - powm1(tpr) * phiu1) / std::sqrt(std::fabs(ee)) * ur;
      lambdau = su * (c.tu01 - powm1(tpr) * tu11) / std::sqrt(std::fabs(ee));
    } else if (u0 >= u3) {
      if (firstpt && (u0 != u3)) {
        p[5] = -2;
        phiu01 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -one / uplus, c.rffu0, u3, u0);
        phiu02 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -umi, c.rffu0, u3, u0);
        c.tu02 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, u3, u0);
        c.tu03 = phiu01;
        c.tu04 = phiu02;
        p[5] = -4;
        c.tu01 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu0, u3, u0);
      } else if (u0 == u3) {
        c.tu01 = c.tu02 = c.tu03 = c.tu04 = 0.0;
        phiu01 = phiu02 = 0.0;
      } else {
        phiu01 = c.tu03;
        phiu02 = c.tu04;
      }
      if (uf != u3) {
        p[5] = -2;
        phiu11 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -one / uplus, c.rffu1, u3, uf);
        phiu12 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, one, -umi, c.rffu1, u3, uf);
        tu12 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu1, u3, uf);
        tu13 = phiu11;
        tu14 = phiu12;
        p[5] = -4;
        tu11 = -ellquarticreal(p, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, one, c.rffu1, u3, uf);
      } else {
        tu11 = tu12 = tu13 = tu14 = 0.0;
        phiu11 = phiu12 = 0.0;
      }
      tu0 = ct1 * c.tu01 + ct2 * c.tu02 - ct3 * c.tu03 + ct4 * c.tu04;
      tu1 = ct1 * tu11 + ct2 * tu12 - ct3 * tu13 + ct4 * tu14;
      tu = su * (tu0 - powm1(tpr) * tu1) / std::sqrt(std::fabs(ee)) * ur;
      phiu0 = -cp1 * phiu01 + cp2 * phiu02;
      phiu1 = -cp1 * phiu11 + cp2 * phiu12;
      phiu = su * (phiu0 - powm1(tpr) * phiu1) / std::sqrt(std::fabs(ee)) * ur;
      lambdau = su * (c.tu01 - powm1(tpr) * tu11) / std::sqrt(std::fabs(ee));
    }
  }
  if (c.ncase > 4) {
    double mneg, mpos;
    mroots(a, q2, ql2, mneg, mpos);
    if (mpos > 1.0) mpos = 1.0;
    double muplus = std::sqrt(mpos);
    a1 = sm;
    a2 = sm * powm1(tpm + 1);
    a3 = 2.0 * (int)((2 * tpm - sm + 1) / 4.0);
    if (mneg < 0.0) {
      if (muplus < mu0) {
        muplus = mu0;
        mpos = muplus * muplus;
      }
      calcphitmusym(a, mneg, mpos, mu0, muf, muplus, c.phimu1, phimu2, c.phimu3, c.tmu1, tmu2, c.tmu3, c.rffmu1,
                    c.rffmu2, c.rffmu3, c.rdc, c.rjc, firstpt);
      tmu = a * a * mneg * iu + (a1 * c.tmu1 + a2 * tmu2 + a3 * c.tmu3);
      phimu = -l * iu + l * (a1 * c.phimu1 + a2 * phimu2 + a3 * c.phimu3);
    } else {
      if (fsign(1.0, mu0) == -1) muplus = -muplus;
      if (std::fabs(muplus) < std::fabs(mu0)) {
        muplus = mu0;
        mpos = muplus * muplus;
      }
      if (std::fabs(mneg) > mu0 * mu0) mneg = mu0 * mu0;
      calcphitmuasym(a, mneg, mpos, mu0, muf, muplus, c.phimu1, phimu2, c.phimu3, c.tmu1, tmu2, c.tmu3, c.rffmu1,
                     c.rffmu2, c.rffmu3, c.rdc, c.rjc, firstpt);
      tmu = (a1 * c.tmu1 + a2 * tmu2 + a3 * c.tmu3);
      phimu = -l * iu + l * (a1 * c.phimu1 + a2 * phimu2 + a3 * c.phimu3);
    }
  }
  lambda = lambdau + tmu;
  if (l == 0.0) phiu = phiu - fsign(1.0, a) * pi * tpm;
}

// ---------------------------------------------------------------- GEOR  geokerr_wrapper.f:2672-3134
void geor(double &u0, double &uf, double mu0, double muf, double a, double l, double l2, double q2, double &imu,
          int &tpm, int &tpr, double su, double sm, GeoCache &c, bool pht, bool firstpt) {
  const double one = 1., two = 2., half = 0.5, third = 0.3333333333333333, three = 3.;
  const double pi = std::acos(-one);
  int parr[6] = {0, 0, 0, 0, 0, 0};
  double uplus = one / (one + std::sqrt(one - a * a));
  double ql2 = q2 + l2;
  double cc = a * a - q2 - l2;
  double dd = two * ((a - l) * (a - l) + q2);
  double ee = -a * a * q2;
  double &u1 = c.u1, &u2 = c.u2, &u3 = c.u3, &u4 = c.u4;
  double a1, a2, a3, muplus, i2mu = 0, sn, cn, dn, dummy = 0, iu1 = 0, iut, iu = 0;
  (void)iu1;
  (void)iu;
  if (q2 == 0.0) {
    if (l2 >= a * a || mu0 == 0.0) {
      imu = 0.0;
      uf = -1.0;
      c.ncase = 0;
      return;
    } else {
      double s1 = fsign(1.0, mu0);
      a1 = s1 * sm;
      a2 = s1 * sm * powm1(tpm + 1);
      muplus = s1 * std::sqrt(one - l2 / a / a);
      c.i1mu = one / std::fabs(a * muplus) * asech(mu0 / muplus);
      i2mu = one / std::fabs(a * muplus) * asech(muf / muplus);
      imu = a1 * c.i1mu + a2 * i2mu;
    }
  } else if (a == 0.0) {
    a1 = sm;
    muplus = std::sqrt(q2 / ql2);
    if (mu0 > muplus) muplus = mu0;
    c.i1mu = (half * pi - std::asin(mu0 / muplus)) / std::sqrt(ql2);
    c.i3mu = pi / std::sqrt(ql2);
    a2 = sm * powm1(tpm);
    a3 = two * (int)((two * double(tpm) + 3.0 - sm) / 4.0) - one;
    i2mu = (half * pi + std::asin(muf / muplus)) / std::sqrt(ql2);
    imu = a1 * c.i1mu + a2 * i2mu + a3 * c.i3mu;
  }
  if ((ee == 0.0) && (dd != 0.0)) {
    parr[1] = -1; parr[2] = -1; parr[3] = -1; parr[4] = 0; parr[5] = 0;
    double qq = cc * cc / dd / dd / 9.0;
    double rr = (two * cc * cc * cc / (dd * dd * dd) + 27.0 / dd) / 54.0;
    double dis = rr * rr - qq * qq * qq;
    if (dis < -1.e-16) {
      double theta = std::acos(rr / std::pow(qq, 1.5));
      u1 = -two * std::sqrt(qq) * std::cos(theta / 3.0) - cc / dd / 3.0;
      u2 = -two * std::sqrt(qq) * std::cos((theta - two * pi) / 3.0) - cc / dd / 3.0;
      u3 = -two * std::sqrt(qq) * std::cos((theta + two * pi) / 3.0) - cc / dd / 3.0;
      for (;;) {  // label 80
        if (u0 <= u2) {
          c.ncase = 1;
          if (firstpt && (u0 != u2)) {
            c.iu0 = ellcubicreal(parr, -u1, one, u2, -one, u3, -one, 0.0, 0.0, c.rffu0, u0, u2);
          } else if (u0 == u2) {
            c.iu0 = 0.0;
          }
          double m1 = (u3 - u2) / (u3 - u1);
          double jarg = std::sqrt((u3 - u1) * dd) * half * (imu - su * c.iu0 / std::sqrt(dd));
          sncndn(jarg, m1, sn, cn, dn);
          double cd2 = cn * cn / (dn * dn);
          uf = u1 + (u2 - u1) * cd2;
          tpr = (int)((fsign(1.0, su * (imu - su * c.iu0 / std::sqrt(dd))) + 1.0) / 2.0);
          if (pht) iu1 = ellcubicreal(parr, -u1, one, u2, -one, u3, -one, 0.0, 0.0, c.rffu1, uf, u2) / std::sqrt(dd);
          break;
        } else if (u0 >= u3) {
          c.ncase = 2;
          if (firstpt && (u0 != u3)) {
            c.iu0 = ellcubicreal(parr, -u1, one, -u2, one, -u3, one, 0.0, 0.0, c.rffu0, u3, u0);
          } else if (u0 != u3) {  // (sic) reference :2808 tests .ne. here
            c.iu0 = 0.0;
          }
          double m1 = (u3 - u2) / (u3 - u1);
          double jarg = std::sqrt((u3 - u1) * dd) * half * (imu + su * c.iu0 / std::sqrt(dd));
          sncndn(jarg, m1, sn, cn, dn);
          double dc2 = dn * dn / (cn * cn);
          dummy = uf;
          uf = u1 + (u3 - u1) * dc2;
          tpr = (int)((-fsign(1.0, su * (imu + c.iu0 / std::sqrt(dd))) + 1.0) / 2.0);
          if (pht) iu1 = ellcubicreal(parr, -u1, one, -u2, one, -u3, one, 0.0, 0.0, c.rffu1, u3, uf) / std::sqrt(dd);
          break;
        } else {
          if (su == 1)
            u0 = u3;
          else
            u0 = u2;
        }
      }
    } else if (std::fabs(dis) < 1.e-16) {
      tpr = 0;
      u1 = -two * std::sqrt(qq) - cc / dd / 3.0;
      u2 = -two * std::sqrt(qq) * std::cos((two * pi) / 3.0) - cc / dd / 3.0;
      u3 = u2;
      tpr = 0;
      if (u0 <= u2) {
        c.ncase = 1;
        if (uf > u2) {
          uf = u2;
        } else {
          double sarg = std::sqrt((u0 - u1) / (u2 - u1));
          double jarg = std::sqrt((u2 - u1) * dd) * half * su * imu + half * std::log((one + sarg) / (one - sarg));
          double th = std::tanh(jarg);
          uf = u1 + (u2 - u1) * th * th;
        }
      } else if (u0 >= u2) {
        c.ncase = 2;
        if (uf < u2) {
          uf = u2;
        } else {
          double sarg = std::sqrt((u2 - u1) / (u0 - u1));
          double jarg = -std::sqrt((u2 - u1) * dd) * half * su * imu + half * std::log((one + sarg) / (one - sarg));
          double th = std::tanh(jarg);
          uf = u1 + (u2 - u1) / (th * th);
        }
      }
    } else {
      c.ncase = 3;
      tpr = 0;
      double aa = -fsign(1.0, rr) * std::pow(std::fabs(rr) + std::sqrt(dis), third);
      double bb;
      if (aa != 0.0)
        bb = qq / aa;
      else
        bb = 0.0;
      u1 = (aa + bb) - cc / dd / 3.0;
      double f = -one / dd / u1;
      double g = f / u1;
      if (su > 0.0) {
        iut = ellcubiccomplex(parr, -u1, one, 0.0, 0.0, f, g, one, dummy, u0, uplus) / std::sqrt(dd);
      } else {
        iut = ellcubiccomplex(parr, -u1, one, 0.0, 0.0, f, g, one, dummy, 0.0, u0) / std::sqrt(dd);
      }
      if (imu > iut) {
        uf = -1.0;
        c.ncase = 0;
        tpr = 0;
        return;
      }
      if (firstpt) c.iu0 = su * ellcubiccomplex(parr, -u1, one, 0.0, 0.0, f, g, one, dummy, u1, u0) / std::sqrt(dd);
      double c3 = -one;
      if (a != 0.0 || l != 0.0) c3 = (a + l) / (a - l);
      double c2 = std::sqrt(u1 * (three * u1 + c3));
      double c1 = std::sqrt(c2 * dd);
      double m1 = half + (6.0 * u1 + c3) / (8.0 * c2);
      double jarg = c1 * (imu + c.iu0);
      sncndn(jarg, m1, sn, cn, dn);
      uf = (c2 + u1 - (c2 - u1) * cn) / (one + cn);
      if (pht) iu = ellcubiccomplex(parr, -u1, one, 0.0, 0.0, f, g, one, c.rffu0, u0, uf) / std::sqrt(dd);
    }
  } else if (ee == 0.0 && dd == 0.0) {
    c.ncase = 4;
    uf = -1.0;
    tpr = 0;
  } else {
    double mneg, mpos;
    mroots(a, q2, ql2, mneg, mpos);
    muplus = std::sqrt(mpos);
    a1 = sm;
    a2 = sm * powm1(tpm + 1);
    a3 = 2.0 * (int)((2 * tpm - sm + 1) / 4.0);
    if (mneg < 0.0) {
      if (mu0 > muplus) {
        muplus = mu0;
        mpos = muplus * muplus;
      }
      if (firstpt) calcimusym(a, mneg, mpos, mu0, muplus, c.i1mu, c.i3mu, c.rffmu1, c.rffmu3);
      calcimusymf(a, mneg, mpos, muf, muplus, i2mu, c.i3mu, c.rffmu2);
      imu = a1 * c.i1mu + a2 * i2mu + a3 * c.i3mu;
    } else {
      if (std::fabs(muf) < std::sqrt(mneg)) {
        uf = -1.0;
        c.ncase = 0;
        tpr = 0;
        return;
      } else {
        if (fsign(1.0, mu0) == -1) muplus = -muplus;
        if (std::fabs(muplus) < std::fabs(mu0)) {
          muplus = mu0;
          mpos = muplus * muplus;
        }
        mneg = std::min(mu0 * mu0, mneg);
        if (firstpt) calcimuasym(a, mneg, mpos, mu0, muplus, c.i1mu, c.i3mu, c.rffmu1, c.rffmu3);
        calcimuasymf(a, mneg, mpos, muf, muplus, i2mu, c.i3mu, c.rffmu2);
      }
      imu = a1 * c.i1mu + a2 * i2mu + a3 * c.i3mu;
    }
    parr[1] = -1; parr[2] = -1; parr[3] = -1; parr[4] = -1; parr[5] = 0;
    if (firstpt) {
      cplx cf[6], root[5];
      cf[1] = cplx(one, 0.0);
      cf[2] = cplx(0.0, 0.0);
      cf[3] = cplx(cc, 0.0);
      cf[4] = cplx(dd, 0.0);
      cf[5] = cplx(ee, 0.0);
      zroots(cf, 4, root, true);
      int nreal = 0;
      for (int i = 1; i <= 4; i++)
        if (root[i].imag() == 0.0) nreal++;
      if (nreal == 2) {
        c.ncase = 5;
        tpr = 0;
        u1 = root[1].real();
        if (root[2].imag() == 0.0)
          u4 = root[2].real();
        else
          u4 = root[4].real();
      } else if (nreal == 0) {
        c.ncase = 6;
        tpr = 0;
      } else {
        u1 = root[1].real();
        u2 = root[2].real();
        u3 = root[3].real();
        u4 = root[4].real();
        for (;;) {  // label 90
          if (u2 > uplus && u3 > uplus) {
            c.ncase = 5;
            break;
          } else if (u0 <= u2) {
            c.ncase = 7;
            break;
          } else if (u0 >= u3) {
            c.ncase = 8;
            break;
          } else {
            if (su == 1)
              u0 = u3;
            else
              u0 = u2;
          }
        }
      }
    }
    if (c.ncase == 5) {
      double qs = fsign(1.0, q2);
      double f = -qs * one / std::fabs(ee) / u1 / u4;
      double g = (u4 + u1) / u1 / u4 * f;
      if (su > 0.0) {
        iut = ellquarticcomplex(parr, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, dummy, u0, uplus) /
              std::sqrt(std::fabs(ee));
      } else {
        iut = ellquarticcomplex(parr, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, dummy, 0.0, u0) /
              std::sqrt(std::fabs(ee));
      }
      if (imu > iut) {
        uf = -1.0;
        c.ncase = 0;
        tpr = 0;
        return;
      }
      double ub;
      if (qs == one) {
        ub = u1;
      } else {
        ub = u4;
      }
      double m = -g * half;
      double n2 = f - g * g / 4.0;
      double c4 = std::sqrt((m - u4) * (m - u4) + n2);
      double c5 = std::sqrt((m - u1) * (m - u1) + n2);
      double c1 = std::sqrt(std::fabs(ee * c4 * c5));
      if (firstpt)
        c.iu0 = su * ellquarticcomplex(parr, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, dummy, ub, u0) /
                std::sqrt(std::fabs(ee));
      double m1 = qs * ((c4 + qs * c5) * (c4 + qs * c5) - (u4 - u1) * (u4 - u1)) / (4.0 * c4 * c5);
      double jarg = c1 * (imu + c.iu0);
      sncndn(jarg, m1, sn, cn, dn);
      uf = (u4 * c5 + qs * u1 * c4 - (qs * u4 * c5 - u1 * c4) * cn) / ((c4 - qs * c5) * cn + (qs * c4 + c5));
      if (pht)
        iu = ellquarticcomplex(parr, -u1, one, qs * u4, -qs * one, 0.0, 0.0, f, g, one, c.rffu0, u0, uf) /
             std::sqrt(std::fabs(ee));
    } else if (c.ncase == 6) {
      c.ncase = 6;
      cplx coefs[8], hroots[7];
      coefs[1] = cplx(one, 0.0);
      coefs[2] = cplx(-cc / std::sqrt(ee), 0.0);
      coefs[3] = cplx(-one, 0.0);
      coefs[4] = cplx(std::sqrt(ee) * (two * cc / ee - (dd / ee) * (dd / ee)), 0.0);
      coefs[5] = cplx(-one, 0.0);
      coefs[6] = cplx(-cc / std::sqrt(ee), 0.0);
      coefs[7] = cplx(one, 0.0);
      zroots(coefs, 6, hroots, true);
      int i = 0;
      c.h1 = 0.0;
      do {
        i = i + 1;
        if (i > 6) break;
        if (hroots[i].imag() == 0.0) c.h1 = hroots[i].real();
      } while (c.h1 == 0.0);
      double h1 = c.h1;
      double h2 = one / h1;
      double g1 = dd / ee / (h2 - h1);
      double g2 = -g1;
      double f1 = one / std::sqrt(ee);
      double f2 = f1;
      if (su > 0.0) {
        iut = elldoublecomplex(parr, f1, g1, h1, f2, g2, h2, 0.0, 0.0, dummy, u0, uplus) / std::sqrt(std::fabs(ee));
      } else {
        iut = elldoublecomplex(parr, f1, g1, h1, f2, g2, h2, 0.0, 0.0, dummy, 0.0, u0) / std::sqrt(std::fabs(ee));
      }
      if (imu > iut) {
        uf = -1.0;
        c.ncase = 0;
        tpr = 0;
        return;
      }
      coefs[1] = std::pow(ee, -three);
      coefs[2] = -cc / (ee * ee * ee);
      coefs[3] = -std::pow(ee, -two);
      coefs[4] = -std::pow(ee, -two) * (dd * dd / ee - two * cc);
      coefs[5] = -one / ee;
      coefs[6] = -cc / ee;
      coefs[7] = one;
      zroots(coefs, 6, hroots, true);
      i = 0;
      double mn2 = 0.0;
      do {
        i = i + 1;
        if (i > 6) break;
        if (hroots[i].imag() == 0.0) mn2 = hroots[i].real();
      } while (mn2 == 0.0);
      double pp = dd / (two * ee * ee * (mn2 * mn2 - one / ee));
      double m = -half * dd / ee - pp;
      if (m < pp) {
        double temp = pp;
        pp = m;
        m = temp;
      }
      double pr2 = one / (mn2 * ee);
      double n = std::sqrt(mn2 - m * m);
      double r = std::sqrt(pr2 - pp * pp);
      double c4 = std::sqrt((m - pp) * (m - pp) + (n + r) * (n + r));
      double c5 = std::sqrt((m - pp) * (m - pp) + (n - r) * (n - r));
      double c1 = (c4 + c5) * half * std::sqrt(std::fabs(ee));
      double c2 = std::sqrt((4.0 * n * n - (c4 - c5) * (c4 - c5)) / ((c4 + c5) * (c4 + c5) - 4.0 * n * n));
      double c3 = m + c2 * n;
      if (firstpt)
        c.iu0 = fsign(1.0, (u0 - c3)) * elldoublecomplex(parr, f1, g1, h1, f2, g2, h2, 0.0, 0.0, dummy, c3, u0) /
                std::sqrt(std::fabs(ee));
      double m1 = ((c4 - c5) / (c4 + c5)) * ((c4 - c5) / (c4 + c5));
      double jarg = c1 * (su * imu + c.iu0);
      sncndn(jarg, m1, sn, cn, dn);
      double sc = sn / cn;
      uf = c3 + (n * (one + c2 * c2) * sc) / (one - c2 * sc);
      if (pht) iu = elldoublecomplex(parr, f1, g1, h1, f2, g2, h2, 0.0, 0.0, c.rffu0, u0, uf) / std::sqrt(-ee);
    } else {
      if (c.ncase == 7) {
        if (firstpt && (u0 != u2))
          c.iu0 = ellquarticreal(parr, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, 0.0, c.rffu0, u0, u2);
        double jarg = std::sqrt(std::fabs(ee) * (u3 - u1) * (u4 - u2)) * half * (imu - su * c.iu0 / std::sqrt(-ee));
        double m1 = (u4 - u1) * (u3 - u2) / ((u4 - u2) * (u3 - u1));
        sncndn(jarg, m1, sn, cn, dn);
        dummy = uf;
        uf = ((u2 - u1) * u3 * sn * sn - u2 * (u3 - u1)) / ((u2 - u1) * sn * sn - (u3 - u1));
        tpr = (int)((fsign(1.0, su * (imu - su * c.iu0 / std::sqrt(std::fabs(ee)))) + 1.0) / 2.0);
        if (pht)
          iu1 = ellquarticreal(parr, -u1, one, u2, -one, u3, -one, u4, -one, 0.0, 0.0, c.rffu1, uf, u2) / std::sqrt(-ee);
      } else if (c.ncase == 8) {
        if (firstpt && (u0 != u3))
          c.iu0 = ellquarticreal(parr, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, 0.0, c.rffu0, u3, u0);
        double jarg = std::sqrt(std::fabs(ee) * (u3 - u1) * (u4 - u2)) * half * (imu + su * c.iu0 / std::sqrt(-ee));
        double m1 = (u4 - u1) * (u3 - u2) / ((u4 - u2) * (u3 - u1));
        sncndn(jarg, m1, sn, cn, dn);
        uf = ((u4 - u3) * u2 * sn * sn - u3 * (u4 - u2)) / ((u4 - u3) * sn * sn - (u4 - u2));
        tpr = (int)((-fsign(1.0, su * (imu + su * c.iu0 / std::sqrt(-ee))) + 1.0) / 2.0);
        if (pht)
          iu1 = ellquarticreal(parr, -u1, one, -u2, one, -u3, one, u4, -one, 0.0, 0.0, c.rffu1, u3, uf) / std::sqrt(-ee);
      }
    }
  }
}

// ---------------------------------------------------------------- FINDMROOTS geokerr_wrapper.f:607-642
static void findmroots(double q2, double l, double a, double mu0, double &muminus, double &muplus) {
  double ql2 = q2 + l * l;
  double mneg, mpos;
  mroots(a, q2, ql2, mneg, mpos);
  mpos = std::min(mpos, 1.0);
  if (mneg > 0.0) {
    muplus = fsign(1.0, mu0) * std::sqrt(mpos);
    muminus = fsign(1.0, mu0) * std::sqrt(mneg);
    if (std::fabs(muminus) > std::fabs(mu0)) muminus = mu0;
    if (std::fabs(mu0) > std::fabs(muplus)) muplus = mu0;
  } else {
    muplus = std::sqrt(mpos);
    if (muplus < mu0) muplus = mu0;
    muminus = -muplus;
  }
}

// ---------------------------------------------------------------- INDEP_MUF geokerr_wrapper.f:645-706
// NB: a1,a2,a3,a1k,a2k,a3k,dmukdmu1 are INTEGER in the reference; several sub-expressions are
// evaluated in default REAL (single precision) there: (-1.)**n, 2.*dtpm+3.-a1, float(npts), 1./a2k.
static void indep_muf(double muminus, double muplus, double mu0, double muf, int tpm0, int tpmf, double sm, int k,
                      int npts, double offset, double &mun, int &tpmk) {
  if (muf <= muplus && muf >= muminus) {
    int dtpm = tpmf - tpm0;
    int a1 = (int)(sm * (double)powm1(tpm0));
    int a2 = (int)(sm * (double)powm1(tpmf));
    int a3 = (int)(2 * (int)std::floor((2.f * (float)dtpm + 3.f - (float)a1) / 4.f) - 1.f);
    double dmu = a1 * (muplus - mu0) + a2 * (muf - muminus) + a3 * (muplus - muminus);
    double mu1 = (double)(((float)a1 + 1.f) / 2.f) * muplus + (double)((1.f - (float)a1) / 2.f) * muminus;
    double mu2 = (double)((1.f - (float)a1) / 2.f) * muplus + (double)(((float)a1 + 1.f) / 2.f) * muminus;
    double dmu1 = std::fabs(mu1 - mu0);
    double dmu2 = std::fabs(mu2 - mu1);
    double dmuk = (k - offset) * dmu / (double)(float)npts;
    int dmukdmu1 = (int)(dmuk / dmu1);
    if (dmukdmu1 < 0) dmukdmu1 = 1;
    tpmk = std::min(dmukdmu1, 1) + std::max((int)((dmuk - dmu1) / dmu2), 0);
    int a1k = (int)(sm * powm1(tpm0));
    int a2k = (int)(sm * powm1(tpmk + tpm0));
    int a3k = (int)(2.f * (float)(int)((2.f * (float)tpmk + 3.f - (float)a1k) / 4.f) - 1.f);
    mun = muminus + (double)(1.f / (float)a2k) * (dmuk - a1k * (muplus - mu0) - a3k * (muplus - muminus));
  } else {
    mun = 0.;
    tpmk = 0;
  }
  tpmk = tpmk + tpm0;
}

// ---------------------------------------------------------------- GEOKERR  geokerr_wrapper.f:357-604
// Arrays are 1-based of length nup+npts-2*kext (caller allocates length+1).
void geokerr(double u0, double uf, double uout, double mu0, double muf, double a, double l, double q2, double alpha,
             double beta, int tpm, int tpr, double su, double sm, int &nup, double offset, bool phit, bool usegeor,
             bool mufill, int &ncase_out, int kext, int npts, double *ufi, double *mufi, double *dti, double *dphi,
             int *tpmi, int *tpri, double *lambdai) {
  (void)alpha;
  const double one = 1., two = 2.;
  double l2 = l * l;
  bool firstpt = false;
  double uplus = one / (one + std::sqrt(one - a * a));
  GeoCache c;
  double iu = 0, un, mun = 0, tu = 0, tmu = 0, phiu = 0, phimu = 0, lambda = 0, ub;
  int k, tpr1, tpm1 = 0, kmax;
  if (usegeor) {
    k = 1;
    mun = mu0 + (k - offset) * (muf - mu0) / double(nup);
    if (mu0 != 0.0 || beta != 0.0) {
      geor(u0, uf, mu0, mun, a, l, l2, q2, iu, tpm, tpr, su, sm, c, phit, true);
      if (phit) geophitime(u0, uf, mu0, mun, a, l, l2, q2, tpm, tpr, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, true);
    } else {
      uf = u0;
    }
    ufi[k] = uf; mufi[k] = mun; dti[k] = tmu + tu; dphi[k] = phimu + phiu; tpmi[k] = tpm; tpri[k] = tpr;
    lambdai[k] = lambda;
    for (k = 2; k <= nup; k++) {
      mun = mu0 + (k - offset) * (muf - mu0) / double(nup);
      tpm = (int)((fsign(1.0, mu0) * sm + one) / two);
      if (mu0 != 0.0 || beta != 0.0) {
        geor(u0, uf, mu0, mun, a, l, l2, q2, iu, tpm, tpr, su, sm, c, phit, false);
        if (phit)
          geophitime(u0, uf, mu0, mun, a, l, l2, q2, tpm, tpr, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, true);
      } else {
        uf = u0;
      }
      ufi[k] = uf; mufi[k] = mun; dti[k] = tmu + tu; dphi[k] = phimu + phiu; tpmi[k] = tpm; tpri[k] = tpr;
      lambdai[k] = lambda;
    }
  } else {
    geomu(u0, uf, mu0, muf, a, l, l2, q2, iu, tpm, tpr, su, sm, c, phit, true);
    int ncase = c.ncase;
    if ((ncase > 2 && ncase < 7) || (ncase == 1 && su < 0.0) || (ncase == 7 && su < 0.0) ||
        (ncase == 2 && su > 0.0) || (ncase == 8 && su > 0.0))
      tpr = 0;
    if (ncase > 2 && ncase < 7) {
      if (su > 0.0)
        ub = uplus;
      else
        ub = 0.0;
    } else if (ncase == 1 || ncase == 7) {
      ub = c.u2;
      if (tpr == 1 && uf == ub) uf = uout;
    } else {
      ub = c.u3;
      if (tpr == 1 && uf == ub) uf = uout;
    }
    tpr1 = 0;
    double du = fsign(one, ub - uout) * su * ((ub - uout) + (two * tpr - one) * (ub - uf));
    du = du / double(nup);
    kmax = std::min((int)((ub - uout) / du + offset), nup);
    if (du == 0.0) kmax = 0;
    if (fsign(1.0, uout - ub) != fsign(1.0, uout - u0) || (uout == u0)) {
      if (kmax != 0) {
        k = 1;
        un = uout + (k - offset) * du;
        geomu(u0, un, mu0, mun, a, l, l2, q2, iu, tpm, tpr1, su, sm, c, phit, false);
        if (phit) geophitime(u0, un, mu0, mun, a, l, l2, q2, tpm, tpr1, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, true);
        ufi[k] = un; mufi[k] = mun; dti[k] = tmu + tu; dphi[k] = phimu + phiu; tpmi[k] = tpm; tpri[k] = 0;
        lambdai[k] = lambda;
      }
      for (k = 2; k <= kmax; k++) {
        un = uout + (k - offset) * du;
        geomu(u0, un, mu0, mun, a, l, l2, q2, iu, tpm, tpr1, su, sm, c, phit, false);
        if (phit)
          geophitime(u0, un, mu0, mun, a, l, l2, q2, tpm, tpr1, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, false);
        ufi[k] = un; mufi[k] = mun; dti[k] = tmu + tu; dphi[k] = phimu + phiu; tpmi[k] = tpm; tpri[k] = 0;
        lambdai[k] = lambda;
      }
      if (tpr == 1) {
        if (nup == 1) firstpt = true;
        if (mufill) {
          k = kmax + kext + 1;
          un = two * ub - (uout + (k - offset) * du);
          geomu(uf, un, mu0, mun, a, l, l2, q2, iu, tpm1, tpr, su, sm, c, phit, false);
          muf = mun;
          double muminus, muplus;
          findmroots(q2, l, a, mu0, muminus, muplus);
          for (k = 1; k <= npts; k++) {
            indep_muf(muminus, muplus, mufi[kmax - kext], muf, tpmi[kmax - kext], tpm1, sm, k, npts, offset, mun, tpm);
            geor(u0, un, mu0, mun, a, l, l2, q2, iu, tpm, tpr, su, sm, c, phit, false);
            if (phit)
              geophitime(u0, un, mu0, mun, a, l, l2, q2, tpm, tpr, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, false);
            int kk = k + kmax - kext;
            ufi[kk] = un; mufi[kk] = mun; dti[kk] = tmu + tu; dphi[kk] = phimu + phiu; tpmi[kk] = tpm; tpri[kk] = tpr;
            lambdai[kk] = lambda;
          }
          npts = npts - 2 * kext;
        }
        for (k = kmax + 1 + kext; k <= nup; k++) {
          un = two * ub - (uout + (k - offset) * du);
          geomu(uf, un, mu0, mun, a, l, l2, q2, iu, tpm, tpr, su, sm, c, phit, false);
          if (phit)
            geophitime(uf, un, mu0, mun, a, l, l2, q2, tpm, tpr, su, sm, iu, phimu, tmu, phiu, tu, lambda, c, firstpt);
          int kk = k + npts;
          ufi[kk] = un; mufi[kk] = mun; dti[kk] = tmu + tu; dphi[kk] = phimu + phiu; tpmi[kk] = tpm; tpri[kk] = 1;
          lambdai[kk] = lambda;
        }
        nup = nup + npts;
      }
    } else {
      nup = 0;
    }
  }
  ncase_out = c.ncase;
}

// ---------------------------------------------------------------- camera  geokerr_wrapper.f:1-354
void initialize_camera_geokerr(int standard, double a1, double a2, double b1, double b2, double rcut, int nrotype,
                               int nro, int nphi, int nup, double &u0, double uout, double mu0, double &a,
                               double &offset, double *ufarr, double *mufarr, double *alarr, double *bearr,
                               double *q2arr, double *larr, double *smarr, double *suarr, int *tpmarr, int *tprarr) {
  (void)uout;
  const double FAC = 100., INPUT_LIMIT = 1.e-5, ZERO = 0., ONE = 1., TWO = 2.;
  const double PI = std::acos(-ONE);
  offset = .5;
  int ngeo = nro * nphi;
  double abmax;
  std::vector<double> ro(nro + 1, 0.0);
  if (nrotype == 1) {
    abmax = rcut * rcut;
    double r1 = a1;
    for (int i = 1; i <= nro; i++) ro[i] = r1 * std::pow(rcut / r1, (double)((float)i / (float)nro));
  } else {
    abmax = std::pow(std::max(a1 * a1, a2 * a2), 2) + std::pow(std::max(b1 * b1, b2 * b2), 2);
  }
  double uplus = ONE / (ONE + std::sqrt(ONE - a * a));
  u0 = std::min(1.e-4, ONE / (FAC * abmax));
  if (nup == 1) offset = 1.e-8;
  for (int ii = 1; ii <= ngeo; ii++) {
    int i = (int)(double(ii - 1) / double(nphi)) + 1;
    int j = (ii - 1) % nphi + 1;
    double alpha, beta;
    if (nrotype > 1) {
      alpha = a1 + (a2 - a1) * (double(i - 1) + 0.5) / double(nro);
      beta = b1 + (b2 - b1) * (double(j - 1) + 0.5) / double(nphi);
    } else {
      double phi;
      if (nphi != 1)
        phi = TWO * PI * (double(j - 1) + 0.5) / double(nphi);
      else
        phi = 0.;
      alpha = ro[i] * std::cos(phi);
      beta = ro[i] * std::sin(phi);
    }
    double l = -alpha * std::sqrt(ONE - mu0 * mu0);
    double q2 = beta * beta - (a * a - alpha * alpha) * mu0 * mu0;
    double su = ONE, sm;
    int k = ii - 1;
    if (standard == 2) {
      if (std::fabs(q2) < INPUT_LIMIT * INPUT_LIMIT) q2 = ZERO;
      if (std::fabs(a) < INPUT_LIMIT) a = ZERO;
      if (std::fabs(l) < INPUT_LIMIT) l = ZERO;
      if (beta > ZERO && mu0 < ONE)
        sm = ONE;
      else
        sm = -ONE;
      int tpm = (int)((fsign(1., mu0) * sm + ONE) / TWO);
      if (nup == 1) offset = 0.;
      alarr[k] = alpha; bearr[k] = beta; q2arr[k] = q2; larr[k] = l; tpmarr[k] = tpm; smarr[k] = sm; suarr[k] = su;
      mufarr[k] = 0.;
    } else {
      double uf = uplus;
      int tpr = 1;
      if (mu0 < ONE && beta >= ZERO)
        sm = ONE;
      else
        sm = -ONE;
      if (std::fabs(q2) < INPUT_LIMIT * INPUT_LIMIT) q2 = ZERO;
      if (std::fabs(a) < INPUT_LIMIT) a = ZERO;
      if (std::fabs(l) < INPUT_LIMIT) l = ZERO;
      bearr[k] = beta; alarr[k] = alpha; q2arr[k] = q2; larr[k] = l; tprarr[k] = tpr; smarr[k] = sm; suarr[k] = su;
      ufarr[k] = uf;
      // (reference leaves TPMARR unset for standard=1: TPM=0 is a local only)
    }
  }
}

}  // namespace orc
