// grtrans_b200 -- semi-analytic Kerr null geodesics (Dexter & Agol 2009) restructured for the GPU.
//
// Replaces geokerr_wrapper.f:357-604 (GEOKERR) and what it calls (GEOMU :771, GEOR :2672,
// GEOPHITIME :1648, FINDMROOTS :607, INDEP_MUF :645, calcimu*/calcphitmu*/ellphitmu :2361-3274) plus
// geodesics.f90:271-377 (geokerr_wrapper).
//
// The reference walks one ray sequentially and uses FIRSTPT flags to cache per-ray integrals inside
// the same routines that evaluate each point.  Here the work is split in two device functions:
//
//   geo_setup()  one thread per RAY: classifies the quartic U(u) (Table 1 case, roots via Laguerre),
//                evaluates every integral that depends only on (u0, mu0, turning points) and the
//                ray's point layout (inbound leg / mu-fill / return leg) -> RayInv
//   geo_point()  one thread per (RAY, POINT): independent of all other points; evaluates the u- and
//                mu- integrals for its own end point, sharing the Carlson argument set between
//                I_u and the four t/phi/lambda integrals (grt_ellint.cuh).
//
// Per-ray constants that the reference recomputes per point (a second degree-6 root solve for
// case 6, the "valid solution" integral iut for cases 3/5/6) are hoisted into geo_setup().
// Equal-root (measure-zero) special branches of the reference are not special-cased.
#pragma once
#include "grt_ellint.cuh"

namespace grt {


struct RayConst {  // one pixel of the camera (geodesics.f90:12-21 geokerr_args, per-ray slice)
  double a, mu0, u0, uout, alpha, beta, q2, l, su, sm, uf, muf, offset, phi0, t0;
  int tpm, tpr;
};

struct GeoFlags {
  int nup, kext, next, mufill, usegeor, phit;
  // 1: leave out the p4 = -2 radial integral that only the coordinate time t_u needs (one R_J + R_C per point).  The images do
  // not depend on t unless several dumps are resident (slow light) or t0 offsets are computed; dt is then meaningless.
  int skip_t;
};

struct RayInv {
  // constants of motion and derived polynomial coefficients
  double a, l, l2, q2, su, sm, mu0, u0, dd, ee, ql2, uplus, umi, ur, qs, sdn;  // sdn = sqrt(dd) or sqrt|ee|
  int ncase, quartic;                                                          // quartic: ee != 0
  double u1, u2, u3, u4, h1;
  // case 5/6 constants (Table 2) and validity bounds
  double f5, g5, iut, c1j, c2j, c3j, c4j, c5j, m1j, nj;
  // cached first-point integrals
  double iu0, tu0c, phiu0c, tu01;
  double i1mu, i3mu, rffmu1, rffmu3, rdc, rjc, tmu1, tmu3, phimu1, phimu3;
  // mu roots (GEOMU/GEOPHITIME convention, clamped) ; musym: symmetric roots (M_- < 0)
  double mneg, mpos, muplus, muminus;
  int musym, mumode;  // mumode: 0 general biquadratic, 1 a==0, 2 q2==0, 3 none (ncase 4)
  // layout (GEOKERR)
  double ub, du, ufleg, offset, uout;
  int tpr, kmax, nkeep, npts, nfill, kext, usegeor, equatorial;
  int skip_t;  // GeoFlags::skip_t
  // mu-fill
  double fmu0, fmuf, fmuminus, fmuplus;
  int ftpm0, ftpmf;
  // standard=2
  double muf_target;
  int tpm_in;
};

struct GeoPoint {
  double u, mu, dt, dphi, lam;
  int tpm, tpr;
};

// ---- roots of M(mu) ------------------------------------------------------------------------------
GRT_HD void m_roots(double a, double q2, double ql2, double &mneg, double &mpos) {
  double d = a * a - ql2;
  double yy = -0.5 * (d + sgn(d) * sqrt(d * d + 4.0 * q2 * a * a));
  if (d < 0.0) {
    mneg = -yy / a / a;
    mpos = q2 / yy;
  } else {
    mneg = q2 / yy;
    mpos = -yy / a / a;
  }
}

GRT_HD double asech_(double x) { return x > 0.0 ? log((1.0 + sqrt(1.0 - x * x)) / x) : 0.0; }

// ---- u-integral families bound to a ray ----------------------------------------------------------
// Evaluate I_u between the ray's reference point and uf and (if PHIT) the Eq.(47)-(49) pieces.
// For turning-point cases (1,2,7,8) "piece" means the integral between uf and the turning point.
struct UPiece {
  double iu1;            // raw I_u piece (not yet divided by sqrt(dd)|sqrt|ee|)
  double t1, t2, t3, t4; // [-4;(0,1)], [-2;(0,1)], [-2;(1,-1/uplus)], [-2;(1,-umi)]
};

template <bool PHIT>
GRT_HDN void upiece_eval(const RayInv &c, double y, double x, int kind, double sgnv, UPiece &o) {
  // kind: 1 row1 (u<=u2), 2 row2 (u>=u3), 3 cubic complex, 5 quartic complex, 6 double complex, 7 row7, 8 row8
  const double iup = 1.0 / c.uplus;
  if (kind == 1 || kind == 2) {
    CubicReal F;
    if (kind == 1)
      F.prep(-c.u1, 1.0, c.u2, -1.0, c.u3, -1.0, y, x);
    else
      F.prep(-c.u1, 1.0, -c.u2, 1.0, -c.u3, 1.0, y, x);
    double r = F.rff();
    o.iu1 = sgnv * F.base(r);
    if (PHIT) {
      o.t1 = sgnv * F.m4(0.0, 1.0, r);
      const double pa[3] = {0.0, 1.0, 1.0}, pb[3] = {1.0, -iup, -c.umi};
      double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int q = c.skip_t; q < 3; q++) t[q] = sgnv * F.m2(pa[q], pb[q], r);  // one copy of m2 (code size, see header)
      o.t2 = t[0]; o.t3 = t[1]; o.t4 = t[2];
    }
  } else if (kind == 3) {
    CubicComplex F;
    double f = -1.0 / c.dd / c.u1, g = f / c.u1;
    F.prep(-c.u1, 1.0, f, g, 1.0, y, x);
    double r = F.rff();
    o.iu1 = sgnv * F.base(r);
    if (PHIT) {
      o.t1 = sgnv * F.m4(0.0, 1.0, r);
      const double pa[3] = {0.0, 1.0, 1.0}, pb[3] = {1.0, -iup, -c.umi};
      double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int q = c.skip_t; q < 3; q++) t[q] = sgnv * F.m2(pa[q], pb[q], r);  // one copy of m2 (code size, see header)
      o.t2 = t[0]; o.t3 = t[1]; o.t4 = t[2];
    }
  } else if (kind == 5) {
    QuarticComplex F;
    F.prep(-c.u1, 1.0, c.qs * c.u4, -c.qs, c.f5, c.g5, 1.0, y, x);
    double r = F.rff();
    o.iu1 = sgnv * F.base(r);
    if (PHIT) {
      o.t1 = sgnv * F.m4(0.0, 1.0, r);
      const double pa[3] = {0.0, 1.0, 1.0}, pb[3] = {1.0, -iup, -c.umi};
      double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int q = c.skip_t; q < 3; q++) t[q] = sgnv * F.m2(pa[q], pb[q], r);  // one copy of m2 (code size, see header)
      o.t2 = t[0]; o.t3 = t[1]; o.t4 = t[2];
    }
  } else if (kind == 6) {
    DoubleComplex F;
    double h2 = 1.0 / c.h1, g1 = c.dd / c.ee / (h2 - c.h1), f1 = 1.0 / sqrt(c.ee);
    F.prep(f1, g1, c.h1, f1, -g1, h2, y, x);
    double r = F.rff();
    o.iu1 = sgnv * F.base(r);
    if (PHIT) {
      o.t1 = sgnv * F.m4(0.0, 1.0, r);
      const double pa[3] = {0.0, 1.0, 1.0}, pb[3] = {1.0, -iup, -c.umi};
      double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int q = c.skip_t; q < 3; q++) t[q] = sgnv * F.m2(pa[q], pb[q], r);  // one copy of m2 (code size, see header)
      o.t2 = t[0]; o.t3 = t[1]; o.t4 = t[2];
    }
  } else {
    QuarticReal F;
    if (kind == 7)
      F.prep(-c.u1, 1.0, c.u2, -1.0, c.u3, -1.0, c.u4, -1.0, y, x);
    else
      F.prep(-c.u1, 1.0, -c.u2, 1.0, -c.u3, 1.0, c.u4, -1.0, y, x);
    double r = F.rff();
    o.iu1 = sgnv * F.base(r);
    if (PHIT) {
      o.t1 = sgnv * F.m4(0.0, 1.0, r);
      const double pa[3] = {0.0, 1.0, 1.0}, pb[3] = {1.0, -iup, -c.umi};
      double t[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
      for (int q = c.skip_t; q < 3; q++) t[q] = sgnv * F.m2(pa[q], pb[q], r);  // one copy of m2 (code size, see header)
      o.t2 = t[0]; o.t3 = t[1]; o.t4 = t[2];
    }
  }
}

// Eq. (47)/(48) coefficient groups
struct TPCoef {
  double ct1, ct2, ct3, ct4, cp1, cp2;
};
GRT_HD TPCoef tp_coef(const RayInv &c) {
  TPCoef k;
  double iup = 1.0 / c.uplus;
  k.ct1 = c.umi - iup;
  k.ct2 = c.umi * c.umi - iup * iup;
  k.ct3 = 2.0 * c.a * (c.a - c.l) + c.a * c.a * iup + iup * iup * iup;
  k.ct4 = 2.0 * c.a * (c.a - c.l) + c.a * c.a * c.umi + c.umi * c.umi * c.umi;
  k.cp1 = c.l * iup + 2.0 * (c.a - c.l);
  k.cp2 = c.l * c.umi + 2.0 * (c.a - c.l);
  return k;
}

// ---- mu pieces -----------------------------------------------------------------------------------
// Legendre pieces for one mu value (the non-FIRSTPT half of calcimu*f + calcphitmu* + ellphitmu)
struct MuPiece {
  double i2mu, tmuf, phimuf, sf;
};
template <bool PHIT>
GRT_HDN void mupiece_eval(const RayInv &c, double muf, MuPiece &o) {
  const double mneg = c.mneg, mpos = c.mpos;
  double fac, m1, n, fac2, sf, phif;
  if (c.musym) {
    fac = fabs(c.a) * sqrt(mpos - mneg);
    m1 = -mneg / (mpos - mneg);
    phif = acos(muf / c.muplus);
    sf = sin(phif);
    if (mpos != 1.0) {
      n = mpos / (1.0 - mpos);
      fac2 = 1.0 - mpos;
    } else {
      n = -1.0;
      fac2 = 1.0;
    }
  } else {
    fac = fabs(c.a) * c.muplus;
    m1 = mneg / mpos;
    sf = sqrt((mpos - muf * muf) / (mpos - mneg));
    phif = asin(sf);
    if (mpos != 1.0) {
      n = (mpos - mneg) / (1.0 - mpos);
      fac2 = 1.0 - mpos;
    } else {
      n = -1.0;
      fac2 = 1.0;
    }
  }
  double ak = sqrt(1.0 - m1);
  double qf = (1.0 - sf * ak) * (1.0 + sf * ak);
  double rff = rf(1.0 - sf * sf, qf, 1.0);
  double imuf = sf * rff;
  if (c.musym && tan(phif) < 0.0) imuf = c.i3mu * fac - imuf;
  o.i2mu = imuf / fac;
  o.sf = sf;
  if (PHIT) {
    double nn = -n;  // Numerical-Recipes sign convention passed to ellphitmu
    double rdf = rd(1.0 - sf * sf, qf, 1.0);
    double ellef = sf * (rff - sq(sf * ak) * rdf / 3.0);
    double ellpif = 0.0;
    if (nn != 1.0) {
      double rjf = rj(1.0 - sf * sf, qf, 1.0, 1.0 - nn * sf * sf);
      ellpif = sf * (rff + nn * sf * sf * rjf / 3.0);
    }
    if (phif > GRT_PI / 2.0) {
      double ellpic = 2.0 * (c.rffmu3 + nn * c.rjc / 3.0);
      double ellec = 2.0 * (c.rffmu3 - ak * ak * c.rdc / 3.0);
      ellpif = ellpic - ellpif;
      ellef = ellec - ellef;
    }
    o.tmuf = fac * ellef;
    o.phimuf = ellpif / fac2 / fac;
  }
}

// mu-side FIRSTPT quantities: i1mu, i3mu, rffmu1, rffmu3, then (phit) rdc, rjc, tmu1, tmu3, phimu1, phimu3
GRT_HD void mu_first(RayInv &c) {
  const double mneg = c.mneg, mpos = c.mpos, mu0 = c.mu0;
  double fac, m1, n, fac2, s0, phi0;
  if (c.musym) {
    fac = fabs(c.a) * sqrt(mpos - mneg);
    m1 = -mneg / (mpos - mneg);
    phi0 = acos(mu0 / c.muplus);
    s0 = sin(phi0);
    n = (mpos != 1.0) ? mpos / (1.0 - mpos) : -1.0;
  } else {
    fac = fabs(c.a) * c.muplus;
    m1 = mneg / mpos;
    s0 = sqrt((mpos - mu0 * mu0) / (mpos - mneg));
    phi0 = asin(s0);
    n = (mpos != 1.0) ? (mpos - mneg) / (1.0 - mpos) : -1.0;
  }
  fac2 = (mpos != 1.0) ? 1.0 - mpos : 1.0;
  double ak = sqrt(1.0 - m1);
  double q0 = (1.0 - s0 * ak) * (1.0 + s0 * ak);
  double rf0 = rf(1.0 - s0 * s0, q0, 1.0);
  double rfc = rf(0.0, m1, 1.0);
  double imum = 2.0 * rfc, imu0 = s0 * rf0;
  if (tan(phi0) < 0.0) imu0 = imum - imu0;
  c.i1mu = imu0 / fac;
  c.i3mu = c.musym ? imum / fac : imum / fac / 2.0;
  c.rffmu1 = rf0;
  c.rffmu3 = rfc;
  // t/phi pieces
  double nn = -n;
  double rd0 = rd(1.0 - s0 * s0, q0, 1.0);
  c.rdc = rd(0.0, m1, 1.0);
  double elle0 = s0 * (rf0 - sq(s0 * ak) * rd0 / 3.0);
  double ellec = 2.0 * (rfc - ak * ak * c.rdc / 3.0);
  double ellpi0 = 0.0, ellpic = 0.0;
  c.rjc = 0.0;
  if (nn != 1.0) {
    double rj0 = rj(1.0 - s0 * s0, q0, 1.0, 1.0 - nn * s0 * s0);
    c.rjc = rj(0.0, m1, 1.0, 1.0 - nn);
    ellpi0 = s0 * (rf0 + nn * s0 * s0 * rj0 / 3.0);
    ellpic = 2.0 * (rfc + nn * c.rjc / 3.0);
  }
  if (phi0 > GRT_PI / 2.0) {
    ellpi0 = ellpic - ellpi0;
    elle0 = ellec - elle0;
  }
  c.tmu1 = fac * elle0;
  c.phimu1 = ellpi0 / fac2 / fac;
  if (c.musym) {
    c.tmu3 = fac * ellec;
    c.phimu3 = ellpic / fac2 / fac;
  } else {
    c.tmu3 = fac * ellec / 2.0;
    c.phimu3 = ellpic / fac2 / fac / 2.0;
  }
}

// number of mu turning points for a given I_u (Eqs. 30,31,33)
GRT_HD int tpm_from_iu(double iu, double i1mu, double i3mu, double sm) {
  if (sm == 1.0) {
    double q = (iu - i1mu) / i3mu;
    return (int)q + (int)((1.0 + sgn(q)) / 2.0);
  }
  return (int)((iu + i1mu) / i3mu);
}

// mu_f and the number of mu turning points from I_u (mu half of GEOMU).  Returns muf.
GRT_HD double mu_of_iu(const RayInv &c, double iu, int &tpm) {
  double muf = 0.0;
  if (c.mumode == 0) {
    tpm = tpm_from_iu(iu, c.i1mu, c.i3mu, c.sm);
    double a2 = c.sm * pm1(tpm);
    double a3 = 2.0 * (int)((2.0 * (double)tpm + 3.0 - c.sm) / 4.0) - 1.0;
    double iarg = fabs(c.a) / a2 * (iu - c.sm * c.i1mu - a3 * c.i3mu);
    double sn, cn, dn;
    if (c.musym) {
      sncndn(sqrt(c.mpos - c.mneg) * iarg, -c.mneg / (c.mpos - c.mneg), sn, cn, dn);
      muf = c.muminus * cn;
    } else {
      sncndn(fabs(c.muplus) * iarg, c.mneg / c.mpos, sn, cn, dn);
      muf = c.muminus / dn;
    }
  } else if (c.mumode == 1) {  // a == 0 (geokerr_wrapper.f:948-967)
    tpm = tpm_from_iu(iu, c.i1mu, c.i3mu, c.sm);
    double a2 = c.sm * pm1(tpm);
    double a3 = 2.0 * (int)((2.0 * (double)tpm + 3.0 - c.sm) / 4.0) - 1.0;
    muf = -c.muplus * cos(sqrt(c.ql2) * (iu - c.sm * c.i1mu - a3 * c.i3mu) / a2);
  } else if (c.mumode == 2) {  // q2 == 0 (:938-947); tpm keeps its input value
    double s1 = sgn(c.mu0);
    if (c.l2 < c.a * c.a && c.mu0 != 0.0) {
      double mup = s1 * sqrt(1.0 - c.l2 / c.a / c.a);
      muf = mup / cosh(fabs(c.a * mup) * iu - s1 * c.sm * asech_(c.mu0 / mup));
    }
  }
  return muf;
}

// t_mu and phi_mu for an end point (muf, tpm) reached at I_u = I_mu = iu (mu half of GEOPHITIME :1924-2349).  One
// out-of-line copy serves the inbound / return legs (muf from mu_of_iu) and the mu-driven points (standard=2, mu-fill).
// mode2: evaluate the q2 == 0 closed forms (the mu-fill segment of the reference leaves them at 0).
GRT_HDN void mu_tphi(const RayInv &c, double muf, int tpm, double iu, bool mode2, double &tmu, double &phimu) {
  tmu = 0.0;
  phimu = 0.0;
  if (c.mumode == 0) {
    MuPiece mp;
    mupiece_eval<true>(c, muf, mp);
    double b1 = c.sm, b2 = c.sm * pm1(tpm + 1), b3 = 2.0 * (int)((2 * tpm - c.sm + 1) / 4.0);
    double tt = b1 * c.tmu1 + b2 * mp.tmuf + b3 * c.tmu3;
    tmu = c.musym ? c.a * c.a * c.mneg * iu + tt : tt;
    phimu = -c.l * iu + c.l * (b1 * c.phimu1 + b2 * mp.phimuf + b3 * c.phimu3);
  } else if (c.mumode == 1) {  // a == 0 (:1941-1966)
    double a2 = c.sm * pm1(tpm);
    double a3 = 2.0 * (int)((2.0 * (double)tpm + 3.0 - c.sm) / 4.0) - 1.0;
    double mup = sqrt(c.q2 / c.ql2);  // (no mu0 clamp in GEOPHITIME)
    double vfm = (muf - mup * mup) / (1.0 - muf) / mup;
    double vfp = -(muf + mup * mup) / (1.0 + muf) / mup;
    if (fabs(vfm) > 1.0) vfm = sgn(vfm);
    if (fabs(vfp) > 1.0) vfp = sgn(vfp);
    double vsm = ((c.mu0 - mup * mup) == 0.0) ? -1.0 : (c.mu0 - mup * mup) / (1.0 - c.mu0) / mup;
    double vsp = ((c.mu0 + mup * mup) == 0.0) ? -1.0 : -(c.mu0 + mup * mup) / (1.0 + c.mu0) / mup;
    if (fabs(vsp) > 1.0) vsp = sgn(vsp);
    if (fabs(vsm) > 1.0) vsm = sgn(vsm);
    double p1 = GRT_PI - asin(vsm) + asin(vsp);
    double p2 = asin(vfm) - asin(vfp) + GRT_PI;
    phimu = -c.l * iu + sgn(c.l) * 0.5 * (c.sm * p1 + a2 * p2 + a3 * 2.0 * GRT_PI);
  } else if (c.mumode == 2 && mode2) {  // q2 == 0 (:1924-1940)
    if (fabs(c.l) < fabs(c.a)) {
      double s1 = sgn(c.mu0);
      double b1 = s1 * c.sm, b2 = s1 * c.sm * pm1(tpm + 1);
      double mup = s1 * sqrt(1.0 - c.l2 / c.a / c.a);
      double sq1 = sqrt(1.0 - mup * mup), sg = sgn(mup * c.l / c.a);
      double t0 = tan(asin(c.mu0 / mup) / 2.0), tf = tan(asin(muf / mup) / 2.0);
      double p1 = sg * (atan((mup - t0) / sq1) + atan((mup + t0) / sq1));
      double p2 = sg * (atan((mup - tf) / sq1) + atan((mup + tf) / sq1));
      phimu = b1 * p1 + b2 * p2;
      tmu = fabs(mup * c.a) * (b2 * sqrt(1.0 - muf * muf / (mup * mup)) + b1 * sqrt(1.0 - c.mu0 * c.mu0 / (mup * mup)));
    }
  }
}

template <bool PHIT>
GRT_HD double mu_from_iu(const RayInv &c, double iu, int &tpm, double &tmu, double &phimu) {
  tmu = 0.0;
  phimu = 0.0;
  double muf = mu_of_iu(c, iu, tpm);
  if (PHIT) mu_tphi(c, muf, tpm, iu, true, tmu, phimu);
  return muf;
}

// I_mu for a given mu_f and tpm (mu half of GEOR).  ok=false: no valid solution (uf=-1).
GRT_HD double imu_from_mu(const RayInv &c, double muf, int tpm, bool &ok) {
  ok = true;
  if (c.mumode == 2) {
    if (c.l2 >= c.a * c.a || c.mu0 == 0.0) {
      ok = false;
      return 0.0;
    }
    double s1 = sgn(c.mu0);
    double mup = s1 * sqrt(1.0 - c.l2 / c.a / c.a);
    double i1 = 1.0 / fabs(c.a * mup) * asech_(c.mu0 / mup), i2 = 1.0 / fabs(c.a * mup) * asech_(muf / mup);
    return s1 * c.sm * i1 + s1 * c.sm * pm1(tpm + 1) * i2;
  }
  if (c.mumode == 1) {
    double sq = sqrt(c.ql2);
    double i1 = (0.5 * GRT_PI - asin(c.mu0 / c.muplus)) / sq, i3 = GRT_PI / sq;
    double a2 = c.sm * pm1(tpm), a3 = 2.0 * (int)((2.0 * (double)tpm + 3.0 - c.sm) / 4.0) - 1.0;
    double i2 = (0.5 * GRT_PI + asin(muf / c.muplus)) / sq;
    return c.sm * i1 + a2 * i2 + a3 * i3;
  }
  if (c.mumode == 3) return 0.0;
  if (!c.musym && fabs(muf) < sqrt(c.mneg)) {
    ok = false;
    return 0.0;
  }
  MuPiece mp;
  mupiece_eval<false>(c, muf, mp);
  double a2 = c.sm * pm1(tpm + 1), a3 = 2.0 * (int)((2 * tpm - c.sm + 1) / 4.0);
  return c.sm * c.i1mu + a2 * mp.i2mu + a3 * c.i3mu;
}

// ---- u side --------------------------------------------------------------------------------------
// Given the end point uf (already clamped) and tpr: I_u and, if PHIT, (t_u, phi_u, lambda_u).
template <bool PHIT>
GRT_HD double u_eval(const RayInv &c, double u0, double uf, int tpr, double &tu, double &phiu, double &lamu) {
  tu = 0.0;
  phiu = 0.0;
  lamu = 0.0;
  const int nc = c.ncase;
  if (nc == 4) {  // U(u) = 1
    if (PHIT) {
      double up = c.uplus, umi = c.umi;
      TPCoef k = tp_coef(c);
      phiu = c.su * c.ur * (c.l * log((uf / up - 1.0) / (u0 / up - 1.0)) - c.l * umi * umi * log((uf * umi - 1.0) / (u0 * umi - 1.0)));
      tu = c.su * c.ur *
           (k.ct1 * (1.0 / u0 - 1.0 / uf) + k.ct2 * log(uf / u0) +
            (c.a * c.a / up + 1.0 / (up * up * up)) * up * log((uf / up - 1.0) / (u0 / up - 1.0)) -
            (c.a * c.a * umi + umi * umi * umi) * umi * log((uf * umi - 1.0) / (u0 * umi - 1.0)));
      lamu = c.su * (1.0 / u0 - 1.0 / uf);
    }
    return c.su * (uf - u0);
  }
  UPiece p;
  p.iu1 = p.t1 = p.t2 = p.t3 = p.t4 = 0.0;
  if (nc == 1 || nc == 7) {
    const double tol = (nc == 1) ? 1e-16 : 1e-12;
    if (uf != c.u2) {
      if (PHIT && !(fabs(uf - c.u2) > tol))
        upiece_eval<false>(c, uf, c.u2, nc, 1.0, p);
      else
        upiece_eval<PHIT>(c, uf, c.u2, nc, 1.0, p);
    }
  } else if (nc == 2 || nc == 8) {
    if (uf != c.u3) upiece_eval<PHIT>(c, c.u3, uf, nc, -1.0, p);
  } else if (nc == 3 || nc == 5 || nc == 6) {
    if (uf > u0)
      upiece_eval<PHIT>(c, u0, uf, nc, 1.0, p);
    else if (uf < u0)
      upiece_eval<PHIT>(c, uf, u0, nc, -1.0, p);
  } else {
    return 0.0;
  }
  if (nc == 3 || nc == 5 || nc == 6) {
    if (PHIT) {
      TPCoef k = tp_coef(c);
      double sd = (nc == 3) ? sqrt(c.dd) : sqrt(-c.qs * c.ee);
      tu = c.su * c.ur / sd * (k.ct1 * p.t1 + k.ct2 * p.t2 - k.ct3 * p.t3 + k.ct4 * p.t4);
      phiu = c.su * c.ur / sd * (k.cp1 * (-p.t3) - k.cp2 * (-p.t4));
      lamu = c.su * p.t1 / ((nc == 6) ? sd : c.sdn);
    }
    return c.su * p.iu1 / c.sdn;
  }
  double s = pm1(tpr);
  if (PHIT) {
    TPCoef k = tp_coef(c);
    double tu1 = k.ct1 * p.t1 + k.ct2 * p.t2 - k.ct3 * p.t3 + k.ct4 * p.t4;
    double phiu1 = -k.cp1 * p.t3 + k.cp2 * p.t4;
    tu = c.su * (c.tu0c - s * tu1) / c.sdn * c.ur;
    phiu = c.su * (c.phiu0c - s * phiu1) / c.sdn * c.ur;
    lamu = c.su * (c.tu01 - s * p.t1) / c.sdn;
  }
  return c.su * (c.iu0 - s * p.iu1) / c.sdn;
}

// u_f from I_mu (u half of GEOR, Table 2).  Returns false when no valid solution exists.
GRT_HDN bool u_from_imu(const RayInv &c, double imu, double &uf, int &tpr) {
  double sn, cn, dn;
  switch (c.ncase) {
    case 1: {
      double arg = imu - c.su * c.iu0 / c.sdn;
      sncndn(sqrt((c.u3 - c.u1) * c.dd) * 0.5 * arg, (c.u3 - c.u2) / (c.u3 - c.u1), sn, cn, dn);
      uf = c.u1 + (c.u2 - c.u1) * (cn * cn / (dn * dn));
      tpr = (int)((sgn(c.su * arg) + 1.0) / 2.0);
      return true;
    }
    case 2: {
      double arg = imu + c.su * c.iu0 / c.sdn;
      sncndn(sqrt((c.u3 - c.u1) * c.dd) * 0.5 * arg, (c.u3 - c.u2) / (c.u3 - c.u1), sn, cn, dn);
      uf = c.u1 + (c.u3 - c.u1) * (dn * dn / (cn * cn));
      tpr = (int)((-sgn(c.su * (imu + c.iu0 / c.sdn)) + 1.0) / 2.0);
      return true;
    }
    case 3: {
      tpr = 0;
      if (imu > c.iut) return false;
      sncndn(c.c1j * (imu + c.iu0), c.m1j, sn, cn, dn);
      uf = (c.c2j + c.u1 - (c.c2j - c.u1) * cn) / (1.0 + cn);
      return true;
    }
    case 5: {
      tpr = 0;
      if (imu > c.iut) return false;
      sncndn(c.c1j * (imu + c.iu0), c.m1j, sn, cn, dn);
      double qs = c.qs;
      uf = (c.u4 * c.c5j + qs * c.u1 * c.c4j - (qs * c.u4 * c.c5j - c.u1 * c.c4j) * cn) /
           ((c.c4j - qs * c.c5j) * cn + (qs * c.c4j + c.c5j));
      return true;
    }
    case 6: {
      tpr = 0;
      if (imu > c.iut) return false;
      sncndn(c.c1j * (c.su * imu + c.iu0), c.m1j, sn, cn, dn);
      double sc = sn / cn;
      uf = c.c3j + (c.nj * (1.0 + c.c2j * c.c2j) * sc) / (1.0 - c.c2j * sc);
      return true;
    }
    case 7: {
      double arg = imu - c.su * c.iu0 / sqrt(-c.ee);
      sncndn(sqrt(fabs(c.ee) * (c.u3 - c.u1) * (c.u4 - c.u2)) * 0.5 * arg,
             (c.u4 - c.u1) * (c.u3 - c.u2) / ((c.u4 - c.u2) * (c.u3 - c.u1)), sn, cn, dn);
      uf = ((c.u2 - c.u1) * c.u3 * sn * sn - c.u2 * (c.u3 - c.u1)) / ((c.u2 - c.u1) * sn * sn - (c.u3 - c.u1));
      tpr = (int)((sgn(c.su * (imu - c.su * c.iu0 / c.sdn)) + 1.0) / 2.0);
      return true;
    }
    case 8: {
      double arg = imu + c.su * c.iu0 / sqrt(-c.ee);
      sncndn(sqrt(fabs(c.ee) * (c.u3 - c.u1) * (c.u4 - c.u2)) * 0.5 * arg,
             (c.u4 - c.u1) * (c.u3 - c.u2) / ((c.u4 - c.u2) * (c.u3 - c.u1)), sn, cn, dn);
      uf = ((c.u4 - c.u3) * c.u2 * sn * sn - c.u3 * (c.u4 - c.u2)) / ((c.u4 - c.u3) * sn * sn - (c.u4 - c.u2));
      tpr = (int)((-sgn(c.su * arg) + 1.0) / 2.0);
      return true;
    }
    default:
      tpr = 0;
      return false;
  }
}

// ---- classification of U(u) (roots + Table 1 case) ------------------------------------------------
// Mutates u0 like the reference's "Unphysical ... Input modified" retry.  geor_mode selects the
// replacement values GEOR uses (u3/u2) instead of GEOMU's (0/uplus) in the quartic branch.
GRT_HDN void classify(RayInv &c, double &u0, bool geor_mode) {
  const double a = c.a, l = c.l, q2 = c.q2;
  double cc = a * a - q2 - c.l2;
  c.dd = 2.0 * (sq(a - l) + q2);
  c.ee = -a * a * q2;
  c.quartic = 0;
  c.u1 = c.u2 = c.u3 = c.u4 = c.h1 = 0.0;
  if (c.ee == 0.0 && c.dd != 0.0) {
    double dd = c.dd;
    double qq = cc * cc / dd / dd / 9.0;
    double rr = (2.0 * cc * cc * cc / (dd * dd * dd) + 27.0 / dd) / 54.0;
    double dis = rr * rr - qq * qq * qq;
    if (dis < -1.e-16 || fabs(dis) < 1.e-16) {
      double theta = (dis < -1.e-16) ? acos(rr / pow(qq, 1.5)) : 0.0;
      c.u1 = -2.0 * sqrt(qq) * cos(theta / 3.0) - cc / dd / 3.0;
      c.u2 = -2.0 * sqrt(qq) * cos((theta - 2.0 * GRT_PI) / 3.0) - cc / dd / 3.0;
      c.u3 = -2.0 * sqrt(qq) * cos((theta + 2.0 * GRT_PI) / 3.0) - cc / dd / 3.0;
      for (int it = 0; it < 2; it++) {
        if (u0 <= c.u2) {
          c.ncase = 1;
          break;
        } else if (u0 >= c.u3) {
          c.ncase = 2;
          break;
        }
        u0 = (c.su == 1.0) ? c.u3 : c.u2;
        c.ncase = (c.su == 1.0) ? 2 : 1;
      }
    } else {
      c.ncase = 3;
      double aa = -sgn(rr) * pow(fabs(rr) + sqrt(dis), 0.3333333333333333);
      double bb = (aa != 0.0) ? qq / aa : 0.0;
      c.u1 = (aa + bb) - cc / dd / 3.0;
    }
    c.sdn = sqrt(dd);
  } else if (c.ee == 0.0 && c.dd == 0.0) {
    c.ncase = 4;
    c.sdn = 1.0;
  } else {
    c.quartic = 1;
    c.sdn = sqrt(fabs(c.ee));
    double coef[5] = {1.0, 0.0, cc, c.dd, c.ee};  // U(u) = 1 + cc u^2 + dd u^3 + ee u^4, constant term first
    cpx root[4];
    zroots(coef, 4, root);
    int nreal = 0;
    for (int i = 0; i < 4; i++)
      if (root[i].im == 0.0) nreal++;
    if (nreal == 2) {
      c.ncase = 5;
      c.u1 = root[0].re;
      c.u4 = (root[1].im == 0.0) ? root[1].re : root[3].re;
    } else if (nreal == 0) {
      c.ncase = 6;
    } else {
      c.u1 = root[0].re; c.u2 = root[1].re; c.u3 = root[2].re; c.u4 = root[3].re;
      for (int it = 0; it < 2; it++) {
        if (c.u2 > c.uplus && c.u3 > c.uplus) {
          c.ncase = 5;
          break;
        } else if (u0 <= c.u2) {
          c.ncase = 7;
          break;
        } else if (u0 >= c.u3) {
          c.ncase = 8;
          break;
        }
        if (geor_mode)
          u0 = (c.su == 1.0) ? c.u3 : c.u2;
        else
          u0 = (c.su == 1.0) ? 0.0 : c.uplus;
      }
    }
    if (c.ncase == 5) {
      c.f5 = -c.qs / fabs(c.ee) / c.u1 / c.u4;
      c.g5 = (c.u4 + c.u1) / c.u1 / c.u4 * c.f5;
    } else if (c.ncase == 6) {
      double se = sqrt(c.ee);
      double co[7] = {1.0, -cc / se, -1.0, se * (2.0 * cc / c.ee - sq(c.dd / c.ee)), -1.0, -cc / se, 1.0};
      // (palindromic: the reference lists coefs(1..7) with coefs(1) the constant term)
      cpx hr[6];
      zroots(co, 6, hr);
      c.h1 = 0.0;
      for (int i = 0; i < 6 && c.h1 == 0.0; i++)
        if (hr[i].im == 0.0) c.h1 = hr[i].re;
    }
  }
}

// Table-2 constants for cases 3,5,6 (needed only when GEOR is used on this ray)
GRT_HDN void geor_consts(RayInv &c, double u0) {
  const double su = c.su;
  UPiece p;
  if (c.ncase == 3) {
    if (su > 0.0)
      upiece_eval<false>(c, u0, c.uplus, 3, 1.0, p);
    else
      upiece_eval<false>(c, 0.0, u0, 3, 1.0, p);
    c.iut = p.iu1 / c.sdn;
    upiece_eval<false>(c, c.u1, u0, 3, 1.0, p);
    c.iu0 = su * p.iu1 / c.sdn;
    double c3 = -1.0;
    if (c.a != 0.0 || c.l != 0.0) c3 = (c.a + c.l) / (c.a - c.l);
    c.c2j = sqrt(c.u1 * (3.0 * c.u1 + c3));
    c.c1j = sqrt(c.c2j * c.dd);
    c.m1j = 0.5 + (6.0 * c.u1 + c3) / (8.0 * c.c2j);
  } else if (c.ncase == 5) {
    if (su > 0.0)
      upiece_eval<false>(c, u0, c.uplus, 5, 1.0, p);
    else
      upiece_eval<false>(c, 0.0, u0, 5, 1.0, p);
    c.iut = p.iu1 / c.sdn;
    double ub = (c.qs == 1.0) ? c.u1 : c.u4;
    double m = -c.g5 * 0.5, n2 = c.f5 - c.g5 * c.g5 / 4.0;
    c.c4j = sqrt(sq(m - c.u4) + n2);
    c.c5j = sqrt(sq(m - c.u1) + n2);
    c.c1j = sqrt(fabs(c.ee * c.c4j * c.c5j));
    upiece_eval<false>(c, ub, u0, 5, 1.0, p);
    c.iu0 = su * p.iu1 / c.sdn;
    c.m1j = c.qs * (sq(c.c4j + c.qs * c.c5j) - sq(c.u4 - c.u1)) / (4.0 * c.c4j * c.c5j);
  } else if (c.ncase == 6) {
    if (su > 0.0)
      upiece_eval<false>(c, u0, c.uplus, 6, 1.0, p);
    else
      upiece_eval<false>(c, 0.0, u0, 6, 1.0, p);
    c.iut = p.iu1 / c.sdn;
    double ee = c.ee, dd = c.dd, cc = c.a * c.a - c.q2 - c.l2;
    double co[7] = {1.0 / (ee * ee * ee), -cc / (ee * ee * ee), -1.0 / (ee * ee), -1.0 / (ee * ee) * (dd * dd / ee - 2.0 * cc),
                    -1.0 / ee, -cc / ee, 1.0};
    cpx hr[6];
    zroots(co, 6, hr);
    double mn2 = 0.0;
    for (int i = 0; i < 6 && mn2 == 0.0; i++)
      if (hr[i].im == 0.0) mn2 = hr[i].re;
    double pp = dd / (2.0 * ee * ee * (mn2 * mn2 - 1.0 / ee));
    double m = -0.5 * dd / ee - pp;
    if (m < pp) {
      double t = pp;
      pp = m;
      m = t;
    }
    double pr2 = 1.0 / (mn2 * ee);
    double n = sqrt(mn2 - m * m), r = sqrt(pr2 - pp * pp);
    c.c4j = sqrt(sq(m - pp) + sq(n + r));
    c.c5j = sqrt(sq(m - pp) + sq(n - r));
    c.c1j = (c.c4j + c.c5j) * 0.5 * sqrt(fabs(ee));
    c.c2j = sqrt((4.0 * n * n - sq(c.c4j - c.c5j)) / (sq(c.c4j + c.c5j) - 4.0 * n * n));
    c.c3j = m + c.c2j * n;
    c.nj = n;
    upiece_eval<false>(c, c.c3j, u0, 6, 1.0, p);
    c.iu0 = sgn(u0 - c.c3j) * p.iu1 / c.sdn;
    c.m1j = sq((c.c4j - c.c5j) / (c.c4j + c.c5j));
  }
}

// ---- INDEP_MUF (geokerr_wrapper.f:645-706) ---------------------------------------------------------
GRT_HD void indep_muf(double muminus, double muplus, double mu0, double muf, int tpm0, int tpmf, double sm, int k,
                      int npts, double offset, double &mun, int &tpmk) {
  if (muf <= muplus && muf >= muminus) {
    int dtpm = tpmf - tpm0;
    int a1 = (int)(sm * pm1(tpm0));
    int a2 = (int)(sm * pm1(tpmf));
    int a3 = (int)(2 * (int)floorf((2.f * (float)dtpm + 3.f - (float)a1) / 4.f) - 1.f);
    double dmu = a1 * (muplus - mu0) + a2 * (muf - muminus) + a3 * (muplus - muminus);
    double w = (double)(((float)a1 + 1.f) / 2.f), wm = (double)((1.f - (float)a1) / 2.f);
    double mu1 = w * muplus + wm * muminus;
    double mu2 = wm * muplus + w * muminus;
    double dmu1 = fabs(mu1 - mu0), dmu2 = fabs(mu2 - mu1);
    double dmuk = (k - offset) * dmu / (double)(float)npts;
    int d1 = (int)(dmuk / dmu1);
    if (d1 < 0) d1 = 1;
    int t2 = (int)((dmuk - dmu1) / dmu2);
    tpmk = (d1 < 1 ? d1 : 1) + (t2 > 0 ? t2 : 0);
    int a1k = a1;
    int a2k = (int)(sm * pm1(tpmk + tpm0));
    int a3k = (int)(2.f * (float)(int)((2.f * (float)tpmk + 3.f - (float)a1k) / 4.f) - 1.f);
    mun = muminus + (double)(1.f / (float)a2k) * (dmuk - a1k * (muplus - mu0) - a3k * (muplus - muminus));
  } else {
    mun = 0.0;
    tpmk = 0;
  }
  tpmk += tpm0;
}

// ---- per-ray setup -------------------------------------------------------------------------------
GRT_HDN void geo_setup(const RayConst &rc, const GeoFlags &fl, RayInv &c) {
  c.a = rc.a; c.l = rc.l; c.l2 = rc.l * rc.l; c.q2 = rc.q2; c.su = rc.su; c.sm = rc.sm; c.mu0 = rc.mu0;
  c.ql2 = c.q2 + c.l2;
  c.uplus = 1.0 / (1.0 + sqrt(1.0 - c.a * c.a));
  c.umi = 1.0 - sqrt(1.0 - c.a * c.a);
  c.ur = -1.0 / (2.0 * sqrt(1.0 - c.a * c.a));
  c.qs = sgn(c.q2);
  c.offset = rc.offset; c.uout = rc.uout; c.kext = fl.kext; c.usegeor = fl.usegeor; c.skip_t = fl.skip_t;
  c.iu0 = c.tu0c = c.phiu0c = c.tu01 = 0.0;
  c.i1mu = c.i3mu = c.rffmu1 = c.rffmu3 = c.rdc = c.rjc = c.tmu1 = c.tmu3 = c.phimu1 = c.phimu3 = 0.0;
  c.f5 = c.g5 = c.iut = c.c1j = c.c2j = c.c3j = c.c4j = c.c5j = c.m1j = c.nj = 0.0;
  c.fmu0 = c.fmuf = c.fmuminus = c.fmuplus = 0.0;
  c.ftpm0 = c.ftpmf = 0; c.nfill = 0; c.npts = 0; c.kmax = 0; c.nkeep = 0; c.tpr = 0;
  c.ub = c.du = c.ufleg = 0.0; c.muf_target = rc.muf; c.tpm_in = rc.tpm;
  c.equatorial = (rc.mu0 == 0.0 && rc.beta == 0.0) ? 1 : 0;
  double u0 = rc.u0;
  classify(c, u0, fl.usegeor != 0);
  // mu roots
  c.mneg = c.mpos = c.muplus = c.muminus = 0.0;
  c.musym = 1;
  if (c.ncase == 4) {
    c.mumode = 3;
  } else if (c.q2 == 0.0) {
    c.mumode = 2;
  } else if (c.a == 0.0) {
    c.mumode = 1;
    c.muplus = sqrt(c.q2 / c.ql2);
    if (c.mu0 > c.muplus) c.muplus = c.mu0;
    c.i1mu = acos(c.mu0 / c.muplus) / sqrt(c.ql2);
    c.i3mu = GRT_PI / sqrt(c.ql2);
  } else {
    c.mumode = 0;
    m_roots(c.a, c.q2, c.ql2, c.mneg, c.mpos);
    if (c.mpos > 1.0) c.mpos = 1.0;
    c.muplus = sqrt(c.mpos);
    if (c.mneg < 0.0) {
      c.musym = 1;
      if (c.muplus < c.mu0) {
        c.muplus = c.mu0;
        c.mpos = c.mu0 * c.mu0;
      }
      c.muminus = -c.muplus;
    } else {
      c.musym = 0;
      c.muplus = sgn(c.mu0) * c.muplus;
      c.muminus = sgn(c.mu0) * sqrt(c.mneg);
      if (fabs(c.muminus) > fabs(c.mu0)) {
        c.muminus = c.mu0;
        c.mneg = c.mu0 * c.mu0;
      }
      if (fabs(c.mu0) > fabs(c.muplus)) {
        c.muplus = c.mu0;
        c.mpos = c.mu0 * c.mu0;
      }
    }
    mu_first(c);
  }
  // u0 pieces for the turning-point cases
  const int nc = c.ncase;
  if (nc == 1 || nc == 7) {
    if (u0 != c.u2) {
      UPiece p;
      upiece_eval<true>(c, u0, c.u2, nc, 1.0, p);
      TPCoef k = tp_coef(c);
      c.iu0 = p.iu1;
      c.tu01 = p.t1;
      c.tu0c = k.ct1 * p.t1 + k.ct2 * p.t2 - k.ct3 * p.t3 + k.ct4 * p.t4;
      c.phiu0c = -k.cp1 * p.t3 + k.cp2 * p.t4;
    }
  } else if (nc == 2 || nc == 8) {
    if (u0 != c.u3) {
      UPiece p;
      upiece_eval<true>(c, c.u3, u0, nc, -1.0, p);
      TPCoef k = tp_coef(c);
      c.iu0 = fl.usegeor ? -p.iu1 : p.iu1;  // GEOR stores +integral, GEOMU stores -integral
      c.tu01 = p.t1;
      c.tu0c = k.ct1 * p.t1 + k.ct2 * p.t2 - k.ct3 * p.t3 + k.ct4 * p.t4;
      c.phiu0c = -k.cp1 * p.t3 + k.cp2 * p.t4;
    }
  }
  c.u0 = u0;
  if (fl.usegeor) {
    geor_consts(c, u0);
    c.npts = fl.nup;
    c.nkeep = fl.nup;
    return;
  }
  // ---- GEOKERR layout for standard=1 (geokerr_wrapper.f:463-598)
  double uf = rc.uf;
  int tpr = rc.tpr;
  if (nc == 1 || nc == 7) {
    if (uf > c.u2) uf = c.u2;
  } else if (nc == 2 || nc == 8) {
    if (uf < c.u3) uf = c.u3;
  }
  if ((nc > 2 && nc < 7) || ((nc == 1 || nc == 7) && c.su < 0.0) || ((nc == 2 || nc == 8) && c.su > 0.0)) tpr = 0;
  double ub;
  if (nc > 2 && nc < 7) {
    ub = (c.su > 0.0) ? c.uplus : 0.0;
  } else if (nc == 1 || nc == 7) {
    ub = c.u2;
    if (tpr == 1 && uf == ub) uf = rc.uout;
  } else {
    ub = c.u3;
    if (tpr == 1 && uf == ub) uf = rc.uout;
  }
  double du = sgn(ub - rc.uout) * c.su * ((ub - rc.uout) + (2.0 * tpr - 1.0) * (ub - uf));
  du = du / (double)fl.nup;
  int kmax = (int)((ub - rc.uout) / du + rc.offset);
  if (kmax > fl.nup) kmax = fl.nup;
  if (du == 0.0) kmax = 0;
  c.ub = ub; c.du = du; c.ufleg = uf; c.tpr = tpr; c.kmax = kmax;
  if (!(sgn(rc.uout - ub) != sgn(rc.uout - rc.u0) || (rc.uout == rc.u0))) {
    c.npts = 0;  // wrong-side UOUT: NUP=0
    return;
  }
  if (tpr == 1) {
    if (fl.mufill) {
      c.nfill = fl.next;
      c.nkeep = kmax - fl.kext;
      c.npts = fl.nup + fl.next - 2 * fl.kext;
      // end points of the mu-fill segment
      double tu, ph, lm, tmu, phm;
      int tpm1 = 0, tpm0 = 0;
      int k = kmax + fl.kext + 1;
      double un = 2.0 * ub - (rc.uout + (k - rc.offset) * du);
      if (un > c.u2 && (nc == 1 || nc == 7)) un = c.u2;
      double iu = u_eval<false>(c, uf, un, 1, tu, ph, lm);
      c.fmuf = mu_from_iu<false>(c, iu, tpm1, tmu, phm);
      c.ftpmf = tpm1;
      k = kmax - fl.kext;
      un = rc.uout + (k - rc.offset) * du;
      iu = u_eval<false>(c, u0, un, 0, tu, ph, lm);
      c.fmu0 = mu_from_iu<false>(c, iu, tpm0, tmu, phm);
      c.ftpm0 = tpm0;
      // FINDMROOTS
      double mneg, mpos;
      m_roots(c.a, c.q2, c.ql2, mneg, mpos);
      mpos = dmin(mpos, 1.0);
      if (mneg > 0.0) {
        c.fmuplus = sgn(c.mu0) * sqrt(mpos);
        c.fmuminus = sgn(c.mu0) * sqrt(mneg);
        if (fabs(c.fmuminus) > fabs(c.mu0)) c.fmuminus = c.mu0;
        if (fabs(c.mu0) > fabs(c.fmuplus)) c.fmuplus = c.mu0;
      } else {
        c.fmuplus = sqrt(mpos);
        if (c.fmuplus < c.mu0) c.fmuplus = c.mu0;
        c.fmuminus = -c.fmuplus;
      }
    } else {
      c.nkeep = kmax;
      c.npts = fl.nup;
    }
  } else {
    c.nkeep = kmax;
    c.npts = fl.nup;  // (slots beyond kmax stay unset in the reference; kmax==nup for non-turning rays)
  }
}

// ---- one point -----------------------------------------------------------------------------------
// slot is the 0-based output slot of the ray.  Returns false for slots the reference leaves unset.
// The four kinds of point (standard=2 disk-plane point, inbound leg, mu-fill, return leg) only differ in how the end point
// is prescribed -- in u (legs) or in mu (standard=2, mu-fill) -- so they share ONE evaluation sequence
//   [mu-driven: I_mu(mu_f) -> u_f]  ->  u integrals  ->  [u-driven: I_u -> mu_f]  ->  mu integrals
// with a single call site per stage: the kernel's code has to stay small enough for the instruction cache.//
// GRT_PHASE_BARRIER (__syncthreads() in grt_k_geo.cu, which defines GRT_USE_PHASE_BARRIER) re-aligns the warps of a thread block between the stages:
// the stages together are ~100 KB of straight-line code, far more than the 32 KB L1.5 instruction cache, and warps that
// drift apart (the Carlson loops have data-dependent trip counts) each stream the whole footprint from L2 on their own
// (ncu: 64% of the stall samples were "no instruction").  With the barriers every warp of the SM is inside the same
// <= 40 KB stage at the same time.  All threads of the block must therefore call geo_point (valid = false for padding).
#if defined(GRT_USE_PHASE_BARRIER) && defined(__CUDA_ARCH__)
#define GRT_PHASE_BARRIER() __syncthreads()
#else
#define GRT_PHASE_BARRIER() ((void)0)
#endif
GRT_HD bool geo_point(const RayInv &c, int nup, int slot, bool valid, GeoPoint &o) {
  o.u = o.mu = o.dt = o.dphi = o.lam = 0.0;
  o.tpm = o.tpr = 0;
  double tu = 0, phiu = 0, lamu = 0, tmu = 0, phimu = 0;
  const int nc = c.ncase;
  int tpm = 0, tpr = 0;
  double un = -1.0, mun = 0.0, ustart = c.u0, ipar = 0.0;
  bool mu_driven = false;
  bool active = valid;  // false: nothing left to compute for this thread (result already final)
  bool ret = valid;
  if (!active) {
  } else if (c.usegeor) {
    // standard=2: GEOR at equal steps in mu (geokerr_wrapper.f:421-461)
    int k = slot + 1;
    mun = c.mu0 + (k - c.offset) * (c.muf_target - c.mu0) / (double)nup;
    tpm = c.tpm_in;
    if (c.equatorial) {  // IF(MU0.NE.0.OR.BETA.NE.0) ... ELSE UF=U0
      o.u = c.u0; o.mu = mun; o.tpm = tpm;
      active = false;
    }
    mu_driven = true;
  } else if (slot < c.nkeep) {  // inbound leg
    int k = slot + 1;
    un = c.uout + (k - c.offset) * c.du;
    if ((nc == 1 || nc == 7) && un > c.u2) un = c.u2;
    if ((nc == 2 || nc == 8) && un < c.u3) un = c.u3;
  } else if (c.tpr == 1 && slot < c.nkeep + c.nfill) {  // mu-fill near the radial turning point
    int k = slot - c.nkeep + 1;
    indep_muf(c.fmuminus, c.fmuplus, c.fmu0, c.fmuf, c.ftpm0, c.ftpmf, c.sm, k, c.nfill, c.offset, mun, tpm);
    mu_driven = true;
  } else if (c.tpr == 1) {  // return leg
    int k = slot + 1 - (c.nfill - 2 * (c.nfill ? c.kext : 0));
    un = 2.0 * c.ub - (c.uout + (k - c.offset) * c.du);
    if ((nc == 1 || nc == 7) && un > c.u2) un = c.u2;
    tpr = 1;
    ustart = c.ufleg;
  } else {
    active = false;
    ret = false;
  }
  if (active && mu_driven) {
    bool ok;
    ipar = imu_from_mu(c, mun, tpm, ok);
    if (ok) ok = u_from_imu(c, ipar, un, tpr);  // (false for ncase 4: no valid solution, uf = -1)
    if (!ok) {
      o.u = -1.0; o.mu = mun; o.tpm = tpm; o.tpr = 0;
      active = false;
    }
  }
  GRT_PHASE_BARRIER();
  double iu = 0.0;
  if (active) iu = u_eval<true>(c, ustart, un, tpr, tu, phiu, lamu);
  GRT_PHASE_BARRIER();
  if (active && !mu_driven) {
    ipar = iu;
    mun = mu_of_iu(c, iu, tpm);
  }
  GRT_PHASE_BARRIER();
  if (active) {
    mu_tphi(c, mun, tpm, ipar, !mu_driven || c.usegeor, tmu, phimu);
    if (c.l == 0.0) phiu -= sgn(c.a) * GRT_PI * tpm;
    o.u = un; o.mu = mun; o.dt = tmu + tu; o.dphi = phimu + phiu; o.lam = lamu + tmu; o.tpm = tpm; o.tpr = tpr;
  }
  return ret;
}

}  // namespace grt
